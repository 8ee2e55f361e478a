// Fused read-out losses on solver output: firing rate of the read-out populations, the
// script-level reduction and the loss, together with the gradient w.r.t. the trajectory, in ONE pass over the selected
// components -- instead of the ~10 elementwise ATen kernels (and as many (T, B, C) temporaries) autograd needs for
//   rate = compute_firing_rate(V - A); loss = smooth_l1_loss(sum_k w_k rate_k, target)
// (reference src/utils.py:74-88 huber_loss_wta; the C4 benchmark applies it to the L2/3e population of every column).
// HBM-bound by construction: reads y_sel once, writes grad_y_sel once.
#include "odecol_common.cuh"

namespace odecol {

// y_sel rows are [V of the G*P read-out populations | A of the same populations]; thread = one (row, group).
__global__ void __launch_bounds__(256) k_huber_rate_loss(const float* __restrict__ y, long long rows, int B, int G, int P,
                                                         const float* __restrict__ w, const float* __restrict__ target,
                                                         long long st_t, long long st_b, long long st_g, float beta,
                                                         float inv_count, float* __restrict__ grad, double* __restrict__ acc) {
    const long long total = rows * G;
    const int GP = G * P;
    double local = 0.0;
    // (row, g) of this thread's first element by one division; when the grid stride is a multiple of G (it is for the
    // launcher's grid whenever G divides 256) g never changes and row advances by a constant -- no division per element
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long e0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long row = e0 / G;
    int g = (int)(e0 - row * G);
    const bool fixed_g = stride % G == 0;
    const long long row_step = stride / G;
    const bool need_tb = (st_t | st_b) != 0;
    for (long long e = e0; e < total; e += stride) {
        if (e != e0) {
            if (fixed_g) row += row_step;
            else { row = e / G; g = (int)(e - row * G); }
        }
        const float* yr = y + row * 2 * GP;
        float* gr = grad + row * 2 * GP;
        float pred = 0.f, drs[8];
        for (int k = 0; k < P; ++k) {
            float r, dr;
            phi_dphi(__fsub_rn(__ldg(yr + g * P + k), __ldg(yr + GP + g * P + k)), r, dr);
            if (k < 8) drs[k] = dr;
            pred = __fadd_rn(pred, w ? __fmul_rn(r, __ldg(w + k)) : r);
        }
        long long toff = (long long)g * st_g;
        if (need_tb) { const long long t = row / B, b = row - t * B; toff += t * st_t + b * st_b; }
        const float d = __fsub_rn(pred, __ldg(target + toff));
        const float ad = fabsf(d);
        float dl;
        if (ad < beta) { local += 0.5 * (double)d * (double)d / (double)beta; dl = d / beta; }
        else { local += (double)ad - 0.5 * (double)beta; dl = d > 0.f ? 1.f : -1.f; }
        dl *= inv_count;
        for (int k = 0; k < P; ++k) {
            float dr;
            if (k < 8) dr = drs[k];
            else { float r; phi_dphi(__fsub_rn(__ldg(yr + g * P + k), __ldg(yr + GP + g * P + k)), r, dr); }
            const float gv = dl * dr * (w ? __ldg(w + k) : 1.f);
            gr[g * P + k] = gv;
            gr[GP + g * P + k] = -gv;
        }
    }
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
        atomicAdd(acc, s);
    }
}

__global__ void k_loss_finalize(const double* __restrict__ acc, double inv_count, float* __restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *loss = (float)(*acc * inv_count);
}

int launch_huber_rate_loss(const float* y_sel, int T, int B, int G, int P, const float* w, const float* target,
                           long long st_t, long long st_b, long long st_g, float beta, float* loss, float* grad,
                           double* acc, cudaStream_t s) {
    const long long rows = (long long)T * B;
    const double count = (double)rows * G;
    if (cudaMemsetAsync(acc, 0, sizeof(double), s) != cudaSuccess) return ODECOL_E_CUDA;
    const long long want = (rows * G + 255) / 256;
    const int blocks = (int)(want < 148LL * 16 ? (want < 1 ? 1 : want) : 148LL * 16);      // 16 resident CTAs on each of 148 SMs
    k_huber_rate_loss<<<blocks, 256, 0, s>>>(y_sel, rows, B, G, P, w, target, st_t, st_b, st_g, beta, (float)(1.0 / count), grad, acc);
    k_loss_finalize<<<1, 32, 0, s>>>(acc, 1.0 / count, loss);
    count_launch(2);
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------
// Window read-out of the XOR and parity tasks (reference scripts/xor_ode.py:120-130, scripts/parity_ode.py:239-249):
//     pred_b = sum_k w_k * mean_{t in the last L grid points} phi(V_k(t, b) - A_k(t, b))        k over the P read-out populations
//     loss   = mean_b | pred_b - target_b |
// XOR: L = 1 (final point), the 8 populations of column C, w = ff_source_mask; parity: L = 100, the output column,
// w = output_weights / output_scale.  One CTA per trial: the window's rates and phi' stay in shared memory between the
// reduction and the gradient, so y_sel is read once; grad_y_sel is zero outside the window (cleared by the launcher).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_window_rate_l1(const float* __restrict__ y, int T, int B, int P, int L,
                                                        const float* __restrict__ w, const float* __restrict__ target,
                                                        float* __restrict__ pred, float* __restrict__ grad,
                                                        float* __restrict__ grad_w, double* __restrict__ acc) {
    extern __shared__ float dph[];                      // [L * P] phi' of the window, then [P] window-mean rates
    __shared__ float red[4];
    const int b = blockIdx.x, n = L * P;
    float* rk = dph + n;
    const float inv_l = 1.0f / (float)L;
    for (int k = threadIdx.x; k < P; k += blockDim.x) rk[k] = 0.f;
    __syncthreads();
    float part = 0.f;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int t = T - L + e / P, k = e % P;
        const float* yr = y + ((size_t)t * B + b) * 2 * P;
        float r, dr;
        phi_dphi(__fsub_rn(__ldg(yr + k), __ldg(yr + P + k)), r, dr);
        const float wk = w ? __ldg(w + k) : 1.f;
        dph[e] = dr * wk * inv_l;
        part += r * wk * inv_l;
        atomicAdd(rk + k, r * inv_l);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    const float pr = red[0] + red[1] + red[2] + red[3];
    const float d = pr - __ldg(target + b);
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    if (threadIdx.x == 0) {
        pred[b] = pr;
        atomicAdd(acc, (double)fabsf(d));
    }
    const float scale = sgn / (float)B;
    for (int k = threadIdx.x; k < P; k += blockDim.x) grad_w[(size_t)b * P + k] = scale * rk[k];     // d loss / d w_k, this trial's share
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int t = T - L + e / P, k = e % P;
        float* gr = grad + ((size_t)t * B + b) * 2 * P;
        const float gv = scale * dph[e];
        gr[k] = gv;
        gr[P + k] = -gv;
    }
}

int launch_window_rate_l1_loss(const float* y_sel, int T, int B, int P, int L, const float* w, const float* target,
                               float* loss, float* pred, float* grad, float* grad_w, double* acc, cudaStream_t s) {
    if (cudaMemsetAsync(acc, 0, sizeof(double), s) != cudaSuccess) return ODECOL_E_CUDA;
    if (cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)T * B * 2 * P, s) != cudaSuccess) return ODECOL_E_CUDA;
    const size_t smem = sizeof(float) * ((size_t)L * P + P);
    if (smem > 200 * 1024) return ODECOL_E_UNSUPPORTED;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_window_rate_l1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_window_rate_l1<<<B, 128, smem, s>>>(y_sel, T, B, P, L, w, target, pred, grad, grad_w, acc);
    k_loss_finalize<<<1, 32, 0, s>>>(acc, 1.0 / (double)B, loss);
    count_launch(2);
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

}  // namespace odecol
