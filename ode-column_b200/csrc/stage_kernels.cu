// Kernel family L ("staged"): N too large for one CTA's registers.  The state of every trial lives in HBM/L2
// and each Runge-Kutta stage is ONE fused launch:
//
//     C[i][b] = sum_k W_aug[i][k] * R_aug[b][k]            (128 x 128 x 16 FFMA tiles, register-prefetched)
//     epilogue: drift -> k_s -> next stage state -> phi -> R_aug of the next stage (written K-major so that it is
//               directly the B operand of the next launch), stimulus columns refreshed by the first row of CTAs
//
// so the only HBM traffic is the per-element RK bookkeeping (about 46 floats per population-step) and
// W_aug / R_aug stream from L2.  The reverse sweep reuses the same contraction core with W^T for J^T lambda and
// a split-K (over trials) contraction with atomics for dW_aug.
//
// Replaces the Python stepping loop of torchdiffeq's rk4 around ColumnNetwork.forward for networks beyond the
// reference's sizes (BASELINE.json configs 4 and 5); same arithmetic as family S.
#include "stage_common.cuh"

namespace odecol {

// ---------------------------------------------------------------------------------------------------------------
// forward stage epilogues
// ---------------------------------------------------------------------------------------------------------------

template <int S>
__global__ void __launch_bounds__(kGemmThreads, 2) k_fwd_stage(FwdStageArgs a) {
    __shared__ __align__(16) GemmSmem sm;
    const int i0 = blockIdx.x * TM, b0 = blockIdx.y * TN;
    float acc[8][8];
    gemm_nt_core(a.Wp + (size_t)i0 * a.KPa, a.KPa, a.Ra_cur + (size_t)b0 * a.KPa, a.KPa, a.KPa, sm, acc);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int N = a.p.N, B = a.p.B;
    const float t0 = __ldg(a.t + a.n), t1 = __ldg(a.t + a.n + 1);
    const float dt = __fsub_rn(t1, t0);
#pragma unroll
    for (int jh = 0; jh < 2; ++jh)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int b = b0 + jh * 64 + 4 * ty + jj;
            if (b >= B) continue;
#pragma unroll
            for (int ih = 0; ih < 2; ++ih) {
                const int i = i0 + ih * 64 + 4 * tx;
                if (i >= N) continue;
                const float tot[4] = {acc[ih * 4 + 0][jh * 4 + jj], acc[ih * 4 + 1][jh * 4 + jj], acc[ih * 4 + 2][jh * 4 + jj],
                                      acc[ih * 4 + 3][jh * 4 + jj]};
                fwd_stage_epilogue<S, 4>(a, i, b, tot, dt);
            }
        }
    // stimulus columns of the next stage's operand: written by the first row of CTAs
    if (blockIdx.x == 0 && a.p.n_in > 0) {
        const float tn = S == 1 ? __fadd_rn(t0, __fmul_rn(dt, kOneThirdL)) : S == 2 ? __fadd_rn(t0, __fmul_rn(dt, kTwoThirdsL)) : t1;
        int idx = 1;
        const float tc = knot_locate(a.p.knot_t, a.p.K, tn, idx);
        const int n_in = a.p.n_in;
        for (int e = tid; e < TN * n_in; e += kGemmThreads) {
            const int bl = e / n_in, ch = e % n_in, b = b0 + bl;
            if (b < B)
                a.Ra_nxt[(size_t)b * a.KPa + N + ch] = knot_value(a.p.knot_t, a.p.knot_u + (size_t)b * a.p.knot_stride_b, n_in, idx, tc, ch);
        }
    }
}

// operand initialisation: Wp (padded copy), Ra[0] = r_aug(t[0], y0), constant-one column in both buffers
__global__ void k_pad_weights(const float* __restrict__ W_aug, int N, int ld_w, int Kaug, float* __restrict__ Wp, int Np, int KPa) {
    const size_t total = (size_t)Np * KPa;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / KPa), k = (int)(e % KPa);
        Wp[e] = (i < N && k < Kaug) ? W_aug[(size_t)i * ld_w + k] : 0.0f;
    }
}

__global__ void k_pad_transpose(const float* __restrict__ W_aug, int N, int ld_w, float* __restrict__ WT, int Np, int NPk) {
    const size_t total = (size_t)Np * NPk;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(e / NPk), i = (int)(e % NPk);
        WT[e] = (i < N && j < N) ? W_aug[(size_t)i * ld_w + j] : 0.0f;
    }
}

__global__ void k_init_operand(DevProblem p, const float* __restrict__ y, const float* __restrict__ t_ptr,
                               float* __restrict__ Ra0, float* __restrict__ Ra1, float* __restrict__ DR, int KPa, int Bp) {
    // one CTA per trial row (padding rows become zero); stimulus taken at t_ptr[0]
    const int b = blockIdx.x, N = p.N, Kaug = N + p.n_in + 1;
    const float tq = __ldg(t_ptr);
    float* r0 = Ra0 + (size_t)b * KPa;
    float* r1 = Ra1 ? Ra1 + (size_t)b * KPa : nullptr;
    if (b >= p.B) {
        for (int k = threadIdx.x; k < KPa; k += blockDim.x) { r0[k] = 0.f; if (r1) r1[k] = 0.f; }
        return;
    }
    const float* yb = y + (size_t)b * 3 * N;
    int idx = 1;
    const float tc = knot_locate(p.knot_t, p.K, tq, idx);
    const float* ku = p.knot_u + (size_t)b * p.knot_stride_b;
    for (int k = threadIdx.x; k < KPa; k += blockDim.x) {
        float v = 0.f, v1 = 0.f;
        if (k < N) {
            if (DR) { float d; phi_dphi(__fsub_rn(yb[k], yb[N + k]), v, d); DR[(size_t)b * N + k] = d; }
            else v = phi(__fsub_rn(yb[k], yb[N + k]));
        } else if (k < N + p.n_in) v = knot_value(p.knot_t, ku, p.n_in, idx, tc, k - N);
        else if (k == Kaug - 1) { v = 1.f; v1 = 1.f; }
        r0[k] = v;
        if (r1) r1[k] = v1;
    }
}

void launch_pad_weights(const float* W_aug, int N, int ld_w, int Kaug, float* Wp, int Np, int KPa, cudaStream_t s) {
    k_pad_weights<<<296, 256, 0, s>>>(W_aug, N, ld_w, Kaug, Wp, Np, KPa);
    count_launch();
}
void launch_pad_transpose(const float* W_aug, int N, int ld_w, float* WT, int Np, int NPk, cudaStream_t s) {
    k_pad_transpose<<<296, 256, 0, s>>>(W_aug, N, ld_w, WT, Np, NPk);
    count_launch();
}
void launch_init_operand(const DevProblem& p, const float* y, const float* t_ptr, float* Ra, float* Ra_extra, float* DR,
                         int KPa, int Bp, cudaStream_t s) {
    k_init_operand<<<Bp, 128, 0, s>>>(p, y, t_ptr, Ra, Ra_extra, DR, KPa, Bp);
    count_launch();
}

__global__ void k_copy(const float* __restrict__ src, float* __restrict__ dst, size_t n4) {
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n4; e += (size_t)gridDim.x * blockDim.x)
        reinterpret_cast<float4*>(dst)[e] = reinterpret_cast<const float4*>(src)[e];
}

void launch_fwd_stage(int S, const FwdStageArgs& a, dim3 grid, cudaStream_t s) {
    switch (S) {
        case 1: k_fwd_stage<1><<<grid, kGemmThreads, 0, s>>>(a); break;
        case 2: k_fwd_stage<2><<<grid, kGemmThreads, 0, s>>>(a); break;
        case 3: k_fwd_stage<3><<<grid, kGemmThreads, 0, s>>>(a); break;
        default: k_fwd_stage<4><<<grid, kGemmThreads, 0, s>>>(a); break;
    }
    count_launch();
}

// ---------------------------------------------------------------------------------------------------------------
// forward driver
// ---------------------------------------------------------------------------------------------------------------
struct FwdLayout {
    int Np, Bp, KPa;
    size_t off_Wp, off_Ra0, off_Ra1, off_k1, off_k2, off_k3, off_ya, off_yb, total;
};

static FwdLayout fwd_layout(const DevProblem& p) {
    FwdLayout L;
    const int Kaug = p.N + p.n_in + 1;
    L.Np = round_up(p.N, TM);
    L.Bp = round_up(p.B, TN);
    L.KPa = round_up(Kaug, TK);
    size_t o = 0;
    auto take = [&](size_t floats) { const size_t r = o; o += (floats * sizeof(float) + 255) / 256 * 256; return r; };
    L.off_Wp = take((size_t)L.Np * L.KPa);
    L.off_Ra0 = take((size_t)L.Bp * L.KPa);
    L.off_Ra1 = take((size_t)L.Bp * L.KPa);
    const size_t st = (size_t)p.B * 3 * p.N;
    L.off_k1 = take(st); L.off_k2 = take(st); L.off_k3 = take(st);
    L.off_ya = take(st); L.off_yb = take(st);
    L.total = o;
    return L;
}

size_t stage_rk4_fwd_workspace_bytes(const DevProblem& p, int) { return fwd_layout(p).total; }

int stage_rk4_fwd(const DevProblem& p, const float* t_dev, int T, const float* y0, float* y_out, int out_every,
                  void* ws, size_t ws_bytes, cudaStream_t s) {
    const FwdLayout L = fwd_layout(p);
    if (!ws || ws_bytes < L.total) return ODECOL_E_WORKSPACE;
    if (p.N % 4 != 0) return ODECOL_E_UNSUPPORTED;
    char* w = static_cast<char*>(ws);
    float* Wp = reinterpret_cast<float*>(w + L.off_Wp);
    float* Ra[2] = {reinterpret_cast<float*>(w + L.off_Ra0), reinterpret_cast<float*>(w + L.off_Ra1)};
    float* k1 = reinterpret_cast<float*>(w + L.off_k1);
    float* k2 = reinterpret_cast<float*>(w + L.off_k2);
    float* k3 = reinterpret_cast<float*>(w + L.off_k3);
    float* ybuf[2] = {reinterpret_cast<float*>(w + L.off_ya), reinterpret_cast<float*>(w + L.off_yb)};
    const int Kaug = p.N + p.n_in + 1;
    const size_t st = (size_t)p.B * 3 * p.N;

    launch_pad_weights(p.W_aug, p.N, p.ld_w, Kaug, Wp, L.Np, L.KPa, s);
    launch_init_operand(p, y0, t_dev, Ra[0], Ra[1], nullptr, L.KPa, L.Bp, s);
    k_copy<<<296, 256, 0, s>>>(y0, y_out, st / 4);
    count_launch();
    const dim3 grid(L.Np / TM, L.Bp / TN);
    const float* ycur = y0;
    int cur = 0;
    for (int n = 0; n < T - 1; ++n) {
        const int j = n + 1;
        const bool emit = (j % out_every == 0) || (j == T - 1);
        const size_t r = (j % out_every == 0) ? (size_t)(j / out_every) : (size_t)((T - 2) / out_every + 1);
        float* ynext = emit ? y_out + r * st : ybuf[n & 1];
        FwdStageArgs a;
        a.p = p; a.Wp = Wp; a.y0 = ycur; a.k1 = k1; a.k2 = k2; a.k3 = k3; a.y1 = ynext; a.y_out_row = nullptr; a.DR_nxt = nullptr; a.Ra_cur_lo = nullptr; a.Ra_nxt_lo = nullptr;
        a.t = t_dev; a.n = n; a.KPa = L.KPa;
        a.Ra_cur = Ra[cur]; a.Ra_nxt = Ra[cur ^ 1];
        for (int S = 1; S <= 4; ++S) {
            launch_fwd_stage(S, a, grid, s);
            cur ^= 1; a.Ra_cur = Ra[cur]; a.Ra_nxt = Ra[cur ^ 1];
        }
        ycur = ynext;
    }
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

// backward and Euler-Maruyama drivers of family L live in stage_bwd.cu / stage_em.cu

}  // namespace odecol
