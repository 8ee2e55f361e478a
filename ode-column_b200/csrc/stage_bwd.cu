// Family L reverse sweep: exact discrete adjoint of the staged RK4 (3/8 rule) solve, state in global memory.
//
// Per step n (from T-2 down to 0), all on the caller's stream:
//   1. recompute   r_aug and phi' of the four stage states from y_traj[n]      (1 elementwise + 3 fused stage launches)
//   2. reverse     Ybar_s = J_s^T kbar_s for s = 4,3,2 then 1: one contraction with W^T per stage,
//                  g[j][b] = sum_i W[i][j] * (gamma * kbar_V)[b][i], epilogue applies phi', the diagonal terms
//                  and forms the next kbar (the Butcher-tableau transposes of the 3/8 rule)
//   3. dW_aug     += sum over the four stages and all trials of (gamma kbar_V) (x) r_aug: a split-K (over trials)
//                  contraction whose tiles are reduced with float atomics
// What it replaces: autograd's replay of every op of every stage of every step (loss.backward() through torchdiffeq,
// reference scripts/xor_ode.py:177, scripts/parity_ode.py:250).
#include "stage_common.cuh"

namespace odecol {

struct BwdStageArgs {
    DevProblem p;
    const float* WT;       // [Np][NPk]   W^T, zero padded
    const float* AV_cur;   // [Bp][NPk]   gamma * kbar_V of this stage (contraction operand)
    float* AV_nxt;         // [Bp][NPk]   operand of the next reverse stage
    float* acur;           // [B][3N]     kbar of this stage (in), of the next reverse stage (out)
    float* lam;            // [B][3N]     adjoint of y_{n+1} (stage 1 writes the adjoint of y_n)
    float* b4;             // [B][3N]     Ybar_4, later Ybar_4 + Ybar_3 + Ybar_2
    float* b3;             // [B][3N]
    const float* DR;       // [B][N]      phi'(x_s)
    const float* grad_y;   // (T, B, G)
    const int* inv;        // [3N] component -> column of grad_y or -1
    const float* t;
    int n, G, NPk;
    float gamma, inv_tau_m, inv_tau_a, inv_tau_s;
};

ODECOL_DEVINL float gather_grad(const BwdStageArgs& a, int n, int b, int comp) {
    const int g = a.inv[comp];
    return g >= 0 ? a.grad_y[((size_t)n * a.p.B + b) * a.G + g] : 0.f;
}

// S = 4, 3, 2, 1: the stage whose Jacobian is applied
template <int S>
ODECOL_DEVINL void bwd_stage_epilogue(const BwdStageArgs& a, int j, int b, const float (&graw)[4], float dt) {
    const int N = a.p.N;
    const size_t base = (size_t)b * 3 * N + j;
    const C4 aV = ldc(a.acur + base), aA = ldc(a.acur + base + N), aF = ldc(a.acur + base + 2 * N);
    const C4 dr = ldc(a.DR + (size_t)b * N + j);
    const C4 kap = ldc(a.p.kappa + j);
    const C4 lV = ldc(a.lam + base), lA = ldc(a.lam + base + N), lF = ldc(a.lam + base + 2 * N);
    C4 p4V, p4A, p4F, p3V, p3A, p3F;
    if (S <= 3) { p4V = ldc(a.b4 + base); p4A = ldc(a.b4 + base + N); p4F = ldc(a.b4 + base + 2 * N); }
    if (S == 2) { p3V = ldc(a.b3 + base); p3A = ldc(a.b3 + base + N); p3F = ldc(a.b3 + base + 2 * N); }
    const float h8 = dt * 0.125f, h38 = 3.0f * h8, h3 = dt * kOneThirdL;
    C4 nV, nA, nF, sV, sA, sF, op;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float g = graw[e] + kap.v[e] * aA.v[e] * a.inv_tau_a + aF.v[e] * a.inv_tau_s;
        const float bV = -aV.v[e] * a.inv_tau_m + dr.v[e] * g;
        const float bA = -aA.v[e] * a.inv_tau_a - dr.v[e] * g;
        const float bF = -aF.v[e] * a.inv_tau_s;
        if (S == 4) {           // store Ybar_4; kbar_3 = 3h/8 lam + h Ybar_4
            sV.v[e] = bV; sA.v[e] = bA; sF.v[e] = bF;
            nV.v[e] = h38 * lV.v[e] + dt * bV; nA.v[e] = h38 * lA.v[e] + dt * bA; nF.v[e] = h38 * lF.v[e] + dt * bF;
        }
        if (S == 3) {           // store Ybar_3; kbar_2 = 3h/8 lam - h Ybar_4 + h Ybar_3
            sV.v[e] = bV; sA.v[e] = bA; sF.v[e] = bF;
            nV.v[e] = h38 * lV.v[e] - dt * p4V.v[e] + dt * bV;
            nA.v[e] = h38 * lA.v[e] - dt * p4A.v[e] + dt * bA;
            nF.v[e] = h38 * lF.v[e] - dt * p4F.v[e] + dt * bF;
        }
        if (S == 2) {           // kbar_1 = h/8 lam + h Ybar_4 - h/3 Ybar_3 + h/3 Ybar_2; keep the running sum in b4
            nV.v[e] = h8 * lV.v[e] + dt * p4V.v[e] - h3 * p3V.v[e] + h3 * bV;
            nA.v[e] = h8 * lA.v[e] + dt * p4A.v[e] - h3 * p3A.v[e] + h3 * bA;
            nF.v[e] = h8 * lF.v[e] + dt * p4F.v[e] - h3 * p3F.v[e] + h3 * bF;
            sV.v[e] = p4V.v[e] + p3V.v[e] + bV; sA.v[e] = p4A.v[e] + p3A.v[e] + bA; sF.v[e] = p4F.v[e] + p3F.v[e] + bF;
        }
        if (S == 1) {           // adjoint of y_n, plus dL/dy_out[n]; then kbar_4 of the previous step
            const float LV = lV.v[e] + p4V.v[e] + bV + gather_grad(a, a.n, b, j + e);
            const float LA = lA.v[e] + p4A.v[e] + bA + gather_grad(a, a.n, b, N + j + e);
            const float LF = lF.v[e] + p4F.v[e] + bF + gather_grad(a, a.n, b, 2 * N + j + e);
            sV.v[e] = LV; sA.v[e] = LA; sF.v[e] = LF;
            float h8p = 0.f;
            if (a.n > 0) h8p = __fsub_rn(__ldg(a.t + a.n), __ldg(a.t + a.n - 1)) * 0.125f;
            nV.v[e] = h8p * LV; nA.v[e] = h8p * LA; nF.v[e] = h8p * LF;
        }
        op.v[e] = a.gamma * nV.v[e];
    }
    if (S == 4) { stc(a.b4 + base, sV); stc(a.b4 + base + N, sA); stc(a.b4 + base + 2 * N, sF); }
    if (S == 3) { stc(a.b3 + base, sV); stc(a.b3 + base + N, sA); stc(a.b3 + base + 2 * N, sF); }
    if (S == 2) { stc(a.b4 + base, sV); stc(a.b4 + base + N, sA); stc(a.b4 + base + 2 * N, sF); }
    if (S == 1) { stc(a.lam + base, sV); stc(a.lam + base + N, sA); stc(a.lam + base + 2 * N, sF); }
    stc(a.acur + base, nV); stc(a.acur + base + N, nA); stc(a.acur + base + 2 * N, nF);
    stc(a.AV_nxt + (size_t)b * a.NPk + j, op);
}

template <int S>
__global__ void __launch_bounds__(kGemmThreads, 2) k_bwd_stage(BwdStageArgs a) {
    __shared__ __align__(16) GemmSmem sm;
    const int j0 = blockIdx.x * TM, b0 = blockIdx.y * TN;
    float acc[8][8];
    gemm_nt_core(a.WT + (size_t)j0 * a.NPk, a.NPk, a.AV_cur + (size_t)b0 * a.NPk, a.NPk, a.NPk, sm, acc);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int N = a.p.N, B = a.p.B;
    const float dt = __fsub_rn(__ldg(a.t + a.n + 1), __ldg(a.t + a.n));
#pragma unroll
    for (int jh = 0; jh < 2; ++jh)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int b = b0 + jh * 64 + 4 * ty + jj;
            if (b >= B) continue;
#pragma unroll
            for (int ih = 0; ih < 2; ++ih) {
                const int j = j0 + ih * 64 + 4 * tx;
                if (j >= N) continue;
                const float g4[4] = {acc[ih * 4 + 0][jh * 4 + jj], acc[ih * 4 + 1][jh * 4 + jj], acc[ih * 4 + 2][jh * 4 + jj],
                                     acc[ih * 4 + 3][jh * 4 + jj]};
                bwd_stage_epilogue<S>(a, j, b, g4, dt);
            }
        }
}

// lam = dL/dy_out[T-1]; kbar_4 of the last step
__global__ void k_bwd_begin(DevProblem p, const float* __restrict__ grad_y, const int* __restrict__ inv, int G,
                            const float* __restrict__ t, int T, float gamma, float* __restrict__ lam,
                            float* __restrict__ acur, float* __restrict__ AV3, int NPk) {
    const int N = p.N;
    const size_t total = (size_t)p.B * 3 * N;
    const float h8 = __fsub_rn(__ldg(t + T - 1), __ldg(t + T - 2)) * 0.125f;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(e / (3 * N)), comp = (int)(e % (3 * N));
        const int g = inv[comp];
        const float L = g >= 0 ? grad_y[((size_t)(T - 1) * p.B + b) * G + g] : 0.f;
        lam[e] = L;
        acur[e] = h8 * L;
        if (comp < N) AV3[(size_t)b * NPk + comp] = gamma * h8 * L;
    }
}

__global__ void k_build_inv(const int* __restrict__ sel, int G, int n3, int* __restrict__ inv) {
    for (int e = threadIdx.x; e < n3; e += blockDim.x) inv[e] = sel ? -1 : e;
    __syncthreads();
    if (sel) for (int g = threadIdx.x; g < G; g += blockDim.x) inv[sel[g]] = g;
}

// dW_aug[i][k] += sum_s sum_b AV_s[b][i] * Ra_s[b][k]; grid (Np/128, ceil(KPa/128), splits); trials split over z
struct DwArgs {
    const float* AV[4];
    const float* Ra[4];
    float* grad_W;
    int N, Kaug, ld_w, NPk, KPa, B, Bp, rows_per_split;
};

__global__ void __launch_bounds__(kGemmThreads, 2) k_dw_accum(DwArgs a) {
    __shared__ __align__(16) GemmSmem sm;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.x * TM, k0 = blockIdx.y * TN;
    const int bs = blockIdx.z * a.rows_per_split, be = min(a.Bp, bs + a.rows_per_split);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    // loader: tile of 16 trial rows x 128 columns = 512 float4, two per thread; row = f / 32, c4 = f % 32
    const int lr0 = tid >> 5, lc = (tid & 31) * 4;
    for (int s = 0; s < 4; ++s) {
        const float* Ag = a.AV[s];
        const float* Bg = a.Ra[s];
        for (int b = bs; b < be; b += TK) {
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int r = lr0 + 8 * m;
                const float4 va = ld4(Ag + (size_t)(b + r) * a.NPk + i0 + lc);
                float4 vb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + lc < a.KPa) vb = ld4(Bg + (size_t)(b + r) * a.KPa + k0 + lc);
                st4(&sm.A[0][r][lc], va);
                st4(&sm.B[0][r][lc], vb);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                const float4 a0 = ld4(&sm.A[0][k][4 * tx]), a1 = ld4(&sm.A[0][k][64 + 4 * tx]);
                const float4 b0 = ld4(&sm.B[0][k][4 * ty]), b1 = ld4(&sm.B[0][k][64 + 4 * ty]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
        const int i = i0 + (ii >> 2) * 64 + 4 * tx + (ii & 3);
        if (i >= a.N) continue;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int k = k0 + (jj >> 2) * 64 + 4 * ty + (jj & 3);
            if (k < a.Kaug) atomicAdd(a.grad_W + (size_t)i * a.ld_w + k, acc[ii][jj]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct BwdLayout {
    int Np, Bp, KPa, NPk;
    size_t off_Wp, off_WT, off_Ra[4], off_DR[4], off_k[3], off_AV[4], off_lam, off_b4, off_b3, off_acur, off_inv, total;
};

static BwdLayout bwd_layout(const DevProblem& p) {
    BwdLayout L;
    const int Kaug = p.N + p.n_in + 1;
    L.Np = round_up(p.N, TM);
    L.Bp = round_up(p.B, TN);
    L.KPa = round_up(Kaug, TK);
    L.NPk = L.Np;                                   // K of the W^T contraction; also the row stride of AV (tile loads of 128)
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
    L.off_Wp = take(sizeof(float) * (size_t)L.Np * L.KPa);
    L.off_WT = take(sizeof(float) * (size_t)L.Np * L.NPk);
    for (int s = 0; s < 4; ++s) L.off_Ra[s] = take(sizeof(float) * (size_t)L.Bp * L.KPa);
    for (int s = 0; s < 4; ++s) L.off_DR[s] = take(sizeof(float) * (size_t)p.B * p.N);
    const size_t st = sizeof(float) * (size_t)p.B * 3 * p.N;
    for (int s = 0; s < 3; ++s) L.off_k[s] = take(st);
    for (int s = 0; s < 4; ++s) L.off_AV[s] = take(sizeof(float) * (size_t)L.Bp * L.NPk);
    L.off_lam = take(st); L.off_b4 = take(st); L.off_b3 = take(st); L.off_acur = take(st);
    L.off_inv = take(sizeof(int) * (size_t)3 * p.N);
    L.total = o;
    return L;
}

size_t stage_rk4_bwd_workspace_bytes(const DevProblem& p, int) { return bwd_layout(p).total; }

int stage_rk4_bwd(const DevProblem& p, const float* t_dev, int T, const float* y_traj, const float* grad_y,
                  const int* sel, int G, float* grad_y0, float* grad_W, void* ws, size_t ws_bytes, cudaStream_t s) {
    const BwdLayout L = bwd_layout(p);
    if (!ws || ws_bytes < L.total) return ODECOL_E_WORKSPACE;
    if (p.N % 4 != 0) return ODECOL_E_UNSUPPORTED;
    char* w = static_cast<char*>(ws);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(w + off); };
    float* Wp = F(L.off_Wp);
    float* WT = F(L.off_WT);
    float *Ra[4], *DR[4], *AV[4], *kk[3];
    for (int i = 0; i < 4; ++i) { Ra[i] = F(L.off_Ra[i]); DR[i] = F(L.off_DR[i]); AV[i] = F(L.off_AV[i]); }
    for (int i = 0; i < 3; ++i) kk[i] = F(L.off_k[i]);
    float *lam = F(L.off_lam), *b4 = F(L.off_b4), *b3 = F(L.off_b3), *acur = F(L.off_acur);
    int* inv = reinterpret_cast<int*>(w + L.off_inv);
    const int Kaug = p.N + p.n_in + 1;
    const size_t st = (size_t)p.B * 3 * p.N;
    const float gamma = p.c.tau_s * p.c.R / p.c.tau_m;

    if (cudaMemsetAsync(w + L.off_AV[0], 0, L.off_lam - L.off_AV[0], s) != cudaSuccess) return ODECOL_E_CUDA;
    launch_pad_weights(p.W_aug, p.N, p.ld_w, Kaug, Wp, L.Np, L.KPa, s);
    launch_pad_transpose(p.W_aug, p.N, p.ld_w, WT, L.Np, L.NPk, s);
    // constant-one column and zero padding of all four operand buffers
    launch_init_operand(p, y_traj, t_dev, Ra[0], Ra[1], nullptr, L.KPa, L.Bp, s);
    launch_init_operand(p, y_traj, t_dev, Ra[2], Ra[3], nullptr, L.KPa, L.Bp, s);
    k_build_inv<<<1, 256, 0, s>>>(sel, G, 3 * p.N, inv);
    k_bwd_begin<<<296, 256, 0, s>>>(p, grad_y, inv, G, t_dev, T, gamma, lam, acur, AV[3], L.NPk);
    count_launch(2);

    const dim3 grid(L.Np / TM, L.Bp / TN);
    // split the trial axis of the dW contraction so that about two waves of CTAs are in flight
    const int tiles = (L.Np / TM) * ((L.KPa + TN - 1) / TN);
    int splits = (2 * 148 + tiles - 1) / tiles;
    const int max_splits = L.Bp / TK;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int rows = (L.Bp / TK + splits - 1) / splits * TK;
    splits = (L.Bp + rows - 1) / rows;

    for (int n = T - 2; n >= 0; --n) {
        const float* yn = y_traj + (size_t)n * st;
        launch_init_operand(p, yn, t_dev + n, Ra[0], nullptr, DR[0], L.KPa, L.Bp, s);
        FwdStageArgs f;
        f.p = p; f.Wp = Wp; f.y0 = yn; f.k1 = kk[0]; f.k2 = kk[1]; f.k3 = kk[2]; f.y1 = nullptr; f.y_out_row = nullptr;
        f.t = t_dev; f.n = n; f.KPa = L.KPa; f.Ra_cur_lo = nullptr; f.Ra_nxt_lo = nullptr;
        for (int S = 1; S <= 3; ++S) {
            f.Ra_cur = Ra[S - 1]; f.Ra_nxt = Ra[S]; f.DR_nxt = DR[S];
            launch_fwd_stage(S, f, grid, s);
        }
        BwdStageArgs a;
        a.p = p; a.WT = WT; a.acur = acur; a.lam = lam; a.b4 = b4; a.b3 = b3; a.grad_y = grad_y; a.inv = inv;
        a.t = t_dev; a.n = n; a.G = G; a.NPk = L.NPk; a.gamma = gamma;
        a.inv_tau_m = 1.0f / p.c.tau_m; a.inv_tau_a = 1.0f / p.c.tau_a; a.inv_tau_s = 1.0f / p.c.tau_s;
        a.AV_cur = AV[3]; a.AV_nxt = AV[2]; a.DR = DR[3];
        k_bwd_stage<4><<<grid, kGemmThreads, 0, s>>>(a);
        a.AV_cur = AV[2]; a.AV_nxt = AV[1]; a.DR = DR[2];
        k_bwd_stage<3><<<grid, kGemmThreads, 0, s>>>(a);
        a.AV_cur = AV[1]; a.AV_nxt = AV[0]; a.DR = DR[1];
        k_bwd_stage<2><<<grid, kGemmThreads, 0, s>>>(a);
        DwArgs d;
        for (int i = 0; i < 4; ++i) { d.AV[i] = AV[i]; d.Ra[i] = Ra[i]; }
        d.grad_W = grad_W; d.N = p.N; d.Kaug = Kaug; d.ld_w = p.ld_w; d.NPk = L.NPk; d.KPa = L.KPa; d.B = p.B; d.Bp = L.Bp;
        d.rows_per_split = rows;
        k_dw_accum<<<dim3(L.Np / TM, (L.KPa + TN - 1) / TN, splits), kGemmThreads, 0, s>>>(d);
        a.AV_cur = AV[0]; a.AV_nxt = AV[3]; a.DR = DR[0];
        k_bwd_stage<1><<<grid, kGemmThreads, 0, s>>>(a);
        count_launch(5);
    }
    if (grad_y0) {
        if (cudaMemcpyAsync(grad_y0, lam, sizeof(float) * st, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return ODECOL_E_CUDA;
    }
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

}  // namespace odecol
