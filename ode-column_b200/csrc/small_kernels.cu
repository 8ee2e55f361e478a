// Kernel family S ("small"): one persistent CTA integrates ONE trial through the whole time loop.
//
//   * thread i owns population i: V_i, A_i, F_i and every Runge-Kutta stage value live in registers;
//   * row i of W_aug = [W | U | bias] lives in registers too (KP floats, compile-time padded), so the
//     recurrent + feedforward + background input is one fully unrolled FFMA dot product against the
//     trial's r_aug = [phi(V-A) ; s(t) ; 1] vector, which is exchanged through (double-buffered) shared
//     memory with ONE barrier per right-hand-side evaluation;
//   * adaptive solvers (dopri5, step-doubling Euler-Maruyama) take their accept/reject decision per trial
//     from a warp-shuffle + shared-memory reduction, exactly like the reference, which solves trials
//     one at a time (scripts/xor_ode.py:104-117);
//   * many such CTAs share an SM (32..128 threads each), trials never communicate.
//
// Used when N <= 128 and N + n_in + 1 <= 128 (the reference's WTA / XOR / parity networks).
// Replaces the Python stepping loops of torchdiffeq / torchsde around ColumnArea*.forward
// (reference src/coupled_columns.py:204-237, 407-442, 753-788).
#include "odecol_common.cuh"

namespace odecol {

constexpr float kOneThird = 0.3333333333333333f;   // float32(1/3), as torch scalar math
constexpr float kTwoThirds = 0.6666666666666666f;

// ---------------------------------------------------------------------------------------------------------------
// block reductions (all threads return the same value)
// ---------------------------------------------------------------------------------------------------------------
ODECOL_DEVINL double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int nw = blockDim.x >> 5;
    if (nw == 1) return v;
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += red[w];
    __syncthreads();
    return s;
}

ODECOL_DEVINL bool block_any(bool flag) { return __syncthreads_or(flag ? 1 : 0) != 0; }

// ---------------------------------------------------------------------------------------------------------------
// per-thread right-hand side with the weight row in registers
// ---------------------------------------------------------------------------------------------------------------
// Transfer function of the on-chip family: the reference's operations in their order with IEEE division, tanhf and expf
// (bit-faithful up to the last ulp of the two libm calls), or -- built with -DODECOL_SMALL_FAST_PHI -- the staged families'
// fast path (1e-7 relative).  Measured (bench.py --workload small, A/B on one box): see DESIGN.md section 5.
#ifdef ODECOL_SMALL_FAST_PHI
ODECOL_DEVINL float small_phi(float x) { return phi_fast(x); }
ODECOL_DEVINL void small_phi_dphi(float x, float& r, float& dr) { phi_dphi_fast(x, r, dr); }
#else
ODECOL_DEVINL float small_phi(float x) { return phi(x); }
ODECOL_DEVINL void small_phi_dphi(float x, float& r, float& dr) { phi_dphi(x, r, dr); }
#endif

template <int KP>
struct RowRhs {
    float w[KP];
    float kappa;
    float* ra;            // shared [2][KP]
    const float* ku;      // this trial's knots
    int idx, buf;
    int N, n_in, K;
    const float* kt;
    Consts c;
    bool act;
    KnotLane kl;          // this lane's stimulus channel (when n_in <= blockDim.x)

    int li, lanes;        // population index of this lane within its trial; lanes per trial

    // One trial per CTA (li = threadIdx.x), or -- packed mode, networks with N, n_in <= 16 -- two trials per warp:
    // lanes 0..15 integrate trial 2*blockIdx.x, lanes 16..31 trial 2*blockIdx.x + 1, each with its own r_aug buffers.
    ODECOL_DEVINL void init(const DevProblem& p, float* ra_smem, int b, int li_ = threadIdx.x, int sub = 0,
                            int lanes_ = blockDim.x) {
        li = li_; lanes = lanes_;
        const int i = li;
        N = p.N; n_in = p.n_in; K = p.K; kt = p.knot_t; c = p.c;
        const bool row = i < N;
        act = row && b < p.B;
        ra = ra_smem + sub * 2 * KP;
        ku = p.knot_u + (size_t)(b < p.B ? b : p.B - 1) * p.knot_stride_b;
        idx = 1; buf = 0;
        kl.reset();
        const int Kaug = p.N + p.n_in + 1;
#pragma unroll
        for (int k = 0; k < KP; ++k) w[k] = (row && k < Kaug) ? __ldg(p.W_aug + (size_t)i * p.ld_w + k) : 0.0f;
        kappa = row ? __ldg(p.kappa + i) : 0.0f;
        for (int k = i; k < 2 * KP; k += lanes) ra[k] = 0.0f;
        __syncthreads();
        if (i == 0) { ra[Kaug - 1] = 1.0f; ra[KP + Kaug - 1] = 1.0f; }
        __syncthreads();
    }

    // publishes r_aug(t, V-A) and returns the total synaptic input of population i; one barrier
    ODECOL_DEVINL float input(float t, float r) {
        buf ^= 1;
        float* cur = ra + buf * KP;
        if (act) cur[li] = r;
        if (li < n_in) {
            if (n_in <= lanes) cur[N + li] = kl.value(kt, ku, K, n_in, li, t);
            else {
                const float tc = knot_locate(kt, K, t, idx);
                for (int ch = li; ch < n_in; ch += lanes) cur[N + ch] = knot_value(kt, ku, n_in, idx, tc, ch);
            }
        }
        __syncthreads();
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float4* r4 = reinterpret_cast<const float4*>(cur);
#pragma unroll
        for (int k = 0; k < KP / 4; ++k) {
            const float4 v = r4[k];
            a0 = fmaf(w[4 * k + 0], v.x, a0);
            a1 = fmaf(w[4 * k + 1], v.y, a1);
            a2 = fmaf(w[4 * k + 2], v.z, a2);
            a3 = fmaf(w[4 * k + 3], v.w, a3);
        }
        return (a0 + a1) + (a2 + a3);
    }

    ODECOL_DEVINL void eval(float t, float V, float A, float F, float& dV, float& dA, float& dF) {
        const float r = small_phi(__fsub_rn(V, A));
        const float tot = input(t, r);
        drift(c, V, A, F, r, kappa, tot, dV, dA, dF);
    }
};

struct Y3 { float V, A, F; };
ODECOL_DEVINL Y3 ld3(const float* base, int N, int i) { return {base[i], base[N + i], base[2 * N + i]}; }
ODECOL_DEVINL void st3(float* base, int N, int i, float V, float A, float F) { base[i] = V; base[N + i] = A; base[2 * N + i] = F; }

// ---------------------------------------------------------------------------------------------------------------
// single RHS evaluation (module.forward parity); generic in N: CTA per trial, r_aug in dynamic smem
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_rhs_generic(DevProblem p, const float* __restrict__ t, const float* __restrict__ y,
                              float* __restrict__ f) {
    extern __shared__ float ra[];
    const int b = blockIdx.x, N = p.N, Kaug = N + p.n_in + 1;
    const float* yb = y + (size_t)b * 3 * N;
    const float* ku = p.knot_u + (size_t)b * p.knot_stride_b;
    const float tb = t[b];
    int idx = 1;
    const float tc = knot_locate(p.knot_t, p.K, tb, idx);
    for (int i = threadIdx.x; i < N; i += blockDim.x) ra[i] = phi(__fsub_rn(yb[i], yb[N + i]));
    for (int ch = threadIdx.x; ch < p.n_in; ch += blockDim.x) ra[N + ch] = knot_value(p.knot_t, ku, p.n_in, idx, tc, ch);
    if (threadIdx.x == 0) ra[Kaug - 1] = 1.0f;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float* wr = p.W_aug + (size_t)i * p.ld_w;
        float acc = 0.f;
        for (int k = 0; k < Kaug; ++k) acc = fmaf(__ldg(wr + k), ra[k], acc);
        float dV, dA, dF;
        drift(p.c, yb[i], yb[N + i], yb[2 * N + i], ra[i], __ldg(p.kappa + i), acc, dV, dA, dF);
        st3(f + (size_t)b * 3 * N, N, i, dV, dA, dF);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// RK4 (3/8 rule) forward
// ---------------------------------------------------------------------------------------------------------------
// occupancy targets of the forward kernels: blocks are N rounded up to a warp (<= 64 threads for KP <= 64, ...); the weight
// row alone takes KP registers, the bounds keep the compiler from trading resident CTAs for a few more
template <int KP> struct FwdBounds {
    static constexpr int threads = KP <= 64 ? 64 : (KP <= 96 ? 96 : 128);
    static constexpr int blocks = KP <= 40 ? 12 : (KP <= 64 ? 8 : (KP <= 96 ? 4 : 3));
};

template <int KP>
__global__ void __launch_bounds__(FwdBounds<KP>::threads, FwdBounds<KP>::blocks) k_rk4_fwd_small(DevProblem p, const float* __restrict__ t, int T,
                                                       const float* __restrict__ y0, float* __restrict__ y_out,
                                                       int out_every, int packed, const unsigned int* __restrict__ run_if) {
    __shared__ __align__(16) float ra[4 * KP];
    if (run_if && *run_if == 0u) return;      // repeat of a 16-bit tensor-core solve (tiny_tc.cu) that met no overflow
    const int sub = packed ? (int)(threadIdx.x >> 4) : 0, i = packed ? (int)(threadIdx.x & 15) : (int)threadIdx.x;
    const int b = packed ? 2 * (int)blockIdx.x + sub : (int)blockIdx.x, N = p.N;
    RowRhs<KP> f;
    f.init(p, ra, b, i, sub, packed ? 16 : (int)blockDim.x);
    const size_t row = (size_t)3 * N;
    float V = 0.f, A = 0.f, F = 0.f;
    if (f.act) {
        const Y3 s = ld3(y0 + b * row, N, i);
        V = s.V; A = s.A; F = s.F;
        st3(y_out + b * row, N, i, V, A, F);
    }
    int since_out = 0, out_row = 0;
    for (int n = 0; n < T - 1; ++n) {
        const float t0 = __ldg(t + n), t1 = __ldg(t + n + 1);
        const float dt = __fsub_rn(t1, t0);
        float k1V, k1A, k1F, k2V, k2A, k2F, k3V, k3A, k3F, k4V, k4A, k4F;
        f.eval(t0, V, A, F, k1V, k1A, k1F);
        // Y2 = y0 + dt*k1*(1/3)
        f.eval(__fadd_rn(t0, __fmul_rn(dt, kOneThird)),
               __fadd_rn(V, __fmul_rn(__fmul_rn(dt, k1V), kOneThird)),
               __fadd_rn(A, __fmul_rn(__fmul_rn(dt, k1A), kOneThird)),
               __fadd_rn(F, __fmul_rn(__fmul_rn(dt, k1F), kOneThird)), k2V, k2A, k2F);
        // Y3 = y0 + dt*(k2 - k1*(1/3))
        f.eval(__fadd_rn(t0, __fmul_rn(dt, kTwoThirds)),
               __fadd_rn(V, __fmul_rn(dt, __fsub_rn(k2V, __fmul_rn(k1V, kOneThird)))),
               __fadd_rn(A, __fmul_rn(dt, __fsub_rn(k2A, __fmul_rn(k1A, kOneThird)))),
               __fadd_rn(F, __fmul_rn(dt, __fsub_rn(k2F, __fmul_rn(k1F, kOneThird)))), k3V, k3A, k3F);
        // Y4 = y0 + dt*(k1 - k2 + k3)
        f.eval(t1,
               __fadd_rn(V, __fmul_rn(dt, __fadd_rn(__fsub_rn(k1V, k2V), k3V))),
               __fadd_rn(A, __fmul_rn(dt, __fadd_rn(__fsub_rn(k1A, k2A), k3A))),
               __fadd_rn(F, __fmul_rn(dt, __fadd_rn(__fsub_rn(k1F, k2F), k3F))), k4V, k4A, k4F);
        // y1 = y0 + (k1 + 3*(k2+k3) + k4)*dt*0.125
        V = __fadd_rn(V, __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(k1V, __fmul_rn(3.f, __fadd_rn(k2V, k3V))), k4V), dt), 0.125f));
        A = __fadd_rn(A, __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(k1A, __fmul_rn(3.f, __fadd_rn(k2A, k3A))), k4A), dt), 0.125f));
        F = __fadd_rn(F, __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(k1F, __fmul_rn(3.f, __fadd_rn(k2F, k3F))), k4F), dt), 0.125f));
        const int j = n + 1;
        if (++since_out == out_every) { since_out = 0; ++out_row; }       // j % out_every == 0 without the division
        if (f.act && (since_out == 0 || j == T - 1)) {
            const size_t r = since_out == 0 ? (size_t)out_row : (size_t)((T - 2) / out_every + 1);
            st3(y_out + (r * p.B + b) * row, N, i, V, A, F);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// RK4 discrete adjoint (reverse sweep over the saved trajectory; stages recomputed per step)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBwdSlots = 7;     // r_aug vectors kept per step: 4 RK4 stages, or the 7 Dormand-Prince stages

template <int KP>
struct BwdShared {
    float* Ws;     // [N][KP+1]   padded rows: row access (thread = row) and column access (thread = col) conflict free
    float* ra;     // [kBwdSlots][KP] r_aug of the stages of one step
    float* av;     // [2][NP]
    int* inv;      // [3N]        state component -> column of grad_y, or -1
};

template <int KP>
ODECOL_DEVINL BwdShared<KP> carve_bwd(float* sm, int N, int NP, int tpc) {
    BwdShared<KP> s;
    s.ra = sm;                                   // 16-byte aligned; [tpc trials][kBwdSlots][KP]
    s.av = s.ra + tpc * kBwdSlots * KP;          // [tpc][2][NP]
    s.Ws = s.av + tpc * 2 * NP;
    s.inv = reinterpret_cast<int*>(s.Ws + (size_t)N * (KP + 1));
    return s;
}

size_t small_bwd_smem_bytes(int N, int KP, int tpc) {
    const int NP = (N + 31) / 32 * 32;           // >= the lanes per trial in either mode
    return sizeof(float) * ((size_t)tpc * kBwdSlots * KP + (size_t)tpc * 2 * NP + (size_t)N * (KP + 1) + 3 * N);
}

// FAST: the transfer function through the staged families' fast path (used for the large batches whose forward solve already
// ran on the tensor cores: tiny_tc.cu -- bit-faithfulness to the reference's libm calls is not the point there)
template <int KP, bool FAST = false>
struct BwdCtx {
    BwdShared<KP> s;
    float dw[KP];
    float wrow[FAST ? KP : 1];   // FAST (N <= 16, two trials per warp): row li of W_aug and column li of its leading block in
    float wcol[FAST ? 16 : 1];   // registers -- the dot products read only the broadcast operand from shared memory (16-byte loads)
    float kappa, gamma, inv_tau_m, inv_tau_a, inv_tau_s;
    const float* ku;
    const float* kt;
    int idx, abuf, N, n_in, K, NP;
    Consts c;
    bool act;
    KnotLane kl;
    int li, lanes;        // population index of this lane within its trial; lanes per trial (see RowRhs::init)
    float* ra_t;          // this trial's [kBwdSlots][KP]
    float* av_t;          // this trial's [2][NP]

    ODECOL_DEVINL void init(const DevProblem& p, float* sm, int b, int li_ = threadIdx.x, int sub = 0,
                            int lanes_ = blockDim.x, int tpc = 1) {
        li = li_; lanes = lanes_;
        const int i = li;
        N = p.N; n_in = p.n_in; K = p.K; kt = p.knot_t; c = p.c;
        NP = (N + 31) / 32 * 32;
        act = i < N && b < p.B;
        s = carve_bwd<KP>(sm, N, NP, tpc);
        ra_t = s.ra + sub * kBwdSlots * KP;
        av_t = s.av + sub * 2 * NP;
        ku = p.knot_u + (size_t)(b < p.B ? b : p.B - 1) * p.knot_stride_b;
        idx = 1; abuf = 0;
        kl.reset();
        const int Kaug = N + n_in + 1;
        for (int e = threadIdx.x; e < N * KP; e += blockDim.x) {
            const int r = e / KP, k = e % KP;
            s.Ws[r * (KP + 1) + k] = k < Kaug ? __ldg(p.W_aug + (size_t)r * p.ld_w + k) : 0.0f;
        }
        for (int e = threadIdx.x; e < tpc * kBwdSlots * KP; e += blockDim.x) s.ra[e] = (e % KP == Kaug - 1) ? 1.0f : 0.0f;
        for (int e = threadIdx.x; e < tpc * 2 * NP; e += blockDim.x) s.av[e] = 0.0f;
#pragma unroll
        for (int k = 0; k < KP; ++k) dw[k] = 0.0f;
        kappa = i < N ? __ldg(p.kappa + i) : 0.0f;
        gamma = c.tau_s * c.R / c.tau_m;
        inv_tau_m = 1.0f / c.tau_m; inv_tau_a = 1.0f / c.tau_a; inv_tau_s = 1.0f / c.tau_s;
        __syncthreads();
        if constexpr (FAST) {
#pragma unroll
            for (int k = 0; k < KP; ++k) wrow[k] = i < N ? s.Ws[i * (KP + 1) + k] : 0.0f;
#pragma unroll
            for (int r = 0; r < 16; ++r) wcol[r] = (i < N && r < N) ? s.Ws[r * (KP + 1) + i] : 0.0f;
        }
    }

    // forward stage: publishes r_aug into ra[stage], returns total input (needs_dot) ; r, dr out
    ODECOL_DEVINL float stage_fwd(int stage, float t, float V, float A, float& r, float& dr, bool needs_dot) {
        if (FAST) phi_dphi_fast(__fsub_rn(V, A), r, dr);
        else small_phi_dphi(__fsub_rn(V, A), r, dr);
        float* cur = ra_t + stage * KP;
        if (act) cur[li] = r;
        if (li < n_in) {
            if (n_in <= lanes) cur[N + li] = kl.value(kt, ku, K, n_in, li, t);
            else {
                const float tc = knot_locate(kt, K, t, idx);
                for (int ch = li; ch < n_in; ch += lanes) cur[N + ch] = knot_value(kt, ku, n_in, idx, tc, ch);
            }
        }
        __syncthreads();
        float acc = 0.f;
        if constexpr (FAST) {
            if (needs_dot && act) {                  // same two accumulators, same order as below
                const float4* c4 = reinterpret_cast<const float4*>(cur);
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int k = 0; k < KP / 4; ++k) {
                    const float4 v = c4[k];
                    a0 = fmaf(wrow[4 * k], v.x, a0); a1 = fmaf(wrow[4 * k + 1], v.y, a1);
                    a0 = fmaf(wrow[4 * k + 2], v.z, a0); a1 = fmaf(wrow[4 * k + 3], v.w, a1);
                }
                acc = a0 + a1;
            }
            return acc;
        }
        if (needs_dot && act) {
            const float* wr = s.Ws + li * (KP + 1);
            float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
            for (int k = 0; k < KP; k += 2) { a0 = fmaf(wr[k], cur[k], a0); a1 = fmaf(wr[k + 1], cur[k + 1], a1); }
            acc = a0 + a1;
        }
        return acc;
    }

    // (Yb) = J(stage)^T (aV, aA, aF);  accumulates dW_aug row
    ODECOL_DEVINL void stage_bwd(int stage, float dr, float aV, float aA, float aF, float& bV, float& bA, float& bF) {
        abuf ^= 1;
        float* a = av_t + abuf * NP;
        const float ga = gamma * aV;
        if (li < NP) a[li] = act ? ga : 0.0f;
        __syncthreads();
        float g = 0.f;
        if constexpr (FAST) {
            if (act) {                               // entries beyond N are zeros on both sides: the sums below are the ones further down
                const float4* a4 = reinterpret_cast<const float4*>(a);
                float g0 = 0.f, g1 = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = a4[q];
                    g0 = fmaf(wcol[4 * q], v.x, g0); g1 = fmaf(wcol[4 * q + 1], v.y, g1);
                    g0 = fmaf(wcol[4 * q + 2], v.z, g0); g1 = fmaf(wcol[4 * q + 3], v.w, g1);
                }
                g = g0 + g1 + kappa * aA * inv_tau_a + aF * inv_tau_s;
            }
        } else if (act) {
            const float* wc = s.Ws + li;
            float g0 = 0.f, g1 = 0.f;
            int r = 0;
            for (; r + 1 < N; r += 2) { g0 = fmaf(wc[r * (KP + 1)], a[r], g0); g1 = fmaf(wc[(r + 1) * (KP + 1)], a[r + 1], g1); }
            if (r < N) g0 = fmaf(wc[r * (KP + 1)], a[r], g0);
            g = g0 + g1 + kappa * aA * inv_tau_a + aF * inv_tau_s;
        }
        const float4* r4 = reinterpret_cast<const float4*>(ra_t + stage * KP);
#pragma unroll
        for (int k = 0; k < KP / 4; ++k) {
            const float4 v = r4[k];
            dw[4 * k + 0] = fmaf(ga, v.x, dw[4 * k + 0]);
            dw[4 * k + 1] = fmaf(ga, v.y, dw[4 * k + 1]);
            dw[4 * k + 2] = fmaf(ga, v.z, dw[4 * k + 2]);
            dw[4 * k + 3] = fmaf(ga, v.w, dw[4 * k + 3]);
        }
        bV = -aV * inv_tau_m + dr * g;
        bA = -aA * inv_tau_a - dr * g;
        bF = -aF * inv_tau_s;
    }

    ODECOL_DEVINL void flush_dw(const DevProblem& p, float* grad_W) {
        if (!act) return;
        const int Kaug = N + n_in + 1;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < Kaug) atomicAdd(grad_W + (size_t)li * p.ld_w + k, dw[k]);
    }
};

// reverse kernels: the dW row (KP registers) rides along; only the smallest configuration has room to trade registers
// for resident CTAs
template <int KP, bool FAST = false> struct BwdBounds {
    static constexpr int threads = KP <= 40 ? 64 : 128;
    static constexpr int blocks = KP <= 40 ? (FAST ? 6 : 10) : 0;      // 0 = unspecified; FAST keeps W in registers (BwdCtx)
};

template <int KP, bool FAST = false>
__global__ void __launch_bounds__(BwdBounds<KP, FAST>::threads, BwdBounds<KP, FAST>::blocks) k_rk4_bwd_small(DevProblem p, const float* __restrict__ t, int T,
                                                       const float* __restrict__ y_traj,
                                                       const float* __restrict__ grad_y, const int* __restrict__ sel,
                                                       int G, float* __restrict__ grad_y0, float* __restrict__ grad_W,
                                                       int packed) {
    extern __shared__ __align__(16) float sm[];
    const int sub = packed ? (int)(threadIdx.x >> 4) : 0, i = packed ? (int)(threadIdx.x & 15) : (int)threadIdx.x;
    const int b = packed ? 2 * (int)blockIdx.x + sub : (int)blockIdx.x, N = p.N, B = p.B;
    BwdCtx<KP, FAST> cx;
    cx.init(p, sm, b, i, sub, packed ? 16 : (int)blockDim.x, packed ? 2 : 1);
    for (int e = threadIdx.x; e < 3 * N; e += blockDim.x) cx.s.inv[e] = sel ? -1 : e;
    __syncthreads();
    if (sel) for (int g = threadIdx.x; g < G; g += blockDim.x) cx.s.inv[sel[g]] = g;
    __syncthreads();
    const int gV = cx.act ? cx.s.inv[i] : -1, gA = cx.act ? cx.s.inv[N + i] : -1, gF = cx.act ? cx.s.inv[2 * N + i] : -1;
    const size_t row = (size_t)3 * N;
    auto add_grad = [&](int n, float& lV, float& lA, float& lF) {
        const float* g = grad_y + ((size_t)n * B + b) * G;
        if (gV >= 0) lV += g[gV];
        if (gA >= 0) lA += g[gA];
        if (gF >= 0) lF += g[gF];
    };
    float lV = 0.f, lA = 0.f, lF = 0.f;
    add_grad(T - 1, lV, lA, lF);
    for (int n = T - 2; n >= 0; --n) {
        const float t0 = __ldg(t + n), t1 = __ldg(t + n + 1);
        const float dt = __fsub_rn(t1, t0);
        float V = 0.f, A = 0.f, F = 0.f;
        if (cx.act) { const Y3 s = ld3(y_traj + ((size_t)n * B + b) * row, N, i); V = s.V; A = s.A; F = s.F; }
        // ---- recompute the stages (same arithmetic as the forward kernel)
        float r, d1, d2, d3, d4, tot, k1V, k1A, k1F, k2V, k2A, k2F, k3V, k3A, k3F;
        tot = cx.stage_fwd(0, t0, V, A, r, d1, true);
        drift(cx.c, V, A, F, r, cx.kappa, tot, k1V, k1A, k1F);
        float sV = __fadd_rn(V, __fmul_rn(__fmul_rn(dt, k1V), kOneThird));
        float sA = __fadd_rn(A, __fmul_rn(__fmul_rn(dt, k1A), kOneThird));
        float sF = __fadd_rn(F, __fmul_rn(__fmul_rn(dt, k1F), kOneThird));
        tot = cx.stage_fwd(1, __fadd_rn(t0, __fmul_rn(dt, kOneThird)), sV, sA, r, d2, true);
        drift(cx.c, sV, sA, sF, r, cx.kappa, tot, k2V, k2A, k2F);
        sV = __fadd_rn(V, __fmul_rn(dt, __fsub_rn(k2V, __fmul_rn(k1V, kOneThird))));
        sA = __fadd_rn(A, __fmul_rn(dt, __fsub_rn(k2A, __fmul_rn(k1A, kOneThird))));
        sF = __fadd_rn(F, __fmul_rn(dt, __fsub_rn(k2F, __fmul_rn(k1F, kOneThird))));
        tot = cx.stage_fwd(2, __fadd_rn(t0, __fmul_rn(dt, kTwoThirds)), sV, sA, r, d3, true);
        drift(cx.c, sV, sA, sF, r, cx.kappa, tot, k3V, k3A, k3F);
        sV = __fadd_rn(V, __fmul_rn(dt, __fadd_rn(__fsub_rn(k1V, k2V), k3V)));
        sA = __fadd_rn(A, __fmul_rn(dt, __fadd_rn(__fsub_rn(k1A, k2A), k3A)));
        (void)cx.stage_fwd(3, t1, sV, sA, r, d4, false);
        // ---- reverse sweep.  y1 = y0 + dt/8 (k1 + 3k2 + 3k3 + k4)
        const float h8 = dt * 0.125f, h38 = 3.0f * h8, h3 = dt * kOneThird;
        float b4V, b4A, b4F, b3V, b3A, b3F, b2V, b2A, b2F, b1V, b1A, b1F;
        cx.stage_bwd(3, d4, h8 * lV, h8 * lA, h8 * lF, b4V, b4A, b4F);
        cx.stage_bwd(2, d3, h38 * lV + dt * b4V, h38 * lA + dt * b4A, h38 * lF + dt * b4F, b3V, b3A, b3F);
        cx.stage_bwd(1, d2, h38 * lV - dt * b4V + dt * b3V, h38 * lA - dt * b4A + dt * b3A,
                     h38 * lF - dt * b4F + dt * b3F, b2V, b2A, b2F);
        cx.stage_bwd(0, d1, h8 * lV + dt * b4V - h3 * b3V + h3 * b2V, h8 * lA + dt * b4A - h3 * b3A + h3 * b2A,
                     h8 * lF + dt * b4F - h3 * b3F + h3 * b2F, b1V, b1A, b1F);
        lV += b4V + b3V + b2V + b1V;
        lA += b4A + b3A + b2A + b1A;
        lF += b4F + b3F + b2F + b1F;
        add_grad(n, lV, lA, lF);
        __syncthreads();    // ra[] of this step fully consumed before the next step overwrites it
    }
    if (grad_y0 && cx.act) st3(grad_y0 + b * row, N, i, lV, lA, lF);
    cx.flush_dw(p, grad_W);
}

// ---------------------------------------------------------------------------------------------------------------
// dopri5 forward with per-trial control and dense output
// ---------------------------------------------------------------------------------------------------------------
// struct DP (the Dormand-Prince tableau as float32) lives in odecol_common.cuh: the staged family uses it too

template <int KP>
__global__ void __launch_bounds__(128) k_dopri5_fwd_small(DevProblem p, const float* __restrict__ t, int T,
                                                          const float* __restrict__ y0, float* __restrict__ y_out,
                                                          float rtol, float atol, int max_steps,
                                                          int* __restrict__ n_accept, int* __restrict__ n_reject,
                                                          int* __restrict__ status, Dopri5Record rec) {
    __shared__ __align__(16) float ra[2 * KP];
    __shared__ double red[4];
    const int b = blockIdx.x, i = threadIdx.x, N = p.N;
    RowRhs<KP> f;
    f.init(p, ra, b);
    const size_t row = (size_t)3 * N;
    const double cnt = 3.0 * N;
    float y[3] = {0.f, 0.f, 0.f};
    if (f.act) {
        const Y3 s = ld3(y0 + b * row, N, i);
        y[0] = s.V; y[1] = s.A; y[2] = s.F;
        st3(y_out + b * row, N, i, y[0], y[1], y[2]);
    }
    float k[7][3];
    auto rms3 = [&](float a0, float a1, float a2) -> float {
        double s = f.act ? ((double)a0 * a0 + (double)a1 * a1 + (double)a2 * a2) : 0.0;
        s = block_sum(s, red);
        return (float)sqrt(s / cnt);
    };
    // ---- before integrate: f0 and Hairer's initial step (order 4 -> exponent 1/5)
    double tc0 = (double)__ldg(t);
    f.eval((float)tc0, y[0], y[1], y[2], k[0][0], k[0][1], k[0][2]);
    double dt;
    {
        float sc[3], q0[3], q1[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            sc[c] = __fadd_rn(atol, __fmul_rn(fabsf(y[c]), rtol));
            q0[c] = __fdiv_rn(y[c], sc[c]);
            q1[c] = __fdiv_rn(k[0][c], sc[c]);
        }
        const float d0 = rms3(q0[0], q0[1], q0[2]);
        const float d1 = rms3(q1[0], q1[1], q1[2]);
        float h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : __fdiv_rn(__fmul_rn(0.01f, d0), d1);
        h0 = fabsf(h0);
        float y1[3], f1[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) y1[c] = __fadd_rn(y[c], __fmul_rn(h0, k[0][c]));
        f.eval((float)(tc0 + (double)h0), y1[0], y1[1], y1[2], f1[0], f1[1], f1[2]);
        const float d2 = fabsf(__fdiv_rn(rms3(__fdiv_rn(__fsub_rn(f1[0], k[0][0]), sc[0]), __fdiv_rn(__fsub_rn(f1[1], k[0][1]), sc[1]),
                                              __fdiv_rn(__fsub_rn(f1[2], k[0][2]), sc[2])), h0));
        float h1;
        if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, __fmul_rn(h0, 1e-3f));
        else h1 = powf(__fdiv_rn(0.01f, fmaxf(d1, d2)), 0.2f);
        dt = (double)fminf(__fmul_rn(100.f, h0), fabsf(h1));
    }
    double st_t0 = tc0, st_t1 = tc0;
    float ce[3], cd[3], cc[3], cb[3], ca[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) ce[c] = cd[c] = cc[c] = cb[c] = ca[c] = y[c];
    int nacc = 0, nrej = 0, st = ODECOL_ST_OK;
    int j = 1;
    for (; j < T && st == ODECOL_ST_OK; ++j) {
        const double next_t = (double)__ldg(t + j);
        while (next_t > st_t1) {
            if (nacc + nrej >= max_steps) { st = ODECOL_ST_MAXSTEPS; break; }
            const double t0 = st_t1, t1 = t0 + dt;
            if (!(t0 + dt > t0)) { st = ODECOL_ST_UNDERFLOW; break; }
            if (block_any(f.act && !(isfinite(y[0]) && isfinite(y[1]) && isfinite(y[2])))) { st = ODECOL_ST_NONFINITE; break; }
            const float t0f = (float)t0, dtf = (float)dt, t1f = (float)t1;
            float ys[3];
#define ODECOL_STAGE(expr, tt, ko)                                                                   \
    _Pragma("unroll") for (int c = 0; c < 3; ++c) ys[c] = __fadd_rn(y[c], (expr));                   \
    f.eval((tt), ys[0], ys[1], ys[2], k[ko][0], k[ko][1], k[ko][2]);
            ODECOL_STAGE(k[0][c] * (DP::b10 * dtf), __fadd_rn(t0f, __fmul_rn(DP::a1, dtf)), 1)
            ODECOL_STAGE(fmaf(k[1][c], DP::b21 * dtf, k[0][c] * (DP::b20 * dtf)), __fadd_rn(t0f, __fmul_rn(DP::a2, dtf)), 2)
            ODECOL_STAGE(fmaf(k[2][c], DP::b32 * dtf, fmaf(k[1][c], DP::b31 * dtf, k[0][c] * (DP::b30 * dtf))),
                         __fadd_rn(t0f, __fmul_rn(DP::a3, dtf)), 3)
            ODECOL_STAGE(fmaf(k[3][c], DP::b43 * dtf, fmaf(k[2][c], DP::b42 * dtf, fmaf(k[1][c], DP::b41 * dtf, k[0][c] * (DP::b40 * dtf)))),
                         __fadd_rn(t0f, __fmul_rn(DP::a4, dtf)), 4)
            ODECOL_STAGE(fmaf(k[4][c], DP::b54 * dtf, fmaf(k[3][c], DP::b53 * dtf, fmaf(k[2][c], DP::b52 * dtf,
                              fmaf(k[1][c], DP::b51 * dtf, k[0][c] * (DP::b50 * dtf))))), t1f, 5)
            ODECOL_STAGE(fmaf(k[5][c], DP::b65 * dtf, fmaf(k[4][c], DP::b64 * dtf, fmaf(k[3][c], DP::b63 * dtf,
                              fmaf(k[2][c], DP::b62 * dtf, k[0][c] * (DP::b60 * dtf))))), t1f, 6)
#undef ODECOL_STAGE
            // ys is y1 (FSAL), k[6] is f1
            float q[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float err = fmaf(k[6][c], DP::e6 * dtf, fmaf(k[5][c], DP::e5 * dtf, fmaf(k[4][c], DP::e4 * dtf,
                                  fmaf(k[3][c], DP::e3 * dtf, fmaf(k[2][c], DP::e2 * dtf, k[0][c] * (DP::e0 * dtf))))));
                const float tol = __fadd_rn(atol, __fmul_rn(rtol, fmaxf(fabsf(y[c]), fabsf(ys[c]))));
                q[c] = __fdiv_rn(err, tol);
            }
            const float ratio = rms3(q[0], q[1], q[2]);
            const bool accept = ratio <= 1.0f;
            if (accept && rec.y) {                 // training mode: remember the step (start state, t0, dt)
                if (nacc >= rec.cap) { st = ODECOL_ST_MAXSTEPS; break; }
                if (f.act) st3(rec.y + ((size_t)nacc * p.B + b) * row, N, i, y[0], y[1], y[2]);
                if (i == 0) { rec.t0[(size_t)b * rec.cap + nacc] = t0; rec.dt[(size_t)b * rec.cap + nacc] = dt; }
            }
            if (accept) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float ymid = __fadd_rn(y[c], fmaf(k[6][c], DP::m6 * dtf, fmaf(k[5][c], DP::m5 * dtf, fmaf(k[4][c], DP::m4 * dtf,
                                       fmaf(k[3][c], DP::m3 * dtf, fmaf(k[2][c], DP::m2 * dtf, k[0][c] * (DP::m0 * dtf)))))));
                    const float f0 = k[0][c], f1 = k[6][c], y0c = y[c], y1c = ys[c];
                    ca[c] = 2.f * dtf * (f1 - f0) - 8.f * (y1c + y0c) + 16.f * ymid;
                    cb[c] = dtf * (5.f * f0 - 3.f * f1) + 18.f * y0c + 14.f * y1c - 32.f * ymid;
                    cc[c] = dtf * (f1 - 4.f * f0) - 11.f * y0c - 5.f * y1c + 16.f * ymid;
                    cd[c] = dtf * f0;
                    ce[c] = y0c;
                    y[c] = y1c;
                    k[0][c] = f1;
                }
                st_t0 = t0; st_t1 = t1;
                ++nacc;
            } else {
                ++nrej;
            }
            // controller in float64 (torchdiffeq keeps dt in float64)
            const double r64 = (double)ratio;
            if (r64 == 0.0) dt = dt * 10.0;
            else {
                const double df = r64 < 1.0 ? 1.0 : 0.2;
                dt = dt * fmin(10.0, fmax(0.9 / pow(r64, 0.2), df));
            }
            if (!(ratio == ratio)) { st = ODECOL_ST_NONFINITE; break; }
        }
        if (st != ODECOL_ST_OK) break;
        const float x = (float)((next_t - st_t0) / (st_t1 - st_t0));
        if (rec.y && i == 0) { rec.out_step[(size_t)b * T + j] = nacc - 1; rec.out_x[(size_t)b * T + j] = x; }
        if (f.act) {
            float o[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float tot = __fadd_rn(ce[c], __fmul_rn(x, cd[c]));
                float xp = __fmul_rn(x, x);
                tot = __fadd_rn(tot, __fmul_rn(xp, cc[c]));
                xp = __fmul_rn(xp, x);
                tot = __fadd_rn(tot, __fmul_rn(xp, cb[c]));
                xp = __fmul_rn(xp, x);
                tot = __fadd_rn(tot, __fmul_rn(xp, ca[c]));
                o[c] = tot;
            }
            st3(y_out + ((size_t)j * p.B + b) * row, N, i, o[0], o[1], o[2]);
        }
    }
    if (st != ODECOL_ST_OK && f.act) {
        const float qnan = __int_as_float(0x7fc00000);
        for (int jj = j; jj < T; ++jj) st3(y_out + ((size_t)jj * p.B + b) * row, N, i, qnan, qnan, qnan);
    }
    if (i == 0) {
        if (n_accept) n_accept[b] = nacc;
        if (n_reject) n_reject[b] = nrej;
        if (status) status[b] = st;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// dopri5 discrete adjoint: reverse sweep over the ACCEPTED steps recorded by the forward kernel (rejected attempts and
// the step-size controller carry no gradient: torchdiffeq computes them under no_grad).  Per step the seven stages are
// recomputed from the recorded start state, then the adjoint runs through the dense-output quartic (every requested
// output is interpolated), the FSAL solution row and the six stage rows of the tableau.
// What it replaces: loss.backward() through torchdiffeq's unrolled adaptive steps (reference scripts/xor_ode.py:177).
// ---------------------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(128) k_dopri5_bwd_small(DevProblem p, int T, Dopri5Record rec,
                                                          const int* __restrict__ n_accept,
                                                          const float* __restrict__ grad_y, const int* __restrict__ sel,
                                                          int G, float* __restrict__ grad_y0, float* __restrict__ grad_W) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.x, i = threadIdx.x, N = p.N, B = p.B;
    BwdCtx<KP> cx;
    cx.init(p, sm, b);
    for (int e = i; e < 3 * N; e += blockDim.x) cx.s.inv[e] = sel ? -1 : e;
    __syncthreads();
    if (sel) for (int g = i; g < G; g += blockDim.x) cx.s.inv[sel[g]] = g;
    __syncthreads();
    int gi[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) gi[c] = cx.act ? cx.s.inv[c * N + i] : -1;
    const size_t row = (size_t)3 * N;
    // tableau as float32, row r = coefficients of stage r+2 (last row = FSAL solution row)
    const float beta[6][6] = {
        {DP::b10, 0.f, 0.f, 0.f, 0.f, 0.f}, {DP::b20, DP::b21, 0.f, 0.f, 0.f, 0.f}, {DP::b30, DP::b31, DP::b32, 0.f, 0.f, 0.f},
        {DP::b40, DP::b41, DP::b42, DP::b43, 0.f, 0.f}, {DP::b50, DP::b51, DP::b52, DP::b53, DP::b54, 0.f},
        {DP::b60, 0.f, DP::b62, DP::b63, DP::b64, DP::b65}};
    const float alpha[4] = {DP::a1, DP::a2, DP::a3, DP::a4};
    const float cmid[7] = {DP::m0, 0.f, DP::m2, DP::m3, DP::m4, DP::m5, DP::m6};
    const int nacc = n_accept[b];
    float lam[3] = {0.f, 0.f, 0.f};        // adjoint of y1 of the step being processed
    float carry[3] = {0.f, 0.f, 0.f};      // adjoint of f1 = k7 handed back by the following step (its k1)
    int j = T - 1;
    for (int n = nacc - 1; n >= 0; --n) {
        const double t0 = rec.t0[(size_t)b * rec.cap + n], dtd = rec.dt[(size_t)b * rec.cap + n];
        const float t0f = (float)t0, dtf = (float)dtd, t1f = (float)(t0 + dtd);
        float y0[3] = {0.f, 0.f, 0.f};
        if (cx.act) { const Y3 q = ld3(rec.y + ((size_t)n * B + b) * row, N, i); y0[0] = q.V; y0[1] = q.A; y0[2] = q.F; }
        // ---- recompute the seven stages
        float k[7][3], dph[7], ys[3] = {y0[0], y0[1], y0[2]};
#pragma unroll
        for (int st = 0; st < 7; ++st) {
            if (st > 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float acc = k[0][c] * (beta[st - 1][0] * dtf);
#pragma unroll
                    for (int m = 1; m < 6; ++m)
                        if (m < st) acc = fmaf(k[m][c], beta[st - 1][m] * dtf, acc);
                    ys[c] = __fadd_rn(y0[c], acc);
                }
            }
            const float ts = st == 0 ? t0f : (st <= 4 ? __fadd_rn(t0f, __fmul_rn(alpha[st - 1], dtf)) : t1f);
            float r;
            const float tot = cx.stage_fwd(st, ts, ys[0], ys[1], r, dph[st], true);
            drift(cx.c, ys[0], ys[1], ys[2], r, cx.kappa, tot, k[st][0], k[st][1], k[st][2]);
        }
        // ys is y1 now.  ---- adjoint of the dense output for every output time inside this step
        float ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0}, cc[3] = {0, 0, 0}, cd[3] = {0, 0, 0}, ce[3] = {0, 0, 0};
        while (j >= 1 && rec.out_step[(size_t)b * T + j] == n) {
            const float x = rec.out_x[(size_t)b * T + j];
            const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float g = gi[c] >= 0 ? grad_y[((size_t)j * B + b) * G + gi[c]] : 0.f;
                ce[c] += g; cd[c] += x * g; cc[c] += x2 * g; cb[c] += x3 * g; ca[c] += x4 * g;
            }
            --j;
        }
        float kb[7][3], yb0[3], yb1[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float ymidb = 16.f * ca[c] - 32.f * cb[c] + 16.f * cc[c];
            yb0[c] = ce[c] - 8.f * ca[c] + 18.f * cb[c] - 11.f * cc[c] + ymidb;
            yb1[c] = lam[c] - 8.f * ca[c] + 14.f * cb[c] - 5.f * cc[c];
#pragma unroll
            for (int m = 0; m < 7; ++m) kb[m][c] = dtf * cmid[m] * ymidb;
            kb[0][c] += dtf * (-2.f * ca[c] + 5.f * cb[c] - 4.f * cc[c] + cd[c]);     // f0 = k1
            kb[6][c] += dtf * (2.f * ca[c] - 3.f * cb[c] + cc[c]) + carry[c];         // f1 = k7 (+ next step's k1)
        }
        // ---- stages 7 .. 2: Ybar_i = [i == 7] ybar1 + J(Y_i)^T kbar_i, then push through the tableau row
#pragma unroll
        for (int st = 6; st >= 1; --st) {
            float bV, bA, bF;
            cx.stage_bwd(st, dph[st], kb[st][0], kb[st][1], kb[st][2], bV, bA, bF);
            float Yb[3] = {bV, bA, bF};
            if (st == 6) { Yb[0] += yb1[0]; Yb[1] += yb1[1]; Yb[2] += yb1[2]; }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                yb0[c] += Yb[c];
#pragma unroll
                for (int m = 0; m < 6; ++m)
                    if (m < st) kb[m][c] += dtf * beta[st - 1][m] * Yb[c];
            }
        }
        if (n > 0) {
            carry[0] = kb[0][0]; carry[1] = kb[0][1]; carry[2] = kb[0][2];
        } else {                                   // first step: k1 = f(t[0], y0) was evaluated, not inherited
            float bV, bA, bF;
            cx.stage_bwd(0, dph[0], kb[0][0], kb[0][1], kb[0][2], bV, bA, bF);
            yb0[0] += bV; yb0[1] += bA; yb0[2] += bF;
        }
        lam[0] = yb0[0]; lam[1] = yb0[1]; lam[2] = yb0[2];
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        if (gi[c] >= 0) lam[c] += grad_y[((size_t)0 * B + b) * G + gi[c]];        // output 0 is y0 itself
    if (grad_y0 && cx.act) st3(grad_y0 + b * row, N, i, lam[0], lam[1], lam[2]);
    cx.flush_dw(p, grad_W);
}

// ---------------------------------------------------------------------------------------------------------------
// Euler-Maruyama in torchsde's integrate loop: fixed step (host increments or Philox) and step doubling
// ---------------------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(128) k_em_fwd_small(DevProblem p, const float* __restrict__ ts, int T,
                                                      const float* __restrict__ y0, float* __restrict__ y_out,
                                                      const float* __restrict__ dWs, unsigned long long seed,
                                                      long long trial_offset, float dt0, int adaptive,
                                                      float rtol, float atol, float dt_min,
                                                      int* __restrict__ n_accept, int* __restrict__ n_reject,
                                                      int* __restrict__ status, float* __restrict__ y_steps,
                                                      long long max_attempts) {
    __shared__ __align__(16) float ra[2 * KP];
    __shared__ double red[4];
    __shared__ float zs[2][kBrownianDepth + 1];
    const int b = blockIdx.x, i = threadIdx.x, N = p.N, B = p.B;
    RowRhs<KP> f;
    f.init(p, ra, b);
    const size_t row = (size_t)3 * N;
    const double cnt = 3.0 * N;
    const Philox px(seed);
    const unsigned long long trial = (unsigned long long)(trial_offset + b);
    float sg[3] = {0.f, 0.f, 0.f};
    float y[3] = {0.f, 0.f, 0.f};
    if (f.act) {
        const Y3 s = ld3(y0 + b * row, N, i);
        y[0] = s.V; y[1] = s.A; y[2] = s.F;
        st3(y_out + b * row, N, i, y[0], y[1], y[2]);
        if (y_steps) st3(y_steps + b * row, N, i, y[0], y[1], y[2]);
        if (p.sigma) { sg[0] = __ldg(p.sigma + i); sg[1] = __ldg(p.sigma + N + i); sg[2] = __ldg(p.sigma + 2 * N + i); }
        if (p.sigma_scale) { const float sc = __ldg(p.sigma_scale + b); sg[0] *= sc; sg[1] *= sc; sg[2] *= sc; }
    }
    const float t_begin = __ldg(ts), t_end = __ldg(ts + T - 1);
    const float span = t_end - t_begin;
    float curr_t = t_begin, prev_t = t_begin;
    float py[3] = {y[0], y[1], y[2]};
    double step = (double)dt0, prev_ratio = 0.0;
    bool has_prev = false;
    long long kstep = 0, attempts = 0;
    int nacc = 0, nrej = 0, st = ODECOL_ST_OK;
    float w_curr = 0.0f;                        // W(curr_t) on the virtual tree (adaptive mode)
    // two tree queries per attempt, computed cooperatively: lanes 0..24 of warp 0 draw the node deviates
    auto tree2 = [&](float ta, float tb, float& wa, float& wb) {
        uint32_t qa, qb; float fa, fb;
        brownian_path(t_begin, span, ta, qa, fa);
        brownian_path(t_begin, span, tb, qb, fb);
        if (i <= kBrownianDepth) {
            const uint32_t lvl = i == kBrownianDepth ? kBrownianRootLevel : (uint32_t)i;
            zs[0][i] = brownian_node(px, trial, lvl, i == kBrownianDepth ? 0u : (qa >> (kBrownianDepth - i)));
            zs[1][i] = brownian_node(px, trial, lvl, i == kBrownianDepth ? 0u : (qb >> (kBrownianDepth - i)));
        }
        __syncthreads();
        wa = brownian_combine(span, qa, fa, zs[0][kBrownianDepth], [&](int l) { return zs[0][l]; });
        wb = brownian_combine(span, qb, fb, zs[1][kBrownianDepth], [&](int l) { return zs[1][l]; });
        __syncthreads();
    };
    int j = 1;
    for (; j < T && st == ODECOL_ST_OK; ++j) {
        const float out_t = __ldg(ts + j);
        while (curr_t < out_t) {
            if (++attempts > max_attempts) { st = ODECOL_ST_MAXSTEPS; break; }
            const float next_t = fminf(__fadd_rn(curr_t, (float)step), t_end);
            float fv[3];
            f.eval(curr_t, y[0], y[1], y[2], fv[0], fv[1], fv[2]);
            if (!adaptive) {
                const float h = __fsub_rn(next_t, curr_t);
                float dw;
                if (dWs) dw = __ldg(dWs + (size_t)kstep * B + b);
                else {
                    const uint4 bits = px((uint32_t)trial, (uint32_t)(trial >> 32), (uint32_t)kstep, 0x80000000u | (uint32_t)(kstep >> 32));
                    dw = sqrtf(h) * normal_from_bits(bits.x, bits.y);
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    py[c] = y[c];
                    y[c] = __fadd_rn(__fadd_rn(y[c], __fmul_rn(fv[c], h)), __fmul_rn(sg[c], dw));
                }
                prev_t = curr_t; curr_t = next_t;
                ++kstep; ++nacc;
                if (y_steps && f.act) st3(y_steps + ((size_t)kstep * B + b) * row, N, i, y[0], y[1], y[2]);
            } else {
                const float mid_t = __fmul_rn(0.5f, __fadd_rn(curr_t, next_t));
                float w_mid, w_next;
                tree2(mid_t, next_t, w_mid, w_next);
                const float h = __fsub_rn(next_t, curr_t), h1 = __fsub_rn(mid_t, curr_t), h2 = __fsub_rn(next_t, mid_t);
                float yf[3], ym[3], yh[3], fm[3], q[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    yf[c] = __fadd_rn(__fadd_rn(y[c], __fmul_rn(fv[c], h)), __fmul_rn(sg[c], w_next - w_curr));
                    ym[c] = __fadd_rn(__fadd_rn(y[c], __fmul_rn(fv[c], h1)), __fmul_rn(sg[c], w_mid - w_curr));
                }
                f.eval(mid_t, ym[0], ym[1], ym[2], fm[0], fm[1], fm[2]);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    yh[c] = __fadd_rn(__fadd_rn(ym[c], __fmul_rn(fm[c], h2)), __fmul_rn(sg[c], w_next - w_mid));
                    const float tol = __fadd_rn(atol, __fmul_rn(rtol, fmaxf(fabsf(yf[c]), fabsf(yh[c]))));
                    q[c] = __fdiv_rn(__fsub_rn(yf[c], yh[c]), tol);
                }
                double s = f.act ? ((double)q[0] * q[0] + (double)q[1] * q[1] + (double)q[2] * q[2]) : 0.0;
                s = block_sum(s, red);
                const double err = (double)(float)sqrt(s / cnt);
                if (!(err == err)) { st = ODECOL_ST_NONFINITE; break; }
                // torchsde adaptive_stepping.update_step_size
                {
                    const double pfac = err > 1.0 ? 0.0 : 0.13, ifac = err > 1.0 ? 1.0 / 1.5 : 1.0 / 4.5;
                    const double ratio = 0.9 / err;
                    const double pr = has_prev ? prev_ratio : ratio;
                    double factor = pow(ratio, ifac) * pow(ratio / pr, pfac);
                    double facmin = 0.2;
                    if (err <= 1.0) { prev_ratio = ratio; has_prev = true; facmin = 1.0; }
                    factor = fmin(1.4, fmax(facmin, factor));
                    step = step * factor;
                }
                if (step < (double)dt_min) { step = (double)dt_min; has_prev = false; }
                if (err <= 1.0 || step <= (double)dt_min) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) { py[c] = y[c]; y[c] = yh[c]; }
                    prev_t = curr_t; curr_t = next_t; w_curr = w_next;
                    ++nacc;
                } else {
                    ++nrej;
                }
            }
        }
        if (st != ODECOL_ST_OK) break;
        if (f.act) {
            const float spn = __fsub_rn(curr_t, prev_t);
            const float w0 = __fdiv_rn(__fsub_rn(curr_t, out_t), spn), w1 = __fdiv_rn(__fsub_rn(out_t, prev_t), spn);
            st3(y_out + ((size_t)j * B + b) * row, N, i,
                __fadd_rn(__fmul_rn(w0, py[0]), __fmul_rn(w1, y[0])),
                __fadd_rn(__fmul_rn(w0, py[1]), __fmul_rn(w1, y[1])),
                __fadd_rn(__fmul_rn(w0, py[2]), __fmul_rn(w1, y[2])));
        }
    }
    if (st != ODECOL_ST_OK && f.act) {
        const float qnan = __int_as_float(0x7fc00000);
        for (int jj = j; jj < T; ++jj) st3(y_out + ((size_t)jj * B + b) * row, N, i, qnan, qnan, qnan);
    }
    if (i == 0) {
        if (n_accept) n_accept[b] = nacc;
        if (n_reject) n_reject[b] = nrej;
        if (status) status[b] = st;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Euler-Maruyama discrete adjoint (fixed step).  k_em_schedule replays the data-independent float32 time loop
// once: for every output j the solver state index it ends on and the two interpolation weights, and the start
// time of every step.  k_em_bwd_small then sweeps the steps in reverse.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_em_schedule(const float* __restrict__ ts, int T, float dt0, int* __restrict__ step_of,
                              float* __restrict__ w, float* __restrict__ tk) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float t_end = ts[T - 1];
    float curr_t = ts[0], prev_t = ts[0];
    int k = 0;
    step_of[0] = 0; w[0] = 0.f; w[1] = 1.f;
    tk[0] = curr_t;
    for (int j = 1; j < T; ++j) {
        const float out_t = ts[j];
        while (curr_t < out_t) {
            const float next_t = fminf(__fadd_rn(curr_t, dt0), t_end);
            prev_t = curr_t; curr_t = next_t;
            tk[++k] = curr_t;
        }
        const float spn = __fsub_rn(curr_t, prev_t);
        step_of[j] = k;
        w[2 * j] = __fdiv_rn(__fsub_rn(curr_t, out_t), spn);
        w[2 * j + 1] = __fdiv_rn(__fsub_rn(out_t, prev_t), spn);
    }
    step_of[T] = k;
}

template <int KP>
__global__ void __launch_bounds__(128) k_em_bwd_small(DevProblem p, int T, const float* __restrict__ y_steps,
                                                      const float* __restrict__ grad_y, const int* __restrict__ sel,
                                                      int G, float* __restrict__ grad_y0, float* __restrict__ grad_W,
                                                      const int* __restrict__ step_of, const float* __restrict__ w,
                                                      const float* __restrict__ tk) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.x, i = threadIdx.x, N = p.N, B = p.B;
    BwdCtx<KP> cx;
    cx.init(p, sm, b);
    for (int e = i; e < 3 * N; e += blockDim.x) cx.s.inv[e] = sel ? -1 : e;
    __syncthreads();
    if (sel) for (int g = i; g < G; g += blockDim.x) cx.s.inv[sel[g]] = g;
    __syncthreads();
    const int gV = cx.act ? cx.s.inv[i] : -1, gA = cx.act ? cx.s.inv[N + i] : -1, gF = cx.act ? cx.s.inv[2 * N + i] : -1;
    const size_t row = (size_t)3 * N;
    const int nsteps = step_of[T];
    auto gcomp = [&](int j, int g) -> float { return g >= 0 ? grad_y[((size_t)j * B + b) * G + g] : 0.f; };
    float lV = 0.f, lA = 0.f, lF = 0.f;      // adjoint of solver state k+1 while processing step k
    float pV = 0.f, pA = 0.f, pF = 0.f;      // contributions destined for state k (from interpolated outputs)
    int j = T - 1;
    for (int k = nsteps - 1; k >= 0; --k) {
        lV += pV; lA += pA; lF += pF;
        pV = pA = pF = 0.f;
        while (j >= 1 && step_of[j] == k + 1) {           // outputs whose "curr" state is k+1, "prev" is k
            const float w0 = w[2 * j], w1 = w[2 * j + 1];
            const float a = gcomp(j, gV), c2 = gcomp(j, gA), d = gcomp(j, gF);
            lV += w1 * a; lA += w1 * c2; lF += w1 * d;
            pV += w0 * a; pA += w0 * c2; pF += w0 * d;
            --j;
        }
        const float t0 = tk[k], h = __fsub_rn(tk[k + 1], tk[k]);
        float V = 0.f, A = 0.f;
        if (cx.act) { const float* yk = y_steps + ((size_t)k * B + b) * row; V = yk[i]; A = yk[N + i]; }
        float r, dr, bV, bA, bF;
        (void)cx.stage_fwd(0, t0, V, A, r, dr, false);
        cx.stage_bwd(0, dr, h * lV, h * lA, h * lF, bV, bA, bF);
        lV += bV; lA += bA; lF += bF;
        __syncthreads();
    }
    lV += pV; lA += pA; lF += pF;
    if (j >= 0) { lV += gcomp(0, gV); lA += gcomp(0, gA); lF += gcomp(0, gF); }   // output 0 is y0 itself
    if (grad_y0 && cx.act) st3(grad_y0 + b * row, N, i, lV, lA, lF);
    cx.flush_dw(p, grad_W);
}

// ---------------------------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------------------------
static inline int small_threads(int N) { return (N + 31) / 32 * 32; }

int small_kp(const DevProblem& p) {
    const int Kaug = p.N + p.n_in + 1;
    if (p.N > 128 || Kaug > 128) return 0;
    if (Kaug <= 40) return 40;
    if (Kaug <= 64) return 64;
    if (Kaug <= 96) return 96;
    return 128;
}

#define ODECOL_KP_SWITCH(kp, ...)                                  \
    switch (kp) {                                                  \
        case 40: { constexpr int KP = 40; __VA_ARGS__; } break;    \
        case 64: { constexpr int KP = 64; __VA_ARGS__; } break;    \
        case 96: { constexpr int KP = 96; __VA_ARGS__; } break;    \
        case 128: { constexpr int KP = 128; __VA_ARGS__; } break;  \
        default: return ODECOL_E_UNSUPPORTED;                      \
    }

int launch_rhs_generic(const DevProblem& p, const float* t, const float* y, float* f, cudaStream_t s) {
    const int Kaug = p.N + p.n_in + 1;
    const size_t smem = sizeof(float) * Kaug;
    if (smem > 200 * 1024) return ODECOL_E_UNSUPPORTED;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_rhs_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_rhs_generic<<<p.B, 256, smem, s>>>(p, t, y, f);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

// two trials per warp when a trial needs at most half a warp (the reference's two-column WTA network)
static inline bool small_packed(const DevProblem& p) { return p.N <= 16 && p.n_in <= 16; }

int launch_rk4_fwd_small(const DevProblem& p, const float* t, int T, const float* y0, float* y_out, int out_every,
                         cudaStream_t s, const unsigned int* run_if) {
    const int kp = small_kp(p);
    const int packed = small_packed(p) ? 1 : 0;
    const int grid = packed ? (p.B + 1) / 2 : p.B;
    ODECOL_KP_SWITCH(kp, (k_rk4_fwd_small<KP><<<grid, small_threads(p.N), 0, s>>>(p, t, T, y0, y_out, out_every, packed, run_if)));
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

int launch_rk4_bwd_small(const DevProblem& p, const float* t, int T, const float* y_traj, const float* grad_y,
                         const int* sel, int G, float* grad_y0, float* grad_W, cudaStream_t s) {
    const int kp = small_kp(p);
    const int packed = small_packed(p) ? 1 : 0;
    const size_t smem = small_bwd_smem_bytes(p.N, kp, packed ? 2 : 1);
    if (kp == 40 && tiny_rk4_applicable(p)) {       // the batches whose forward solve ran on the tensor cores (tiny_tc.cu)
        cudaFuncSetAttribute(k_rk4_bwd_small<40, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_rk4_bwd_small<40, true><<<packed ? (p.B + 1) / 2 : p.B, small_threads(p.N), smem, s>>>(p, t, T, y_traj, grad_y, sel, G,
                                                                                              grad_y0, grad_W, packed);
        count_launch();
        return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
    }
    ODECOL_KP_SWITCH(kp, {
        cudaFuncSetAttribute(k_rk4_bwd_small<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_rk4_bwd_small<KP><<<packed ? (p.B + 1) / 2 : p.B, small_threads(p.N), smem, s>>>(p, t, T, y_traj, grad_y, sel, G,
                                                                                        grad_y0, grad_W, packed);
    });
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

int launch_dopri5_fwd_small(const DevProblem& p, const float* t, int T, const float* y0, float* y_out, float rtol,
                            float atol, int max_steps, int* n_accept, int* n_reject, int* status, const Dopri5Record& rec,
                            cudaStream_t s) {
    const int kp = small_kp(p);
    ODECOL_KP_SWITCH(kp, (k_dopri5_fwd_small<KP><<<p.B, small_threads(p.N), 0, s>>>(p, t, T, y0, y_out, rtol, atol, max_steps,
                                                                                   n_accept, n_reject, status, rec)));
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

int launch_dopri5_bwd_small(const DevProblem& p, int T, const Dopri5Record& rec, const int* n_accept, const float* grad_y,
                            const int* sel, int G, float* grad_y0, float* grad_W, cudaStream_t s) {
    const int kp = small_kp(p);
    const size_t smem = small_bwd_smem_bytes(p.N, kp, 1);
    ODECOL_KP_SWITCH(kp, {
        cudaFuncSetAttribute(k_dopri5_bwd_small<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_dopri5_bwd_small<KP><<<p.B, small_threads(p.N), smem, s>>>(p, T, rec, n_accept, grad_y, sel, G, grad_y0, grad_W);
    });
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

int launch_em_fwd_small(const DevProblem& p, const float* ts, int T, const float* y0, float* y_out, const float* dW,
                        uint64_t seed, int64_t trial_offset, float dt, int adaptive, float rtol, float atol,
                        float dt_min, int* n_accept, int* n_reject, int* status, float* y_steps,
                        long long max_attempts, cudaStream_t s) {
    const int kp = small_kp(p);
    ODECOL_KP_SWITCH(kp, (k_em_fwd_small<KP><<<p.B, small_threads(p.N), 0, s>>>(
                             p, ts, T, y0, y_out, dW, (unsigned long long)seed, (long long)trial_offset, dt, adaptive, rtol,
                             atol, dt_min, n_accept, n_reject, status, y_steps, max_attempts)));
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

int launch_em_schedule(const float* ts, int T, float dt, int* step_of, float* w, float* tk, cudaStream_t s) {
    k_em_schedule<<<1, 32, 0, s>>>(ts, T, dt, step_of, w, tk);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

int launch_em_bwd_small(const DevProblem& p, const float* ts, int T, const float* y_steps, const float* grad_y,
                        const int* sel, int G, float* grad_y0, float* grad_W, const int* step_of, const float* w,
                        const float* tk, cudaStream_t s) {
    (void)ts;
    const int kp = small_kp(p);
    const size_t smem = small_bwd_smem_bytes(p.N, kp, 1);
    ODECOL_KP_SWITCH(kp, {
        cudaFuncSetAttribute(k_em_bwd_small<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_em_bwd_small<KP><<<p.B, small_threads(p.N), smem, s>>>(p, T, y_steps, grad_y, sel, G, grad_y0, grad_W, step_of, w, tk);
    });
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------
// torchsde method='srk' for scalar noise: Roessler's SRI2 ("SRID2" tableau), fixed step, in torchsde's integrate loop.
// What every committed sdeint call of the reference names (scripts/wta_ode.py:174,200; plotting_results.py:391,506,594).
// The diffusion of the column networks is state independent (src/coupled_columns.py:239-249, 444-454, 790-800), so the
// H1 stage values (which only feed g) are never needed and g_prod = sigma * g_weight; the drift sees the space-time
// Levy area U through the third stage.  Every float32 operation of diagonal_or_scalar_step is replayed in its order.
// The fourth stage has alpha = 0 and H0 = y0: its drift contributes an exact zero and is not evaluated.
// ---------------------------------------------------------------------------------------------------------------
// Srid2, srk_increments, srk_g_weights, srk_h2: odecol_common.cuh (shared with the staged family)

template <int KP>
__global__ void __launch_bounds__(FwdBounds<KP>::threads, FwdBounds<KP>::blocks) k_srk_fwd_small(DevProblem p, const float* __restrict__ ts, int T,
                                                       const float* __restrict__ y0, float* __restrict__ y_out,
                                                       const float* __restrict__ dWs, const float* __restrict__ dUs,
                                                       unsigned long long seed, long long trial_offset, float dt0,
                                                       int* __restrict__ status, float* __restrict__ y_steps, int packed) {
    __shared__ __align__(16) float ra[4 * KP];
    const int sub = packed ? (int)(threadIdx.x >> 4) : 0, i = packed ? (int)(threadIdx.x & 15) : (int)threadIdx.x;
    const int b = packed ? 2 * (int)blockIdx.x + sub : (int)blockIdx.x, N = p.N, B = p.B;
    const int bs = b < B ? b : B - 1;            // in-bounds trial index for the (uniform) increment loads of an idle half
    RowRhs<KP> f;
    f.init(p, ra, b, i, sub, packed ? 16 : (int)blockDim.x);
    const size_t row = (size_t)3 * N;
    const Philox px(seed);
    const unsigned long long trial = (unsigned long long)(trial_offset + b);
    float sg[3] = {0.f, 0.f, 0.f};
    float y[3] = {0.f, 0.f, 0.f};
    if (f.act) {
        const Y3 s = ld3(y0 + b * row, N, i);
        y[0] = s.V; y[1] = s.A; y[2] = s.F;
        st3(y_out + b * row, N, i, y[0], y[1], y[2]);
        if (y_steps) st3(y_steps + b * row, N, i, y[0], y[1], y[2]);
        if (p.sigma) { sg[0] = __ldg(p.sigma + i); sg[1] = __ldg(p.sigma + N + i); sg[2] = __ldg(p.sigma + 2 * N + i); }
        if (p.sigma_scale) { const float sc = __ldg(p.sigma_scale + b); sg[0] *= sc; sg[1] *= sc; sg[2] *= sc; }
    }
    const float t_end = __ldg(ts + T - 1);
    float curr_t = __ldg(ts), prev_t = curr_t;
    float py[3] = {y[0], y[1], y[2]};
    long long kstep = 0;
    for (int j = 1; j < T; ++j) {
        const float out_t = __ldg(ts + j);
        while (curr_t < out_t) {
            const float next_t = fminf(__fadd_rn(curr_t, dt0), t_end);
            const float h = __fsub_rn(next_t, curr_t), rdt = __fdiv_rn(1.0f, h);
            float dw, du, gw[4];
            srk_increments(dWs, dUs, px, trial, kstep, B, bs, h, dw, du);
            srk_g_weights(h, dw, du, gw);
            float f0[3], f1[3], f2[3], H[3];
            f.eval(curr_t, y[0], y[1], y[2], f0[0], f0[1], f0[2]);
#pragma unroll
            for (int c = 0; c < 3; ++c) H[c] = __fadd_rn(y[c], __fmul_rn(f0[c], h));
            f.eval(__fadd_rn(curr_t, h), H[0], H[1], H[2], f1[0], f1[1], f1[2]);
#pragma unroll
            for (int c = 0; c < 3; ++c) H[c] = srk_h2(y[c], f0[c], f1[c], sg[c], h, du, rdt);
            f.eval(__fadd_rn(curr_t, __fmul_rn(Srid2::half, h)), H[0], H[1], H[2], f2[0], f2[1], f2[2]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                py[c] = y[c];
                float v = __fadd_rn(__fadd_rn(y[c], __fmul_rn(__fmul_rn(Srid2::a0, f0[c]), h)), __fmul_rn(sg[c], gw[0]));
                v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(Srid2::a1, f1[c]), h)), __fmul_rn(sg[c], gw[1]));
                v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(Srid2::a2, f2[c]), h)), __fmul_rn(sg[c], gw[2]));
                y[c] = __fadd_rn(v, __fmul_rn(sg[c], gw[3]));
            }
            prev_t = curr_t; curr_t = next_t;
            ++kstep;
            if (y_steps && f.act) st3(y_steps + ((size_t)kstep * B + b) * row, N, i, y[0], y[1], y[2]);
        }
        if (f.act) {
            const float spn = __fsub_rn(curr_t, prev_t);
            const float w0 = __fdiv_rn(__fsub_rn(curr_t, out_t), spn), w1 = __fdiv_rn(__fsub_rn(out_t, prev_t), spn);
            st3(y_out + ((size_t)j * B + b) * row, N, i,
                __fadd_rn(__fmul_rn(w0, py[0]), __fmul_rn(w1, y[0])),
                __fadd_rn(__fmul_rn(w0, py[1]), __fmul_rn(w1, y[1])),
                __fadd_rn(__fmul_rn(w0, py[2]), __fmul_rn(w1, y[2])));
        }
    }
    const bool bad_lane = f.act && !(isfinite(y[0]) && isfinite(y[1]) && isfinite(y[2]));
    bool bad;
    if (packed) {                                   // one warp, two trials: the verdict of this lane's half
        const unsigned m = __ballot_sync(0xffffffffu, bad_lane);
        bad = ((sub ? (m >> 16) : m) & 0xffffu) != 0u;
    } else {
        bad = block_any(bad_lane);
    }
    if (i == 0 && status && b < B) status[b] = bad ? ODECOL_ST_NONFINITE : ODECOL_ST_OK;
}

// Discrete adjoint of the fixed-step SRI2 solve.  With additive noise the step is
//   H1 = y + h f0,  H2 = y + h/4 (f0 + f1) + 3/2 sigma U / h,  y1 = y + h (f0/6 + f1/6 + 2/3 f2) + sigma sum(g_weight)
// so the Brownian inputs enter the Jacobians only through H2: the reverse sweep recomputes the three stages from the
// saved solver states and the same (W, U) (host tables or the same Philox stream) and pushes the adjoint through them.
// Replaces loss.backward() through torchsde's unrolled srk steps (reference scripts/wta_ode.py:174-181).
template <int KP>
__global__ void __launch_bounds__(BwdBounds<KP>::threads, BwdBounds<KP>::blocks) k_srk_bwd_small(DevProblem p, int T, const float* __restrict__ y_steps,
                                                       const float* __restrict__ dWs, const float* __restrict__ dUs,
                                                       unsigned long long seed, long long trial_offset,
                                                       const float* __restrict__ grad_y, const int* __restrict__ sel,
                                                       int G, float* __restrict__ grad_y0, float* __restrict__ grad_W,
                                                       const int* __restrict__ step_of, const float* __restrict__ w,
                                                       const float* __restrict__ tk, int packed) {
    extern __shared__ __align__(16) float sm[];
    const int sub = packed ? (int)(threadIdx.x >> 4) : 0, i = packed ? (int)(threadIdx.x & 15) : (int)threadIdx.x;
    const int b = packed ? 2 * (int)blockIdx.x + sub : (int)blockIdx.x, N = p.N, B = p.B;
    const int bs = b < B ? b : B - 1;
    BwdCtx<KP> cx;
    cx.init(p, sm, b, i, sub, packed ? 16 : (int)blockDim.x, packed ? 2 : 1);
    for (int e = threadIdx.x; e < 3 * N; e += blockDim.x) cx.s.inv[e] = sel ? -1 : e;
    __syncthreads();
    if (sel) for (int g = threadIdx.x; g < G; g += blockDim.x) cx.s.inv[sel[g]] = g;
    __syncthreads();
    int gi[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) gi[c] = cx.act ? cx.s.inv[c * N + i] : -1;
    const size_t row = (size_t)3 * N;
    const int nsteps = step_of[T];
    const Philox px(seed);
    const unsigned long long trial = (unsigned long long)(trial_offset + b);
    float sg[3] = {0.f, 0.f, 0.f};
    if (cx.act && p.sigma) { sg[0] = __ldg(p.sigma + i); sg[1] = __ldg(p.sigma + N + i); sg[2] = __ldg(p.sigma + 2 * N + i); }
    if (p.sigma_scale) { const float sc = __ldg(p.sigma_scale + bs); sg[0] *= sc; sg[1] *= sc; sg[2] *= sc; }
    auto gcomp = [&](int j, int c) -> float { return gi[c] >= 0 ? grad_y[((size_t)j * B + b) * G + gi[c]] : 0.f; };
    float lam[3] = {0.f, 0.f, 0.f};      // adjoint of solver state k+1 while processing step k
    float pend[3] = {0.f, 0.f, 0.f};     // contributions destined for state k (from interpolated outputs)
    int j = T - 1;
    for (int k = nsteps - 1; k >= 0; --k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { lam[c] += pend[c]; pend[c] = 0.f; }
        while (j >= 1 && step_of[j] == k + 1) {
            const float w0 = w[2 * j], w1 = w[2 * j + 1];
#pragma unroll
            for (int c = 0; c < 3; ++c) { const float g = gcomp(j, c); lam[c] += w1 * g; pend[c] += w0 * g; }
            --j;
        }
        const float t0 = tk[k], h = __fsub_rn(tk[k + 1], tk[k]), rdt = __fdiv_rn(1.0f, h);
        float dw, du;
        srk_increments(dWs, dUs, px, trial, (long long)k, B, bs, h, dw, du);
        float y[3] = {0.f, 0.f, 0.f};
        if (cx.act) { const Y3 q = ld3(y_steps + ((size_t)k * B + b) * row, N, i); y[0] = q.V; y[1] = q.A; y[2] = q.F; }
        // ---- recompute the stages (same arithmetic as the forward kernel)
        float r, d0, d1, d2, tot, f0[3], f1[3], H[3];
        tot = cx.stage_fwd(0, t0, y[0], y[1], r, d0, true);
        drift(cx.c, y[0], y[1], y[2], r, cx.kappa, tot, f0[0], f0[1], f0[2]);
#pragma unroll
        for (int c = 0; c < 3; ++c) H[c] = __fadd_rn(y[c], __fmul_rn(f0[c], h));
        tot = cx.stage_fwd(1, __fadd_rn(t0, h), H[0], H[1], r, d1, true);
        drift(cx.c, H[0], H[1], H[2], r, cx.kappa, tot, f1[0], f1[1], f1[2]);
#pragma unroll
        for (int c = 0; c < 3; ++c) H[c] = srk_h2(y[c], f0[c], f1[c], sg[c], h, du, rdt);
        (void)cx.stage_fwd(2, __fadd_rn(t0, __fmul_rn(Srid2::half, h)), H[0], H[1], r, d2, false);
        // ---- reverse
        const float h23 = h * Srid2::a2, h6 = h * Srid2::a0, h4 = h * 0.25f;
        float Y2[3], Y1[3], Y0[3];
        cx.stage_bwd(2, d2, h23 * lam[0], h23 * lam[1], h23 * lam[2], Y2[0], Y2[1], Y2[2]);
        cx.stage_bwd(1, d1, h6 * lam[0] + h4 * Y2[0], h6 * lam[1] + h4 * Y2[1], h6 * lam[2] + h4 * Y2[2], Y1[0], Y1[1], Y1[2]);
        cx.stage_bwd(0, d0, h6 * lam[0] + h4 * Y2[0] + h * Y1[0], h6 * lam[1] + h4 * Y2[1] + h * Y1[1],
                     h6 * lam[2] + h4 * Y2[2] + h * Y1[2], Y0[0], Y0[1], Y0[2]);
#pragma unroll
        for (int c = 0; c < 3; ++c) lam[c] += Y2[c] + Y1[c] + Y0[c];
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) lam[c] += pend[c];
    if (j >= 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) lam[c] += gcomp(0, c);      // output 0 is y0 itself
    }
    if (grad_y0 && cx.act) st3(grad_y0 + b * row, N, i, lam[0], lam[1], lam[2]);
    cx.flush_dw(p, grad_W);
}

// ---------------------------------------------------------------------------------------------------------------
// torchsde sdeint(method='srk', adaptive=True): step doubling with the SRI2 step on the Levy-area-consistent virtual
// Brownian tree (odecol_common.cuh) and torchsde's controller, per trial -- what the reference's own "avoid the
// artefacts" option literally is (scripts/parity_ode.py:234, README.md:28-29).  Per attempt: f(t0, y) once (shared by the
// full step and the first half step), two more drift evaluations for the full step, two for the first and three for the
// second half step; the error estimate, the controller and the interpolated outputs are those of k_em_fwd_small.
// ---------------------------------------------------------------------------------------------------------------
template <int KP>
struct SrkStepper {
    RowRhs<KP>& f;
    const float (&sg)[3];
    // y1 = SRI2 step from (t0, y) over h with increments (dw, du); f0 = f(t0, y) is given
    ODECOL_DEVINL void step(float t0, float h, const float (&y)[3], const float (&f0)[3], float dw, float du, float (&y1)[3]) {
        const float rdt = __fdiv_rn(1.0f, h);
        float gw[4], f1[3], f2[3], H[3];
        srk_g_weights(h, dw, du, gw);
#pragma unroll
        for (int c = 0; c < 3; ++c) H[c] = __fadd_rn(y[c], __fmul_rn(f0[c], h));
        f.eval(__fadd_rn(t0, h), H[0], H[1], H[2], f1[0], f1[1], f1[2]);
#pragma unroll
        for (int c = 0; c < 3; ++c) H[c] = srk_h2(y[c], f0[c], f1[c], sg[c], h, du, rdt);
        f.eval(__fadd_rn(t0, __fmul_rn(Srid2::half, h)), H[0], H[1], H[2], f2[0], f2[1], f2[2]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __fadd_rn(__fadd_rn(y[c], __fmul_rn(__fmul_rn(Srid2::a0, f0[c]), h)), __fmul_rn(sg[c], gw[0]));
            v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(Srid2::a1, f1[c]), h)), __fmul_rn(sg[c], gw[1]));
            v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(Srid2::a2, f2[c]), h)), __fmul_rn(sg[c], gw[2]));
            y1[c] = __fadd_rn(v, __fmul_rn(sg[c], gw[3]));
        }
    }
};

template <int KP>
__global__ void __launch_bounds__(128) k_srk_adaptive_small(DevProblem p, const float* __restrict__ ts, int T,
                                                            const float* __restrict__ y0, float* __restrict__ y_out,
                                                            unsigned long long seed, long long trial_offset, float dt0,
                                                            float rtol, float atol, float dt_min, int* __restrict__ n_accept,
                                                            int* __restrict__ n_reject, int* __restrict__ status,
                                                            long long max_attempts) {
    __shared__ __align__(16) float ra[2 * KP];
    __shared__ double red[4];
    __shared__ float zs[2][kBrownianDepth + 1][2];
    const int b = blockIdx.x, i = threadIdx.x, N = p.N, B = p.B;
    RowRhs<KP> f;
    f.init(p, ra, b);
    const size_t row = (size_t)3 * N;
    const double cnt = 3.0 * N;
    const Philox px(seed);
    const unsigned long long trial = (unsigned long long)(trial_offset + b);
    float sg[3] = {0.f, 0.f, 0.f};
    float y[3] = {0.f, 0.f, 0.f};
    if (f.act) {
        const Y3 s = ld3(y0 + b * row, N, i);
        y[0] = s.V; y[1] = s.A; y[2] = s.F;
        st3(y_out + b * row, N, i, y[0], y[1], y[2]);
        if (p.sigma) { sg[0] = __ldg(p.sigma + i); sg[1] = __ldg(p.sigma + N + i); sg[2] = __ldg(p.sigma + 2 * N + i); }
        if (p.sigma_scale) { const float sc = __ldg(p.sigma_scale + b); sg[0] *= sc; sg[1] *= sc; sg[2] *= sc; }
    }
    SrkStepper<KP> srk{f, sg};
    const float t_begin = __ldg(ts), t_end = __ldg(ts + T - 1);
    const float span = t_end - t_begin;
    float curr_t = t_begin, prev_t = t_begin;
    float py[3] = {y[0], y[1], y[2]};
    double step = (double)dt0, prev_ratio = 0.0;
    bool has_prev = false;
    long long attempts = 0;
    int nacc = 0, nrej = 0, st = ODECOL_ST_OK;
    double w_curr = 0.0, i_curr = 0.0;          // W(curr_t), I(curr_t) on the Levy tree
    // two tree queries per attempt, computed cooperatively: lanes 0..24 draw the node deviates of both paths
    auto tree2 = [&](float ta, float tb, double& wa, double& ia, double& wb, double& ib) {
        uint32_t qa, qb; float fa, fb;
        brownian_path(t_begin, span, ta, qa, fa);
        brownian_path(t_begin, span, tb, qb, fb);
        if (i <= kBrownianDepth) {
            const bool root = i == kBrownianDepth;
            const uint32_t lvl = root ? kLevyRootLevel : (uint32_t)i;
            levy_node(px, trial, lvl, root || i == 0 ? 0u : (qa >> (kBrownianDepth - i)), zs[0][i][0], zs[0][i][1]);
            levy_node(px, trial, lvl, root || i == 0 ? 0u : (qb >> (kBrownianDepth - i)), zs[1][i][0], zs[1][i][1]);
        }
        __syncthreads();
        levy_combine(span, qa, fa, [&](int l, float& z, float& n) { z = zs[0][l][0]; n = zs[0][l][1]; }, wa, ia);
        levy_combine(span, qb, fb, [&](int l, float& z, float& n) { z = zs[1][l][0]; n = zs[1][l][1]; }, wb, ib);
        __syncthreads();
    };
    int j = 1;
    for (; j < T && st == ODECOL_ST_OK; ++j) {
        const float out_t = __ldg(ts + j);
        while (curr_t < out_t) {
            if (++attempts > max_attempts) { st = ODECOL_ST_MAXSTEPS; break; }
            const float next_t = fminf(__fadd_rn(curr_t, (float)step), t_end);
            const float mid_t = __fmul_rn(0.5f, __fadd_rn(curr_t, next_t));
            double w_mid, i_mid, w_next, i_next;
            tree2(mid_t, next_t, w_mid, i_mid, w_next, i_next);
            const float h = __fsub_rn(next_t, curr_t), h1 = __fsub_rn(mid_t, curr_t), h2 = __fsub_rn(next_t, mid_t);
            // (W, U) of the full step and of the two halves: U(a, b) = I(b) - I(a) - (b - a) W(a)
            const float dwf = (float)(w_next - w_curr), duf = (float)(i_next - i_curr - (double)h * w_curr);
            const float dw1 = (float)(w_mid - w_curr), du1 = (float)(i_mid - i_curr - (double)h1 * w_curr);
            const float dw2 = (float)(w_next - w_mid), du2 = (float)(i_next - i_mid - (double)h2 * w_mid);
            float f0[3], fm[3], yf[3], ym[3], yh[3], q[3];
            f.eval(curr_t, y[0], y[1], y[2], f0[0], f0[1], f0[2]);
            srk.step(curr_t, h, y, f0, dwf, duf, yf);
            srk.step(curr_t, h1, y, f0, dw1, du1, ym);
            f.eval(mid_t, ym[0], ym[1], ym[2], fm[0], fm[1], fm[2]);
            srk.step(mid_t, h2, ym, fm, dw2, du2, yh);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float tol = __fadd_rn(atol, __fmul_rn(rtol, fmaxf(fabsf(yf[c]), fabsf(yh[c]))));
                q[c] = __fdiv_rn(__fsub_rn(yf[c], yh[c]), tol);
            }
            double s = f.act ? ((double)q[0] * q[0] + (double)q[1] * q[1] + (double)q[2] * q[2]) : 0.0;
            s = block_sum(s, red);
            const double err = (double)(float)sqrt(s / cnt);
            if (!(err == err)) { st = ODECOL_ST_NONFINITE; break; }
            {                                    // torchsde adaptive_stepping.update_step_size
                const double pfac = err > 1.0 ? 0.0 : 0.13, ifac = err > 1.0 ? 1.0 / 1.5 : 1.0 / 4.5;
                const double ratio = 0.9 / err;
                const double pr = has_prev ? prev_ratio : ratio;
                double factor = pow(ratio, ifac) * pow(ratio / pr, pfac);
                double facmin = 0.2;
                if (err <= 1.0) { prev_ratio = ratio; has_prev = true; facmin = 1.0; }
                factor = fmin(1.4, fmax(facmin, factor));
                step = step * factor;
            }
            if (step < (double)dt_min) { step = (double)dt_min; has_prev = false; }
            if (err <= 1.0 || step <= (double)dt_min) {
#pragma unroll
                for (int c = 0; c < 3; ++c) { py[c] = y[c]; y[c] = yh[c]; }
                prev_t = curr_t; curr_t = next_t; w_curr = w_next; i_curr = i_next;
                ++nacc;
            } else {
                ++nrej;
            }
        }
        if (st != ODECOL_ST_OK) break;
        if (f.act) {
            const float spn = __fsub_rn(curr_t, prev_t);
            const float w0 = __fdiv_rn(__fsub_rn(curr_t, out_t), spn), w1 = __fdiv_rn(__fsub_rn(out_t, prev_t), spn);
            st3(y_out + ((size_t)j * B + b) * row, N, i,
                __fadd_rn(__fmul_rn(w0, py[0]), __fmul_rn(w1, y[0])),
                __fadd_rn(__fmul_rn(w0, py[1]), __fmul_rn(w1, y[1])),
                __fadd_rn(__fmul_rn(w0, py[2]), __fmul_rn(w1, y[2])));
        }
    }
    if (st != ODECOL_ST_OK && f.act) {
        const float qnan = __int_as_float(0x7fc00000);
        for (int jj = j; jj < T; ++jj) st3(y_out + ((size_t)jj * B + b) * row, N, i, qnan, qnan, qnan);
    }
    if (i == 0) {
        if (n_accept) n_accept[b] = nacc;
        if (n_reject) n_reject[b] = nrej;
        if (status) status[b] = st;
    }
}

int launch_srk_adaptive_small(const DevProblem& p, const float* ts, int T, const float* y0, float* y_out, uint64_t seed,
                              int64_t trial_offset, float dt, float rtol, float atol, float dt_min, int* n_accept,
                              int* n_reject, int* status, long long max_attempts, cudaStream_t s) {
    const int kp = small_kp(p);
    ODECOL_KP_SWITCH(kp, (k_srk_adaptive_small<KP><<<p.B, small_threads(p.N), 0, s>>>(
                             p, ts, T, y0, y_out, (unsigned long long)seed, (long long)trial_offset, dt, rtol, atol, dt_min,
                             n_accept, n_reject, status, max_attempts)));
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

// W(t[m]), I(t[m]) of the Levy-area-consistent tree (odecol_brownian_levy_query): one thread per (query time, trial)
__global__ void k_brownian_levy_query(unsigned long long seed, long long trial_offset, int B, float t_begin, float span,
                                      const float* __restrict__ t, int M, double* __restrict__ w, double* __restrict__ iw) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= (long long)M * B) return;
    const int m = (int)(e / B), b = (int)(e % B);
    const Philox px(seed);
    levy_tree(px, (unsigned long long)(trial_offset + b), t_begin, span, __ldg(t + m), w[e], iw[e]);
}

int launch_brownian_levy_query(uint64_t seed, int64_t trial_offset, int B, float t_begin, float span, const float* t, int M,
                               double* w, double* iw, cudaStream_t s) {
    const long long total = (long long)M * B;
    k_brownian_levy_query<<<(unsigned)((total + 127) / 128), 128, 0, s>>>((unsigned long long)seed, (long long)trial_offset, B,
                                                                           t_begin, span, t, M, w, iw);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

int launch_srk_fwd_small(const DevProblem& p, const float* ts, int T, const float* y0, float* y_out, const float* dW,
                         const float* dU, uint64_t seed, int64_t trial_offset, float dt, int* status, float* y_steps,
                         cudaStream_t s) {
    const int kp = small_kp(p);
    const int packed = small_packed(p) ? 1 : 0;
    ODECOL_KP_SWITCH(kp, (k_srk_fwd_small<KP><<<packed ? (p.B + 1) / 2 : p.B, small_threads(p.N), 0, s>>>(
                             p, ts, T, y0, y_out, dW, dU, (unsigned long long)seed, (long long)trial_offset, dt, status, y_steps,
                             packed)));
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

int launch_srk_bwd_small(const DevProblem& p, int T, const float* y_steps, const float* dW, const float* dU, uint64_t seed,
                         int64_t trial_offset, const float* grad_y, const int* sel, int G, float* grad_y0, float* grad_W,
                         const int* step_of, const float* w, const float* tk, cudaStream_t s) {
    const int kp = small_kp(p);
    const int packed = small_packed(p) ? 1 : 0;
    const size_t smem = small_bwd_smem_bytes(p.N, kp, packed ? 2 : 1);
    ODECOL_KP_SWITCH(kp, {
        cudaFuncSetAttribute(k_srk_bwd_small<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_srk_bwd_small<KP><<<packed ? (p.B + 1) / 2 : p.B, small_threads(p.N), smem, s>>>(
            p, T, y_steps, dW, dU, (unsigned long long)seed, (long long)trial_offset, grad_y, sel, G, grad_y0, grad_W, step_of, w,
            tk, packed);
    });
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------
// W(t) of the virtual Brownian tree the adaptive Euler-Maruyama solvers draw from (odecol_brownian_query): one thread
// per (query time, trial), the sequential walk every kernel family uses.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_brownian_query(unsigned long long seed, long long trial_offset, int B, float t_begin, float span,
                                 const float* __restrict__ t, int M, float* __restrict__ w) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= (long long)M * B) return;
    const int m = (int)(e / B), b = (int)(e % B);
    const Philox px(seed);
    w[e] = brownian_tree(px, (unsigned long long)(trial_offset + b), t_begin, span, __ldg(t + m));
}

int launch_brownian_query(uint64_t seed, int64_t trial_offset, int B, float t_begin, float span, const float* t, int M,
                          float* w, cudaStream_t s) {
    const long long total = (long long)M * B;
    k_brownian_query<<<(unsigned)((total + 127) / 128), 128, 0, s>>>((unsigned long long)seed, (long long)trial_offset, B, t_begin,
                                                                      span, t, M, w);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

}  // namespace odecol
