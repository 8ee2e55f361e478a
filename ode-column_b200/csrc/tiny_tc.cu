// Kernel family S, large batches of the smallest networks (N <= 16, n_in <= 16: the reference's two-column WTA network):
// rk4 forward with the TRIALS on the M axis of the tensor core.
//
// The on-chip kernels give a trial half a warp and spend ~140 warp instructions per stage on two trials (the dot product
// against r_aug in shared memory, the transfer function in IEEE arithmetic, one barrier).  At tens of thousands of trials
// the better shape is the transpose: one CTA = 128 trials = the 128 lanes of tensor memory,
//
//     D[trial][i] = sum_k R_aug[trial][k] * W_aug[i][k]        tcgen05.mma, M = 128 trials, N = 16 populations, K = 48
//
// with BOTH operands written by the CTA itself into shared memory (no TMA: the operand of a stage is 10 bytes per trial
// and exists only in registers) in the K-major 128-byte-swizzle layout, as FP16 pairs (stage_tc.cuh: x = xh + xl / 2048,
// three kind::f16 products, FP32 accumulation).  Thread (trial, half) owns populations 8 half .. 8 half + 7 of its trial
// for the whole time loop -- state and Runge-Kutta slopes in registers -- and writes exactly one 16-byte chunk of rates
// and one of stimulus channels per stage; K layout: [0,16) rates, [16,32) stimulus channels, 32 the constant one.
// Per stage: phi -> operand chunks -> barrier -> nine MMAs by one thread -> commit -> everyone reads its 8 + 8 accumulator
// columns (warps w and w + 4 share the lane quarter w) -> slopes.  ~25 warp instructions per trial and stage.
//
// Values beyond FP16's range raise *ovf; the caller then runs the on-chip kernel over the same solve (launch_rk4_fwd_small
// with run_if), so results never depend on the format holding the data.
#include "stage_tc.cuh"

namespace odecol {
namespace tc {

constexpr int TY_TRIALS = 128;           // trials per CTA = TMEM lanes
constexpr int TY_THREADS = 256;          // two threads per trial (8 populations each)
constexpr int TY_KSTEPS = 3;             // K = 48 halves used of the 64-half row (33 columns of W_aug)
constexpr uint32_t TY_A_BYTES = TY_TRIALS * 128;     // one operand plane: 128 rows of 128 bytes
constexpr uint32_t TY_B_BYTES = 16 * 128;

// byte offset of 16-byte chunk c of row r in a K-major tile with the 128-byte swizzle (what TMA would have written)
ODECOL_DEVINL uint32_t sw128(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

ODECOL_DEVINL void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
ODECOL_DEVINL void tmem_ld8(uint32_t taddr, float (&f)[8]) {
    uint32_t u[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(u[i]);
}

// eight values -> one chunk of the high plane and one of the low plane
ODECOL_DEVINL void pack8(const float (&x)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = x[2 * j], b = x[2 * j + 1];
        const __half2 hh = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn((a - hf.x) * 2048.0f, (b - hf.y) * 2048.0f);
        h[j] = *reinterpret_cast<const uint32_t*>(&hh);
        l[j] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(TY_THREADS, 2)
k_rk4_fwd_tiny(DevProblem p, const float* __restrict__ t, int T, const float* __restrict__ y0, float* __restrict__ y_out,
               int out_every, unsigned int* __restrict__ ovf) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float red[TY_THREADS / 32];
    int tid;                                                        // read %tid.x ONCE (opaque move: no re-reads of the special register)
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int warp = tid >> 5, lane = tid & 31;
    const int tr = tid & (TY_TRIALS - 1), half = tid >> 7;          // trial within the CTA, population half
    int b = blockIdx.x * TY_TRIALS + tr;
    asm volatile("mov.u32 %0, %1;" : "=r"(b) : "r"(b));              // likewise %ctaid.x
    const bool live = b < p.B;
    const int N = p.N, n_in = p.n_in;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_hi = base, a_lo = base + TY_A_BYTES, b_hi = base + 2 * TY_A_BYTES, b_lo = b_hi + TY_B_BYTES;
    float2* stim = reinterpret_cast<float2*>(smem_raw + (base - smem_u32(smem_raw)) + 2 * TY_A_BYTES + 2 * TY_B_BYTES);   // [16 ch][128 trials] (y0, slope)
    const uint32_t mb = smem_u32(&bar);

    // ---- weights: max |W_aug| -> power-of-two scale -> FP16 pair in the tile layout, columns permuted to the fixed K layout
    const int Kaug = N + n_in + 1;
    float wm = 0.f;
    for (int e = tid; e < N * Kaug; e += TY_THREADS) {
        const float w = fabsf(__ldg(p.W_aug + (size_t)(e / Kaug) * p.ld_w + e % Kaug));
        if (w == w && w < 3.0e38f) wm = fmaxf(wm, w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    if (lane == 0) red[warp] = wm;
    if (tid == 0) { mbar_init(mb, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // zero both trial planes and the weight tiles (padding rows / columns must be zeros, not stale shared memory)
    for (uint32_t o = tid * 16; o < 2 * TY_A_BYTES + 2 * TY_B_BYTES; o += TY_THREADS * 16) sts128(base + o, make_uint4(0, 0, 0, 0));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
#pragma unroll
    for (int w = 0; w < TY_THREADS / 32; ++w) wm = fmaxf(wm, red[w]);
    int ex = 0;
    if (wm > 0.f) frexpf(wm, &ex);
    const int sh = wm > 0.f ? max(-100, min(100, 14 - ex)) : 0;
    const float wsc = ldexpf(1.0f, sh), ws_inv = ldexpf(1.0f, -sh);
    bool big = false;
    for (int e = tid; e < 16 * 40; e += TY_THREADS) {           // tile element (population i, K position k): k = 0..39
        const int i = e / 40, k = e % 40;
        int col = -1;                                           // column of W_aug that sits at K position k
        if (k < 16) col = k < N ? k : -1;
        else if (k < 32) col = (k - 16) < n_in ? N + (k - 16) : -1;
        else if (k == 32) col = N + n_in;
        const float w = (i < N && col >= 0) ? __ldg(p.W_aug + (size_t)i * p.ld_w + col) * wsc : 0.f;
        big |= !(fabsf(w) <= kF16Limit);
        const F16x2 s2 = f16_split2(w);
        const uint32_t off = sw128(i, k >> 3) + (uint32_t)(k & 7) * 2;
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(b_hi + off), "h"(s2.h) : "memory");
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(b_lo + off), "h"(s2.l) : "memory");
    }
    if (half == 0) {                                            // the constant-one column: K position 32 = chunk 4, element 0
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(a_hi + sw128(tr, 4)), "h"((unsigned short)0x3C00) : "memory");
    }

    // ---- per-thread state: populations i0 .. i0 + 7 of trial b
    const int i0 = 8 * half;
    const size_t row = (size_t)3 * N;
    float V[8], A[8], F[8], kap[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool in = live && i0 + j < N;
        V[j] = in ? y0[b * row + i0 + j] : 0.f;
        A[j] = in ? y0[b * row + N + i0 + j] : 0.f;
        F[j] = in ? y0[b * row + 2 * N + i0 + j] : 0.f;
        kap[j] = (i0 + j < N) ? __ldg(p.kappa + i0 + j) : 0.f;
        if (in) { y_out[b * row + i0 + j] = V[j]; y_out[b * row + N + i0 + j] = A[j]; y_out[b * row + 2 * N + i0 + j] = F[j]; }
    }
    const float inv_tm = 1.0f / p.c.tau_m, inv_ta = 1.0f / p.c.tau_a, inv_ts = 1.0f / p.c.tau_s, gain = p.c.tau_s * p.c.R;
    const float* ku = p.knot_u + (size_t)(live ? b : 0) * p.knot_stride_b;
    int kidx = 1;
    const float kt_lo = __ldg(p.knot_t), kt_hi = __ldg(p.knot_t + p.K - 1);
    float kx0 = 1.0f, kx1 = 0.0f, kbase = 0.0f;                 // empty interval: the first evaluation locates
    uint32_t phase = 0;
    float vmax = 0.f, nanchk = 0.f;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // F16 x F16 -> F32, N = 16, M = 128
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)i0;

    // this thread's four operand chunks and its stimulus slots: fixed for the whole solve.  Routed through an opaque move so
    // that the compiler keeps them in registers (it re-derived them from %tid.x at every stage: S2R was 8 % of the stall samples)
    auto pin = [](uint32_t v) { uint32_t o; asm volatile("mov.u32 %0, %1;" : "=r"(o) : "r"(v)); return o; };
    const uint32_t ar_hi = pin(a_hi + sw128(tr, half)), ar_lo = pin(a_lo + sw128(tr, half));
    const uint32_t au_hi = pin(a_hi + sw128(tr, 2 + half)), au_lo = pin(a_lo + sw128(tr, 2 + half));
    float2* const stim_t = stim + pin((uint32_t)(i0 * TY_TRIALS + tr));

    // one right-hand side: r = phi(Vs - As) -> operand -> contraction -> tot; returns r and tot of this thread's populations
    auto rhs = [&](float tq, const float (&Vs)[8], const float (&As)[8], float (&r)[8], float (&tot)[8]) {
        const float tc = fminf(fmaxf(tq, kt_lo), kt_hi);
        if (!(tc >= kx0 && tc < kx1)) {                         // new knot interval (all threads at once: tq is shared)
            knot_locate(p.knot_t, p.K, tq, kidx);
            const float x0 = __ldg(p.knot_t + kidx - 1), x1 = __ldg(p.knot_t + kidx);
            kx0 = kidx == 1 ? -INFINITY : x0;                   // the interval the lookup stays valid on (searchsorted right)
            kx1 = kidx == p.K - 1 ? INFINITY : x1;
            kbase = x0;
            const float dx = __fsub_rn(x1, x0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int ch = i0 + j;
                float yl = 0.f, sl = 0.f;
                if (live && ch < n_in) {
                    yl = __ldg(ku + (size_t)(kidx - 1) * n_in + ch);
                    const float yh = __ldg(ku + (size_t)kidx * n_in + ch);
                    sl = __fdiv_rn(__fsub_rn(yh, yl), dx);
                    vmax = fmaxf(vmax, fmaxf(fabsf(yl), fabsf(yh)));      // the interpolant stays between its knots
                }
                stim_t[j * TY_TRIALS] = make_float2(yl, sl);
            }
        }
        const float dtc = __fsub_rn(tc, kbase);
        float xr[8], xu[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            r[j] = (live && i0 + j < N) ? phi_fast(Vs[j] - As[j]) : 0.f;
            xr[j] = r[j];
            vmax = fmaxf(vmax, fabsf(r[j]));                     // overflow check; fmaxf drops a NaN, the product with zero keeps it
            nanchk = fmaf(r[j], 0.0f, nanchk);
            const float2 s2 = stim_t[j * TY_TRIALS];
            xu[j] = __fadd_rn(s2.x, __fmul_rn(s2.y, dtc));       // knot_value's arithmetic
        }
        uint4 hi, lo;
        pack8(xr, hi, lo);
        sts128(ar_hi, hi); sts128(ar_lo, lo);
        pack8(xu, hi, lo);
        sts128(au_hi, hi); sts128(au_lo, lo);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> tensor-core (async proxy) reads
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint64_t dah = make_smem_desc(a_hi), dal = make_smem_desc(a_lo), dbh = make_smem_desc(b_hi), dbl = make_smem_desc(b_lo);
#pragma unroll
            for (int k = 0; k < TY_KSTEPS; ++k) {
                const uint64_t adv = (uint64_t)(k * 32 >> 4);
                umma_f16(tmem + 16, dal + adv, dbh + adv, idesc, k != 0);
                umma_f16(tmem + 16, dah + adv, dbl + adv, idesc, 1);
                umma_f16(tmem, dah + adv, dbh + adv, idesc, k != 0);
            }
            umma_commit(mb);
        }
        mbar_wait(mb, phase);
        phase ^= 1;
        tc_fence_after();
        float cm[8], cx[8];
        tmem_ld8(lane_addr, cm);
        tmem_ld8(lane_addr + 16, cx);
#pragma unroll
        for (int j = 0; j < 8; ++j) tot[j] = (cx[j] * 4.8828125e-4f + cm[j]) * ws_inv;
    };

    int since_out = 0, out_row = 0;
    for (int n = 0; n < T - 1; ++n) {
        const float t0 = __ldg(t + n), t1 = __ldg(t + n + 1), dt = __fsub_rn(t1, t0);
        const float h3 = dt * kOneThirdL;
        float r[8], tot[8], Vs[8], As[8];
        // slopes: P = k1, Q = k2 per component; after stage 3 they collapse into S = k1 + 3 (k2 + k3) and the stage-4 state
        float PV[8], PA[8], PF[8], QV[8], QA[8], QF[8], Fs[8];
        rhs(t0, V, A, r, tot);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            PV[j] = (tot[j] * gain - V[j]) * inv_tm; PA[j] = (kap[j] * r[j] - A[j]) * inv_ta; PF[j] = (r[j] - F[j]) * inv_ts;
            Vs[j] = fmaf(h3, PV[j], V[j]); As[j] = fmaf(h3, PA[j], A[j]); Fs[j] = fmaf(h3, PF[j], F[j]);
        }
        rhs(__fadd_rn(t0, __fmul_rn(dt, kOneThirdL)), Vs, As, r, tot);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            QV[j] = (tot[j] * gain - Vs[j]) * inv_tm; QA[j] = (kap[j] * r[j] - As[j]) * inv_ta; QF[j] = (r[j] - Fs[j]) * inv_ts;
            Vs[j] = fmaf(dt, QV[j] - PV[j] * kOneThirdL, V[j]); As[j] = fmaf(dt, QA[j] - PA[j] * kOneThirdL, A[j]);
            Fs[j] = fmaf(dt, QF[j] - PF[j] * kOneThirdL, F[j]);
        }
        rhs(__fadd_rn(t0, __fmul_rn(dt, kTwoThirdsL)), Vs, As, r, tot);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float k3V = (tot[j] * gain - Vs[j]) * inv_tm, k3A = (kap[j] * r[j] - As[j]) * inv_ta, k3F = (r[j] - Fs[j]) * inv_ts;
            Vs[j] = fmaf(dt, PV[j] - QV[j] + k3V, V[j]); As[j] = fmaf(dt, PA[j] - QA[j] + k3A, A[j]); Fs[j] = fmaf(dt, PF[j] - QF[j] + k3F, F[j]);
            PV[j] = PV[j] + 3.f * (QV[j] + k3V); PA[j] = PA[j] + 3.f * (QA[j] + k3A); PF[j] = PF[j] + 3.f * (QF[j] + k3F);
        }
        rhs(t1, Vs, As, r, tot);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float k4V = (tot[j] * gain - Vs[j]) * inv_tm, k4A = (kap[j] * r[j] - As[j]) * inv_ta, k4F = (r[j] - Fs[j]) * inv_ts;
            V[j] = fmaf((PV[j] + k4V) * dt, 0.125f, V[j]); A[j] = fmaf((PA[j] + k4A) * dt, 0.125f, A[j]);
            F[j] = fmaf((PF[j] + k4F) * dt, 0.125f, F[j]);
        }
        const int jn = n + 1;
        if (++since_out == out_every) { since_out = 0; ++out_row; }
        if (since_out == 0 || jn == T - 1) {
            const size_t ro = since_out == 0 ? (size_t)out_row : (size_t)((T - 2) / out_every + 1);
            float* yo = y_out + (ro * p.B + b) * row;
            if (N == 16 && live && (reinterpret_cast<uintptr_t>(y_out) & 15) == 0) {        // 8 consecutive floats per component
                float4* yv = reinterpret_cast<float4*>(yo + i0);
                yv[0] = make_float4(V[0], V[1], V[2], V[3]); yv[1] = make_float4(V[4], V[5], V[6], V[7]);
                yv[4] = make_float4(A[0], A[1], A[2], A[3]); yv[5] = make_float4(A[4], A[5], A[6], A[7]);
                yv[8] = make_float4(F[0], F[1], F[2], F[3]); yv[9] = make_float4(F[4], F[5], F[6], F[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (live && i0 + j < N) { yo[i0 + j] = V[j]; yo[N + i0 + j] = A[j]; yo[2 * N + i0 + j] = F[j]; }
            }
        }
    }
    if (big || !(vmax <= kF16Limit) || !(nanchk == 0.0f)) *ovf = 1u;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

}  // namespace tc

bool tiny_rk4_applicable(const DevProblem& p) {
    static const bool on = [] { const char* e = getenv("ODECOL_TINY_TC"); return e ? atoi(e) != 0 : true; }();
    return on && p.N <= 16 && p.n_in <= 16 && p.B >= 4096;
}

// *ovf (device, 4 bytes, caller-provided) is cleared, then raised by the kernel when a value did not fit FP16
int launch_rk4_fwd_tiny(const DevProblem& p, const float* t, int T, const float* y0, float* y_out, int out_every,
                        unsigned int* ovf, cudaStream_t s) {
    using namespace tc;
    if (cudaMemsetAsync(ovf, 0, sizeof(unsigned int), s) != cudaSuccess) return ODECOL_E_CUDA;
    const size_t smem = 2 * TY_A_BYTES + 2 * TY_B_BYTES + 16 * TY_TRIALS * sizeof(float2) + 1024;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_rk4_fwd_tiny, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return ODECOL_E_CUDA;
        configured = true;
    }
    k_rk4_fwd_tiny<<<(p.B + TY_TRIALS - 1) / TY_TRIALS, TY_THREADS, smem, s>>>(p, t, T, y0, y_out, out_every, ovf);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

}  // namespace odecol
