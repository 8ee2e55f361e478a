// Shared pieces of kernel family L: tile constants, the FP32-FFMA contraction core, small vector helpers.
#pragma once
#include "odecol_common.cuh"

namespace odecol {

constexpr float kOneThirdL = 0.3333333333333333f;
constexpr float kTwoThirdsL = 0.6666666666666666f;

constexpr int TM = 128, TN = 128, TK = 16;     // CTA tile: TM rows of W (populations) x TN trials x TK
constexpr int kGemmThreads = 256;

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------------------------
// contraction core: acc[8][8] per thread.  A rows = populations (ldA floats, K contiguous), B rows = trials.
// thread (tx = tid % 16, ty = tid / 16): rows {4tx..4tx+3} u {64+4tx..}, cols {4ty..4ty+3} u {64+4ty..}
// ---------------------------------------------------------------------------------------------------------------
struct GemmSmem {
    float A[2][TK][TM];
    float B[2][TK][TN];
};

ODECOL_DEVINL void gemm_nt_core(const float* __restrict__ Ag, int ldA, const float* __restrict__ Bg, int ldB, int Kdim,
                                GemmSmem& sm, float (&acc)[8][8]) {
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    // loader mapping: float4 index f = tid + 256*m (m = 0,1): row = f % 128, kq = f / 128
    const int lrow = tid & 127, lkq = tid >> 7;       // lkq in {0,1}; second float4 uses kq + 2
    const float4* a_src0 = reinterpret_cast<const float4*>(Ag + (size_t)lrow * ldA) + lkq;
    const float4* b_src0 = reinterpret_cast<const float4*>(Bg + (size_t)lrow * ldB) + lkq;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float4 ra0 = __ldg(a_src0), ra1 = __ldg(a_src0 + 2), rb0 = __ldg(b_src0), rb1 = __ldg(b_src0 + 2);
    const int nk = Kdim / TK;
    int buf = 0;
    auto stash = [&](int bf) {
        const int k0 = 4 * lkq, k1 = 4 * (lkq + 2);
        sm.A[bf][k0 + 0][lrow] = ra0.x; sm.A[bf][k0 + 1][lrow] = ra0.y; sm.A[bf][k0 + 2][lrow] = ra0.z; sm.A[bf][k0 + 3][lrow] = ra0.w;
        sm.A[bf][k1 + 0][lrow] = ra1.x; sm.A[bf][k1 + 1][lrow] = ra1.y; sm.A[bf][k1 + 2][lrow] = ra1.z; sm.A[bf][k1 + 3][lrow] = ra1.w;
        sm.B[bf][k0 + 0][lrow] = rb0.x; sm.B[bf][k0 + 1][lrow] = rb0.y; sm.B[bf][k0 + 2][lrow] = rb0.z; sm.B[bf][k0 + 3][lrow] = rb0.w;
        sm.B[bf][k1 + 0][lrow] = rb1.x; sm.B[bf][k1 + 1][lrow] = rb1.y; sm.B[bf][k1 + 2][lrow] = rb1.z; sm.B[bf][k1 + 3][lrow] = rb1.w;
    };
    stash(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) {
            const int off = (kt + 1) * (TK / 4);
            ra0 = __ldg(a_src0 + off); ra1 = __ldg(a_src0 + off + 2);
            rb0 = __ldg(b_src0 + off); rb1 = __ldg(b_src0 + off + 2);
        }
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sm.A[buf][k][4 * tx]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sm.A[buf][k][64 + 4 * tx]);
            const float4 b0 = *reinterpret_cast<const float4*>(&sm.B[buf][k][4 * ty]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sm.B[buf][k][64 + 4 * ty]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
}

ODECOL_DEVINL float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
ODECOL_DEVINL void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// scratch that is streamed once per launch: L2 only.  (With ~200 KB of the SM's array carved out as shared memory the
// remaining L1 is too small to hold a line for every outstanding load, and allocating loads stall on it.)
ODECOL_DEVINL float4 ld4s(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
ODECOL_DEVINL void st4s(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }

// one state component of 4 consecutive populations
struct C4 { float v[4]; };
ODECOL_DEVINL C4 ldc(const float* p) { const float4 q = ld4(p); return {{q.x, q.y, q.z, q.w}}; }
ODECOL_DEVINL void stc(float* p, const C4& c) { st4(p, make_float4(c.v[0], c.v[1], c.v[2], c.v[3])); }


// ---- forward stage launch arguments (stage_kernels.cu) ---------------------------------------------------------
struct FwdStageArgs {
    DevProblem p;
    const float* Wp;       // [Np][KPa]
    const float* Ra_cur;   // [Bp][KPa]
    float* Ra_nxt;         // [Bp][KPa]
    const float* Ra_cur_lo; // tensor family only: low TF32 parts of the operands (else NULL)
    float* Ra_nxt_lo;
    const float* y0;       // [B][3N] state at the start of the step
    float* k1;             // [B][3N]
    float* k2;
    float* k3;
    float* y1;             // [B][3N] state at the end of the step (stage 4)
    float* y_out_row;      // optional second destination of y1 (trajectory row), may be NULL
    float* DR_nxt;         // optional [B][N]: phi'(x) of the NEXT stage state (reverse sweep recompute)
    const float* t;        // device time grid
    int n;                 // step index: t0 = t[n], t1 = t[n+1]
    int KPa;
};

template <int VW> struct Pk { float v[VW]; };
template <int VW> ODECOL_DEVINL Pk<VW> ldp(const float* p) {
    Pk<VW> r;
    if constexpr (VW == 4) { const float4 q = ld4(p); r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w; }
    else {
#pragma unroll
        for (int e = 0; e < VW; ++e) r.v[e] = p[e];
    }
    return r;
}
template <int VW> ODECOL_DEVINL void stp(float* p, const Pk<VW>& x) {
    if constexpr (VW == 4) st4(p, make_float4(x.v[0], x.v[1], x.v[2], x.v[3]));
    else {
#pragma unroll
        for (int e = 0; e < VW; ++e) p[e] = x.v[e];
    }
}
ODECOL_DEVINL float tf32_round(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// fused epilogue of forward stage S (1..4) for VW consecutive populations of trial b, shared by the FFMA and the
// tensor-core contraction kernels
template <int S, int VW>
ODECOL_DEVINL void fwd_stage_epilogue(const FwdStageArgs& a, int i, int b, const float (&tot)[VW], float dt) {
    const int N = a.p.N;
    const size_t base = (size_t)b * 3 * N + i;
    const Pk<VW> V0 = ldp<VW>(a.y0 + base), A0 = ldp<VW>(a.y0 + base + N), F0 = ldp<VW>(a.y0 + base + 2 * N);
    Pk<VW> rs = ldp<VW>(a.Ra_cur + (size_t)b * a.KPa + i);
    if (a.Ra_cur_lo) {                       // tensor family: the operand is stored split, r = hi + lo
        const Pk<VW> rl = ldp<VW>(a.Ra_cur_lo + (size_t)b * a.KPa + i);
#pragma unroll
        for (int e = 0; e < VW; ++e) rs.v[e] += rl.v[e];
    }
    const Pk<VW> kap = ldp<VW>(a.p.kappa + i);
    Pk<VW> k1V, k1A, k1F, k2V, k2A, k2F, k3V, k3A, k3F;
    if (S >= 2) { k1V = ldp<VW>(a.k1 + base); k1A = ldp<VW>(a.k1 + base + N); k1F = ldp<VW>(a.k1 + base + 2 * N); }
    if (S >= 3) { k2V = ldp<VW>(a.k2 + base); k2A = ldp<VW>(a.k2 + base + N); k2F = ldp<VW>(a.k2 + base + 2 * N); }
    if (S >= 4) { k3V = ldp<VW>(a.k3 + base); k3A = ldp<VW>(a.k3 + base + N); k3F = ldp<VW>(a.k3 + base + 2 * N); }
    Pk<VW> oV, oA, oF, oR, oD, kV, kA, kF;
#pragma unroll
    for (int e = 0; e < VW; ++e) {
        // the stage state this launch's contraction belongs to (same expressions as family S)
        float V, A, F;
        if (S == 1) { V = V0.v[e]; A = A0.v[e]; F = F0.v[e]; }
        if (S == 2) {
            V = __fadd_rn(V0.v[e], __fmul_rn(__fmul_rn(dt, k1V.v[e]), kOneThirdL));
            A = __fadd_rn(A0.v[e], __fmul_rn(__fmul_rn(dt, k1A.v[e]), kOneThirdL));
            F = __fadd_rn(F0.v[e], __fmul_rn(__fmul_rn(dt, k1F.v[e]), kOneThirdL));
        }
        if (S == 3) {
            V = __fadd_rn(V0.v[e], __fmul_rn(dt, __fsub_rn(k2V.v[e], __fmul_rn(k1V.v[e], kOneThirdL))));
            A = __fadd_rn(A0.v[e], __fmul_rn(dt, __fsub_rn(k2A.v[e], __fmul_rn(k1A.v[e], kOneThirdL))));
            F = __fadd_rn(F0.v[e], __fmul_rn(dt, __fsub_rn(k2F.v[e], __fmul_rn(k1F.v[e], kOneThirdL))));
        }
        if (S == 4) {
            V = __fadd_rn(V0.v[e], __fmul_rn(dt, __fadd_rn(__fsub_rn(k1V.v[e], k2V.v[e]), k3V.v[e])));
            A = __fadd_rn(A0.v[e], __fmul_rn(dt, __fadd_rn(__fsub_rn(k1A.v[e], k2A.v[e]), k3A.v[e])));
            F = __fadd_rn(F0.v[e], __fmul_rn(dt, __fadd_rn(__fsub_rn(k1F.v[e], k2F.v[e]), k3F.v[e])));
        }
        float dV, dA, dF;
        drift(a.p.c, V, A, F, rs.v[e], kap.v[e], tot[e], dV, dA, dF);
        kV.v[e] = dV; kA.v[e] = dA; kF.v[e] = dF;
        // next stage state
        float nV, nA, nF;
        if (S == 1) {
            nV = __fadd_rn(V0.v[e], __fmul_rn(__fmul_rn(dt, dV), kOneThirdL));
            nA = __fadd_rn(A0.v[e], __fmul_rn(__fmul_rn(dt, dA), kOneThirdL));
            nF = 0.f;
        }
        if (S == 2) {
            nV = __fadd_rn(V0.v[e], __fmul_rn(dt, __fsub_rn(dV, __fmul_rn(k1V.v[e], kOneThirdL))));
            nA = __fadd_rn(A0.v[e], __fmul_rn(dt, __fsub_rn(dA, __fmul_rn(k1A.v[e], kOneThirdL))));
            nF = 0.f;
        }
        if (S == 3) {
            nV = __fadd_rn(V0.v[e], __fmul_rn(dt, __fadd_rn(__fsub_rn(k1V.v[e], k2V.v[e]), dV)));
            nA = __fadd_rn(A0.v[e], __fmul_rn(dt, __fadd_rn(__fsub_rn(k1A.v[e], k2A.v[e]), dA)));
            nF = 0.f;
        }
        if (S == 4) {
            nV = __fadd_rn(V0.v[e], __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(k1V.v[e], __fmul_rn(3.f, __fadd_rn(k2V.v[e], k3V.v[e]))), dV), dt), 0.125f));
            nA = __fadd_rn(A0.v[e], __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(k1A.v[e], __fmul_rn(3.f, __fadd_rn(k2A.v[e], k3A.v[e]))), dA), dt), 0.125f));
            nF = __fadd_rn(F0.v[e], __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(k1F.v[e], __fmul_rn(3.f, __fadd_rn(k2F.v[e], k3F.v[e]))), dF), dt), 0.125f));
        }
        oV.v[e] = nV; oA.v[e] = nA; oF.v[e] = nF;
        if (a.DR_nxt) phi_dphi(__fsub_rn(nV, nA), oR.v[e], oD.v[e]);
        else oR.v[e] = phi(__fsub_rn(nV, nA));
    }
    if (a.DR_nxt) stp<VW>(a.DR_nxt + (size_t)b * N + i, oD);
    if (S == 1) { stp<VW>(a.k1 + base, kV); stp<VW>(a.k1 + base + N, kA); stp<VW>(a.k1 + base + 2 * N, kF); }
    if (S == 2) { stp<VW>(a.k2 + base, kV); stp<VW>(a.k2 + base + N, kA); stp<VW>(a.k2 + base + 2 * N, kF); }
    if (S == 3) { stp<VW>(a.k3 + base, kV); stp<VW>(a.k3 + base + N, kA); stp<VW>(a.k3 + base + 2 * N, kF); }
    if (S == 4) {
        stp<VW>(a.y1 + base, oV); stp<VW>(a.y1 + base + N, oA); stp<VW>(a.y1 + base + 2 * N, oF);
        if (a.y_out_row) { stp<VW>(a.y_out_row + base, oV); stp<VW>(a.y_out_row + base + N, oA); stp<VW>(a.y_out_row + base + 2 * N, oF); }
    }
    if (a.Ra_nxt_lo) {
        Pk<VW> oH, oL;
#pragma unroll
        for (int e = 0; e < VW; ++e) { oH.v[e] = tf32_round(oR.v[e]); oL.v[e] = tf32_round(oR.v[e] - oH.v[e]); }
        stp<VW>(a.Ra_nxt + (size_t)b * a.KPa + i, oH);
        stp<VW>(a.Ra_nxt_lo + (size_t)b * a.KPa + i, oL);
    } else {
        stp<VW>(a.Ra_nxt + (size_t)b * a.KPa + i, oR);
    }
}


// launchers defined in stage_kernels.cu (kernels cannot be launched across translation units)
void launch_pad_weights(const float* W_aug, int N, int ld_w, int Kaug, float* Wp, int Np, int KPa, cudaStream_t s);
void launch_pad_transpose(const float* W_aug, int N, int ld_w, float* WT, int Np, int NPk, cudaStream_t s);
// Ra = r_aug(t_ptr[0], y) (+ phi' into DR if given); Ra_extra (optional) gets the constant-one column and zero padding
void launch_init_operand(const DevProblem& p, const float* y, const float* t_ptr, float* Ra, float* Ra_extra, float* DR,
                         int KPa, int Bp, cudaStream_t s);
void launch_fwd_stage(int S, const FwdStageArgs& a, dim3 grid, cudaStream_t s);

}  // namespace odecol
