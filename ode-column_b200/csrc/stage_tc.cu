// Kernel family T ("tensor"): the staged solver's contraction on the 5th-generation tensor cores.
//
//   C[i][b] = sum_k W_aug[i][k] * R_aug[b][k]   as   3xTF32:  W = Whi + Wlo, R = Rhi + Rlo (each rounded to TF32),
//   C = Wlo.Rhi + Whi.Rlo + Whi.Rhi accumulated in FP32 in tensor memory -- FP32-level accuracy at tensor-core rate.
//
// One persistent CTA per SM, warp specialised:
//   warp 0      TMA producer: cp.async.bulk.tensor (128-byte swizzle) of the four operand tiles of a 32-wide K block into
//               a 3-stage shared-memory ring, completion on mbarriers
//   warp 1      tcgen05.mma issuer (one elected thread; M = 128 populations x N = tile of trials x K = 8 per instruction),
//               accumulators double-buffered in TMEM; tcgen05.commit releases ring slots and publishes finished tiles
//   warps 2..9  epilogue: tcgen05.ld the accumulator rows (TMEM lane = population), then the same fused RK-stage epilogue
//               as the FFMA family -- drift, next stage state, phi, next operand (already split into hi/lo) -- while
//               warp 1 is contracting the next tile
// The trial tile is chosen so that the number of tiles is a multiple of the SM count (148) where possible.
#include <cuda.h>
#include "stage_common.cuh"

namespace odecol {
namespace tc {

constexpr int BM = 128;          // populations per tile = TMEM lanes
constexpr int BK = 32;           // floats per K block = one 128-byte swizzle row
constexpr int STAGES = 3;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr uint32_t kSpinLimit = 1u << 24;

ODECOL_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

ODECOL_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
ODECOL_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
ODECOL_DEVINL void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug must trap, never hang the GPU
ODECOL_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (++spins > kSpinLimit) __trap();
    }
}
ODECOL_DEVINL void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
ODECOL_DEVINL void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
ODECOL_DEVINL void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
ODECOL_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
ODECOL_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
ODECOL_DEVINL void tmem_ld16(uint32_t taddr, float (&r)[16]) {
    uint32_t u[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
                   "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart
ODECOL_DEVINL uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);     // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                           // leading byte offset      bits [16,30)  (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset       bits [32,46)
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}
ODECOL_DEVINL uint32_t make_idesc(int tile_n) {
    // c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(tile_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

ODECOL_DEVINL float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

struct TileShape {
    int MT, NT, TN, KB;      // m tiles, trial tiles, trials per tile (multiple of 16, <= 128), K blocks of 32
};

// ---------------------------------------------------------------------------------------------------------------
// the persistent warp-specialised contraction; Epi supplies element(i, b, acc) and tile_done(m_tile, n0, tid, nthreads)
// ---------------------------------------------------------------------------------------------------------------
template <class Epi>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_contract(const __grid_constant__ CUtensorMap mA_hi, const __grid_constant__ CUtensorMap mA_lo,
              const __grid_constant__ CUtensorMap mB_hi, const __grid_constant__ CUtensorMap mB_lo, TileShape ts, Epi epi) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_bytes = BM * BK * 4, b_bytes = (uint32_t)ts.TN * BK * 4;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
    const uint32_t tfull0 = smem_u32(&bars[2 * STAGES]), tempty0 = smem_u32(&bars[2 * STAGES + 2]);
    const int tiles = ts.MT * ts.NT;
    const uint32_t acc_stride = (uint32_t)ts.TN;
    uint32_t ncols = 32;
    while (ncols < 2 * acc_stride) ncols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int m0 = (tile % ts.MT) * BM, n0 = (tile / ts.MT) * ts.TN;
                for (int kb = 0; kb < ts.KB; ++kb) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t base = ring + stage * stage_bytes, fb = full0 + 8 * stage;
                    mbar_expect_tx(fb, stage_bytes);
                    tma_load_2d(base, &mA_hi, fb, kb * BK, m0);
                    tma_load_2d(base + a_bytes, &mA_lo, fb, kb * BK, m0);
                    tma_load_2d(base + 2 * a_bytes, &mB_hi, fb, kb * BK, n0);
                    tma_load_2d(base + 2 * a_bytes + b_bytes, &mB_lo, fb, kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(ts.TN);
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * acc_stride;
                for (int kb = 0; kb < ts.KB; ++kb) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t base = ring + stage * stage_bytes;
                    const uint64_t a_hi = make_smem_desc(base), a_lo = make_smem_desc(base + a_bytes);
                    const uint64_t b_hi = make_smem_desc(base + 2 * a_bytes), b_lo = make_smem_desc(base + 2 * a_bytes + b_bytes);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4);       // 8 TF32 = 32 bytes along the swizzled row
                        umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
                        umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
                        umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
                    }
                    umma_commit(empty0 + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull0 + 8 * acc);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int ew = warp - 2;
        const int quarter = warp & 3;                 // a warp may only read TMEM lanes [32*(warpid%4), +32)
        const int half = ew >> 2;
        const int etid = ew * 32 + lane;
        epi.prepare();
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int m_tile = tile % ts.MT, n0 = (tile / ts.MT) * ts.TN;
            const int row = m_tile * BM + quarter * 32 + lane;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            for (int c = half; c < ts.TN / 16; c += 2) {
                float r[16];
                tmem_ld16(tmem_base + acc * acc_stride + ((uint32_t)(quarter * 32) << 16) + c * 16, r);
#pragma unroll
                for (int j = 0; j < 16; ++j) epi.element(row, n0 + c * 16 + j, r[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            epi.tile_done(m_tile, n0, ts.TN, etid, kEpiWarps * 32);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: tensor maps, tile shape, launch
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// rows x cols float32 matrix, row-major with `ld` floats per row; box = box_rows x 32 floats, 128-byte swizzle
static bool make_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld * sizeof(float)};
    const cuuint32_t box[2] = {BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// trials per tile: multiple of 16, <= 128, minimising (waves x per-tile cost) -- lands on a multiple of the SM count
static int pick_tile_n(int MT, int B) {
    const int sms = num_sms();
    int best = 16;
    double best_cost = 1e30;
    for (int tn = 16; tn <= 128; tn += 16) {
        const int nt = (B + tn - 1) / tn;
        const long tiles = (long)MT * nt;
        const long waves = (tiles + sms - 1) / sms;
        const double cost = (double)waves * (tn + 24.0);      // 24: per-tile fixed cost in "trial" units (pipeline fill)
        if (cost < best_cost - 1e-9) { best_cost = cost; best = tn; }
    }
    return best;
}

template <class Epi>
static int launch_contract(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                           const TileShape& ts, const Epi& epi, cudaStream_t s) {
    const size_t smem = (size_t)STAGES * (2 * BM * BK * 4 + 2 * (size_t)ts.TN * BK * 4) + 1024;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_tc_contract<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return ODECOL_E_CUDA;
        configured = true;
    }
    const int tiles = ts.MT * ts.NT;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    k_tc_contract<Epi><<<grid, kThreads, smem, s>>>(a_hi, a_lo, b_hi, b_lo, ts, epi);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_split_pad(const float* __restrict__ src, int rows, int cols, int ld, float* __restrict__ hi,
                            float* __restrict__ lo, int rows_p, int cols_p) {
    const size_t total = (size_t)rows_p * cols_p;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / cols_p), c = (int)(e % cols_p);
        const float x = (r < rows && c < cols) ? src[(size_t)r * ld + c] : 0.0f;
        const float h = tf32_rna(x);
        hi[e] = h;
        lo[e] = tf32_rna(x - h);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// diagnostic: the contraction core alone, C[n][m] = sum_k A[m][k] B[n][k]
// ---------------------------------------------------------------------------------------------------------------
struct StoreEpi {
    float* C; int M, N, ldc;
    ODECOL_DEVINL void prepare() {}
    ODECOL_DEVINL void element(int i, int b, float acc) const { if (i < M && b < N) C[(size_t)b * ldc + i] = acc; }
    ODECOL_DEVINL void tile_done(int, int, int, int, int) const {}
};

struct ContractLayout { int Mp, Np, Kp, TN; size_t off_Ahi, off_Alo, off_Bhi, off_Blo, total; };
static ContractLayout contract_layout(int M, int N, int K) {
    ContractLayout L;
    L.Mp = round_up(M, BM); L.Kp = round_up(K, BK);
    L.TN = pick_tile_n(L.Mp / BM, N);
    L.Np = round_up(N, L.TN);
    size_t o = 0;
    auto take = [&](size_t floats) { const size_t r = o; o += (floats * 4 + 1023) / 1024 * 1024; return r; };
    L.off_Ahi = take((size_t)L.Mp * L.Kp); L.off_Alo = take((size_t)L.Mp * L.Kp);
    L.off_Bhi = take((size_t)L.Np * L.Kp); L.off_Blo = take((size_t)L.Np * L.Kp);
    L.total = o;
    return L;
}

// ---------------------------------------------------------------------------------------------------------------
// forward RK4 stages on the tensor cores
// ---------------------------------------------------------------------------------------------------------------
template <int S>
struct FwdEpi {
    FwdStageArgs a;
    float t0, t1, dt;
    ODECOL_DEVINL void prepare() {
        t0 = __ldg(a.t + a.n); t1 = __ldg(a.t + a.n + 1);
        dt = __fsub_rn(t1, t0);
    }
    ODECOL_DEVINL void element(int i, int b, float acc) const {
        if (i < a.p.N && b < a.p.B) {
            const float tot[1] = {acc};
            fwd_stage_epilogue<S, 1>(a, i, b, tot, dt);
        }
    }
    // stimulus columns of the next operand, written once per trial tile (by the CTA that owns population tile 0)
    ODECOL_DEVINL void tile_done(int m_tile, int n0, int tile_n, int etid, int nthr) const {
        if (m_tile != 0 || a.p.n_in == 0) return;
        const float tn = S == 1 ? __fadd_rn(t0, __fmul_rn(dt, kOneThirdL)) : S == 2 ? __fadd_rn(t0, __fmul_rn(dt, kTwoThirdsL)) : t1;
        int idx = 1;
        const float tc = knot_locate(a.p.knot_t, a.p.K, tn, idx);
        const int n_in = a.p.n_in, N = a.p.N;
        for (int e = etid; e < tile_n * n_in; e += nthr) {
            const int b = n0 + e / n_in, ch = e % n_in;
            if (b < a.p.B) {
                const float v = knot_value(a.p.knot_t, a.p.knot_u + (size_t)b * a.p.knot_stride_b, n_in, idx, tc, ch);
                const float h = tf32_rna(v);
                a.Ra_nxt[(size_t)b * a.KPa + N + ch] = h;
                a.Ra_nxt_lo[(size_t)b * a.KPa + N + ch] = tf32_rna(v - h);
            }
        }
    }
};

// r_aug(t_ptr[0], y) split into hi/lo for buffer 0; constant-one column + zero padding for buffer 1
__global__ void k_init_operand_split(DevProblem p, const float* __restrict__ y, const float* __restrict__ t_ptr,
                                     float* __restrict__ hi0, float* __restrict__ lo0, float* __restrict__ hi1,
                                     float* __restrict__ lo1, float* __restrict__ DR, int KPa) {
    const int b = blockIdx.x, N = p.N, Kaug = N + p.n_in + 1;
    const size_t ro = (size_t)b * KPa;
    if (b >= p.B) {
        for (int k = threadIdx.x; k < KPa; k += blockDim.x) {
            hi0[ro + k] = 0.f; lo0[ro + k] = 0.f;
            if (hi1) { hi1[ro + k] = 0.f; lo1[ro + k] = 0.f; }
        }
        return;
    }
    const float* yb = y + (size_t)b * 3 * N;
    int idx = 1;
    const float tc = knot_locate(p.knot_t, p.K, __ldg(t_ptr), idx);
    const float* ku = p.knot_u + (size_t)b * p.knot_stride_b;
    for (int k = threadIdx.x; k < KPa; k += blockDim.x) {
        float v = 0.f, one = 0.f;
        if (k < N) {
            if (DR) { float d; phi_dphi(__fsub_rn(yb[k], yb[N + k]), v, d); DR[(size_t)b * N + k] = d; }
            else v = phi(__fsub_rn(yb[k], yb[N + k]));
        } else if (k < N + p.n_in) v = knot_value(p.knot_t, ku, p.n_in, idx, tc, k - N);
        else if (k == Kaug - 1) { v = 1.f; one = 1.f; }
        const float h = tf32_rna(v);
        hi0[ro + k] = h; lo0[ro + k] = tf32_rna(v - h);
        if (hi1) { hi1[ro + k] = one; lo1[ro + k] = 0.f; }
    }
}

struct TcFwdLayout {
    int Np, Bp, KPa, TN;
    size_t off_Whi, off_Wlo, off_Rhi[2], off_Rlo[2], off_k[3], off_y[2], total;
};

static TcFwdLayout tc_fwd_layout(const DevProblem& p) {
    TcFwdLayout L;
    L.Np = round_up(p.N, BM);
    L.KPa = round_up(p.N + p.n_in + 1, BK);
    L.TN = pick_tile_n(L.Np / BM, p.B);
    L.Bp = round_up(p.B, L.TN);
    size_t o = 0;
    auto take = [&](size_t floats) { const size_t r = o; o += (floats * 4 + 1023) / 1024 * 1024; return r; };
    L.off_Whi = take((size_t)L.Np * L.KPa); L.off_Wlo = take((size_t)L.Np * L.KPa);
    for (int i = 0; i < 2; ++i) { L.off_Rhi[i] = take((size_t)L.Bp * L.KPa); L.off_Rlo[i] = take((size_t)L.Bp * L.KPa); }
    const size_t st = (size_t)p.B * 3 * p.N;
    for (int i = 0; i < 3; ++i) L.off_k[i] = take(st);
    for (int i = 0; i < 2; ++i) L.off_y[i] = take(st);
    L.total = o;
    return L;
}

__global__ void k_copy4(const float* __restrict__ src, float* __restrict__ dst, size_t n4) {
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n4; e += (size_t)gridDim.x * blockDim.x)
        reinterpret_cast<float4*>(dst)[e] = reinterpret_cast<const float4*>(src)[e];
}

}  // namespace tc

size_t tc_rk4_fwd_workspace_bytes(const DevProblem& p, int) { return tc::tc_fwd_layout(p).total; }

int tc_rk4_fwd(const DevProblem& p, const float* t_dev, int T, const float* y0, float* y_out, int out_every, void* ws,
               size_t ws_bytes, cudaStream_t s) {
    using namespace tc;
    const TcFwdLayout L = tc_fwd_layout(p);
    if (!ws || ws_bytes < L.total) return ODECOL_E_WORKSPACE;
    if (p.N % 4 != 0) return ODECOL_E_UNSUPPORTED;
    char* w = static_cast<char*>(ws);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(w + off); };
    float *Whi = F(L.off_Whi), *Wlo = F(L.off_Wlo);
    float* Rhi[2] = {F(L.off_Rhi[0]), F(L.off_Rhi[1])};
    float* Rlo[2] = {F(L.off_Rlo[0]), F(L.off_Rlo[1])};
    float* kk[3] = {F(L.off_k[0]), F(L.off_k[1]), F(L.off_k[2])};
    float* ybuf[2] = {F(L.off_y[0]), F(L.off_y[1])};
    const int Kaug = p.N + p.n_in + 1;
    const size_t st = (size_t)p.B * 3 * p.N;

    k_split_pad<<<296, 256, 0, s>>>(p.W_aug, p.N, Kaug, p.ld_w, Whi, Wlo, L.Np, L.KPa);
    k_init_operand_split<<<L.Bp, 128, 0, s>>>(p, y0, t_dev, Rhi[0], Rlo[0], Rhi[1], Rlo[1], nullptr, L.KPa);
    k_copy4<<<296, 256, 0, s>>>(y0, y_out, st / 4);
    count_launch(3);
    CUtensorMap mWhi, mWlo, mRhi[2], mRlo[2];
    bool ok = make_map(&mWhi, Whi, L.Np, L.KPa, L.KPa, BM) && make_map(&mWlo, Wlo, L.Np, L.KPa, L.KPa, BM);
    for (int i = 0; i < 2; ++i)
        ok = ok && make_map(&mRhi[i], Rhi[i], L.Bp, L.KPa, L.KPa, L.TN) && make_map(&mRlo[i], Rlo[i], L.Bp, L.KPa, L.KPa, L.TN);
    if (!ok) return ODECOL_E_CUDA;
    const TileShape ts{L.Np / BM, L.Bp / L.TN, L.TN, L.KPa / BK};

    const float* ycur = y0;
    int cur = 0;
    for (int n = 0; n < T - 1; ++n) {
        const int j = n + 1;
        const bool emit = (j % out_every == 0) || (j == T - 1);
        const size_t r = (j % out_every == 0) ? (size_t)(j / out_every) : (size_t)((T - 2) / out_every + 1);
        float* ynext = emit ? y_out + r * st : ybuf[n & 1];
        FwdStageArgs a;
        a.p = p; a.Wp = nullptr; a.y0 = ycur; a.k1 = kk[0]; a.k2 = kk[1]; a.k3 = kk[2]; a.y1 = ynext; a.y_out_row = nullptr;
        a.DR_nxt = nullptr; a.t = t_dev; a.n = n; a.KPa = L.KPa;
        for (int S = 1; S <= 4; ++S) {
            a.Ra_cur = Rhi[cur]; a.Ra_cur_lo = Rlo[cur]; a.Ra_nxt = Rhi[cur ^ 1]; a.Ra_nxt_lo = Rlo[cur ^ 1];
            int rc;
            switch (S) {
                case 1: rc = launch_contract(mWhi, mWlo, mRhi[cur], mRlo[cur], ts, FwdEpi<1>{a}, s); break;
                case 2: rc = launch_contract(mWhi, mWlo, mRhi[cur], mRlo[cur], ts, FwdEpi<2>{a}, s); break;
                case 3: rc = launch_contract(mWhi, mWlo, mRhi[cur], mRlo[cur], ts, FwdEpi<3>{a}, s); break;
                default: rc = launch_contract(mWhi, mWlo, mRhi[cur], mRlo[cur], ts, FwdEpi<4>{a}, s); break;
            }
            if (rc != ODECOL_OK) return rc;
            cur ^= 1;
        }
        ycur = ynext;
    }
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

size_t tc_contract_workspace_bytes(int M, int N, int K) { return tc::contract_layout(M, N, K).total; }

int tc_contract(const float* A, const float* B, float* C, int M, int N, int K, void* ws, size_t ws_bytes, cudaStream_t s) {
    using namespace tc;
    const ContractLayout L = contract_layout(M, N, K);
    if (!ws || ws_bytes < L.total) return ODECOL_E_WORKSPACE;
    char* w = static_cast<char*>(ws);
    float* Ahi = reinterpret_cast<float*>(w + L.off_Ahi); float* Alo = reinterpret_cast<float*>(w + L.off_Alo);
    float* Bhi = reinterpret_cast<float*>(w + L.off_Bhi); float* Blo = reinterpret_cast<float*>(w + L.off_Blo);
    k_split_pad<<<296, 256, 0, s>>>(A, M, K, K, Ahi, Alo, L.Mp, L.Kp);
    k_split_pad<<<296, 256, 0, s>>>(B, N, K, K, Bhi, Blo, L.Np, L.Kp);
    count_launch(2);
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    if (!make_map(&ma_hi, Ahi, L.Mp, L.Kp, L.Kp, BM) || !make_map(&ma_lo, Alo, L.Mp, L.Kp, L.Kp, BM) ||
        !make_map(&mb_hi, Bhi, L.Np, L.Kp, L.Kp, L.TN) || !make_map(&mb_lo, Blo, L.Np, L.Kp, L.Kp, L.TN))
        return ODECOL_E_CUDA;
    TileShape ts{L.Mp / BM, L.Np / L.TN, L.TN, L.Kp / BK};
    StoreEpi epi{C, M, N, M};
    return launch_contract(ma_hi, ma_lo, mb_hi, mb_lo, ts, epi, s);
}

}  // namespace odecol
