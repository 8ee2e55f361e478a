// Microbenchmark 2: epilogue access pattern + synthetic compute between loads and stores; plain vs software-pipelined loads.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int L>
__device__ __forceinline__ void load_group(const float* base, size_t plane, size_t off, float4 (&v)[L]) {
#pragma unroll
    for (int l = 0; l < L; ++l) v[l] = __ldcg(reinterpret_cast<const float4*>(base + l * plane + off));
}

template <int L, int M, int PIPE>
__global__ void __launch_bounds__(576, 1)
k(float* __restrict__ base, size_t plane, int NT, int Np, int nq, int scatter, float* rowmajor, int ld, int C) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= 16) return;
    const int quarter = warp & 3, g = warp >> 2;
    const int tiles = 4 * NT;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int m = tile & 3, nt = tile >> 2;
        const int i = m * 128 + quarter * 32 + lane;
        float4 cur[L], nxt[L];
        auto off_of = [&](int q) { return ((size_t)(nt * 4 * nq + g * nq + q) * Np + i) * 4; };
        if (PIPE) load_group<L>(base, plane, off_of(0), nxt);
        for (int q = 0; q < nq; ++q) {
            const size_t off = off_of(q);
            if (PIPE) {
#pragma unroll
                for (int l = 0; l < L; ++l) cur[l] = nxt[l];
                if (q + 1 < nq) load_group<L>(base, plane, off_of(q + 1), nxt);
            } else {
                load_group<L>(base, plane, off, cur);
            }
            float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int l = 0; l < L; ++l) { acc.x += cur[l].x; acc.y += cur[l].y; acc.z += cur[l].z; acc.w += cur[l].w; }
            for (int c = 0; c < C; ++c) {                 // dependent chain per element: 4 independent chains per thread
                acc.x = fmaf(acc.x, 1.0001f, 0.5f); acc.y = fmaf(acc.y, 1.0001f, 0.5f);
                acc.z = fmaf(acc.z, 1.0001f, 0.5f); acc.w = fmaf(acc.w, 1.0001f, 0.5f);
            }
#pragma unroll
            for (int s = 0; s < M; ++s) __stcg(reinterpret_cast<float4*>(base + (16 + s) * plane + off), acc);
            if (scatter) {
                const int b0 = (nt * 4 * nq + g * nq + q) * 4;
                for (int e = 0; e < 4; ++e)
                    for (int c = 0; c < scatter; ++c) rowmajor[(size_t)(b0 + e) * ld + c * Np + i] = (&acc.x)[e];
            }
        }
    }
}

int main() {
    const int NT = 74, Np = 512, nq = 7;
    const size_t plane = (size_t)NT * 4 * nq * Np * 4;
    float* base; float* rm;
    cudaMalloc(&base, plane * 4 * 24);
    cudaMemset(base, 0, plane * 4 * 24);
    const int ld = 8 * Np;
    cudaMalloc(&rm, (size_t)NT * 112 * ld * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int L : {3, 10})
        for (int pipe = 0; pipe < 2; ++pipe)
            for (int C : {0, 40, 80, 160, 320}) {
                const int M = L == 3 ? 2 : 4, scatter = L == 3 ? 2 : 5;
                auto launch = [&]() {
#define CASE(LL, MM, PP) if (L == LL && pipe == PP) { cudaFuncSetAttribute(k<LL, MM, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k<LL, MM, PP><<<148, 576, 190 * 1024>>>(base, plane, NT, Np, nq, scatter, rm, ld, C); }
                    CASE(3, 2, 0) CASE(3, 2, 1) CASE(10, 4, 0) CASE(10, 4, 1)
                };
                for (int it = 0; it < 3; ++it) launch();
                cudaEventRecord(e0);
                const int reps = 20;
                for (int it = 0; it < reps; ++it) launch();
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double us = ms * 1000.0 / reps;
                const double bytes = (double)plane * 4 * (L + M) + (double)NT * 112 * Np * 4.0 * scatter;
                printf("L %2d M %d scatter %d pipe %d C %3d : %7.1f us  %6.2f TB/s  %5.1f GB/s/SM  err=%s\n", L, M, scatter, pipe, C, us,
                       bytes / us * 1e-6, bytes / us * 1e-3 / 148, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
