// Batched Wong-Wang (2006) decision-model generator: the training targets of the WTA task.
// Replaces the per-sample numpy loop DM.run_sim -> DM.simulate -> DM.update (reference src/ww_model.py:92-127) inside
// make_ds_wwp (reference scripts/wta_ode.py:56-93): 3 x 5001 sequential float64 updates per sample, 3010 samples.
// One thread integrates one sample through all three phases; every float64 operation is issued in the reference's order
// with explicit round-to-nearest intrinsics (no FMA contraction), so the only difference to numpy is the last ulp of exp().
// Output is the dataset tensor itself: r of every `every`-th update, first `time_steps` of them, as float32 (B, time_steps, 2).
#include "odecol_common.cuh"

namespace odecol {

struct WwParams {
    double gamma, tau_s, tau_ampa, j_within, j_between, j_ext, i_0, dt, sigma_noise;
};

ODECOL_DEVINL double ww_f(double x) {
    // (270. * x - 108) / (1. - np.exp(-0.154 * (270. * x - 108.)))
    const double a = __dsub_rn(__dmul_rn(270.0, x), 108.0);
    return __ddiv_rn(a, __dsub_rn(1.0, exp(__dmul_rn(-0.154, a))));
}

__global__ void __launch_bounds__(32) k_ww_generate(const double* __restrict__ mu, const double* __restrict__ i_noise0,
                                                    int B, int steps_per_phase, int every, int time_steps, WwParams q,
                                                    unsigned long long seed, long long trial_offset,
                                                    float* __restrict__ states) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double s0 = 0.1, s1 = 0.1;
    double n0 = i_noise0 ? i_noise0[2 * b] : 0.0, n1 = i_noise0 ? i_noise0[2 * b + 1] : 0.0;
    const double dsig = __dmul_rn(sqrt(__ddiv_rn(q.dt, q.tau_ampa)), q.sigma_noise);
    const Philox px(seed);
    const unsigned long long trial = (unsigned long long)(trial_offset + b);
    float* out = states + (size_t)b * time_steps * 2;
    long long k = 0;
    for (int phase = 0; phase < 3; ++phase) {
        const double e0 = phase == 1 ? __dmul_rn(q.j_ext, mu[2 * b]) : __dmul_rn(q.j_ext, 0.0);
        const double e1 = phase == 1 ? __dmul_rn(q.j_ext, mu[2 * b + 1]) : __dmul_rn(q.j_ext, 0.0);
        for (int it = 0; it < steps_per_phase; ++it, ++k) {
            // I_rec = W . s
            const double r0 = __dadd_rn(__dmul_rn(q.j_within, s0), __dmul_rn(-q.j_between, s1));
            const double r1 = __dadd_rn(__dmul_rn(-q.j_between, s0), __dmul_rn(q.j_within, s1));
            // I_noise += dt * (I_0 - I_noise) / tau_ampa + dsig * randn(2)
            double z0 = 0.0, z1 = 0.0;
            if (q.sigma_noise != 0.0) {
                const uint4 bits = px((uint32_t)trial, (uint32_t)(trial >> 32), (uint32_t)k, 0x40000000u | (uint32_t)(k >> 32));
                z0 = (double)normal_from_bits(bits.x, bits.y);
                z1 = (double)normal_from_bits(bits.z, bits.w);
            }
            n0 = __dadd_rn(n0, __dadd_rn(__ddiv_rn(__dmul_rn(q.dt, __dsub_rn(q.i_0, n0)), q.tau_ampa), __dmul_rn(dsig, z0)));
            n1 = __dadd_rn(n1, __dadd_rn(__ddiv_rn(__dmul_rn(q.dt, __dsub_rn(q.i_0, n1)), q.tau_ampa), __dmul_rn(dsig, z1)));
            const double x0 = __dadd_rn(__dadd_rn(r0, e0), n0), x1 = __dadd_rn(__dadd_rn(r1, e1), n1);
            const double f0 = ww_f(x0), f1 = ww_f(x1);
            // s += dt * (-s / tau_s + (1 - s) * gamma * r)
            s0 = __dadd_rn(s0, __dmul_rn(q.dt, __dadd_rn(__ddiv_rn(-s0, q.tau_s), __dmul_rn(__dmul_rn(__dsub_rn(1.0, s0), q.gamma), f0))));
            s1 = __dadd_rn(s1, __dmul_rn(q.dt, __dadd_rn(__ddiv_rn(-s1, q.tau_s), __dmul_rn(__dmul_rn(__dsub_rn(1.0, s1), q.gamma), f1))));
            if (k % every == 0) {
                const long long col = k / every;
                if (col < time_steps) { out[2 * col] = (float)f0; out[2 * col + 1] = (float)f1; }
            }
        }
    }
}

int launch_ww_generate(const double* mu, const double* i_noise0, int B, int steps_per_phase, int every, int time_steps,
                       double sigma_noise, uint64_t seed, int64_t trial_offset, float* states, cudaStream_t s) {
    // constants of DM.__init__ (reference src/ww_model.py:57-71)
    const WwParams q{0.641, 0.1, 0.002, 0.2609, 0.0497, 5.2e-4, 0.3255, 1e-3, sigma_noise};
    k_ww_generate<<<(B + 31) / 32, 32, 0, s>>>(mu, i_noise0, B, steps_per_phase, every, time_steps, q,
                                              (unsigned long long)seed, (long long)trial_offset, states);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? ODECOL_OK : ODECOL_E_CUDA;
}

}  // namespace odecol
