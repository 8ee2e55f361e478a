// PyTorch C++ extension over the C ABI (include/odecol.h).  Torch is plumbing here: it owns the device memory
// (outputs and workspaces come from its caching allocator) and names the stream; every computation happens behind
// the extern "C" entry points of libodecol.so.  Nothing in this file has a CPU path: tensors must be CUDA.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <optional>
#include <string>
#include "../../include/odecol.h"

namespace {

void check(int rc, const char* what) {
    if (rc == ODECOL_OK) return;
    // the message is assembled with std::string only: streaming an integer through an ostream from inside this module
    // crashed (facet lookup across two copies of libstdc++ state), turning every error code into a segmentation fault
    const std::string msg = std::string("odecol ") + what + " failed: " + odecol_strerror(rc) + " (" + std::to_string(rc) + ")";
    TORCH_CHECK(false, msg);
}

void want(const torch::Tensor& t, const char* name, c10::ScalarType dt = torch::kFloat32) {
    TORCH_CHECK(t.is_cuda(), "odecol: ", name, " must be a CUDA tensor (there is no CPU path)");
    TORCH_CHECK(t.scalar_type() == dt, "odecol: ", name, " has the wrong dtype");
    TORCH_CHECK(t.is_contiguous(), "odecol: ", name, " must be contiguous");
}

struct Problem {
    torch::Tensor W_aug, kappa, sigma, knot_t, knot_u, sigma_scale, lat_gain, W_local;
    bool has_sigma = false;
    odecol_problem p{};

    Problem(torch::Tensor W_aug_, torch::Tensor kappa_, std::optional<torch::Tensor> sigma_, torch::Tensor knot_t_,
            torch::Tensor knot_u_, int64_t n_in, int64_t B, double tau_s, double tau_m, double tau_a, double R,
            int64_t flags)
        : W_aug(std::move(W_aug_)), kappa(std::move(kappa_)), knot_t(std::move(knot_t_)), knot_u(std::move(knot_u_)) {
        want(W_aug, "W_aug"); want(kappa, "kappa"); want(knot_t, "knot_t"); want(knot_u, "knot_u");
        TORCH_CHECK(W_aug.dim() == 2 && kappa.dim() == 1 && knot_t.dim() == 1 && knot_u.dim() == 3, "odecol: bad ranks");
        const int64_t N = W_aug.size(0);
        TORCH_CHECK(kappa.size(0) == N, "odecol: kappa must have N entries");
        TORCH_CHECK(W_aug.size(1) >= N + n_in + 1 && W_aug.size(1) % 4 == 0, "odecol: W_aug needs >= N+n_in+1 columns, multiple of 4");
        TORCH_CHECK(knot_u.size(1) == knot_t.size(0) && knot_u.size(2) == n_in, "odecol: knot_u must be (B or 1, K, n_in)");
        TORCH_CHECK(knot_u.size(0) == B || knot_u.size(0) == 1, "odecol: knot_u batch must be B or 1");
        if (sigma_.has_value()) {
            sigma = *sigma_;
            want(sigma, "sigma");
            TORCH_CHECK(sigma.numel() == 3 * N, "odecol: sigma must have 3N entries");
            has_sigma = true;
        }
        p.N = (int32_t)N; p.n_in = (int32_t)n_in; p.B = (int32_t)B; p.K = (int32_t)knot_t.size(0);
        p.ld_w = (int32_t)W_aug.size(1); p.flags = (int32_t)flags;
        p.W_aug = W_aug.data_ptr<float>(); p.kappa = kappa.data_ptr<float>();
        p.sigma = has_sigma ? sigma.data_ptr<float>() : nullptr;
        p.knot_t = knot_t.data_ptr<float>(); p.knot_u = knot_u.data_ptr<float>();
        p.knot_stride_b = knot_u.size(0) == 1 ? 0 : knot_u.size(1) * knot_u.size(2);
        p.tau_s = (float)tau_s; p.tau_m = (float)tau_m; p.tau_a = (float)tau_a; p.resistance = (float)R;
    }

    // per-trial factor on sigma (the noise axis of a parameter sweep); undefined tensor = none
    void set_sigma_scale(std::optional<torch::Tensor> sc) {
        if (!sc.has_value()) { sigma_scale = torch::Tensor(); p.sigma_scale = nullptr; return; }
        want(*sc, "sigma_scale");
        TORCH_CHECK(sc->numel() == p.B, "odecol: sigma_scale must have B entries");
        sigma_scale = *sc;
        p.sigma_scale = sigma_scale.data_ptr<float>();
    }

    // per-trial gain on the dense recurrent input + the within-column weights that stay outside it (include/odecol.h)
    void set_lateral_gain(std::optional<torch::Tensor> gain, std::optional<torch::Tensor> w_local) {
        if (!gain.has_value()) { lat_gain = torch::Tensor(); W_local = torch::Tensor(); p.lat_gain = nullptr; p.W_local = nullptr; return; }
        want(*gain, "lateral_gain");
        TORCH_CHECK(gain->numel() == p.B, "odecol: lateral_gain must have B entries");
        lat_gain = *gain;
        p.lat_gain = lat_gain.data_ptr<float>();
        p.W_local = nullptr;
        if (w_local.has_value()) {
            want(*w_local, "W_local");
            TORCH_CHECK(w_local->dim() == 2 && w_local->size(0) == p.N && w_local->size(1) == 8, "odecol: W_local must be (N, 8)");
            W_local = *w_local;
            p.W_local = W_local.data_ptr<float>();
        }
    }

    torch::TensorOptions fopts() const { return W_aug.options(); }
    torch::TensorOptions iopts() const { return W_aug.options().dtype(torch::kInt32); }
    int64_t N() const { return p.N; }
    int64_t B() const { return p.B; }

    torch::Tensor workspace(int op, int64_t T, int64_t n_steps = 0) const {
        const size_t bytes = odecol_workspace_bytes(&p, op, (int32_t)T, n_steps);
        return torch::empty({(int64_t)((bytes + 15) / 16 * 16)}, W_aug.options().dtype(torch::kUInt8));
    }
    void* stream() const { return at::cuda::getCurrentCUDAStream(W_aug.device().index()).stream(); }
};

void check_state(const Problem& pr, const torch::Tensor& y, const char* name) {
    want(y, name);
    TORCH_CHECK(y.dim() == 2 && y.size(0) == pr.B() && y.size(1) == 3 * pr.N(), "odecol: ", name, " must be (B, 3N)");
}

torch::Tensor rhs(const Problem& pr, const torch::Tensor& t, const torch::Tensor& y) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    check_state(pr, y, "y");
    want(t, "t");
    TORCH_CHECK(t.numel() == pr.B(), "odecol: t must have one entry per trial");
    auto f = torch::empty_like(y);
    check(odecol_rhs(&pr.p, t.data_ptr<float>(), y.data_ptr<float>(), f.data_ptr<float>(), pr.stream()), "rhs");
    return f;
}

torch::Tensor drift_staged(const Problem& pr, const torch::Tensor& t, const torch::Tensor& y) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    check_state(pr, y, "y");
    want(t, "t");
    TORCH_CHECK(t.numel() == pr.B(), "odecol: t must have one entry per trial");
    auto f = torch::empty_like(y);
    Problem staged = pr;                                   // workspace of the staged family whatever the size
    staged.p.flags |= ODECOL_FLAG_FORCE_STAGED;
    auto ws = staged.workspace(ODECOL_OP_EM_FWD, 2);
    check(odecol_drift_staged(&pr.p, t.data_ptr<float>(), y.data_ptr<float>(), f.data_ptr<float>(), ws.data_ptr(),
                              (size_t)ws.numel(), pr.stream()), "drift_staged");
    return f;
}

torch::Tensor rk4_fwd(const Problem& pr, const torch::Tensor& t, const torch::Tensor& y0, int64_t out_every) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    check_state(pr, y0, "y0");
    want(t, "t");
    const int64_t T = t.numel();
    TORCH_CHECK(T >= 2 && out_every >= 1, "odecol: need at least two time points");
    const int64_t rows = (T - 2) / out_every + 2;
    auto y = torch::empty({rows, pr.B(), 3 * pr.N()}, pr.fopts());
    auto ws = pr.workspace(ODECOL_OP_RK4_FWD, T);
    check(odecol_rk4_fwd(&pr.p, t.data_ptr<float>(), (int32_t)T, y0.data_ptr<float>(), y.data_ptr<float>(),
                         (int32_t)out_every, ws.data_ptr(), (size_t)ws.numel(), pr.stream()), "rk4_fwd");
    return y;
}

std::vector<torch::Tensor> rk4_bwd(const Problem& pr, const torch::Tensor& t, const torch::Tensor& y_traj,
                                   const torch::Tensor& grad_y, std::optional<torch::Tensor> sel) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    want(t, "t"); want(y_traj, "y_traj"); want(grad_y, "grad_y");
    const int64_t T = t.numel();
    TORCH_CHECK(y_traj.dim() == 3 && y_traj.size(0) == T && y_traj.size(1) == pr.B() && y_traj.size(2) == 3 * pr.N(),
                "odecol: y_traj must be (T, B, 3N)");
    TORCH_CHECK(grad_y.dim() == 3 && grad_y.size(0) == T && grad_y.size(1) == pr.B(), "odecol: grad_y must be (T, B, G)");
    const int64_t G = grad_y.size(2);
    const int32_t* selp = nullptr;
    if (sel.has_value()) {
        want(*sel, "sel", torch::kInt32);
        TORCH_CHECK(sel->numel() == G, "odecol: sel must have G entries");
        selp = sel->data_ptr<int32_t>();
    } else {
        TORCH_CHECK(G == 3 * pr.N(), "odecol: dense grad_y must have 3N components");
    }
    auto gy0 = torch::empty({pr.B(), 3 * pr.N()}, pr.fopts());
    auto gW = torch::empty_like(pr.W_aug);
    auto ws = pr.workspace(ODECOL_OP_RK4_BWD, T);
    check(odecol_rk4_bwd(&pr.p, t.data_ptr<float>(), (int32_t)T, y_traj.data_ptr<float>(), grad_y.data_ptr<float>(), selp,
                         (int32_t)G, gy0.data_ptr<float>(), gW.data_ptr<float>(), ws.data_ptr(), (size_t)ws.numel(),
                         pr.stream()), "rk4_bwd");
    return {gy0, gW};
}

// checkpoint mode (include/odecol.h): bytes of the checkpoint buffer, 0 if the problem's kernel family has no such mode
int64_t rk4_ckpt_bytes(const Problem& pr, int64_t T) { return (int64_t)odecol_rk4_ckpt_bytes(&pr.p, (int32_t)T); }

std::vector<torch::Tensor> rk4_fwd_ckpt(const Problem& pr, const torch::Tensor& t, const torch::Tensor& y0,
                                        const torch::Tensor& sel) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    check_state(pr, y0, "y0");
    want(t, "t"); want(sel, "sel", torch::kInt32);
    const int64_t T = t.numel(), G = sel.numel();
    const size_t cb = odecol_rk4_ckpt_bytes(&pr.p, (int32_t)T);
    TORCH_CHECK(cb > 0, "odecol: checkpoint mode is not available for this problem");
    auto y_sel = torch::empty({T, pr.B(), G}, pr.fopts());
    auto ckpt = torch::empty({(int64_t)cb}, torch::TensorOptions().dtype(torch::kUInt8).device(pr.W_aug.device()));
    auto ws = pr.workspace(ODECOL_OP_RK4_FWD, T);
    check(odecol_rk4_fwd_ckpt(&pr.p, t.data_ptr<float>(), (int32_t)T, y0.data_ptr<float>(), sel.data_ptr<int32_t>(), (int32_t)G,
                              y_sel.data_ptr<float>(), ckpt.data_ptr(), cb, ws.data_ptr(), (size_t)ws.numel(), pr.stream()),
          "rk4_fwd_ckpt");
    return {y_sel, ckpt};
}

std::vector<torch::Tensor> rk4_bwd_ckpt(const Problem& pr, const torch::Tensor& t, const torch::Tensor& ckpt,
                                        const torch::Tensor& grad_y, const torch::Tensor& sel) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    want(t, "t"); want(grad_y, "grad_y"); want(sel, "sel", torch::kInt32);
    TORCH_CHECK(ckpt.is_cuda() && ckpt.is_contiguous() && ckpt.scalar_type() == torch::kUInt8, "odecol: bad checkpoint buffer");
    const int64_t T = t.numel(), G = sel.numel();
    TORCH_CHECK(grad_y.dim() == 3 && grad_y.size(0) == T && grad_y.size(1) == pr.B() && grad_y.size(2) == G,
                "odecol: grad_y must be (T, B, G)");
    auto gy0 = torch::empty({pr.B(), 3 * pr.N()}, pr.fopts());
    auto gW = torch::empty_like(pr.W_aug);
    auto ws = pr.workspace(ODECOL_OP_RK4_BWD, T);
    check(odecol_rk4_bwd_ckpt(&pr.p, t.data_ptr<float>(), (int32_t)T, ckpt.data_ptr(), (size_t)ckpt.numel(),
                              grad_y.data_ptr<float>(), sel.data_ptr<int32_t>(), (int32_t)G, gy0.data_ptr<float>(),
                              gW.data_ptr<float>(), ws.data_ptr(), (size_t)ws.numel(), pr.stream()), "rk4_bwd_ckpt");
    return {gy0, gW};
}

std::vector<torch::Tensor> dopri5_fwd(const Problem& pr, const torch::Tensor& t, const torch::Tensor& y0, double rtol,
                                      double atol, int64_t max_steps) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    check_state(pr, y0, "y0");
    want(t, "t");
    const int64_t T = t.numel();
    auto y = torch::empty({T, pr.B(), 3 * pr.N()}, pr.fopts());
    auto na = torch::zeros({pr.B()}, pr.iopts()), nr = torch::zeros({pr.B()}, pr.iopts()), st = torch::zeros({pr.B()}, pr.iopts());
    auto ws = pr.workspace(ODECOL_OP_DOPRI5_FWD, T);
    check(odecol_dopri5_fwd(&pr.p, t.data_ptr<float>(), (int32_t)T, y0.data_ptr<float>(), y.data_ptr<float>(), (float)rtol,
                            (float)atol, (int32_t)std::min<int64_t>(max_steps, INT32_MAX), na.data_ptr<int32_t>(),
                            nr.data_ptr<int32_t>(), st.data_ptr<int32_t>(), ws.data_ptr(), (size_t)ws.numel(), pr.stream()),
          "dopri5_fwd");
    return {y, na, nr, st};
}

// training mode: (y, n_accept, n_reject, status, rec_y, rec_t0, rec_dt, out_step, out_x)
std::vector<torch::Tensor> dopri5_fwd_record(const Problem& pr, const torch::Tensor& t, const torch::Tensor& y0, double rtol,
                                             double atol, int64_t max_steps, int64_t cap) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    check_state(pr, y0, "y0");
    want(t, "t");
    const int64_t T = t.numel();
    auto y = torch::empty({T, pr.B(), 3 * pr.N()}, pr.fopts());
    auto na = torch::zeros({pr.B()}, pr.iopts()), nr = torch::zeros({pr.B()}, pr.iopts()), st = torch::zeros({pr.B()}, pr.iopts());
    auto rec_y = torch::empty({cap, pr.B(), 3 * pr.N()}, pr.fopts());
    auto rec_t0 = torch::zeros({pr.B(), cap}, pr.fopts().dtype(torch::kFloat64));
    auto rec_dt = torch::zeros({pr.B(), cap}, pr.fopts().dtype(torch::kFloat64));
    auto out_step = torch::zeros({pr.B(), T}, pr.iopts());
    auto out_x = torch::zeros({pr.B(), T}, pr.fopts());
    auto ws = pr.workspace(ODECOL_OP_DOPRI5_FWD, T);
    check(odecol_dopri5_fwd_record(&pr.p, t.data_ptr<float>(), (int32_t)T, y0.data_ptr<float>(), y.data_ptr<float>(), (float)rtol,
                                   (float)atol, (int32_t)std::min<int64_t>(max_steps, INT32_MAX), na.data_ptr<int32_t>(),
                                   nr.data_ptr<int32_t>(), st.data_ptr<int32_t>(), rec_y.data_ptr<float>(),
                                   rec_t0.data_ptr<double>(), rec_dt.data_ptr<double>(), out_step.data_ptr<int32_t>(),
                                   out_x.data_ptr<float>(), (int32_t)cap, ws.data_ptr(), (size_t)ws.numel(), pr.stream()),
          "dopri5_fwd_record");
    return {y, na, nr, st, rec_y, rec_t0, rec_dt, out_step, out_x};
}

std::vector<torch::Tensor> dopri5_bwd(const Problem& pr, int64_t T, const torch::Tensor& rec_y, const torch::Tensor& rec_t0,
                                      const torch::Tensor& rec_dt, const torch::Tensor& out_step, const torch::Tensor& out_x,
                                      const torch::Tensor& n_accept, const torch::Tensor& grad_y,
                                      std::optional<torch::Tensor> sel) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    want(rec_y, "rec_y"); want(rec_t0, "rec_t0", torch::kFloat64); want(rec_dt, "rec_dt", torch::kFloat64);
    want(out_step, "out_step", torch::kInt32); want(out_x, "out_x"); want(n_accept, "n_accept", torch::kInt32);
    want(grad_y, "grad_y");
    TORCH_CHECK(grad_y.dim() == 3 && grad_y.size(0) == T && grad_y.size(1) == pr.B(), "odecol: grad_y must be (T, B, G)");
    const int64_t G = grad_y.size(2), cap = rec_y.size(0);
    const int32_t* selp = nullptr;
    if (sel.has_value()) {
        want(*sel, "sel", torch::kInt32);
        TORCH_CHECK(sel->numel() == G, "odecol: sel must have G entries");
        selp = sel->data_ptr<int32_t>();
    } else {
        TORCH_CHECK(G == 3 * pr.N(), "odecol: dense grad_y must have 3N components");
    }
    auto gy0 = torch::empty({pr.B(), 3 * pr.N()}, pr.fopts());
    auto gW = torch::empty_like(pr.W_aug);
    auto ws = pr.workspace(ODECOL_OP_DOPRI5_BWD, T);
    check(odecol_dopri5_bwd(&pr.p, (int32_t)T, rec_y.data_ptr<float>(), rec_t0.data_ptr<double>(), rec_dt.data_ptr<double>(),
                            out_step.data_ptr<int32_t>(), out_x.data_ptr<float>(), (int32_t)cap, n_accept.data_ptr<int32_t>(),
                            grad_y.data_ptr<float>(), selp, (int32_t)G, gy0.data_ptr<float>(), gW.data_ptr<float>(), ws.data_ptr(),
                            (size_t)ws.numel(), pr.stream()), "dopri5_bwd");
    return {gy0, gW};
}

int64_t em_num_steps(const torch::Tensor& ts_cpu, double dt) {
    TORCH_CHECK(!ts_cpu.is_cuda() && ts_cpu.scalar_type() == torch::kFloat32 && ts_cpu.is_contiguous(),
                "odecol: em_num_steps wants a contiguous float32 CPU tensor");
    return odecol_em_num_steps(ts_cpu.data_ptr<float>(), (int32_t)ts_cpu.numel(), (float)dt);
}

std::vector<torch::Tensor> em_fwd(const Problem& pr, const torch::Tensor& ts, const torch::Tensor& y0,
                                  std::optional<torch::Tensor> dW, int64_t seed, int64_t trial_offset, double dt,
                                  bool adaptive, double rtol, double atol, double dt_min, int64_t save_steps) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    check_state(pr, y0, "y0");
    want(ts, "ts");
    const int64_t T = ts.numel();
    const float* dWp = nullptr;
    if (dW.has_value()) {
        want(*dW, "dW");
        TORCH_CHECK(dW->dim() == 2 && dW->size(1) == pr.B(), "odecol: dW must be (n_steps, B)");
        dWp = dW->data_ptr<float>();
    }
    auto y = torch::empty({T, pr.B(), 3 * pr.N()}, pr.fopts());
    auto na = torch::zeros({pr.B()}, pr.iopts()), nr = torch::zeros({pr.B()}, pr.iopts()), st = torch::zeros({pr.B()}, pr.iopts());
    torch::Tensor ysteps;
    float* ysp = nullptr;
    if (save_steps > 0) {
        ysteps = torch::empty({save_steps + 1, pr.B(), 3 * pr.N()}, pr.fopts());
        ysp = ysteps.data_ptr<float>();
    } else {
        ysteps = torch::empty({0}, pr.fopts());
    }
    auto ws = pr.workspace(ODECOL_OP_EM_FWD, T);
    check(odecol_em_fwd(&pr.p, ts.data_ptr<float>(), (int32_t)T, y0.data_ptr<float>(), y.data_ptr<float>(), dWp,
                        (uint64_t)seed, trial_offset, (float)dt, adaptive ? 1 : 0, (float)rtol, (float)atol, (float)dt_min,
                        na.data_ptr<int32_t>(), nr.data_ptr<int32_t>(), st.data_ptr<int32_t>(), ysp, ws.data_ptr(),
                        (size_t)ws.numel(), pr.stream()), "em_fwd");
    return {y, na, nr, st, ysteps};
}

std::vector<torch::Tensor> em_bwd(const Problem& pr, const torch::Tensor& ts, const torch::Tensor& y_steps,
                                  const torch::Tensor& grad_y, std::optional<torch::Tensor> sel, double dt) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    want(ts, "ts"); want(y_steps, "y_steps"); want(grad_y, "grad_y");
    const int64_t T = ts.numel();
    TORCH_CHECK(y_steps.dim() == 3 && y_steps.size(1) == pr.B() && y_steps.size(2) == 3 * pr.N(), "odecol: y_steps must be (n_steps+1, B, 3N)");
    const int64_t n_steps = y_steps.size(0) - 1;
    TORCH_CHECK(grad_y.dim() == 3 && grad_y.size(0) == T && grad_y.size(1) == pr.B(), "odecol: grad_y must be (T, B, G)");
    const int64_t G = grad_y.size(2);
    const int32_t* selp = nullptr;
    if (sel.has_value()) {
        want(*sel, "sel", torch::kInt32);
        TORCH_CHECK(sel->numel() == G, "odecol: sel must have G entries");
        selp = sel->data_ptr<int32_t>();
    } else {
        TORCH_CHECK(G == 3 * pr.N(), "odecol: dense grad_y must have 3N components");
    }
    auto gy0 = torch::empty({pr.B(), 3 * pr.N()}, pr.fopts());
    auto gW = torch::empty_like(pr.W_aug);
    auto ws = pr.workspace(ODECOL_OP_EM_BWD, T, n_steps);
    check(odecol_em_bwd(&pr.p, ts.data_ptr<float>(), (int32_t)T, y_steps.data_ptr<float>(), n_steps, grad_y.data_ptr<float>(),
                        selp, (int32_t)G, (float)dt, gy0.data_ptr<float>(), gW.data_ptr<float>(), ws.data_ptr(),
                        (size_t)ws.numel(), pr.stream()), "em_bwd");
    return {gy0, gW};
}

// (dW, dU) tables of the srk entry points: both or neither
static void srk_tables(const Problem& pr, const std::optional<torch::Tensor>& dW, const std::optional<torch::Tensor>& dU,
                       const float*& dWp, const float*& dUp) {
    dWp = dUp = nullptr;
    TORCH_CHECK(dW.has_value() == dU.has_value(), "odecol: srk wants both dW and dU, or neither (Philox)");
    if (!dW.has_value()) return;
    want(*dW, "dW"); want(*dU, "dU");
    TORCH_CHECK(dW->dim() == 2 && dW->size(1) == pr.B() && dU->sizes() == dW->sizes(), "odecol: dW and dU must be (n_steps, B)");
    dWp = dW->data_ptr<float>();
    dUp = dU->data_ptr<float>();
}

std::vector<torch::Tensor> srk_fwd(const Problem& pr, const torch::Tensor& ts, const torch::Tensor& y0,
                                   std::optional<torch::Tensor> dW, std::optional<torch::Tensor> dU, int64_t seed,
                                   int64_t trial_offset, double dt, int64_t save_steps) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    check_state(pr, y0, "y0");
    want(ts, "ts");
    const int64_t T = ts.numel();
    const float *dWp, *dUp;
    srk_tables(pr, dW, dU, dWp, dUp);
    auto y = torch::empty({T, pr.B(), 3 * pr.N()}, pr.fopts());
    auto st = torch::zeros({pr.B()}, pr.iopts());
    torch::Tensor ysteps = save_steps > 0 ? torch::empty({save_steps + 1, pr.B(), 3 * pr.N()}, pr.fopts()) : torch::empty({0}, pr.fopts());
    auto ws = pr.workspace(ODECOL_OP_SRK_FWD, T);
    check(odecol_srk_fwd(&pr.p, ts.data_ptr<float>(), (int32_t)T, y0.data_ptr<float>(), y.data_ptr<float>(), dWp, dUp,
                         (uint64_t)seed, trial_offset, (float)dt, st.data_ptr<int32_t>(),
                         save_steps > 0 ? ysteps.data_ptr<float>() : nullptr, ws.data_ptr(), (size_t)ws.numel(), pr.stream()),
          "srk_fwd");
    return {y, st, ysteps};
}

// adaptive srk (on-chip family): (y, n_accept, n_reject, status)
std::vector<torch::Tensor> srk_fwd_adaptive(const Problem& pr, const torch::Tensor& ts, const torch::Tensor& y0, int64_t seed,
                                            int64_t trial_offset, double dt, double rtol, double atol, double dt_min) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    check_state(pr, y0, "y0");
    want(ts, "ts");
    const int64_t T = ts.numel();
    auto y = torch::empty({T, pr.B(), 3 * pr.N()}, pr.fopts());
    auto na = torch::zeros({pr.B()}, pr.iopts()), nr = torch::zeros({pr.B()}, pr.iopts()), st = torch::zeros({pr.B()}, pr.iopts());
    check(odecol_srk_fwd_adaptive(&pr.p, ts.data_ptr<float>(), (int32_t)T, y0.data_ptr<float>(), y.data_ptr<float>(), (uint64_t)seed,
                                  trial_offset, (float)dt, (float)rtol, (float)atol, (float)dt_min, na.data_ptr<int32_t>(),
                                  nr.data_ptr<int32_t>(), st.data_ptr<int32_t>(), nullptr, 0, pr.stream()), "srk_fwd_adaptive");
    return {y, na, nr, st};
}

// cumulative (W, int W) of the Levy-area-consistent tree: t (M,) -> two float64 tensors (M, B)
std::vector<torch::Tensor> brownian_levy_query(int64_t seed, int64_t trial_offset, int64_t B, double t_begin, double t_end,
                                               const torch::Tensor& t) {
    c10::cuda::CUDAGuard g(t.device());
    want(t, "t");
    auto w = torch::empty({t.numel(), B}, t.options().dtype(torch::kFloat64));
    auto iw = torch::empty_like(w);
    check(odecol_brownian_levy_query((uint64_t)seed, trial_offset, (int32_t)B, (float)t_begin, (float)t_end, t.data_ptr<float>(),
                                     (int32_t)t.numel(), w.data_ptr<double>(), iw.data_ptr<double>(),
                                     at::cuda::getCurrentCUDAStream(t.device().index()).stream()), "brownian_levy_query");
    return {w, iw};
}

std::vector<torch::Tensor> srk_bwd(const Problem& pr, const torch::Tensor& ts, const torch::Tensor& y_steps,
                                   std::optional<torch::Tensor> dW, std::optional<torch::Tensor> dU, int64_t seed,
                                   int64_t trial_offset, const torch::Tensor& grad_y, std::optional<torch::Tensor> sel,
                                   double dt) {
    c10::cuda::CUDAGuard g(pr.W_aug.device());
    want(ts, "ts"); want(y_steps, "y_steps"); want(grad_y, "grad_y");
    const int64_t T = ts.numel();
    TORCH_CHECK(y_steps.dim() == 3 && y_steps.size(1) == pr.B() && y_steps.size(2) == 3 * pr.N(), "odecol: y_steps must be (n_steps+1, B, 3N)");
    const int64_t n_steps = y_steps.size(0) - 1;
    TORCH_CHECK(grad_y.dim() == 3 && grad_y.size(0) == T && grad_y.size(1) == pr.B(), "odecol: grad_y must be (T, B, G)");
    const int64_t G = grad_y.size(2);
    const float *dWp, *dUp;
    srk_tables(pr, dW, dU, dWp, dUp);
    TORCH_CHECK(dWp == nullptr || dW->size(0) == n_steps, "odecol: dW/dU must have n_steps rows");
    const int32_t* selp = nullptr;
    if (sel.has_value()) {
        want(*sel, "sel", torch::kInt32);
        TORCH_CHECK(sel->numel() == G, "odecol: sel must have G entries");
        selp = sel->data_ptr<int32_t>();
    } else {
        TORCH_CHECK(G == 3 * pr.N(), "odecol: dense grad_y must have 3N components");
    }
    auto gy0 = torch::empty({pr.B(), 3 * pr.N()}, pr.fopts());
    auto gW = torch::empty_like(pr.W_aug);
    auto ws = pr.workspace(ODECOL_OP_SRK_BWD, T, n_steps);
    check(odecol_srk_bwd(&pr.p, ts.data_ptr<float>(), (int32_t)T, y_steps.data_ptr<float>(), n_steps, dWp, dUp, (uint64_t)seed,
                         trial_offset, grad_y.data_ptr<float>(), selp, (int32_t)G, (float)dt, gy0.data_ptr<float>(),
                         gW.data_ptr<float>(), ws.data_ptr(), (size_t)ws.numel(), pr.stream()), "srk_bwd");
    return {gy0, gW};
}

// W(t) - W(t_begin) on the virtual Brownian tree of the adaptive solvers: t (M,) -> (M, B)
torch::Tensor brownian_query(int64_t seed, int64_t trial_offset, int64_t B, double t_begin, double t_end, const torch::Tensor& t) {
    c10::cuda::CUDAGuard g(t.device());
    want(t, "t");
    auto w = torch::empty({t.numel(), B}, t.options());
    check(odecol_brownian_query((uint64_t)seed, trial_offset, (int32_t)B, (float)t_begin, (float)t_end, t.data_ptr<float>(),
                                (int32_t)t.numel(), w.data_ptr<float>(), at::cuda::getCurrentCUDAStream(t.device().index()).stream()),
          "brownian_query");
    return w;
}

torch::Tensor ww_generate(const torch::Tensor& mu, std::optional<torch::Tensor> i_noise0, int64_t steps_per_phase,
                          int64_t every, int64_t time_steps, double sigma_noise, int64_t seed, int64_t trial_offset) {
    c10::cuda::CUDAGuard g(mu.device());
    want(mu, "mu", torch::kFloat64);
    TORCH_CHECK(mu.dim() == 2 && mu.size(1) == 2, "odecol: mu must be (B, 2) float64");
    const double* i0 = nullptr;
    if (i_noise0.has_value()) {
        want(*i_noise0, "i_noise0", torch::kFloat64);
        TORCH_CHECK(i_noise0->sizes() == mu.sizes(), "odecol: i_noise0 must be (B, 2) float64");
        i0 = i_noise0->data_ptr<double>();
    }
    auto states = torch::empty({mu.size(0), time_steps, 2}, mu.options().dtype(torch::kFloat32));
    check(odecol_ww_generate(mu.data_ptr<double>(), i0, (int32_t)mu.size(0), (int32_t)steps_per_phase, (int32_t)every,
                             (int32_t)time_steps, sigma_noise, (uint64_t)seed, trial_offset, states.data_ptr<float>(),
                             at::cuda::getCurrentCUDAStream(mu.device().index()).stream()), "ww_generate");
    return states;
}

// fused read-out loss: returns {loss (0-d), grad_y_sel}
std::vector<torch::Tensor> huber_rate_loss(const torch::Tensor& y_sel, const torch::Tensor& target, int64_t P,
                                           std::optional<torch::Tensor> w, double beta) {
    c10::cuda::CUDAGuard g(y_sel.device());
    want(y_sel, "y_sel");
    TORCH_CHECK(y_sel.dim() == 3 && P >= 1 && y_sel.size(2) % (2 * P) == 0, "odecol: y_sel must be (T, B, 2*G*P)");
    const int64_t T = y_sel.size(0), B = y_sel.size(1), G = y_sel.size(2) / (2 * P);
    TORCH_CHECK(target.is_cuda() && target.scalar_type() == torch::kFloat32 && target.dim() == 3, "odecol: target must be a 3-d float32 CUDA tensor");
    TORCH_CHECK((target.size(0) == T || target.size(0) == 1) && (target.size(1) == B || target.size(1) == 1) &&
                (target.size(2) == G || target.size(2) == 1), "odecol: target must broadcast to (T, B, G)");
    const float* wp = nullptr;
    if (w.has_value()) { want(*w, "w"); TORCH_CHECK(w->numel() == P, "odecol: w must have P entries"); wp = w->data_ptr<float>(); }
    auto st = [&](int d) { return target.size(d) == 1 ? (int64_t)0 : target.stride(d); };
    auto loss = torch::empty({}, y_sel.options());
    auto grad = torch::empty_like(y_sel);
    auto ws = torch::empty({1}, y_sel.options().dtype(torch::kFloat64));
    check(odecol_huber_rate_loss(y_sel.data_ptr<float>(), (int32_t)T, (int32_t)B, (int32_t)G, (int32_t)P, wp,
                                 target.data_ptr<float>(), st(0), st(1), st(2), (float)beta, loss.data_ptr<float>(),
                                 grad.data_ptr<float>(), ws.data_ptr(), sizeof(double),
                                 at::cuda::getCurrentCUDAStream(y_sel.device().index()).stream()), "huber_rate_loss");
    return {loss, grad};
}

// fused window read-out (XOR final point / parity last-100 mean): returns {loss (0-d), pred (B), grad_y_sel, grad_w (P)}
std::vector<torch::Tensor> window_rate_l1_loss(const torch::Tensor& y_sel, const torch::Tensor& target, int64_t last,
                                               std::optional<torch::Tensor> w) {
    c10::cuda::CUDAGuard g(y_sel.device());
    want(y_sel, "y_sel"); want(target, "target");
    TORCH_CHECK(y_sel.dim() == 3 && y_sel.size(2) % 2 == 0, "odecol: y_sel must be (T, B, 2*P)");
    const int64_t T = y_sel.size(0), B = y_sel.size(1), P = y_sel.size(2) / 2;
    TORCH_CHECK(target.numel() == B, "odecol: target must have one entry per trial");
    const float* wp = nullptr;
    if (w.has_value()) { want(*w, "w"); TORCH_CHECK(w->numel() == P, "odecol: w must have P entries"); wp = w->data_ptr<float>(); }
    auto loss = torch::empty({}, y_sel.options());
    auto pred = torch::empty({B}, y_sel.options());
    auto grad = torch::empty_like(y_sel);
    auto grad_w = torch::empty({B, P}, y_sel.options());
    auto ws = torch::empty({1}, y_sel.options().dtype(torch::kFloat64));
    check(odecol_window_rate_l1_loss(y_sel.data_ptr<float>(), (int32_t)T, (int32_t)B, (int32_t)P, (int32_t)last, wp,
                                     target.data_ptr<float>(), loss.data_ptr<float>(), pred.data_ptr<float>(),
                                     grad.data_ptr<float>(), grad_w.data_ptr<float>(), ws.data_ptr(), sizeof(double),
                                     at::cuda::getCurrentCUDAStream(y_sel.device().index()).stream()), "window_rate_l1_loss");
    return {loss, pred, grad, grad_w.sum(0)};
}

torch::Tensor tc_contract(const torch::Tensor& A, const torch::Tensor& B) {
    c10::cuda::CUDAGuard g(A.device());
    want(A, "A"); want(B, "B");
    TORCH_CHECK(A.dim() == 2 && B.dim() == 2 && A.size(1) == B.size(1), "odecol: tc_contract wants A (M,K), B (N,K)");
    const int64_t M = A.size(0), N = B.size(0), K = A.size(1);
    auto C = torch::empty({N, M}, A.options());
    const size_t bytes = odecol_tc_contract_workspace_bytes((int32_t)M, (int32_t)N, (int32_t)K);
    auto ws = torch::empty({(int64_t)bytes}, A.options().dtype(torch::kUInt8));
    check(odecol_tc_contract(A.data_ptr<float>(), B.data_ptr<float>(), C.data_ptr<float>(), (int32_t)M, (int32_t)N, (int32_t)K,
                             ws.data_ptr(), bytes, at::cuda::getCurrentCUDAStream(A.device().index()).stream()), "tc_contract");
    return C;
}

torch::Tensor tc_contract_tn(const torch::Tensor& A, const torch::Tensor& B) {
    c10::cuda::CUDAGuard g(A.device());
    want(A, "A"); want(B, "B");
    TORCH_CHECK(A.dim() == 2 && B.dim() == 2 && A.size(0) == B.size(0), "odecol: tc_contract_tn wants A (K,M), B (K,N)");
    const int64_t K = A.size(0), M = A.size(1), N = B.size(1);
    auto C = torch::empty({M, N}, A.options());
    const size_t bytes = odecol_tc_contract_tn_workspace_bytes((int32_t)M, (int32_t)N, (int32_t)K);
    auto ws = torch::empty({(int64_t)bytes}, A.options().dtype(torch::kUInt8));
    check(odecol_tc_contract_tn(A.data_ptr<float>(), B.data_ptr<float>(), C.data_ptr<float>(), (int32_t)M, (int32_t)N, (int32_t)K,
                                ws.data_ptr(), bytes, at::cuda::getCurrentCUDAStream(A.device().index()).stream()), "tc_contract_tn");
    return C;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "odecol: PyTorch binding of the sm_100a fused column-ODE integrator (C ABI in include/odecol.h)";
    py::class_<Problem>(m, "Problem")
        .def(py::init<torch::Tensor, torch::Tensor, std::optional<torch::Tensor>, torch::Tensor, torch::Tensor, int64_t,
                      int64_t, double, double, double, double, int64_t>(),
             py::arg("W_aug"), py::arg("kappa"), py::arg("sigma"), py::arg("knot_t"), py::arg("knot_u"), py::arg("n_in"),
             py::arg("B"), py::arg("tau_s"), py::arg("tau_m"), py::arg("tau_a"), py::arg("resistance"), py::arg("flags") = 0)
        .def("set_sigma_scale", &Problem::set_sigma_scale)
        .def("set_lateral_gain", &Problem::set_lateral_gain)
        .def_property_readonly("N", &Problem::N)
        .def_property_readonly("B", &Problem::B)
        .def("kernel_family", [](const Problem& pr, int op) { return odecol_kernel_family(&pr.p, op); })
        .def("workspace_bytes", [](const Problem& pr, int op, int64_t T, int64_t n_steps) {
            return (int64_t)odecol_workspace_bytes(&pr.p, op, (int32_t)T, n_steps);
        });
    m.def("abi_version", &odecol_abi_version);
    m.def("check_code", [](int rc) { check(rc, "self-test"); }, "raise the RuntimeError a failing entry point would raise");
    m.def("strerror", [](int c) { return std::string(odecol_strerror(c)); });
    m.def("last_launch_count", &odecol_last_launch_count);
    m.def("rhs", &rhs);
    m.def("drift_staged", &drift_staged);
    m.def("rk4_fwd", &rk4_fwd);
    m.def("rk4_bwd", &rk4_bwd);
    m.def("rk4_ckpt_bytes", &rk4_ckpt_bytes);
    m.def("rk4_fwd_ckpt", &rk4_fwd_ckpt);
    m.def("rk4_bwd_ckpt", &rk4_bwd_ckpt);
    m.def("dopri5_fwd", &dopri5_fwd);
    m.def("dopri5_fwd_record", &dopri5_fwd_record);
    m.def("dopri5_bwd", &dopri5_bwd);
    m.def("em_num_steps", &em_num_steps);
    m.def("em_fwd", &em_fwd);
    m.def("em_bwd", &em_bwd);
    m.def("srk_fwd", &srk_fwd);
    m.def("srk_bwd", &srk_bwd);
    m.def("srk_fwd_adaptive", &srk_fwd_adaptive);
    m.def("brownian_levy_query", &brownian_levy_query);
    m.def("brownian_query", &brownian_query);
    m.def("ww_generate", &ww_generate);
    m.def("huber_rate_loss", &huber_rate_loss);
    m.def("window_rate_l1_loss", &window_rate_l1_loss);
    m.def("tc_contract", &tc_contract);
    m.def("tc_contract_tn", &tc_contract_tn);
    m.attr("OP_RHS") = (int)ODECOL_OP_RHS;
    m.attr("OP_RK4_FWD") = (int)ODECOL_OP_RK4_FWD;
    m.attr("OP_RK4_BWD") = (int)ODECOL_OP_RK4_BWD;
    m.attr("OP_DOPRI5_FWD") = (int)ODECOL_OP_DOPRI5_FWD;
    m.attr("OP_EM_FWD") = (int)ODECOL_OP_EM_FWD;
    m.attr("OP_EM_BWD") = (int)ODECOL_OP_EM_BWD;
    m.attr("OP_SRK_FWD") = (int)ODECOL_OP_SRK_FWD;
    m.attr("OP_SRK_BWD") = (int)ODECOL_OP_SRK_BWD;
    m.attr("OP_DOPRI5_BWD") = (int)ODECOL_OP_DOPRI5_BWD;
    m.attr("FLAG_FORCE_STAGED") = (int)ODECOL_FLAG_FORCE_STAGED;
    m.attr("FLAG_FORCE_TENSOR") = (int)ODECOL_FLAG_FORCE_TENSOR;
    m.attr("FLAG_DETERMINISTIC") = (int)ODECOL_FLAG_DETERMINISTIC;
}
