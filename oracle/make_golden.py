"""Generate tests/golden/*.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container, where /root/reference exists:

    python -m oracle.make_golden            # writes tests/golden/{wta,xor,parity}.npz

What is pinned by the reference's own code: model construction from config/model.toml, ``forward`` and
``diffusion`` of the three networks, the loss helpers.  What drives them here is the restated solver in
``oracle/solvers.py`` (torchdiffeq / torchsde are not installable offline) -- solver parity is unpinned.

The reference is imported from ``$ODECOL_REFERENCE`` (default /root/reference); nothing is copied.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

REF = os.environ.get("ODECOL_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}")
    sys.path.insert(0, REF)
    from src import coupled_columns as cc          # noqa
    from src import utils as ru                     # noqa
    return cc, ru


def _np(x):
    return x.detach().cpu().numpy().copy()


def _rand_states(gen, n_states, N):
    V = torch.rand(n_states, N, generator=gen) * 44.0 - 30.0      # [-30, 14]
    A = torch.rand(n_states, N, generator=gen) * 2.0
    F = torch.rand(n_states, N, generator=gen) * 20.0
    return torch.cat((V, A, F), dim=1)


def _rhs_samples(net, tv, N, gen, n_states=8):
    ys = _rand_states(gen, n_states, N)
    ts = torch.rand(n_states, generator=gen) * float(tv[-1]) * 1.1 - 0.02 * float(tv[-1])   # also outside the ends
    ts[0] = tv[3]                      # exactly on a knot
    ts[1] = tv[len(tv) // 2]           # on the switching knot (xor/parity)
    out = torch.stack([net.forward(ts[i], ys[i:i + 1])[0] for i in range(n_states)])
    return ys, ts, out


def make_wta(cc, ru, cfg):
    from oracle import solvers, stimuli
    torch.manual_seed(0)
    net = cc.ColumnAreaWTA(cfg, "mt")
    tv = stimuli.time_vec(1500, 1e-4)
    stim = stimuli.wta_stimulus(tv, (20.0, 30.0))
    net.time_vec, net.stim = tv, stim
    y0 = torch.zeros(1, 48)
    gen = torch.Generator().manual_seed(11)
    out = {}
    ys, ts, f = _rhs_samples(net, tv, 16, gen)
    out.update(rhs_y=_np(ys), rhs_t=_np(ts), rhs_f=_np(f), rhs_g=_np(net.diffusion(ts[0], ys[:1])))
    out.update(orig_weights=extract_orig_weights())
    out.update(recurrent_weights=_np(net.recurrent_weights), feedforward_weights=_np(net.feedforward_weights),
               background_weights=_np(net.background_weights), adaptation_strength=_np(net.adaptation_strength),
               lat_in_mask=_np(net.lat_in_mask), time_vec=_np(tv), stim=_np(stim))
    # rk4 forward + gradient of the script's loss (utils.py:74-88) against a fixed synthetic target
    t0 = time.time()
    traj = solvers.odeint_rk4(net, y0, tv)                         # (T,1,48)
    target = torch.linspace(0, 1, 1500).reshape(1, 1500, 1).repeat(1, 1, 2) * torch.tensor([0.6, 0.3])
    loss = ru.huber_loss_wta(traj.unsqueeze(0), target, net)
    loss.backward()
    out.update(rk4_traj=_np(traj), rk4_loss=_np(loss), rk4_target=_np(target),
               rk4_grad_recurrent_weights=_np(net.recurrent_weights.grad))
    net.recurrent_weights.grad = None
    print(f"  wta rk4 fwd+bwd {time.time() - t0:.1f}s loss={float(loss):.6f}")
    with torch.no_grad():
        st = {}
        trajd = solvers.odeint_dopri5(net, y0, tv, stats=st)
        out.update(dopri5_traj=_np(trajd[::10]), dopri5_final=_np(trajd[-1]),
                   dopri5_counts=np.array([st["n_accept"], st["n_reject"]]))
        print("  wta dopri5", st)
        st = {}
        trajl = solvers.odeint_dopri5(net, y0, tv, rtol=1e-3, atol=1e-4, stats=st)
        out.update(dopri5_loose_final=_np(trajl[-1]), dopri5_loose_counts=np.array([st["n_accept"], st["n_reject"]]))
        print("  wta dopri5 loose", st)
        sched = solvers.em_step_schedule(tv, 1e-3)
        g = torch.Generator().manual_seed(1234)
        dW = torch.randn(len(sched), 1, 1, generator=g) * float(np.sqrt(1e-3))
        trajs = solvers.sdeint_euler(net, y0, tv, solvers.TabulatedBrownian(dW), dt=1e-3)
        out.update(em_dW=_np(dW), em_traj=_np(trajs), em_nsteps=np.array(len(sched)))
    return out


def make_xor(cc, ru, cfg):
    from oracle import solvers, stimuli
    torch.manual_seed(0)
    nd = {"nr_areas": 2, "areas": ["mt", "mt"], "nr_columns_per_area": [2, 1], "nr_input_units": 2}
    net = cc.ColumnNetworkXOR(cfg, nd)
    tv = stimuli.time_vec(1000, 1e-3)
    net.time_vec = tv
    y0 = torch.zeros(1, 72)
    gen = torch.Generator().manual_seed(12)
    out = {}
    net.stim = stimuli.xor_stimulus(tv, (20.0, 7.0))
    ys, ts, f = _rhs_samples(net, tv, 24, gen)
    out.update(rhs_y=_np(ys), rhs_t=_np(ts), rhs_f=_np(f), rhs_stim=_np(net.stim),
               rhs_g=_np(net.diffusion(ts[0], ys[:1])))
    for a in ("0", "1"):
        for i in range(2):
            out[f"ffw_{a}_{i}"] = _np(net.feedforward_target_weights[a][i])
    for a in ("0", "1"):
        out[f"area{a}_recurrent"] = _np(net.areas[a].recurrent_weights)
        out[f"area{a}_background"] = _np(net.areas[a].background_weights)
    out.update(time_vec=_np(tv), kappa=_np(net.network_as_area.adaptation_strength))
    stims = [stimuli.xor_stimulus(tv, c) for c in stimuli.XOR_CONDITIONS]
    out["stims"] = _np(torch.stack(stims))                           # (4,T,2,16)
    targets = torch.tensor([1.0, 1.0, 0.25, 0.25])

    def run(method, **kw):
        trajs, stats = [], []
        for s in stims:
            net.stim = s
            st = {}
            trajs.append(solvers.odeint(net, y0, tv, method=method, stats=st, **kw))
            stats.append([st.get("n_accept", 0), st.get("n_reject", 0)])
        batch = torch.stack(trajs)                                   # (4,T,1,72)
        fr = ru.compute_firing_rate(batch[:, :, :, :24] - batch[:, :, :, 24:48]).squeeze(2)
        final_c = torch.sum(fr[:, -1, 16:] * net.ff_source_mask, dim=1)     # xor_ode.py:120-130
        loss = torch.mean(abs(final_c - targets))
        return batch, loss, np.array(stats)

    t0 = time.time()
    batch, loss, _ = run("rk4")
    loss.backward()
    out.update(rk4_traj=_np(batch[:, ::5, 0, :]), rk4_final=_np(batch[:, -1, 0, :]), rk4_loss=_np(loss))
    for a in ("0", "1"):
        for i in range(2):
            p = net.feedforward_target_weights[a][i]
            out[f"rk4_grad_ffw_{a}_{i}"] = _np(p.grad)
            p.grad = None
    print(f"  xor rk4 fwd+bwd {time.time() - t0:.1f}s loss={float(loss):.6f}")
    t0 = time.time()
    batch, loss, counts = run("dopri5")
    loss.backward()
    out.update(dopri5_traj=_np(batch[:, ::10, 0, :]), dopri5_final=_np(batch[:, -1, 0, :]), dopri5_loss=_np(loss),
               dopri5_counts=counts)
    for a in ("0", "1"):
        for i in range(2):
            p = net.feedforward_target_weights[a][i]
            out[f"dopri5_grad_ffw_{a}_{i}"] = _np(p.grad)
            p.grad = None
    print(f"  xor dopri5 fwd+bwd {time.time() - t0:.1f}s loss={float(loss):.6f} counts={counts.tolist()}")
    with torch.no_grad():
        batch, loss, counts = run("dopri5", rtol=1e-5, atol=1e-6)
        out.update(dopri5_loose_final=_np(batch[:, -1, 0, :]), dopri5_loose_counts=counts)
        print(f"  xor dopri5 loose counts={counts.tolist()}")
        sched = solvers.em_step_schedule(tv, 1e-3)
        g = torch.Generator().manual_seed(1234)
        dW = torch.randn(len(sched), 4, 1, generator=g) * float(np.sqrt(1e-3))
        trajs = []
        for b, s in enumerate(stims):
            net.stim = s
            trajs.append(solvers.sdeint_euler(net, y0, tv, solvers.TabulatedBrownian(dW[:, b:b + 1]), dt=1e-3))
        out.update(em_dW=_np(dW), em_traj=_np(torch.stack(trajs)[:, ::5, 0, :]), em_nsteps=np.array(len(sched)))
    return out


def make_parity(cc, ru, cfg):
    from oracle import solvers, stimuli
    torch.manual_seed(0)
    nd = {"nr_areas": 3, "areas": ["mt"] * 3, "nr_columns_per_area": [8, 4, 1], "nr_input_units": 4}
    net = cc.ColumnNetwork(cfg, nd, torch.device("cpu"))
    tv = stimuli.time_vec(1000, 1e-3)
    net.time_vec = tv
    N = 104
    y0 = torch.zeros(1, 3 * N)
    gen = torch.Generator().manual_seed(13)
    out = {}
    net.stim = stimuli.parity_stimulus(tv, (0.3, 1.0, 0.0, 0.7))
    ys, ts, f = _rhs_samples(net, tv, N, gen)
    out.update(rhs_y=_np(ys), rhs_t=_np(ts), rhs_f=_np(f), rhs_stim=_np(net.stim),
               rhs_g=_np(net.diffusion(ts[0], ys[:1])))
    for k in ("0", "1", "2"):
        a = net.areas[k]
        out[f"lateral_{k}"] = _np(a.lateral_weights)
        out[f"inner_{k}"] = _np(a.inner_weights)
        out[f"background_{k}"] = _np(a.background_weights)
        out[f"lateral_mask_{k}"] = _np(a.lateral_mask)
        if k != "0":
            out[f"feedforward_{k}"] = _np(a.feedforward_weights)
            out[f"feedforward_mask_{k}"] = _np(a.feedforward_mask)
    out.update(input_weights=_np(net.areas["0"].input_weights), input_mask=_np(net.areas["0"].input_mask),
               output_weights=_np(net.output_weights), time_vec=_np(tv),
               kappa=_np(net.network_as_area.adaptation_strength))
    stims = [stimuli.parity_stimulus(tv, p) for p in stimuli.PARITY_PATTERNS]
    out["stims"] = _np(torch.stack(stims))                           # (4,T,4)
    raw = torch.tensor(stimuli.PARITY_PATTERNS) * 15.0
    targets = (raw.sum(dim=1) % 30 == 0).float() * 20.0              # parity_ode.py:245-246

    t0 = time.time()
    trajs = []
    for s in stims:
        net.stim = s
        trajs.append(solvers.odeint_rk4(net, y0, tv))
    batch = torch.stack(trajs)                                       # (4,T,1,312)
    fr = ru.compute_firing_rate(batch[:, :, :, :N] - batch[:, :, :, N:2 * N])
    final = torch.mean(fr[:, -100:, 0, -8:], dim=1)                  # parity_ode.py:239-249
    summed = torch.sum(final * net.output_weights / net.output_scale, dim=-1)
    loss = torch.mean(abs(summed - targets))
    loss.backward()
    out.update(rk4_traj=_np(batch[:, ::10, 0, :]), rk4_final=_np(batch[:, -1, 0, :]), rk4_loss=_np(loss),
               rk4_targets=_np(targets))
    for name, p in net.named_parameters():
        if p.grad is not None:
            out["rk4_grad_" + name.replace(".", "_")] = _np(p.grad)
            p.grad = None
    print(f"  parity rk4 fwd+bwd {time.time() - t0:.1f}s loss={float(loss):.6f}")
    with torch.no_grad():
        sched = solvers.em_step_schedule(tv, 1e-3)
        g = torch.Generator().manual_seed(1234)
        dW = torch.randn(len(sched), 4, 1, generator=g) * float(np.sqrt(1e-3))
        trajs = []
        for b, s in enumerate(stims):
            net.stim = s
            trajs.append(solvers.sdeint_euler(net, y0, tv, solvers.TabulatedBrownian(dW[:, b:b + 1]), dt=1e-3))
        out.update(em_dW=_np(dW), em_traj=_np(torch.stack(trajs)[:, ::10, 0, :]), em_nsteps=np.array(len(sched)))
        st = {}
        net.stim = stims[1]
        trajd = solvers.odeint_dopri5(net, y0, tv, rtol=1e-5, atol=1e-6, stats=st)
        out.update(dopri5_loose_final=_np(trajd[-1]), dopri5_loose_counts=np.array([st["n_accept"], st["n_reject"]]))
        print("  parity dopri5 loose", st)
    return out


def extract_orig_weights() -> np.ndarray:
    """The one numeric fixture the reference holds for this path: the hard-coded 16x16 ``orig_weights`` literal of
    scripts/plotting_results.py:36-99 (units are 1000x today's; six entries are stale).  Parsed
    from the source text (the script itself cannot be imported: matplotlib is missing)."""
    import ast
    import re
    src = open(os.path.join(REF, "scripts", "plotting_results.py")).read()
    m = re.search(r"orig_weights\s*=\s*torch\.tensor\((\[\[.*?\]\])\)", src, flags=re.S)
    mat = np.asarray(ast.literal_eval(m.group(1)), dtype=np.float64)
    assert mat.shape == (16, 16)
    return mat


def add_fp64_gradients():
    """float64 'truth' for the three gradient fixtures: the oracle's unified form (validated against the reference's
    forward to 1e-7) integrated and differentiated in float64 with the scripts' losses.  Two float32 evaluations of
    these gradients (the reference's autograd and any other summation order) differ from each other by up to 5e-4
    relative on the ill-conditioned final-time XOR loss, so the tests measure the CUDA result against this truth and
    against the reference's own float32 error."""
    import tomllib
    sys.path.insert(0, os.path.join(os.path.dirname(OUT)))
    from oracle import column_model as cm, rhs as orhs, solvers
    cfg = tomllib.load(open(os.path.join(os.path.dirname(os.path.dirname(OUT)), "config", "model.toml"), "rb"))
    f64 = torch.float64

    def solve(lf, tv, table, N):
        ode = orhs.UnifiedColumnODE(lf, tv, table, dtype=f64, requires_grad=True)
        y = solvers.odeint_rk4(ode, torch.zeros(table.shape[0], 3 * N, dtype=f64), torch.tensor(tv, dtype=f64))
        return ode, y

    # wta
    path = os.path.join(OUT, "wta.npz")
    g = dict(np.load(path))
    lf = cm.wta_linear_form(cfg, g["recurrent_weights"])
    ode, y = solve(lf, g["time_vec"], g["stim"][None], 16)
    r = orhs.phi(y[:, 0, :16] - y[:, 0, 16:32])
    both = torch.stack((r[:, 0], r[:, 8]), dim=1)[None]
    loss = torch.nn.functional.smooth_l1_loss(both, torch.tensor(g["rk4_target"], dtype=f64), beta=1.0)
    loss.backward()
    g["rk4_grad64_recurrent_weights"] = _np(ode.W.grad)
    np.savez_compressed(path, **g)
    print("wta fp64 loss", float(loss), "vs fp32", float(g["rk4_loss"]))
    # xor
    path = os.path.join(OUT, "xor.npz")
    g = dict(np.load(path))
    lf = cm.xor_linear_form(cfg, [[g["ffw_0_0"], g["ffw_0_1"]], [g["ffw_1_0"], g["ffw_1_1"]]])
    ode, y = solve(lf, g["time_vec"], g["stims"].reshape(4, 1000, 32), 24)
    fc = orhs.phi(y[-1, :, 16:24] - y[-1, :, 40:48])[:, 0]
    loss = torch.mean(abs(fc - torch.tensor([1.0, 1.0, 0.25, 0.25], dtype=f64)))
    loss.backward()
    for i in range(2):
        g[f"rk4_grad64_ffw_0_{i}"] = _np(torch.diagonal(ode.U.grad[:16, 16 * i:16 * (i + 1)]))
        g[f"rk4_grad64_ffw_1_{i}"] = _np(10.0 * ode.W.grad[16:24, 8 * i])
    np.savez_compressed(path, **g)
    print("xor fp64 loss", float(loss), "vs fp32", float(g["rk4_loss"]))
    # parity
    path = os.path.join(OUT, "parity.npz")
    g = dict(np.load(path))
    lf = cm.parity_linear_form(cfg, [g[f"lateral_{k}"] for k in range(3)], {1: g["feedforward_1"], 2: g["feedforward_2"]},
                               g["input_weights"])
    ode, y = solve(lf, g["time_vec"], g["stims"], 104)
    N = 104
    r = orhs.phi(y[-100:, :, N - 8:N] - y[-100:, :, 2 * N - 8:2 * N])
    ow = torch.tensor(g["output_weights"], dtype=f64, requires_grad=True)
    summed = (r.mean(0) * ow).sum(-1)
    loss = torch.mean(abs(summed - torch.tensor(g["rk4_targets"], dtype=f64)))
    loss.backward()
    gW, gU = ode.W.grad, ode.U.grad
    g["rk4_grad64_output_weights"] = _np(ow.grad)
    g["rk4_grad64_areas_0_lateral_weights"] = _np(gW[0:64, 0:64])
    g["rk4_grad64_areas_1_lateral_weights"] = _np(gW[64:96, 64:96])
    g["rk4_grad64_areas_1_feedforward_weights"] = _np(gW[64:96, 0:64])
    g["rk4_grad64_areas_2_feedforward_weights"] = _np(gW[96:104, 64:96])
    g["rk4_grad64_areas_0_input_weights"] = _np(gU[0:64, :])
    np.savez_compressed(path, **g)
    print("parity fp64 loss", float(loss), "vs fp32", float(g["rk4_loss"]))


def add_srk():
    """method='srk' fixtures (what scripts/wta_ode.py:174,200 call): the UNMODIFIED reference modules driven by the
    restated SRI2 stepper with tabulated (W, U), seed 4321; for the WTA network also the script's Huber loss and its
    autograd gradient (wta_ode.py:176-181).  Adds keys srk_* to the existing wta / xor files."""
    from oracle import solvers, stimuli
    cc, ru = _import_reference()
    cfg = ru.load_config(os.path.join(REF, "config", "model.toml"))
    torch.set_num_threads(1)
    # wta: one trial, the full 1500-point grid, sigma = 100 on all 48 components
    path = os.path.join(OUT, "wta.npz")
    g = dict(np.load(path))
    torch.manual_seed(0)
    net = cc.ColumnAreaWTA(cfg, "mt")
    assert np.array_equal(_np(net.recurrent_weights), g["recurrent_weights"])
    tv = torch.tensor(g["time_vec"])
    net.time_vec, net.stim = tv, torch.tensor(g["stim"])
    n_steps = len(solvers.em_step_schedule(tv, 1e-3))
    gen = torch.Generator().manual_seed(4321)
    W, U = solvers.sample_w_u(n_steps, 1, 1e-3, gen)
    W, U = 0.1 * W, 0.1 * U                       # sigma = 100: keep V - A away from the pole of phi
    traj = solvers.sdeint_srk(net, torch.zeros(1, 48), tv, solvers.TabulatedBrownianU(W, U), dt=1e-3)
    target = torch.tensor(g["rk4_target"])
    loss = ru.huber_loss_wta(traj.unsqueeze(0), target, net)
    loss.backward()
    g.update(srk_dW=_np(W), srk_dU=_np(U), srk_traj=_np(traj), srk_loss=_np(loss),
             srk_grad_recurrent_weights=_np(net.recurrent_weights.grad))
    np.savez_compressed(path, **g)
    print(f"wta srk: {n_steps} steps, loss {float(loss):.6f}, |V|max {float(traj[..., :16].abs().max()):.2f}")
    # xor: the four patterns, sigma = 10 on V
    path = os.path.join(OUT, "xor.npz")
    g = dict(np.load(path))
    torch.manual_seed(0)
    nd = {"nr_areas": 2, "areas": ["mt", "mt"], "nr_columns_per_area": [2, 1], "nr_input_units": 2}
    net = cc.ColumnNetworkXOR(cfg, nd)
    assert np.array_equal(_np(net.feedforward_target_weights["0"][0]), g["ffw_0_0"])
    tv = torch.tensor(g["time_vec"])
    net.time_vec = tv
    n_steps = len(solvers.em_step_schedule(tv, 1e-3))
    W, U = solvers.sample_w_u(n_steps, 4, 1e-3, gen)
    W, U = 0.3 * W, 0.3 * U
    trajs = []
    with torch.no_grad():
        for b in range(4):
            net.stim = torch.tensor(g["stims"][b])
            trajs.append(solvers.sdeint_srk(net, torch.zeros(1, 72), tv, solvers.TabulatedBrownianU(W[:, b:b + 1], U[:, b:b + 1]), dt=1e-3))
    tr = torch.stack(trajs)[:, ::5, 0, :]
    g.update(srk_dW=_np(W), srk_dU=_np(U), srk_traj=_np(tr))
    np.savez_compressed(path, **g)
    print(f"xor srk: {n_steps} steps, |V|max {float(tr[..., :24].abs().max()):.2f}")


def add_ww():
    """tests/golden/ww.npz: the UNMODIFIED reference Wong-Wang class (src/ww_model.py) inside the reference's dataset
    loop (scripts/wta_ode.py:56-93, restated here only as far as the loop body: the script itself imports matplotlib),
    numpy seed 7, 5 samples -- stimuli AND states, so that both the generator's arithmetic and its consumption of the
    global random stream are pinned."""
    sys.path.insert(0, REF)
    from src.ww_model import DM
    nr_samples, time_steps = 5, 1500
    np.random.seed(7)
    states = torch.Tensor(nr_samples, time_steps, 2)
    stims = torch.Tensor(nr_samples, 2)
    dm = DM()
    full_first = None
    for i in range(nr_samples):
        muA = np.random.uniform(15.0, 25.0)
        muB = muA + np.random.uniform(10., 20.)
        mu_vals = [muA, muB]
        np.random.shuffle(mu_vals)
        muA, muB = mu_vals
        R = dm.run_sim(muA, muB)
        if full_first is None:
            full_first = R.copy()
        R = R[:, ::10]
        R = R[:, :time_steps]
        states[i, :, :] = torch.tensor(R).transpose(0, 1)
        stims[i, :] = torch.tensor([muA, muB])
    path = os.path.join(OUT, "ww.npz")
    np.savez_compressed(path, states=_np(states), stims=_np(stims), first_full=full_first[:, ::25], seed=np.array(7))
    print(f"wrote {path} ({os.path.getsize(path) / 1e3:.0f} kB); r range {float(states.min()):.3f}..{float(states.max()):.3f}")


def main():
    if "--ww-only" in sys.argv:
        add_ww()
        return
    if "--fp64-only" in sys.argv:
        add_fp64_gradients()
        return
    if "--srk-only" in sys.argv:
        add_srk()
        return
    if "--orig-weights-only" in sys.argv:      # cheap refresh of one key without re-running the solvers
        path = os.path.join(OUT, "wta.npz")
        data = dict(np.load(path))
        data["orig_weights"] = extract_orig_weights()
        np.savez_compressed(path, **data)
        print("updated", path)
        return
    cc, ru = _import_reference()
    cfg = ru.load_config(os.path.join(REF, "config", "model.toml"))
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    for name, fn in (("wta", make_wta), ("xor", make_xor), ("parity", make_parity)):
        print(name)
        data = fn(cc, ru, cfg)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **{k: np.asarray(v) for k, v in data.items()})
        print(f"  wrote {path} ({os.path.getsize(path) / 1e3:.0f} kB)")
    add_fp64_gradients()
    add_srk()
    add_ww()


if __name__ == "__main__":
    main()
