"""The oracle's unified linear form against the golden vectors the UNMODIFIED reference produced
(oracle/make_golden.py): weight construction, forward(t, y), diffusion, and a stretch of the rk4 / Euler-Maruyama
trajectories driven by the reference's own nn.Modules."""
import numpy as np
import pytest
import torch

from oracle import column_model as cm, rhs as orhs, solvers as S
from helpers import oracle_form, stim_table, rel_err

NAMES = ("wta", "xor", "parity")


def _rhs_stim(name, g):
    return g["stim"] if name == "wta" else g["rhs_stim"]


@pytest.mark.parametrize("name", NAMES)
def test_rhs_matches_reference_forward(name, cfg, golden):
    g = golden[name]
    lf = oracle_form(name, cfg, g)
    ode = orhs.UnifiedColumnODE(lf, g["time_vec"], stim_table(name, _rhs_stim(name, g)))
    ys, ts, f = (torch.tensor(g[k]) for k in ("rhs_y", "rhs_t", "rhs_f"))
    out = torch.stack([ode(ts[i], ys[i:i + 1])[0] for i in range(len(ts))])
    # fp32 reassociation of ~1e5-sized currents: a few 1e-3 absolute, 1e-7 relative
    assert float((out - f).abs().max()) <= 2e-7 * float(f.abs().max())


@pytest.mark.parametrize("name", NAMES)
def test_diffusion_quirks(name, cfg, golden):
    g = golden[name]
    lf = oracle_form(name, cfg, g)
    assert np.array_equal(lf.sigma, g["rhs_g"].reshape(-1))          # wta: 100 everywhere; xor: 10 on V; parity: 10 everywhere


def test_construction_against_reference_constants(cfg, golden):
    gx, gp = golden["xor"], golden["parity"]
    a0 = cm.area_constants(cfg, "mt", 2, small_network=True)
    a1 = cm.area_constants(cfg, "mt", 1, small_network=True)
    assert np.array_equal(a0.recurrent_weights * a0.internal_mask, gx["area0_recurrent"])
    assert np.array_equal(a1.recurrent_weights, gx["area1_recurrent"])
    assert np.array_equal(a0.background_weights, gx["area0_background"])
    for k, cols in zip("012", (8, 4, 1)):
        a = cm.area_constants(cfg, "mt", cols)
        assert np.array_equal(a.recurrent_weights * a.internal_mask, gp[f"inner_{k}"])
        assert np.array_equal(a.background_weights, gp[f"background_{k}"])
    assert np.array_equal(cm.area_constants(cfg, "mt", 13).adaptation_strength, gp["kappa"])


def test_orig_weights_known_answer(cfg, golden):
    """The reference's only numeric fixture for this path: scripts/plotting_results.py:36-99 hard-codes the 16x16
    WTA matrix (units /1000).  It pins today's construction on 106 of its 112 non-zeros (5 printed digits); the six stale entries are a
    trained self-excitation (0,0),(8,8), trained lateral weights (1,8),(9,0) and a probability that has since been
    edited in config/model.toml:6 (0,2),(8,10) -- SURVEY.md section 4."""
    ref = golden["wta"]["orig_weights"]
    a = cm.area_constants(cfg, "mt", 2, small_network=True)
    W = a.recurrent_weights.astype(np.float64) / 1000.0
    stale = {(0, 0), (8, 8), (1, 8), (9, 0), (0, 2), (8, 10)}
    checked = 0
    for i in range(16):
        for j in range(16):
            if (i, j) in stale:
                continue
            if ref[i, j] != 0:
                assert abs(W[i, j] - ref[i, j]) <= 6e-5 * abs(ref[i, j]) + 1e-9, (i, j, W[i, j], ref[i, j])
                checked += 1
            else:
                assert W[i, j] == 0
    assert checked == 106


@pytest.mark.parametrize("name,steps", [("wta", 400), ("xor", 120), ("parity", 60)])
def test_rk4_prefix_matches_reference_driven_solve(name, steps, cfg, golden):
    g = golden[name]
    lf = oracle_form(name, cfg, g)
    tv = torch.tensor(g["time_vec"])
    if name == "wta":
        stim, ref = g["stim"], torch.tensor(g["rk4_traj"][:steps + 1, 0])
        every = 1
    else:
        stim, every = g["stims"][1], (5 if name == "xor" else 10)
        ref = torch.tensor(g["rk4_traj"][1, :steps // every + 1])
    ode = orhs.UnifiedColumnODE(lf, tv, stim_table(name, stim))
    y = S.odeint_rk4(ode, torch.zeros(1, 3 * lf.n), tv[:steps + 1])[::every, 0]
    assert rel_err(y, ref) < 2e-6


def test_parity_rk4_noise_floor(cfg, golden):
    """How far two float32 CPU implementations of the same arithmetic drift apart over the WHOLE parity solve (1000 steps):
    the oracle's unified linear form against the golden trajectory of the unmodified reference modules.  The GPU tests
    quote this floor where they judge the tensor family on this network (tests/test_gpu_parity.py)."""
    g = golden["parity"]
    lf = oracle_form("parity", cfg, g)
    tv = torch.tensor(g["time_vec"])
    worst = 0.0
    for b in (1, 3):                                    # the two patterns with the largest drift
        ode = orhs.UnifiedColumnODE(lf, tv, stim_table("parity", g["stims"][b]))
        y = S.odeint_rk4(ode, torch.zeros(1, 3 * lf.n), tv)[::10, 0]
        worst = max(worst, rel_err(y, torch.tensor(g["rk4_traj"][b])))
    print(f"\nparity rk4, float32 CPU oracle vs reference golden over 1000 steps: {worst:.2e}")
    assert 2e-6 < worst < 8e-6


def test_em_prefix_matches_reference_driven_solve(cfg, golden):
    g = golden["wta"]
    lf = oracle_form("wta", cfg, g)
    tv = torch.tensor(g["time_vec"])
    ode = orhs.UnifiedColumnODE(lf, tv, stim_table("wta", g["stim"]))
    dW = torch.tensor(g["em_dW"])
    y = S.sdeint_euler(ode, torch.zeros(1, 48), tv, S.TabulatedBrownian(dW), dt=1e-3)
    assert rel_err(y[:, 0], torch.tensor(g["em_traj"][:, 0])) < 5e-6


def test_srk_matches_reference_driven_solve(cfg, golden):
    # unified-form oracle under the restated SRI2 stepper vs the UNMODIFIED reference module under the same stepper
    g = golden["wta"]
    lf = oracle_form("wta", cfg, g)
    tv = torch.tensor(g["time_vec"])
    ode = orhs.UnifiedColumnODE(lf, tv, stim_table("wta", g["stim"]))
    bm = S.TabulatedBrownianU(torch.tensor(g["srk_dW"]), torch.tensor(g["srk_dU"]))
    y = S.sdeint_srk(ode, torch.zeros(1, 48), tv, bm, dt=1e-3)
    assert rel_err(y[:, 0], torch.tensor(g["srk_traj"][:, 0])) < 5e-6
