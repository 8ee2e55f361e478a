"""Host side of the boundary on the CPU: drop-in modules (names, shapes, seed-for-seed initial parameters), their
torch forward against the reference's golden RHS, the linear form against the oracle's, knot compression."""
import io
import pickle

import numpy as np
import pytest
import torch

import odecol
from oracle import rhs as orhs
from helpers import PARITY_DICT, XOR_DICT, oracle_form, product_network, stim_table

NAMES = ("wta", "xor", "parity")


def test_seed_for_seed_initial_parameters(cfg, golden):
    torch.manual_seed(0)
    w = odecol.ColumnAreaWTA(cfg, "mt")
    assert np.array_equal(w.recurrent_weights.detach().numpy(), golden["wta"]["recurrent_weights"])
    torch.manual_seed(0)
    x = odecol.ColumnNetworkXOR(cfg, XOR_DICT)
    for a in "01":
        for i in range(2):
            assert np.array_equal(x.feedforward_target_weights[a][i].detach().numpy(), golden["xor"][f"ffw_{a}_{i}"])
    torch.manual_seed(0)
    p = odecol.ColumnNetwork(cfg, PARITY_DICT, torch.device("cpu"))
    g = golden["parity"]
    for k in "012":
        assert np.array_equal(p.areas[k].lateral_weights.detach().numpy(), g[f"lateral_{k}"])
        assert np.array_equal(p.areas[k].lateral_mask.numpy(), g[f"lateral_mask_{k}"])
    for k in "12":
        assert np.array_equal(p.areas[k].feedforward_weights.detach().numpy(), g[f"feedforward_{k}"])
        assert np.array_equal(p.areas[k].feedforward_mask.numpy(), g[f"feedforward_mask_{k}"])
    assert np.array_equal(p.areas["0"].input_weights.detach().numpy(), g["input_weights"])
    assert np.array_equal(p.areas["0"].input_mask.numpy(), g["input_mask"])
    assert np.array_equal(p.output_weights.detach().numpy(), g["output_weights"])


def test_parameter_names_match_the_reference(cfg):
    torch.manual_seed(1)
    assert [n for n, _ in odecol.ColumnAreaWTA(cfg, "mt").named_parameters()] == ["recurrent_weights"]
    x = odecol.ColumnNetworkXOR(cfg, XOR_DICT)
    assert [n for n, _ in x.named_parameters()] == [f"feedforward_target_weights.{a}.{i}" for a in "01" for i in "01"]
    p = odecol.ColumnNetwork(cfg, PARITY_DICT, torch.device("cpu"))
    names = {n: q.requires_grad for n, q in p.named_parameters()}
    assert names == {"output_weights": True, "areas.0.lateral_weights": True, "areas.0.input_weights": True,
                     "areas.1.lateral_weights": True, "areas.1.feedforward_weights": True,
                     "areas.2.lateral_weights": False, "areas.2.feedforward_weights": True}
    for attr in ("lat_in_mask", "output_weights", "noise_type", "sde_type", "set_stim", "set_time_vec"):
        assert hasattr(odecol.ColumnAreaWTA(cfg, "mt"), attr)
    assert x.network_as_area.num_populations == 24 and p.nr_areas == 3 and p.output_scale == 1.0
    assert x.noise_type == "scalar" and p.sde_type == "ito"


@pytest.mark.parametrize("name", NAMES)
def test_forward_and_diffusion_match_reference(name, cfg, golden):
    g = golden[name]
    net = product_network(name, cfg, g)
    net.stim = torch.tensor(g["stim"] if name == "wta" else g["rhs_stim"])
    ys, ts, f = (torch.tensor(g[k]) for k in ("rhs_y", "rhs_t", "rhs_f"))
    out = torch.stack([net.forward(ts[i], ys[i:i + 1])[0] for i in range(len(ts))])
    assert out.shape == f.shape
    assert float((out - f).abs().max()) <= 2e-7 * float(f.abs().max())
    gd = net.diffusion(ts[0], ys[:1])
    assert gd.shape == (1, ys.shape[1], 1) and np.array_equal(gd.numpy(), g["rhs_g"])


@pytest.mark.parametrize("name", NAMES)
def test_linear_form_equals_oracle(name, cfg, golden):
    g = golden[name]
    net = product_network(name, cfg, g)
    lf, lo = net.export_linear_form(), oracle_form(name, cfg, g)
    N, n_in = lo.n, lo.n_in
    Wa = lf.W_aug.detach().numpy()
    assert lf.n_in == n_in and Wa.shape[1] % 4 == 0 and Wa.shape[1] >= N + n_in + 1
    assert np.array_equal(Wa[:, :N], lo.W) and np.array_equal(Wa[:, N:N + n_in], lo.U)
    assert np.array_equal(Wa[:, N + n_in], lo.bias) and not Wa[:, N + n_in + 1:].any()
    assert np.array_equal(lf.kappa.numpy(), lo.kappa) and np.array_equal(lf.sigma.numpy(), lo.sigma)
    assert (lf.tau_s, lf.tau_m, lf.tau_a, lf.resistance) == (lo.tau_s, lo.tau_m, lo.tau_a, lo.resistance)


@pytest.mark.parametrize("name", NAMES)
def test_linear_form_is_differentiable_in_the_reference_parameters(name, cfg, golden):
    net = product_network(name, cfg, golden[name])
    lf = net.export_linear_form()
    lf.W_aug.sum().backward()
    got = {n for n, p in net.named_parameters() if p.grad is not None and p.grad.abs().sum() > 0}
    want = {n for n, p in net.named_parameters() if p.requires_grad and n != "output_weights"}
    assert got == want


def test_batched_forward_equals_per_trial_forward(cfg, golden):
    g = golden["xor"]
    net = product_network("xor", cfg, g)
    stims = torch.tensor(g["stims"])                      # (4,T,2,16)
    y = torch.tensor(g["rhs_y"][:4])
    t = torch.tensor(0.7312)
    net.stim = stims
    batched = net.forward(t, y)
    for b in range(4):
        net.stim = stims[b]
        assert torch.allclose(batched[b], net.forward(t, y[b:b + 1])[0], rtol=1e-6, atol=1e-3)


def test_modules_pickle(cfg, golden):
    for name in NAMES:
        net = product_network(name, cfg, golden[name])
        clone = pickle.loads(pickle.dumps(net))
        for (n1, p1), (n2, p2) in zip(net.named_parameters(), clone.named_parameters()):
            assert n1 == n2 and torch.equal(p1, p2)


def test_compress_knots_is_exact(cfg, golden):
    from oracle.rhs import interp_knots
    for name in NAMES:
        g = golden[name]
        tv = torch.tensor(g["time_vec"])
        table = stim_table(name, g["stims"] if name != "wta" else g["stim"]).float()
        kt, ku = odecol.compress_knots(tv, table)
        assert 2 <= len(kt) <= 8 and ku.shape == (table.shape[0], len(kt), table.shape[2])
        gen = torch.Generator().manual_seed(3)
        probe = torch.cat((tv[::37], tv[len(tv) // 2 - 2:len(tv) // 2 + 2], tv[len(tv) // 3 - 2:len(tv) // 3 + 2],
                           torch.rand(300, generator=gen) * float(tv[-1]) * 1.2 - 0.1 * float(tv[-1])))
        for t in probe:
            assert torch.equal(interp_knots(t, tv, table), interp_knots(t, kt, ku))
    # a smooth stimulus keeps every knot
    tv = torch.linspace(0, 1, 50)
    kt, ku = odecol.compress_knots(tv, torch.sin(tv)[None, :, None])
    assert len(kt) == 50


def test_solvers_refuse_cpu_and_foreign_functions(cfg, golden):
    net = product_network("wta", cfg, golden["wta"])
    net.stim = torch.tensor(golden["wta"]["stim"])
    with pytest.raises(RuntimeError, match="CUDA"):
        odecol.odeint(net, torch.zeros(1, 48), net.time_vec, method="rk4")
    with pytest.raises(TypeError, match="fallback"):
        odecol.odeint(lambda t, y: -y, torch.zeros(1, 48), net.time_vec)
    with pytest.raises(RuntimeError, match="CUDA"):
        odecol.sdeint(net, torch.zeros(1, 48), net.time_vec, method="srk")
    with pytest.raises(RuntimeError, match="CUDA"):               # adaptive srk is fused too (on-chip family): same CPU refusal
        odecol.sdeint(net, torch.zeros(1, 48), net.time_vec, method="srk", adaptive=True)
    with pytest.raises(NotImplementedError):
        odecol.sdeint(net, torch.zeros(1, 48), net.time_vec, method="milstein")


def test_synthetic_sheet_matches_oracle_structure(cfg):
    sheet = odecol.SyntheticColumnSheet(cfg, 4, seed=0)
    lf = sheet.export_linear_form()
    assert lf.N == 32 and lf.n_in == 4
    W = lf.W_aug.detach()[:, :32]
    from oracle.column_model import area_constants
    one = area_constants(cfg, "mt", 1)
    for c in range(4):
        assert np.array_equal(W[8 * c:8 * c + 8, 8 * c:8 * c + 8].numpy(), one.recurrent_weights)
    off = W[0:8, 8:16]
    assert (off[0, 1] < 0) and (off[4, 5] < 0) and int((off != 0).sum()) == 2


def test_networks_pickle_like_the_reference_scripts_do(cfg, golden):
    # the reference pickles whole networks (scripts/wta_ode.py:214-215, parity_ode.py:210-211,281-282): the drop-in
    # modules hold no native handles, a round trip keeps parameters, masks, stimulus and the exported linear form
    import pickle
    for name in ("wta", "xor", "parity"):
        net = product_network(name, cfg, golden[name])
        key = "stim" if name == "wta" else "stims"
        net.stim = torch.tensor(golden[name][key]) if name == "wta" else torch.tensor(golden[name][key][1])
        clone = pickle.loads(pickle.dumps(net))
        assert [n for n, _ in clone.named_parameters()] == [n for n, _ in net.named_parameters()]
        for (_, a), (_, b) in zip(net.named_parameters(), clone.named_parameters()):
            assert torch.equal(a, b) and b.requires_grad == a.requires_grad
        assert torch.equal(clone.time_vec, net.time_vec) and torch.equal(clone.stim, net.stim)
        a, b = net.export_linear_form(), clone.export_linear_form()
        assert torch.equal(a.W_aug, b.W_aug) and torch.equal(a.sigma, b.sigma) and a.n_in == b.n_in
        y = torch.zeros(1, 3 * a.N)
        assert torch.equal(net.forward(net.time_vec[3], y), clone.forward(clone.time_vec[3], y))


def test_compress_knots_property_random_tables():
    # property: for ANY table, looking the stimulus up on the compressed knots equals looking it up on the full table,
    # bit for bit, at any time (inside, on and outside the grid) -- the contract the kernels rely on
    from hypothesis import given, settings, strategies as st
    from oracle.rhs import interp_knots

    @settings(max_examples=60, deadline=None)
    @given(st.integers(3, 40), st.integers(1, 3), st.integers(1, 4), st.integers(0, 2 ** 31 - 1), st.floats(0.0, 1.0))
    def check(T, B, n_in, seed, p_change):
        g = torch.Generator().manual_seed(seed)
        tv = torch.cumsum(torch.rand(T, generator=g) * 0.01 + 1e-4, 0)
        # runs of constant values with occasional jumps / ramps: what step stimuli sampled on a grid look like
        jumps = (torch.rand(T, generator=g) < p_change).float().cumsum(0).long()
        levels = torch.randn(int(jumps.max()) + 1, B, n_in, generator=g) * 20
        table = levels[jumps].permute(1, 0, 2).contiguous()                 # (B, T, n_in)
        kt, ku = odecol.compress_knots(tv, table)
        assert kt[0] == tv[0] and kt[-1] == tv[-1] and len(kt) <= T and ku.shape == (B, len(kt), n_in)
        probes = torch.cat((tv, (tv[1:] + tv[:-1]) / 2, tv[:1] - 1.0, tv[-1:] + 1.0, tv[0] + torch.rand(16, generator=g) * (tv[-1] - tv[0])))
        for t in probes:
            assert torch.equal(interp_knots(t, tv, table), interp_knots(t, kt, ku))

    check()


def test_the_reference_scripts_import_lines_work_against_odecol():
    # scripts/wta_ode.py:9-14, xor_ode.py:2-7, parity_ode.py:10-16 with the package names swapped
    from odecol import sdeint, sdeint_adjoint                      # from torchsde import sdeint, sdeint_adjoint
    from odecol import odeint, odeint_adjoint                      # from torchdiffeq import odeint, odeint_adjoint
    from odecol import ColumnAreaWTA, ColumnNetworkXOR, ColumnNetwork, ColumnArea      # from src.coupled_columns import ...
    from odecol import load_config, compute_firing_rate, soft_clamp, torch_interp, min_max, fr_to_binary, huber_loss_wta  # src.utils
    from odecol import make_ds_wwp, get_data                       # scripts/wta_ode.py:56-107 (uses src.ww_model.DM)
    import inspect
    for fn, names in ((odeint, ("func", "y0", "t", "rtol", "atol", "method", "options", "event_fn")),
                      (odeint_adjoint, ("func", "y0", "t", "adjoint_rtol", "adjoint_atol", "adjoint_method", "adjoint_options", "adjoint_params")),
                      (sdeint, ("sde", "y0", "ts", "bm", "method", "dt", "adaptive", "rtol", "atol", "dt_min", "options", "names",
                                "logqp", "extra", "extra_solver_state")),
                      (sdeint_adjoint, ("sde", "y0", "ts", "bm", "method", "adjoint_method", "dt", "adaptive", "adjoint_adaptive",
                                        "adjoint_params", "names"))):
        params = inspect.signature(fn).parameters
        assert all(n in params for n in names), (fn.__name__, [n for n in names if n not in params])
    assert list(inspect.signature(sdeint).parameters)[:4] == ["sde", "y0", "ts", "bm"]


def test_move_to_moves_plain_tensor_attributes(cfg, golden):
    net = product_network("xor", cfg, golden["xor"])
    net.stim = torch.tensor(golden["xor"]["stims"][0])
    moved = odecol.move_to(net, torch.device("cpu"))
    assert moved is net and moved.ff_source_mask.device.type == "cpu" and moved.stim.device.type == "cpu"
    assert all(torch.is_tensor(v) for v in (moved.time_vec, moved.ff_target_mask))
