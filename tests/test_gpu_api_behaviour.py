"""Behaviour of the drop-in entry points that torchdiffeq / torchsde users rely on: solver failures
raise, default Brownian paths are fresh per call and reproducible under torch.manual_seed, and the dopri5 record is
sized from the accepted steps."""
import numpy as np
import pytest
import torch

import odecol
from helpers import product_network

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _xor(cfg, golden, B=4):
    net = product_network("xor", cfg, golden["xor"], DEV)
    net.stim = torch.tensor(golden["xor"]["stims"]).to(DEV)[:B]
    return net, torch.zeros(B, 72, device=DEV)


def test_failed_adaptive_solve_raises_like_torchdiffeq(cfg, golden):
    net, y0 = _xor(cfg, golden)
    ts = net.time_vec[:200]
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="step budget"):
            odecol.odeint(net, y0, ts, options={"max_num_steps": 5})
        st = {}
        y = odecol.odeint(net, y0, ts, options={"max_num_steps": 5, "check_status": False}, stats=st)
    assert int(st["status"].min()) == 2 and bool(torch.isnan(y[-1]).all())
    y0g = y0.clone().requires_grad_(True)
    with pytest.raises(RuntimeError, match="step budget"):
        odecol.odeint(net, y0g, ts, options={"max_num_steps": 5})


def test_default_brownian_path_is_fresh_per_call_and_reproducible(cfg, golden):
    """torchsde draws a new BrownianInterval per sdeint call (reference scripts/wta_ode.py:174,200 pass neither bm nor a
    seed); torch.manual_seed must still reproduce the run."""
    net = product_network("wta", cfg, golden["wta"], DEV)
    net.stim = torch.tensor(golden["wta"]["stim"]).to(DEV)
    y0 = torch.zeros(1, 48, device=DEV)
    ts = net.time_vec[:300]
    kw = dict(method="euler", dt=1e-3, options={"sigma_scale": [0.05]})
    torch.manual_seed(123)
    with torch.no_grad():
        a1 = odecol.sdeint(net, y0, ts, **kw)
        a2 = odecol.sdeint(net, y0, ts, **kw)
    torch.manual_seed(123)
    with torch.no_grad():
        b1 = odecol.sdeint(net, y0, ts, **kw)
        b2 = odecol.sdeint(net, y0, ts, **kw)
    assert not torch.equal(a1, a2)                       # successive calls: independent paths
    assert torch.equal(a1, b1) and torch.equal(a2, b2)   # same global seed: same sequence of paths


def test_dopri5_record_is_sized_from_the_accepted_steps(cfg, golden, monkeypatch):
    """Large batches: the record holds max(n_accept) + 1 steps (a forward-only pre-pass counts them), not 4096."""
    from ode_column_b200 import dopri5_adjoint as da
    net, y0 = _xor(cfg, golden)
    ts = net.time_vec[:120]
    caps = []
    orig = da._Dopri5Function.forward

    def spy(ctx, y0_, W, setup, rtol, atol, max_steps, cap, *rest):
        caps.append(cap)
        return orig(ctx, y0_, W, setup, rtol, atol, max_steps, cap, *rest)

    monkeypatch.setattr(da._Dopri5Function, "forward", staticmethod(spy))
    y0g = y0.clone().requires_grad_(True)
    st = {}
    y = odecol.odeint(net, y0g, ts, rtol=1e-5, atol=1e-6, stats=st)
    y[-1, :, :24].sum().backward()
    g_small = y0g.grad.clone()
    assert caps[-1] == 4096
    monkeypatch.setattr(da, "_SMALL_RECORD_BYTES", 0)    # force the pre-pass
    y0h = y0.clone().requires_grad_(True)
    y2 = odecol.odeint(net, y0h, ts, rtol=1e-5, atol=1e-6)
    y2[-1, :, :24].sum().backward()
    assert caps[-1] < 4096 and caps[-1] > int(st["n_accept"].max())
    assert torch.equal(y, y2) and torch.equal(g_small, y0h.grad)


@pytest.mark.parametrize("task", ["xor", "parity"])
def test_fused_window_readout_matches_the_script_expressions(task, cfg, golden):
    """odecol.window_rate_l1_loss against the reference scripts' read-out + loss written in torch (scripts/xor_ode.py:120-130,
    scripts/parity_ode.py:239-249) and autograd's gradient w.r.t. the trajectory."""
    gen = torch.Generator().manual_seed(3)
    if task == "xor":
        net = product_network("xor", cfg, golden["xor"], DEV)
        T, B, N, last = 40, 8, 24, 1
        pops = torch.arange(16, 24)
        w = net.ff_source_mask
        target = torch.where(torch.rand(B, generator=gen) > 0.5, 1.0, 0.25)
    else:
        net = product_network("parity", cfg, golden["parity"], DEV)
        T, B, N, last = 130, 6, 104, 100
        pops = torch.arange(N - 8, N)
        w = net.output_weights / net.output_scale
        target = torch.where(torch.rand(B, generator=gen) > 0.5, 20.0, 0.0)
    traj = torch.cat((torch.rand(T, B, N, generator=gen) * 8 - 6, torch.rand(T, B, N, generator=gen), torch.rand(T, B, N, generator=gen)), 2).to(DEV)
    sel = odecol.readout_components(pops, N).to(DEV)
    # reference expression on the full trajectory
    tr = traj.clone().requires_grad_(True)
    wr = w.detach().clone().to(DEV).requires_grad_(True)
    rates = odecol.compute_firing_rate(tr[:, :, :N] - tr[:, :, N:2 * N])
    if task == "xor":
        pred_ref = torch.sum(rates[-1, :, 16:] * wr, dim=1)
    else:
        pred_ref = torch.sum(rates[-100:, :, -8:].mean(dim=0) * wr, dim=-1)
    loss_ref = torch.mean(abs(pred_ref - target.to(DEV)))
    loss_ref.backward()
    # fused
    ys = traj[:, :, sel].clone().requires_grad_(True)
    wf = w.detach().clone().to(DEV).requires_grad_(True)
    loss, pred = odecol.window_rate_l1_loss(ys, target, weights=wf, last=last)
    (2.0 * loss).backward()
    torch.cuda.synchronize()
    g_ref = tr.grad[:, :, sel]
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    print(f"\n[{task} read-out] loss {float(loss):.6f} vs {float(loss_ref):.6f}; pred {rel(pred, pred_ref.detach()):.1e}; "
          f"grad y {rel(ys.grad, 2 * g_ref):.1e}; grad w {rel(wf.grad, 2 * wr.grad):.1e}")
    assert abs(float(loss) - float(loss_ref)) < 2e-6 * max(1.0, abs(float(loss_ref)))
    assert rel(pred, pred_ref.detach()) < 2e-6 and rel(ys.grad, 2 * g_ref) < 2e-5 and rel(wf.grad, 2 * wr.grad) < 2e-5
    assert float(ys.grad[:T - last].abs().max()) == 0.0
