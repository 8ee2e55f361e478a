"""Diagnostic: on-chip vs staged dopri5 gradients on the parity network against oracle autograd (short horizons)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, odecol
from oracle import rhs as orhs, solvers as S
from helpers import product_network, oracle_form
cfg = odecol.load_config(os.path.join(ROOT, "config", "model.toml"))
golden = {k: np.load(os.path.join(ROOT, "tests", "golden", k + ".npz")) for k in ("xor", "parity")}
DEV = "cuda"
for name, N in (("parity", 104), ("xor", 24)):
    g = golden[name]
    net = product_network(name, cfg, g, DEV)
    stims = torch.tensor(g["stims"])
    net.stim = stims.to(DEV)
    lf = oracle_form(name, cfg, g)
    for npts, stride in ((9, 5), (17, 10), (41, 25)):
        tv = net.time_vec[::stride][:npts].contiguous()
        B = 4
        y0 = torch.zeros(B, 3 * N)
        wN = torch.linspace(0.5, 1.5, N)
        res = {}
        for fam in (None, "staged"):
            net.zero_grad()
            y0g = y0.to(DEV).requires_grad_(True)
            st = {}
            y = odecol.odeint(net, y0g, tv, rtol=1e-5, atol=1e-6, options={"family": fam} if fam else None, stats=st)
            (y[-1, :, :N] * wN.to(DEV)).sum().backward()
            res[fam] = (y0g.grad.cpu().clone(), st["n_accept"].tolist(), y[-1].detach().cpu())
        # oracle, trial by trial
        g0 = []
        table = net.stimulus_channels().cpu()
        nacc = []
        for b in range(B):
            ode = orhs.UnifiedColumnODE(lf, net.time_vec.cpu().numpy(), table[b:b + 1].numpy(), requires_grad=True)
            y0o = y0[b:b + 1].clone().requires_grad_(True)
            so = {}
            yo = S.odeint_dopri5(ode, y0o, tv.cpu(), rtol=1e-5, atol=1e-6, stats=so)
            (yo[-1, :, :N] * wN).sum().backward()
            g0.append(y0o.grad); nacc.append(so["n_accept"])
        g0 = torch.cat(g0)
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        print(f"{name} T={npts} horizon {float(tv[-1]):.3f}: accepted on-chip {res[None][1]} staged {res['staged'][1]} oracle {nacc}; "
              f"dy0 on-chip vs oracle {rel(res[None][0], g0):.2e}, staged vs oracle {rel(res['staged'][0], g0):.2e}, per trial on-chip "
              f"{[f'{rel(res[None][0][b], g0[b]):.1e}' for b in range(B)]}", flush=True)
