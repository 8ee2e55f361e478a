"""Diagnostic: the on-chip and the staged dopri5 reverse sweeps on the SAME recorded accepted steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, odecol
from helpers import product_network
from ode_column_b200.solvers import _Setup
cfg = odecol.load_config(os.path.join(ROOT, "config", "model.toml"))
golden = {k: np.load(os.path.join(ROOT, "tests", "golden", k + ".npz")) for k in ("parity",)}
DEV = "cuda"
ext = odecol._native.ext()
g = golden["parity"]; N = 104
net = product_network("parity", cfg, g, DEV)
net.stim = torch.tensor(g["stims"]).to(DEV)
tv = net.time_vec[::25][:41].contiguous()
y0 = torch.zeros(4, 3 * N, device=DEV)
for fwd_fam in (None, "staged"):
    sf = _Setup(net, y0, tv, fwd_fam)
    pf = sf.problem(sf.lf.W_aug)
    y, na, nr, st, rec_y, rec_t0, rec_dt, out_step, out_x = ext.dopri5_fwd_record(pf, sf.t, y0, 1e-5, 1e-6, 4000000, 1024)
    grad = torch.zeros_like(y)
    grad[-1, :, :N] = torch.linspace(0.5, 1.5, N, device=DEV)
    out = {}
    for bwd_fam in (None, "staged"):
        sb = _Setup(net, y0, tv, bwd_fam)
        pb = sb.problem(sb.lf.W_aug)
        gy0, gW = ext.dopri5_bwd(pb, tv.numel(), rec_y, rec_t0, rec_dt, out_step, out_x, na, grad.contiguous(), None)
        out[bwd_fam] = (gy0.cpu(), gW.cpu())
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print(f"forward {fwd_fam}: accepted {na.tolist()}; reverse on-chip vs staged on the same record: dy0 per trial "
          f"{[f'{rel(out[None][0][b], out['staged'][0][b]):.1e}' for b in range(4)]}, dW {rel(out[None][1], out['staged'][1]):.1e}; "
          f"last step dt per trial {[float(rec_dt[b, na[b]-1]) for b in range(4)]}, t0 {[float(rec_t0[b, na[b]-1]) for b in range(4)]} "
          f"out_step tail {out_step[:, -3:].tolist()} out_x tail {out_x[:, -3:].tolist()}", flush=True)
