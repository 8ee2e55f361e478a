"""A/B of the operand format of the staged dopri5 / srk forward solves (ODECOL_DP16 / ODECOL_SRK16, read once per process):
python scratch/ab_dp16.py [columns] [trials]  -> one line per solver with the solve time (CUDA events, best of three runs)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odecol
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg = odecol.load_config(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config", "model.toml"))
N = 8 * cols
sheet = odecol.SyntheticColumnSheet(cfg, cols, seed=0, device="cuda")
gen = torch.Generator().manual_seed(5)
amp = torch.rand(B, cols, generator=gen) * 30.0
kt, ku = odecol.step_knots(0.002, 0.006, 0.01, amp, 1e-3)
sheet.set_knots(kt.cuda(), ku.cuda())
ts = torch.linspace(0.0, 0.01, 11).cuda()
y0 = torch.zeros(B, 3 * N, device="cuda")
sel = torch.arange(0, N, 8)
def timed(fn):
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best, out
st = {}
with torch.no_grad():
    ms_d, yd = timed(lambda: odecol.odeint(sheet, y0, ts, rtol=1e-5, atol=1e-6, stats=st, components=sel))
    ms_s, ys = timed(lambda: odecol.sdeint(sheet, y0, ts, method="srk", dt=1e-4, seed=3, components=sel,
                                            options={"sigma_scale": torch.full((B,), 0.1)}))
print(json.dumps({"DP16": os.environ.get("ODECOL_DP16", "1"), "SRK16": os.environ.get("ODECOL_SRK16", "1"), "N": N, "trials": B,
                  "dopri5_ms": ms_d, "dopri5_accepted_mean": float(st["n_accept"].float().mean()),
                  "dopri5_rejected_mean": float(st["n_reject"].float().mean()), "dopri5_checksum": float(yd.double().abs().sum()),
                  "srk_ms": ms_s, "srk_steps": 100, "srk_checksum": float(ys.double().abs().sum())}))
