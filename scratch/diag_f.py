import sys, torch
sys.path.insert(0, "/root/repo")
import odecol
DEV = "cuda"
cfg = odecol.load_config("/root/repo/config/model.toml")
torch.manual_seed(3)
B, T = 301, 24
net = odecol.SyntheticColumnSheet(cfg, 32, seed=5, device=DEV)
kt = torch.tensor([0.0, 5e-4, 1e-3, 3e-3], device=DEV)
ku = torch.rand(B, 4, 32, device=DEV) * 20.0
net.set_knots(kt, ku)
tv = torch.linspace(0, (T - 1) * 1e-4, T, device=DEV)
y0 = torch.zeros(B, 3 * 256, device=DEV)
comps = list(range(2 * 256, 2 * 256 + 256, 8)) + [0, 300]
out = {}
for ck in (False, True):
    y0r = y0.clone().requires_grad_(True)
    y = odecol.odeint(net, y0r, tv, method="rk4", components=comps, options={"family": "tensor", "checkpoint": ck})
    out[ck] = y.detach().clone()
    print(ck, type(y.grad_fn).__name__)
d = (out[False] - out[True]).abs()
print("max diff", float(d.max()))
idx = (d > 1e-3).nonzero()
print("n bad", len(idx), "of", d.numel())
print("bad columns", sorted(set(idx[:, 2].tolist()))[:40])
print("bad rows(t)", sorted(set(idx[:, 0].tolist()))[:30])
print("bad trials", sorted(set(idx[:, 1].tolist()))[:20], "...")
print(out[False][5, :3, :4], out[True][5, :3, :4])
