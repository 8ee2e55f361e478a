"""tiny_tc.cu against the on-chip kernel: the same WTA trials as one batch of 4224 (tensor-core path, ragged last CTA) and as
three batches below the 4096-trial threshold (on-chip path)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odecol
cfg = odecol.load_config(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config", "model.toml"))
dev = "cuda"
net = odecol.ColumnAreaWTA(cfg, "mt").to(dev)
for m in [net] + list(net.modules()):
    for k, v in list(vars(m).items()):
        if torch.is_tensor(v) and not isinstance(v, torch.nn.Parameter):
            setattr(m, k, v.to(dev))
B, T, dt = 4224, 400, 1e-4
g = torch.Generator().manual_seed(1)
amp = torch.zeros(B, 16); a = torch.rand(B, 2, generator=g) * 30.0
amp[:, 2] = amp[:, 3] = a[:, 0]; amp[:, 10] = amp[:, 11] = a[:, 1]
t_end = T * dt; grid = t_end / (T - 1)
kt, ku = odecol.step_knots((T // 3) * grid, (2 * (T // 3)) * grid, t_end, amp, grid)
tv = torch.linspace(0.0, t_end, T, device=dev)
y0 = torch.rand(B, 48, generator=g).to(dev) * 2 - 1
def solve(lo, hi, every=1):
    net.time_vec, net.stim = kt.to(dev), ku[lo:hi].to(dev)
    with torch.no_grad():
        return odecol.odeint(net, y0[lo:hi], tv[::every] if every > 1 else tv, method="rk4")
big = solve(0, B)
ref = torch.cat([solve(lo, min(B, lo + 2000)) for lo in range(0, B, 2000)], 1)
torch.cuda.synchronize()
print("finite", bool(torch.isfinite(big).all()), "shape", tuple(big.shape))
for c, nm in enumerate("VAF"):
    d = (big[..., 16 * c:16 * c + 16] - ref[..., 16 * c:16 * c + 16]).abs().max() / ref[..., 16 * c:16 * c + 16].abs().max()
    print(nm, "rel diff", float(d))
worst = ((big - ref).abs().amax(dim=(0, 2)) / ref.abs().amax(dim=(0, 2))).max()
print("worst trial", float(worst))
