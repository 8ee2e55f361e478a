"""Profiling driver for the kernels added in session 2 (run plain first, then under ncu): the packed on-chip rk4 / srk
kernels on the WTA network, the fused Huber read-out and the Wong-Wang generator.
   python scratch/prof_small.py [trials] [time_points]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odecol

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda")
cfg = odecol.load_config(os.path.join(ROOT, "config", "model.toml"))
torch.manual_seed(0)
net = odecol.ColumnAreaWTA(cfg, "mt").to(dev)
for m in [net] + list(net.modules()):
    for k, v in list(vars(m).items()):
        if torch.is_tensor(v) and not isinstance(v, torch.nn.Parameter):
            setattr(m, k, v.to(dev))
amp = torch.zeros(B, 16)
a = torch.rand(B, 2) * 30
amp[:, 2] = amp[:, 3] = a[:, 0]; amp[:, 10] = amp[:, 11] = a[:, 1]
dt = 1e-4
grid = T * dt / (T - 1)
kt, ku = odecol.step_knots((T // 3) * grid, (2 * (T // 3)) * grid, T * dt, amp, grid)
net.time_vec, net.stim = kt.to(dev), ku.to(dev)
tv = torch.linspace(0, T * dt, T, device=dev)
y0 = torch.zeros(B, 48, device=dev)
sel = torch.tensor([0, 8, 16, 24], device=dev)
target = torch.full((1, 1, 2), 0.5, device=dev)
for _ in range(2):
    net.zero_grad()
    y = odecol.odeint(net, y0, tv, method="rk4", components=sel)
    odecol.huber_rate_loss(y, target, 1).backward()
    net.zero_grad()
    y = odecol.sdeint(net, y0, tv, method="srk", dt=1e-3, seed=0, components=sel, options={"sigma_scale": torch.full((B,), 0.1)})
    odecol.huber_rate_loss(y, target, 1).backward()
    big = torch.randn(300, 8192, 128, device=dev) * 5 - 5
    big.requires_grad_(True)
    odecol.huber_rate_loss(big, torch.full((1, 1, 64), 0.5, device=dev), 1).backward()
    del big
    np.random.seed(0)
    states = odecol.wongwang.generate_states(odecol.wongwang.sample_stimuli(64).repeat(47, axis=0)[:3010], 1500)
torch.cuda.synchronize()
print("ok", float(states.mean()))
