import sys, os, numpy as np, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import odecol
from helpers import product_network
cfg = odecol.load_config(os.path.join(ROOT,'config/model.toml'))
g = np.load(os.path.join(ROOT,'tests/golden/parity.npz'))
T = int(sys.argv[1]) if len(sys.argv) > 1 else 6
res = {}
for fam in ('staged', 'tensor'):
    net = product_network('parity', cfg, g, 'cuda')
    stims = torch.tensor(g['stims']).cuda()
    net.time_vec = net.time_vec[:T]; net.stim = stims[:, :T]
    gen = torch.Generator().manual_seed(1)
    y0 = (torch.rand(4, 312, generator=gen) * 2 - 1).cuda().requires_grad_(True)
    y = odecol.odeint(net, y0, net.time_vec, method='rk4', options={'family': fam})
    w = torch.randn(T, 4, 312, generator=gen).cuda()
    (y * w).sum().backward()
    lf_grad = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    res[fam] = (y.detach(), y0.grad.clone(), lf_grad)
ys, g0s, gs = res['staged']; yt, g0t, gt = res['tensor']
print('traj diff', float((ys-yt).abs().max()/ys.abs().max()))
print('grad_y0 rel diff', float((g0s-g0t).abs().max()/g0s.abs().max()), 'norms', float(g0s.abs().max()), float(g0t.abs().max()))
for n in gs:
    print(n, 'rel diff', float((gs[n]-gt[n]).abs().max()/gs[n].abs().max()), 'max staged', float(gs[n].abs().max()), 'max tensor', float(gt[n].abs().max()))
