import sys, torch
sys.path.insert(0, "/root/repo")
import odecol
ext = odecol._native.ext()
DEV = "cuda"
N, B, T = 8, 64, 9
for flags in (0, ext.FLAG_FORCE_STAGED):
    mk = lambda: ext.Problem(torch.zeros(N, 12, device=DEV), torch.zeros(N, device=DEV), torch.ones(3 * N, device=DEV),
                             torch.tensor([0.0, 1.0], device=DEV), torch.zeros(1, 2, 1, device=DEV), 1, B, 1e30, 1e30, 1e30, 1.0, flags)
    ts = torch.linspace(0, 1, T, device=DEV)
    y0 = torch.zeros(B, 3 * N, device=DEV)
    scale = torch.tensor([0.0, 0.5, 1.0, 2.0], device=DEV).repeat(B // 4)
    base, scaled = mk(), mk()
    scaled.set_sigma_scale(scale)
    for adaptive in (False, True):
        a, *_ = ext.em_fwd(base, ts, y0, None, 11, 0, 1.0 / 32, adaptive, 1e-3, 1e-3, 1e-4, 0)
        b, *_ = ext.em_fwd(scaled, ts, y0, None, 11, 0, 1.0 / 32, adaptive, 1e-3, 1e-3, 1e-4, 0)
        d = (b - a * scale[None, :, None])
        print(flags, adaptive, float(d.abs().max()), torch.equal(a, b), a[-1, :4, 0].tolist(), b[-1, :4, 0].tolist(), b[-1, :4, 9].tolist(), b[-1,:4,17].tolist())
