import sys, torch
sys.path.insert(0, "/root/repo")
import odecol
ext = odecol._native.ext()
cfg = odecol.load_config("/root/repo/config/model.toml")
for cols in (64, 1024):
    sheet = odecol.SyntheticColumnSheet(cfg, cols, seed=0, device="cuda")
    lf = sheet.export_linear_form()
    W = lf.W_aug.detach()[:, :8 * cols + cols + 1].contiguous()
    g = torch.Generator().manual_seed(1)
    B = 256
    r = (torch.rand(B, 8 * cols, generator=g) * 4.0)
    s = torch.rand(B, cols, generator=g) * 30
    R = torch.cat((r, s, torch.ones(B, 1)), 1).cuda().contiguous()
    C = ext.tc_contract(W, R)                       # (B, N): input current
    ref = R.double() @ W.double().T
    mag = R.double().abs() @ W.double().abs().T
    e = (C.double() - ref)
    print(f"columns {cols}: |I| typical {float(ref.abs().median()):.3f}; err/|I|: max {float((e/ref).abs().max()):.2e} median {float((e/ref).abs().median()):.2e} mean signed {float((e/ref).mean()):+.2e};"
          f" err/sum|a||b|: max {float((e/mag).abs().max()):.2e} mean signed {float((e/mag).mean()):+.2e}; fp32 matmul err/|I| max {float((((R @ W.T).double()-ref)/ref).abs().max()):.2e}")
