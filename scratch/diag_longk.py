import sys, torch
sys.path.insert(0, "/root/repo")
import odecol
ext = odecol._native.ext()
for K in (608, 2048, 8192):
    g = torch.Generator().manual_seed(K)
    for label, A, B in (("mixed-sign A, positive B", torch.randn(256, K, generator=g) * 3, torch.rand(384, K, generator=g) * 5),
                        ("all positive", torch.rand(256, K, generator=g) * 3, torch.rand(384, K, generator=g) * 5)):
        A, B = A.cuda(), B.cuda()
        C = ext.tc_contract(A, B)
        ref = B.double() @ A.double().T
        mag = B.double().abs() @ A.double().abs().T
        err = ((C.double() - ref) / mag)
        print(f"K={K:5d} {label:26s} max |err|/sum|a||b| {float(err.abs().max()):.2e}  mean signed {float(err.mean()):+.2e}   cuBLAS fp32 {float((((B @ A.T).double() - ref) / mag).abs().max()):.2e}")
