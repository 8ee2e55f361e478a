import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odecol
ext = odecol._native.ext()
def tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
g = torch.Generator().manual_seed(0)
for K in (64, 608, 2432):
    M, N = 512, 1024
    A = (torch.randn(M, K, generator=g) * 30).cuda(); B = (torch.randn(N, K, generator=g).abs() * 5).cuda()
    for name, (a, b) in {'general': (A, B), 'tf32-exact inputs': (tf32(A), tf32(B)), 'all-positive exact': (tf32(A.abs()), tf32(B))}.items():
        C = ext.tc_contract(a.contiguous(), b.contiguous())
        ref = b.double() @ a.double().T
        mag = b.double().abs() @ a.double().abs().T
        d = (C.double() - ref)
        print(f'K={K:5d} {name:22s} max|err|/mag {float((d.abs()/mag).max()):.2e}  mean signed err/mag {float((d/mag).mean()):+.2e}  rms {float(((d/mag)**2).mean().sqrt()):.2e}')
