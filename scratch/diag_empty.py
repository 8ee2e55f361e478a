import sys, faulthandler, torch
faulthandler.enable()
sys.path.insert(0, "/root/repo")
import odecol
ext = odecol._native.ext()
DEV = "cuda"
N = 16
def mk(B):
    return ext.Problem(torch.zeros(N, 36, device=DEV), torch.zeros(N, device=DEV), torch.ones(3 * N, device=DEV),
                       torch.tensor([0.0, 1.0], device=DEV), torch.zeros(1, 2, 16, device=DEV), 16, B, 5e-4, 0.02, 10.0, 80.0, 0)
ts = torch.linspace(0, 1e-3, 11, device=DEV)
print("B=2", flush=True)
y = ext.rk4_fwd(mk(2), ts, torch.zeros(2, 48, device=DEV), 1); print(y.shape, flush=True)
print("workspace_bytes B=0", mk(0).workspace_bytes(ext.OP_RK4_FWD, 11, 0), flush=True)
print("kernel_family B=0", mk(0).kernel_family(ext.OP_RK4_FWD), flush=True)
try:
    print("T=1", flush=True)
    ext.rk4_fwd(mk(2), ts[:1].contiguous(), torch.zeros(2, 48, device=DEV), 1)
except Exception as e:
    print("raised", type(e).__name__, str(e)[:80], flush=True)
try:
    print("B=0 call", flush=True)
    ext.rk4_fwd(mk(0), ts, torch.zeros(0, 48, device=DEV), 1)
except Exception as e:
    print("raised", type(e).__name__, str(e)[:80], flush=True)
print("done", flush=True)
