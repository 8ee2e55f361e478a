"""Where the 12 ms of the C1 forward+backward go: kernels vs host-side setup (run on the GPU box)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odecol
from oracle import stimuli
dev = torch.device("cuda")
cfg = odecol.load_config(os.path.join(ROOT, "config", "model.toml"))
torch.manual_seed(0)
net = odecol.ColumnAreaWTA(cfg, "mt").to(dev)
for m in [net] + list(net.modules()):
    for k, v in list(vars(m).items()):
        if torch.is_tensor(v) and not isinstance(v, torch.nn.Parameter):
            setattr(m, k, v.to(dev))
tv = stimuli.time_vec(1500, 1e-4).to(dev)
net.time_vec, net.stim = tv, stimuli.wta_stimulus(tv.cpu(), (20.0, 30.0)).to(dev)
y0 = torch.zeros(1, 48, device=dev)
target = (torch.linspace(0, 1, 1500).reshape(1, 1500, 1).repeat(1, 1, 2) * torch.tensor([0.6, 0.3])).to(dev)

def wall(fn, n=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n

def full():
    net.zero_grad()
    y = odecol.odeint(net, y0, tv, method="rk4")
    odecol.huber_loss_wta(y.unsqueeze(0), target, net).backward()

def fwd_only():
    with torch.no_grad():
        odecol.odeint(net, y0, tv, method="rk4")

from ode_column_b200.solvers import _Setup
def setup_only():
    s = _Setup(net, y0, tv, None)
    s.problem(s.lf.W_aug)

s = _Setup(net, y0, tv, None)
prob = s.problem(s.lf.W_aug)
ext = odecol._native.ext()
def k_fwd():
    return ext.rk4_fwd(prob, s.t, y0, 1)
y = k_fwd()
g = torch.randn_like(y)
def k_bwd():
    ext.rk4_bwd(prob, s.t, y, g, None)
def loss_only():
    yy = y.detach().clone().requires_grad_(True)
    odecol.huber_loss_wta(yy.unsqueeze(0), target, net).backward()

for name, fn in (("full fwd+loss+bwd", full), ("odeint forward (no grad)", fwd_only), ("_Setup + Problem", setup_only),
                 ("rk4_fwd kernel call", k_fwd), ("rk4_bwd kernel call", k_bwd), ("huber_loss_wta fwd+bwd (torch ops)", loss_only)):
    print(f"{name:40s} {wall(fn):8.3f} ms")
