"""dW of one C4-shaped pass (64 columns, 8192 trials, 200 grid points) saved to gpurun_out/dw_<tag>.pt for comparison
between ODECOL_DW_WAVES settings (accumulation-chain length of the dW contraction)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odecol
import bench
dev = torch.device("cuda")
cfg = odecol.load_config(os.path.join(ROOT, "config", "model.toml"))
columns, B, T = 64, 8192, 200
net = odecol.SyntheticColumnSheet(cfg, columns, seed=0, device=dev)
kt, ku, _ = bench.make_stimulus(torch, B, columns, T, 1e-4, 0, "cpu")
net.set_knots(kt.to(dev), ku.to(dev))
tv = torch.linspace(0.0, T * 1e-4, T, device=dev)
sel = bench.loss_components(torch, columns).to(dev)
target = torch.full((1, 1, columns), 0.5, device=dev)
y = odecol.odeint(net, torch.zeros(B, 3 * 512, device=dev), tv, method="rk4", components=sel)
odecol.huber_rate_loss(y, target, 1).backward()
torch.save({"W": net.recurrent_weights.grad.cpu(), "U": net.input_weights.grad.cpu()}, os.path.join(ROOT, "gpurun_out", f"dw_{sys.argv[1]}.pt"))
print("saved", sys.argv[1], float(net.recurrent_weights.grad.abs().max()))
