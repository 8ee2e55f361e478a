import os, sys, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import odecol, bench
class A: pass
args=A(); args.columns=64; args.trials_per_gpu=8192; args.time_points=8; args.dt=1e-4
cfg=odecol.load_config(os.path.join(ROOT,'config/model.toml'))
dev=torch.device('cuda')
net=odecol.SyntheticColumnSheet(cfg,64,seed=0,device=dev)
kt,ku,_=bench.make_stimulus(torch,8192,64,1500,1e-4,0,'cpu')
net.set_knots(kt.to(dev),ku.to(dev))
tv=torch.linspace(0,1500*1e-4,1500,device=dev)[:8].contiguous()
y0=torch.zeros(8192,1536,device=dev)
with torch.no_grad():
    for _ in range(2):
        y=odecol.odeint(net,y0,tv,method='rk4')
torch.cuda.synchronize()
