import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import odecol
from helpers import product_network
cfg = odecol.load_config('config/model.toml')
golden = {k: np.load(f'tests/golden/{k}.npz') for k in ('wta','xor','parity')}
for name in ('wta','xor','parity'):
    g = golden[name]
    net = product_network(name, cfg, g, 'cuda')
    stims = torch.tensor(g['stim'])[None] if name=='wta' else torch.tensor(g['stims'])
    stims = stims.cuda()
    net.stim = stims if name!='wta' else stims[0]
    B = stims.shape[0]; N = net.export_linear_form().N
    y0 = torch.zeros(B, 3*N, device='cuda')
    dW = torch.tensor(g['em_dW'])[:, :, 0]
    with torch.no_grad():
        y = odecol.sdeint(net, y0, net.time_vec, bm=dW, method='euler', dt=1e-3).cpu()
    every = 1 if name=='wta' else (5 if name=='xor' else 10)
    ref = torch.tensor(g['em_traj'])
    if name=='wta': ref = ref.permute(1,0,2)   # (1,T,48)
    got = y[::every].permute(1,0,2)
    for b in range(B):
        for c,blk in enumerate('VAF'):
            d = (got[b,:,c*N:(c+1)*N]-ref[b,:,c*N:(c+1)*N]).abs()
            sc = ref[b,:,c*N:(c+1)*N].abs().max()
            tmax = int(d.max(1).values.argmax())
            first = (d.max(1).values > 1e-4*sc).nonzero()
            print(name, 'trial', b, blk, 'rel', float(d.max()/sc), 'scale', float(sc), 'at t-index', tmax, 'first>1e-4', int(first[0]) if len(first) else None, 'nan', bool(torch.isnan(got[b]).any()), bool(torch.isnan(ref[b]).any()))
