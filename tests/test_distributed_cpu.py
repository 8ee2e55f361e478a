"""Trial-parallel sharding on the CPU (gloo, world_size 2): shard bounds, the single flattened all-reduce of parameter
gradients, and the property the multi-GPU path relies on -- the sum over shards of per-shard dW equals the full-batch
dW (checked with the CPU oracle's autograd so that it runs without a GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_every_trial_once():
    import odecol
    for B in (1, 7, 8, 65536, 1000):
        for W in (1, 2, 3, 8):
            spans = [odecol.distributed.shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[r][1] == spans[r + 1][0] for r in range(W - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import odecol
    from oracle import rhs as orhs, solvers as S
    from helpers import oracle_form, stim_table
    cfg = odecol.load_config(os.path.join(ROOT, "config", "model.toml"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "xor.npz"))
    T = 40
    tv = g["time_vec"][:T]
    stims = g["stims"][:, :T]
    lo, hi = odecol.distributed.shard_bounds(4, rank, world)
    assert torch.equal(odecol.distributed.shard_trials(torch.arange(4), rank, world), torch.arange(4)[lo:hi])
    net = torch.nn.ParameterList([torch.nn.Parameter(torch.tensor(oracle_form("xor", cfg, g).W))])

    def grad_of(trials):
        lf = oracle_form("xor", cfg, g)
        ode = orhs.UnifiedColumnODE(lf, tv, stim_table("xor", stims[trials]), requires_grad=True)
        y = S.odeint_rk4(ode, torch.zeros(len(trials), 72), torch.tensor(tv))
        y[-1, :, :24].sum().backward()
        return ode.W.grad

    net[0].grad = grad_of(list(range(lo, hi)))
    sent = odecol.distributed.allreduce_gradients(net)
    assert sent == 24 * 24
    if rank == 0:
        full = grad_of([0, 1, 2, 3])
        out.put(float((net[0].grad - full).abs().max() / full.abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_of_sharded_gradients_equals_full_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert out.get(timeout=5) < 1e-5
