"""The example training loops (the reference's three scripts on the drop-in API, batched) run end to end on the GPU:
data -> fused solve -> the script's loss -> backward through the fused adjoint -> the script's masks / optimizer."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))
pytestmark = pytest.mark.gpu


def test_wta_training_loop_runs_and_updates_only_masked_weights():
    import train_wta
    net, losses = train_wta.train(nr_samples=22, batch_size=16, iters=2, sigma_scale=0.05, verbose=False)
    assert len(losses) == 2 and all(l == l and v == v for l, v in losses)                  # finite
    cfg_net = train_wta.ColumnAreaWTA(train_wta.load_config(train_wta.CONFIG), area="mt")
    changed = (net.recurrent_weights.detach().cpu() != cfg_net.recurrent_weights.detach()).float()
    mask = net.lat_in_mask.cpu()
    assert float((changed * (1 - mask)).sum()) == 0 and float((changed * mask).sum()) >= 1  # only lateral / self-excitation entries move


def test_xor_training_loop_decreases_the_loss():
    import train_xor
    net, losses = train_xor.train(iters=4, verbose=False)
    assert all(l == l for l in losses) and min(losses[1:]) < losses[0]


def test_parity_training_loop_runs_with_clamps():
    import train_parity
    net, losses = train_parity.train(iters=2, verbose=False)
    assert all(l == l for l in losses)
    for name, p in net.named_parameters():
        if "lateral" in name:
            assert float(p.max()) <= 0.0
        else:
            assert float(p.min()) >= 0.0
