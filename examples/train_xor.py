"""The reference's XOR training (scripts/xor_ode.py:93-200) on the drop-in API: the four input patterns of a batch are
ONE fused dopri5 solve (the script's default `odeint(network, initial_state, time_vec)`) with the discrete adjoint; the
read-out, loss, gradient masks, RMSprop and scheduler are the script's.  No plotting.

    python examples/train_xor.py --iters 5
"""
import argparse

import torch

from common import CONFIG, to_device
from odecol import ColumnNetworkXOR, compute_firing_rate, load_config, min_max, odeint

CONDITIONS = ((20.0, 0.0), (0.0, 20.0), (20.0, 20.0), (0.0, 0.0))


def make_stim(device):
    """xor_ode.py:52-74: the four conditions as 16-vectors (input on L4e / L4i of columns A and B), shuffled."""
    stims = torch.zeros(4, 16, device=device)
    for k, (a, b) in enumerate(CONDITIONS):
        stims[k, 2] = stims[k, 3] = a
        stims[k, 10] = stims[k, 11] = b
    return stims[torch.randperm(4, device=device)]


def prep_stim_ode(stims, time_vec):
    """Batched xor_ode.py:76-91: (S, 16) -> (S, T, 2, 16): second half of the window carries the pattern; the second
    channel is the column-swapped copy."""
    T = len(time_vec)
    half = int(T / 2)
    whole = torch.zeros(stims.shape[0], T, 16, device=stims.device)
    whole[:, half:half * 2] = stims[:, None, :]
    mirror = torch.cat((whole[..., 8:], whole[..., :8]), dim=-1)
    return torch.stack((whole, mirror), dim=2)


def run_four_xor_samples(network, time_vec, batch_size=4):
    stim_batch = torch.cat([make_stim(time_vec.device) for _ in range(batch_size // 4)])
    network.stim = prep_stim_ode(stim_batch, time_vec)
    y = odeint(network, torch.zeros(batch_size, 72, device=time_vec.device), time_vec)          # default dopri5, as the script
    firing_rates = compute_firing_rate(y[:, :, :24] - y[:, :, 24:48]).permute(1, 0, 2)            # (S, T, 24)
    final_fr_C = torch.sum(firing_rates[:, -1, 16:] * network.ff_source_mask, dim=1)
    xor_targets = torch.where(stim_batch[:, 2] != stim_batch[:, 10], 1.0, 0.25)
    loss = torch.mean(abs(final_fr_C - xor_targets))
    return min_max(final_fr_C), loss, firing_rates, stim_batch, xor_targets


def train(iters=5, batch_size=4, device="cuda", seed=0, verbose=True):
    torch.manual_seed(seed)
    nd = {"nr_areas": 2, "areas": ["mt", "mt"], "nr_columns_per_area": [2, 1], "nr_input_units": 2}
    network = to_device(ColumnNetworkXOR(load_config(CONFIG), nd), device)
    dt, stim_duration = 1e-3, 0.5
    time_steps = int(stim_duration * 2 / dt)
    time_vec = torch.linspace(0., time_steps * dt, time_steps, device=device)
    network.time_vec = time_vec
    optimizer = torch.optim.RMSprop(network.parameters(), lr=0.5, alpha=0.95)
    scheduler = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=0.8)
    losses = []
    for itr in range(iters):
        optimizer.zero_grad()
        _, loss, *_ = run_four_xor_samples(network, time_vec, batch_size)
        loss.backward()
        with torch.no_grad():                                                                   # xor_ode.py:179-183
            network.feedforward_target_weights["0"][0].grad *= torch.tile(network.ff_target_mask, (2,))
            network.feedforward_target_weights["0"][1].grad *= torch.tile(network.ff_target_mask, (2,))
            network.feedforward_target_weights["1"][0].grad *= network.ff_target_mask
            network.feedforward_target_weights["1"][1].grad *= network.ff_target_mask
        optimizer.step()
        scheduler.step()
        losses.append(float(loss.detach()))
        if verbose:
            print("Iter {:02d} | Total Loss {:.5f}".format(itr + 1, losses[-1]))
    return network, losses


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4)
    a = ap.parse_args()
    train(a.iters, a.batch)
