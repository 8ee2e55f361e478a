"""The reference's parity training (scripts/parity_ode.py:198-283) on the drop-in API: one fused rk4 solve per batch
instead of the per-sample loop over `odeint` (parity_ode.py:223-236), read-out over the last 100 grid points, Adam, the
script's gradient masks and weight clamps.  No plotting / pickling.

    python examples/train_parity.py --iters 3 --batch 4
"""
import argparse

import torch

from common import CONFIG, to_device
from odecol import ColumnNetwork, compute_firing_rate, load_config, odeint


def make_ds(batch_size, device):
    """parity_ode.py:116-137 (fixed position): the four fixed-position patterns x 15, tiled and shuffled."""
    base = torch.tensor([[0., 0., 0., 1.], [0., 0., 1., 1.], [0., 1., 1., 1.], [1., 1., 1., 1.]], device=device) * 15.
    allc = torch.tile(base, (max(1, batch_size // 4), 1))
    return allc[torch.randperm(allc.size(0), device=device)][:batch_size]


def prep_stim_ode(stims, time_vec):
    """Batched parity_ode.py:139-153: (S, 4) -> (S, T, 4): zeros for the first half, the pattern for the second."""
    T = len(time_vec)
    half = int(T / 2)
    out = torch.zeros(stims.shape[0], 2 * half, stims.shape[1], device=stims.device)
    out[:, half:] = stims[:, None, :]
    return out


def mask_weights(network):
    network.output_weights.grad *= network.output_mask
    network.areas["0"].input_weights.grad *= network.areas["0"].input_mask
    for a in range(1, network.nr_areas):
        network.areas[str(a)].feedforward_weights.grad *= network.areas[str(a)].feedforward_mask
    for a in range(network.nr_areas - 1):
        network.areas[str(a)].lateral_weights.grad *= network.areas[str(a)].lateral_mask


def train(iters=3, batch_size=4, device="cuda", seed=0, verbose=True):
    torch.manual_seed(seed)
    nd = {"nr_areas": 3, "areas": ["mt", "mt", "mt"], "nr_columns_per_area": [8, 4, 1], "nr_input_units": 4}
    network = to_device(ColumnNetwork(load_config(CONFIG), nd, torch.device("cpu")), device)
    dt, stim_duration = 1e-3, 0.5
    time_steps = int(stim_duration * 2 / dt)
    time_vec = torch.linspace(0., time_steps * dt, time_steps, device=device)
    network.time_vec = time_vec
    N = network.network_as_area.num_populations
    optimizer = torch.optim.Adam(network.parameters(), lr=0.1, betas=(0.9, 0.999), eps=1e-08)
    losses = []
    for it in range(iters):
        optimizer.zero_grad()
        train_set = make_ds(batch_size, device)
        network.stim = prep_stim_ode(train_set, time_vec)
        y = odeint(network, torch.zeros(batch_size, 3 * N, device=device), time_vec, method="rk4")     # (T, S, 3N)
        firing_rates = compute_firing_rate(y[:, :, :N] - y[:, :, N:2 * N]).permute(1, 0, 2)           # (S, T, N)
        final_fr = torch.mean(firing_rates[:, -100:, -8:], dim=1)                                    # parity_ode.py:239-243
        summed = torch.sum(final_fr * network.output_weights / network.output_scale, dim=-1)
        targets = (train_set.sum(dim=1) % 30 == 0).float() * 20.0                                    # parity_ode.py:245-246
        loss = torch.mean(abs(summed - targets))
        loss.backward()
        mask_weights(network)
        optimizer.step()
        with torch.no_grad():                                                                        # parity_ode.py:262-268
            for name, param in network.named_parameters():
                if "lateral" in name:
                    param.clamp_(max=0.0)
                if "lateral" not in name:
                    param.clamp_(min=0.0)
                if "output" in name:
                    param.clamp_(min=0.0, max=float(network.output_scale))
        losses.append(float(loss.detach()))
        if verbose:
            print("Iter {:02d} | Total Loss {:.5f}".format(it + 1, losses[-1]))
    return network, losses


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4)
    a = ap.parse_args()
    train(a.iters, a.batch)
