"""The reference's WTA training (scripts/wta_ode.py:139-213) on the drop-in API, batched.

What changes against the reference script: the Wong-Wang targets of all samples come from one kernel launch
(`odecol.get_data`, same numbers for the same numpy seed), and the per-sample Python loop over `sdeint(..., method='srk')`
(wta_ode.py:167-176) becomes ONE fused solve over the batch; the loss, the gradient mask, RMSprop and the scheduler are
the script's.  No plotting.

    python examples/train_wta.py --samples 54 --batch 16 --iters 4
"""
import argparse

import numpy as np
import torch

from common import CONFIG, to_device
import odecol
from odecol import ColumnAreaWTA, huber_loss_wta, load_config, sdeint


def set_stim_three_phases(num_populations, time_vec, raw_stims):
    """Batched wta_ode.py:109-122: raw_stims (S, 2) -> (S, T, num_populations), input on L4e / L4i of both columns
    during the middle third of the window."""
    S, T = raw_stims.shape[0], len(time_vec)
    vec = torch.zeros(S, num_populations, device=raw_stims.device)
    vec[:, 2] = vec[:, 3] = raw_stims[:, 0]
    vec[:, 10] = vec[:, 11] = raw_stims[:, 1]
    out = torch.zeros(S, T, num_populations, device=raw_stims.device)
    on = int(T / 3)
    off = int(on + T / 3)
    out[:, on:off] = vec[:, None, :]
    return out


def train(nr_samples=54, batch_size=16, iters=None, device="cuda", sigma_scale=1.0, seed=0, verbose=True):
    torch.manual_seed(seed)
    np.random.seed(seed)
    dt, stim_phase = 1e-4, 0.05
    time_steps = int((stim_phase * 3) / dt)
    loader = odecol.get_data(nr_samples, batch_size, time_steps, None, device=device)          # wta_ode.py:150
    network = to_device(ColumnAreaWTA(load_config(CONFIG), area="mt"), device)
    time_vec = torch.linspace(0., time_steps * dt, time_steps, device=device)
    network.time_vec = time_vec
    optimizer = torch.optim.RMSprop([network.recurrent_weights], lr=10.0, alpha=0.9)
    scheduler = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=0.99)
    losses = []
    for it, (true_states, stim_batch) in enumerate(loader):
        if iters is not None and it >= iters:
            break
        optimizer.zero_grad()
        true_states, stim_batch = true_states.to(device), stim_batch.to(device)
        S = true_states.shape[0] - 1                                                           # last sample: validation
        network.stim = set_stim_three_phases(network.num_populations, time_vec, stim_batch[:S])
        y = sdeint(network, torch.zeros(S, 48, device=device), time_vec, names={"drift": "forward", "diffusion": "diffusion"},
                   method="srk", seed=seed + it, options={"sigma_scale": torch.full((S,), float(sigma_scale))})
        pred_states = y.permute(1, 0, 2).unsqueeze(2)                                          # (S, T, 1, 48) as the script stacks them
        loss = huber_loss_wta(pred_states, true_states[:-1], network)
        loss.backward()
        with torch.no_grad():
            network.recurrent_weights.grad *= network.lat_in_mask                              # wta_ode.py:182-183
        optimizer.step()
        scheduler.step()
        with torch.no_grad():
            network.stim = set_stim_three_phases(network.num_populations, time_vec, stim_batch[-1:])
            pred = sdeint(network, torch.zeros(1, 48, device=device), time_vec, method="srk", seed=10_000 + it,
                          options={"sigma_scale": [float(sigma_scale)]})
            test_loss = huber_loss_wta(pred.permute(1, 0, 2).unsqueeze(2), true_states[-1:], network)
        losses.append((float(loss.detach()), float(test_loss)))
        if verbose:
            print("Iter {:02d} | Total Loss {:.5f} | Validation {:.5f}".format(it + 1, *losses[-1]))
    return network, losses


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=54)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=None)
    ap.add_argument("--sigma-scale", type=float, default=1.0, help="1.0 = the reference's diffusion (sigma = 100)")
    a = ap.parse_args()
    train(a.samples, a.batch, a.iters, sigma_scale=a.sigma_scale)
