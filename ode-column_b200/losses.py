"""Read-outs and losses the reference scripts apply to solver output (plain torch on the trajectory).

Mirrors reference src/utils.py:51-88 (min_max, fr_to_binary, huber_loss_wta) and the script-level read-outs of
scripts/xor_ode.py:120-130 and scripts/parity_ode.py:239-249, batched as (T, B, 3N) trajectories.
"""
from __future__ import annotations

import torch

from .model import compute_firing_rate


def min_max(firing_rates: torch.Tensor) -> torch.Tensor:
    lo, hi = firing_rates.min(), firing_rates.max()
    return (firing_rates - lo) / (hi - lo)


def fr_to_binary(firing_rates: torch.Tensor, scaling_factor: float = 1.0) -> torch.Tensor:
    z = (firing_rates - firing_rates.mean()) / (firing_rates.std() / scaling_factor)
    return torch.sigmoid(z)


def huber_loss_wta(pred_states: torch.Tensor, true: torch.Tensor, network) -> torch.Tensor:
    """pred_states (S, T, 1, 48) as the reference stacks them, true (S, T, 2)."""
    rate = compute_firing_rate(pred_states[:, :, 0, :16] - pred_states[:, :, 0, 16:32])
    w = network.output_weights.to(rate.device)
    both = torch.stack(((rate[:, :, :8] * w).sum(2), (rate[:, :, 8:] * w).sum(2)), dim=2)
    return torch.nn.functional.smooth_l1_loss(both, true, beta=1.0)


def xor_readout(traj: torch.Tensor, network) -> torch.Tensor:
    """traj (T, B, 72) -> final L2/3e rate of column C per trial (scripts/xor_ode.py:120-125)."""
    rate = compute_firing_rate(traj[-1, :, 16:24] - traj[-1, :, 40:48])
    return (rate * network.ff_source_mask.to(rate.device)).sum(dim=1)


def parity_readout(traj: torch.Tensor, network, last: int = 100) -> torch.Tensor:
    """traj (T, B, 3N) -> weighted mean rate of the output column over the last `last` points
    (scripts/parity_ode.py:239-243)."""
    N = traj.shape[2] // 3
    rate = compute_firing_rate(traj[-last:, :, N - 8:N] - traj[-last:, :, 2 * N - 8:2 * N])
    return (rate.mean(dim=0) * network.output_weights / network.output_scale).sum(dim=-1)
