"""Stimulus tables -> knot lists (host logic of the boundary).

The reference hands the solver a dense table sampled on ``time_vec`` (``network.stim``; scripts/wta_ode.py:109-122,
scripts/xor_ode.py:76-91, scripts/parity_ode.py:139-153) and looks it up with ``torch_interp`` (src/utils.py:31-46).
The kernels take per-trial knot lists instead.  ``compress_knots`` drops every sample whose two neighbours carry the
same value for all trials and channels: inside such a run the reference's interpolation has slope exactly zero, so
the lookup is unchanged bit for bit, and a (T, n_in) step stimulus shrinks to a handful of knots.
"""
from __future__ import annotations

from typing import Tuple

import torch


def compress_knots(time_vec: torch.Tensor, table: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """time_vec (T,), table (B, T, n_in) -> knot_t (K,), knot_u (B, K, n_in), K >= 2."""
    assert table.dim() == 3 and table.shape[1] == time_vec.shape[0]
    T = time_vec.shape[0]
    if T <= 2:
        return time_vec.contiguous(), table.contiguous()
    change = (table[:, 1:, :] != table[:, :-1, :]).any(dim=2).any(dim=0)      # (T-1,) change between j and j+1
    keep = torch.zeros(T, dtype=torch.bool, device=table.device)
    keep[0] = keep[T - 1] = True
    keep[:-1] |= change
    keep[1:] |= change
    idx = keep.nonzero().flatten()
    return time_vec[idx].contiguous(), table[:, idx, :].contiguous()


def step_knots(t_on: float, t_off: float, t_end: float, amplitudes: torch.Tensor, ramp: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Knots of an off/on/off stimulus with one-interval linear ramps (what sampling a step on the reference's
    time grid produces).  amplitudes (B, n_in) -> knot_t (6,), knot_u (B, 6, n_in)."""
    dev = amplitudes.device
    kt = torch.tensor([0.0, t_on - ramp, t_on, t_off - ramp, t_off, t_end], dtype=torch.float32, device=dev)
    z = torch.zeros_like(amplitudes)
    ku = torch.stack((z, z, amplitudes, amplitudes, z, z), dim=1)
    return kt, ku.contiguous()
