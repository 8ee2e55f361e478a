"""Oracle restatement of the reference scripts' stimulus / time-grid builders (TEST INFRASTRUCTURE ONLY).

Follows
  * /root/reference/scripts/wta_ode.py:109-122  set_stim_three_phases     (T,16)
  * /root/reference/scripts/wta_ode.py:124-137  init_network              time_vec = linspace(0, T*dt, T)
  * /root/reference/scripts/xor_ode.py:52-91    make_stim / prep_stim_ode (T,2,16)
  * /root/reference/scripts/xor_ode.py:137-158  init_xor
  * /root/reference/scripts/parity_ode.py:116-153 make_ds / prep_stim_ode (T,4)
  * /root/reference/scripts/parity_ode.py:156-183 init_network
The tables are used as knot lists with K = T (no compression) so that the oracle does not share the
product's knot-compression logic.
"""
from __future__ import annotations

import torch


def time_vec(time_steps: int, dt: float) -> torch.Tensor:
    return torch.linspace(0.0, time_steps * dt, time_steps)


def wta_stimulus(tv: torch.Tensor, raw) -> torch.Tensor:
    """(T,16): raw[0] on L4e/L4i of column A, raw[1] on column B, during the middle third."""
    T = len(tv)
    vec = torch.zeros(16)
    vec[2] = vec[3] = float(raw[0])
    vec[10] = vec[11] = float(raw[1])
    out = torch.zeros(T, 16)
    on = int(T / 3)
    off = int(on + T / 3)
    out[on:off] = vec
    return out


XOR_CONDITIONS = ((20.0, 0.0), (0.0, 20.0), (20.0, 20.0), (0.0, 0.0))


def xor_stimulus(tv: torch.Tensor, cond) -> torch.Tensor:
    """(T,2,16): second half carries the pattern; channel 1 is the column-swapped copy."""
    T = len(tv)
    vec = torch.zeros(16)
    vec[2] = vec[3] = float(cond[0])
    vec[10] = vec[11] = float(cond[1])
    half = int(T / 2)
    whole = torch.cat((torch.zeros(half, 16), vec.expand(half, 16)), dim=0)
    mirror = torch.cat((whole[:, 8:], whole[:, :8]), dim=1)
    return torch.stack((whole, mirror), dim=1)


PARITY_PATTERNS = ((0., 0., 0., 1.), (0., 0., 1., 1.), (0., 1., 1., 1.), (1., 1., 1., 1.))


def parity_stimulus(tv: torch.Tensor, pattern, amp: float = 15.0) -> torch.Tensor:
    """(T,4): zeros for the first half, then amp * pattern."""
    T = len(tv)
    half = int(T / 2)
    raw = torch.tensor(pattern, dtype=torch.float32) * amp
    return torch.cat((torch.zeros(half, 4), raw.repeat(half, 1)), dim=0)
