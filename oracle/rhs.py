"""Oracle RHS of the column ODE in the unified linear form (TEST INFRASTRUCTURE ONLY).

Restates, with torch CPU ops in the reference's operation order,
  * /root/reference/src/utils.py:13-25   compute_firing_rate  (phi)
  * /root/reference/src/utils.py:27-28   soft_clamp
  * /root/reference/src/utils.py:31-46   torch_interp         (stimulus lookup)
  * /root/reference/src/coupled_columns.py:204-237 / 407-442 / 753-788   forward
  * /root/reference/src/coupled_columns.py:239-249 / 444-454 / 790-800   diffusion
batched over trials: y is (B, 3N), stimulus knots are (B, K, n_in).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .column_model import LinearForm


def phi(x: torch.Tensor) -> torch.Tensor:
    """utils.py:13-25 -- x_nom / (1 - exp(clamp80(-d x_nom))), pole (NaN) at x = 981/48."""
    a, b, d = 48.0, 981.0, 0.0089
    x_nom = a * x - b
    z = -d * x_nom
    z = 80 * torch.tanh(z / 80)          # utils.py:27-28
    return x_nom / (1 - torch.exp(z))


def interp_knots(t: torch.Tensor, knot_t: torch.Tensor, knot_u: torch.Tensor) -> torch.Tensor:
    """utils.py:31-46 with fp = knot_u[b] for every trial b.  t is a 0-d tensor (one time for the batch)
    or (B,) (one time per trial).  Returns (B, n_in)."""
    K = knot_t.shape[0]
    tc = torch.clamp(t, knot_t[0], knot_t[-1])
    idx = torch.searchsorted(knot_t, tc.reshape(-1).contiguous(), right=True).clamp(1, K - 1)
    x0 = knot_t[idx - 1]
    x1 = knot_t[idx]
    if idx.numel() == 1:
        i = int(idx)
        y0 = knot_u[:, i - 1, :]
        y1 = knot_u[:, i, :]
        slope = (y1 - y0) / (x1 - x0)
        return y0 + slope * (tc.reshape(-1) - x0)
    ar = torch.arange(knot_u.shape[0])
    y0 = knot_u[ar, idx - 1, :]
    y1 = knot_u[ar, idx, :]
    slope = (y1 - y0) / (x1 - x0).unsqueeze(-1)
    return y0 + slope * (tc.reshape(-1) - x0).unsqueeze(-1)


class UnifiedColumnODE:
    """func(t, y) / diffusion(t, y) for the oracle solvers, over a batch of trials.

    Parameters are torch tensors (so autograd gives reference-style discretise-then-optimise
    gradients); ``dtype`` float32 reproduces the reference arithmetic, float64 is the "clean" truth
    used to judge which of two fp32 implementations is closer.
    """

    def __init__(self, lf: LinearForm, knot_t, knot_u, dtype=torch.float32, requires_grad: bool = False):
        as_t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype).clone()
        self.dtype = dtype
        self.N = lf.n
        self.W = as_t(lf.W).requires_grad_(requires_grad)
        self.U = as_t(lf.U).requires_grad_(requires_grad)
        self.bias = as_t(lf.bias).requires_grad_(requires_grad)
        self.kappa = as_t(lf.kappa)
        self.sigma = as_t(lf.sigma)
        self.tau_s = torch.tensor(lf.tau_s, dtype=dtype)
        self.tau_m = torch.tensor(lf.tau_m, dtype=dtype)
        self.tau_a = torch.tensor(lf.tau_a, dtype=dtype)
        self.R = torch.tensor(lf.resistance, dtype=dtype)
        self.knot_t = torch.as_tensor(np.asarray(knot_t), dtype=dtype)
        self.knot_u = torch.as_tensor(np.asarray(knot_u), dtype=dtype)
        self.noise_type = "scalar"
        self.sde_type = "ito"
        self.nfe = 0

    def select_trials(self, sl) -> "UnifiedColumnODE":
        other = object.__new__(UnifiedColumnODE)
        other.__dict__.update(self.__dict__)
        other.knot_u = self.knot_u[sl]
        return other

    def __call__(self, t, y):
        return self.forward(t, y)

    def forward(self, t, y):
        self.nfe += 1
        N = self.N
        t = torch.as_tensor(t, dtype=self.dtype)
        V, A, F = y[..., :N], y[..., N:2 * N], y[..., 2 * N:]
        r = phi(V - A)
        s = interp_knots(t, self.knot_t, self.knot_u)                 # (B, n_in)
        ff = s @ self.U.T
        rec = r @ self.W.T
        total = (ff + self.bias + rec) * self.tau_s                    # coupled_columns.py:225
        dV = (-V + total * self.R) / self.tau_m                        # :228-229
        dA = (-A + self.kappa * r) / self.tau_a                        # :230-231
        dF = (-F + r) / self.tau_s                                     # :232-233
        return torch.cat((dV, dA, dF), dim=-1)

    def diffusion(self, t, y):
        return (torch.zeros_like(y) + self.sigma).unsqueeze(-1)       # (B, 3N, 1)
