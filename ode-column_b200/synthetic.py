"""Synthetic large networks of BASELINE.json configs 4 and 5: ``num_columns`` copies of the
full-size ``mt`` column on the block diagonal plus dense lateral inhibition between all column pairs at the TOML's
lateral-mask positions; one stimulus channel per column into L4e/L4i.  Assembled with vectorised ``kron`` products on
the host (no O(N^2) Python loops, reference src/coupled_columns.py:125-140 would need them) and moved to ``device``,
with W and U as trainable parameters."""
from __future__ import annotations

import torch
import torch.nn as nn

from .model import ColumnArea, LinearForm, _LinearFormNetwork, pack_w_aug, POPS


class SyntheticColumnSheet(_LinearFormNetwork):
    def __init__(self, column_parameters: dict, num_columns: int, area: str = "mt", seed: int = 0,
                 lateral_mean: float = 0.1, lateral_std: float = 0.01, sigma_v: float = 10.0, device="cpu"):
        super().__init__()
        one = ColumnArea(column_parameters, area, 1)
        n = POPS * num_columns
        self.num_columns, self.num_populations = num_columns, n
        gen = torch.Generator(device="cpu").manual_seed(seed)
        eye = torch.eye(num_columns)
        W = torch.kron(eye, one.recurrent_weights)
        lat8 = torch.tensor(column_parameters["connection_masks"]["lateral"])
        self.lateral_mask = torch.kron(1 - eye, lat8).to(device)
        # lateral weights are drawn per column pair and broadcast over the 8x8 block to keep N = 8192 cheap
        pair = -(torch.randn(num_columns, num_columns, generator=gen) * lateral_std + lateral_mean).abs()
        W = W + torch.kron(pair, torch.ones(POPS, POPS)) * torch.kron(1 - eye, lat8)
        gains = torch.tensor(column_parameters["connection_inits"]["input"])[:, 0]
        self.recurrent_weights = nn.Parameter(W.to(device))
        self.input_weights = nn.Parameter(torch.kron(eye, gains[:, None]).to(device))
        self.register_buffer("bias", (one.background_weights * one.background_drive).repeat(num_columns).to(device))
        self.register_buffer("kappa", one.adaptation_strength.repeat(num_columns).to(device))
        sigma = torch.zeros(3 * n)
        sigma[:n] = sigma_v
        self.register_buffer("sigma", sigma.to(device))
        self._scalars = one.scalars()
        self.output_weights = torch.tensor([1.0, 0, 0, 0, 0, 0, 0, 0])

    def _weights(self):
        return self.recurrent_weights, self.input_weights, self.bias, self.kappa, self.sigma, self._scalars

    def _channels(self, stim):
        return stim if stim.dim() == 3 else stim.unsqueeze(0)

    def lateral_split(self):
        """(W_lateral (N, N), W_local (N, 8)): the between-column part of the recurrent weights (what a sweep's global
        lateral gain multiplies, BASELINE.json configs[4]) and the within-column 8 x 8 blocks, row by row."""
        W = self.recurrent_weights
        c = self.num_columns
        blocks = W.reshape(c, POPS, c, POPS)
        idx = torch.arange(c, device=W.device)
        local = blocks[idx, :, idx, :].reshape(c * POPS, POPS)            # row i -> its own column's 8 sources
        eye = torch.eye(c, device=W.device, dtype=W.dtype)
        off = (1 - eye)[:, None, :, None].expand(c, POPS, c, POPS).reshape(c * POPS, c * POPS)
        return W * off, local

    def set_knots(self, knot_t: torch.Tensor, knot_u: torch.Tensor):
        """Stimulus given directly as knots: time_vec = knot_t (K,), stim = knot_u (B, K, n_in)."""
        self.time_vec, self.stim = knot_t, knot_u
