"""Drop-in column-network modules (host side of the boundary).

Same class names, constructor signatures, attribute and trainable-parameter names as the reference's
``src/coupled_columns.py`` (ColumnArea :8-141, ColumnAreaWTA :143-249, ColumnNetworkXOR :254-454,
ColumnNetwork :458-800) and the same ``config/model.toml`` drives them -- but every network also exposes

    export_linear_form()  ->  LinearForm(W_aug, kappa, sigma, n_in, taus)   (differentiable in the parameters)
    stimulus_channels()   ->  (B or 1, T, n_in) table of the stimulus the solver interpolates

which is what the fused CUDA solvers in ``solvers.py`` integrate.  ``forward(t, y)`` / ``diffusion(t, y)`` keep the
reference signatures (so the modules still plug into torchdiffeq / torchsde) and are written against the same
linear form with ordinary torch ops; they accept a batch of trials, y of shape (B, 3N).

Random initialisation draws from torch's global generator in the same order and with the same shapes as the
reference, so ``torch.manual_seed(s)`` reproduces the reference's initial parameters.
"""
from __future__ import annotations

import dataclasses
import tomllib
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

POPS = 8  # populations per column: L2/3e, L2/3i, L4e, L4i, L5e, L5i, L6e, L6i


def load_config(filepath: str) -> dict:
    """reference src/utils.py:5-10"""
    with open(filepath, "rb") as fh:
        return tomllib.load(fh)


# ------------------------------------------------------------------------------------------------------------------
# math helpers with the reference's names (src/utils.py:13-46)
# ------------------------------------------------------------------------------------------------------------------
def soft_clamp(x: torch.Tensor, max_val: float = 80) -> torch.Tensor:
    return max_val * torch.tanh(x / max_val)


def compute_firing_rate(x: torch.Tensor) -> torch.Tensor:
    gain, threshold, noise = 48.0, 981.0, 0.0089
    drive = gain * x - threshold
    return drive / (1 - torch.exp(soft_clamp(-noise * drive)))


def torch_interp(x: torch.Tensor, xp: torch.Tensor, fp: torch.Tensor) -> torch.Tensor:
    """Piecewise-linear lookup of ``fp`` (leading dim = len(xp)) at the scalar time x, ends held."""
    x = torch.clamp(x, xp[0], xp[-1])
    hi = torch.clamp(torch.searchsorted(xp, x, right=True), 1, len(xp) - 1)
    lo = hi - 1
    frac_num = x - xp[lo]
    slope = (fp[hi] - fp[lo]) / (xp[hi] - xp[lo])
    return fp[lo] + slope * frac_num


@dataclasses.dataclass
class LinearForm:
    """I = W_aug . [r ; s(t) ; 1];  W_aug = [W | U | bias | 0-pad] with ld_w a multiple of 4."""

    W_aug: torch.Tensor      # (N, ld_w) differentiable w.r.t. the module parameters
    kappa: torch.Tensor      # (N,)
    sigma: torch.Tensor      # (3N,)
    n_in: int
    tau_s: float
    tau_m: float
    tau_a: float
    resistance: float

    @property
    def N(self) -> int:
        return self.W_aug.shape[0]


def pack_w_aug(W: torch.Tensor, U: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    N, n_in = W.shape[0], U.shape[1]
    ld = (N + n_in + 1 + 3) // 4 * 4
    pad = W.new_zeros(N, ld - (N + n_in + 1))
    return torch.cat((W, U, bias.reshape(N, 1), pad), dim=1)


# ------------------------------------------------------------------------------------------------------------------
# one cortical area: constants from the TOML (reference :8-141)
# ------------------------------------------------------------------------------------------------------------------
class ColumnArea(nn.Module):
    def __init__(self, column_parameters: dict, area: str, num_columns: int, small_network: bool = False):
        super().__init__()
        self.num_columns = num_columns
        self.area = area.lower()
        self.num_populations = POPS * num_columns
        cp = column_parameters
        f32 = torch.float32

        tc = cp["time_constants"]
        self.time_constants = tc
        self.register_buffer("background_drive", torch.tensor(cp["background_drive"], dtype=f32))
        self.register_buffer("adaptation_strength",
                             torch.tensor(cp["adaptation_strength"], dtype=f32).repeat(num_columns))
        self.register_buffer("synapse_time_constant", torch.tensor(tc["synapse"], dtype=f32))
        self.register_buffer("membrane_time_constant", torch.tensor(tc["membrane"], dtype=f32))
        self.register_buffer("adapt_time_constant", torch.tensor(tc["adaptation"], dtype=f32))
        self.register_buffer("resistance", torch.tensor(tc["membrane"] / cp["capacitance"], dtype=f32))

        sizes = np.tile(np.asarray(cp["population_size"][self.area], dtype=np.float64), num_columns)
        if small_network:
            sizes = sizes / num_columns
        self.population_sizes = sizes

        col = torch.arange(self.num_populations) // POPS
        self.internal_mask = (col[:, None] == col[None, :]).to(f32)
        self.external_mask = 1 - self.internal_mask

        p8 = torch.tensor(cp["connection_probabilities"]["internal"], dtype=f32)
        self.internal_connection_probabilities = p8
        self.connection_probabilities = torch.block_diag(*([p8] * num_columns)).numpy()

        bg = [2510] * POPS if small_network else cp["synapse_counts"]["background"]
        self.background_synapse_counts = torch.tensor(bg).repeat(num_columns)
        self.feedforward_synapse_counts = torch.tensor(cp["synapse_counts"]["feedforward"]).repeat(num_columns)
        self.baseline_synaptic_strength = cp["synaptic_strength"]["baseline"]

        # K_ij = log(1 - p_ij) / log(1 - 1/(N_i N_j)) / N_i ; numerator in float32 like the reference (:94-98)
        num = np.log(1 - self.connection_probabilities)
        den = np.log(1 - 1 / np.outer(sizes, sizes))
        self.recurrent_synapse_counts = torch.tensor(num / den / sizes[:, None], dtype=f32)

        j_col = torch.full((self.num_populations,), self.baseline_synaptic_strength, dtype=f32)
        j_col[1::2] = torch.tensor(-sizes[0::2] / sizes[1::2] * self.baseline_synaptic_strength).to(f32)
        self.recurrent_synaptic_strength = j_col.repeat(self.num_populations, 1) * self.internal_mask

        self.recurrent_weights = self.recurrent_synapse_counts * self.recurrent_synaptic_strength
        self.register_buffer("background_weights", self.background_synapse_counts * self.baseline_synaptic_strength)
        self.feedforward_weights = self.feedforward_synapse_counts * self.baseline_synaptic_strength

    def scalars(self) -> Dict[str, float]:
        return dict(tau_s=float(self.synapse_time_constant), tau_m=float(self.membrane_time_constant),
                    tau_a=float(self.adapt_time_constant), resistance=float(self.resistance))


# ------------------------------------------------------------------------------------------------------------------
# shared behaviour of the three networks
# ------------------------------------------------------------------------------------------------------------------
class _LinearFormNetwork(nn.Module):
    noise_type = "scalar"
    sde_type = "ito"

    def set_time_vec(self, time_vec):
        self.time_vec = time_vec

    def set_stim(self, stim):
        self.stim = stim

    # --- to be provided by subclasses -------------------------------------------------------------------------
    def _weights(self):  # -> W (N,N), U (N,n_in), bias (N,), kappa (N,), sigma (3N,), scalars dict
        raise NotImplementedError

    def _channels(self, stim: torch.Tensor) -> torch.Tensor:  # stim (..., T, *) -> (B or 1, T, n_in)
        raise NotImplementedError

    # --- public ---------------------------------------------------------------------------------------------------
    def export_linear_form(self) -> LinearForm:
        W, U, bias, kappa, sigma, sc = self._weights()
        return LinearForm(W_aug=pack_w_aug(W, U, bias), kappa=kappa, sigma=sigma, n_in=U.shape[1], **sc)

    def stimulus_channels(self) -> torch.Tensor:
        return self._channels(self.stim)

    def forward(self, t, state):
        """f(t, y) with the reference signature; y is (B, 3N) (the reference uses B = 1)."""
        W, U, bias, kappa, _, sc = self._weights()
        N = W.shape[0]
        V, A, F = state[..., :N], state[..., N:2 * N], state[..., 2 * N:]
        rate = compute_firing_rate(V - A)
        table = self._channels(self.stim)                                   # (B|1, T, n_in)
        s = torch_interp(torch.as_tensor(t, dtype=state.dtype, device=state.device),
                         self.time_vec.to(state.device), table.transpose(0, 1))   # (B|1, n_in)
        total = (s @ U.T + bias + rate @ W.T) * sc["tau_s"]
        dV = (-V + total * sc["resistance"]) / sc["tau_m"]
        dA = (-A + kappa * rate) / sc["tau_a"]
        dF = (-F + rate) / sc["tau_s"]
        return torch.cat((dV, dA, dF), dim=-1)

    def diffusion(self, t, y):
        sigma = self._weights()[4].to(y.device)
        return (torch.zeros_like(y) + sigma).unsqueeze(-1)


class ColumnAreaWTA(ColumnArea, _LinearFormNetwork):
    """Two columns with trainable lateral inhibition / self excitation (reference :143-249)."""

    def __init__(self, column_parameters: dict, area: str):
        ColumnArea.__init__(self, column_parameters, area, 2, small_network=True)
        n = self.num_populations
        mask = torch.zeros(n, n)
        mask[1, 8] = mask[9, 0] = 1.0      # lateral inhibition L2/3e -> L2/3i of the other column
        mask[0, 0] = mask[8, 8] = 1.0      # L2/3e self excitation
        self.lat_in_mask = mask
        base = self.recurrent_weights.clone().detach()
        noisy = torch.normal(mean=base, std=0.0001).abs()
        self.recurrent_weights = nn.Parameter(noisy * (mask * self.external_mask) + base, requires_grad=True)
        self.output_weights = torch.tensor([1.0, 0, 0, 0, 0, 0, 0, 0])

    def _weights(self):
        dev = self.recurrent_weights.device
        U = torch.diag(self.feedforward_weights.to(dev))                                 # elementwise gains (:221)
        bias = self.background_weights * self.background_drive                          # (:222)
        sigma = torch.full((3 * self.num_populations,), 100.0, device=dev)              # (:244-247) hits every component
        return self.recurrent_weights, U, bias, self.adaptation_strength, sigma, self.scalars()

    def _channels(self, stim):
        return stim.reshape(-1, stim.shape[-2], stim.shape[-1])                          # (T,16) -> (1,T,16)


class ColumnNetworkXOR(_LinearFormNetwork):
    """2+1 columns, trainable feedforward target weights (reference :254-454)."""

    def __init__(self, column_parameters: dict, network_dict: dict):
        super().__init__()
        nd = network_dict
        self.areas = nn.ModuleDict({
            str(a): ColumnArea(column_parameters, nd["areas"][a], nd["nr_columns_per_area"][a], small_network=True)
            for a in range(nd["nr_areas"])})
        self.network_as_area = ColumnArea(column_parameters, "mt", sum(nd["nr_columns_per_area"]))
        self.nr_input_units = nd["nr_input_units"]
        self.nr_columns_per_area = nd["nr_columns_per_area"]
        for area in self.areas.values():
            area.recurrent_weights = area.recurrent_weights * area.internal_mask
        self.ff_source_mask = torch.tensor([1., 0, 0, 0, 0, 0, 0, 0])      # from L2/3e
        self.ff_target_mask = torch.tensor([0., 0, 1, 1, 0, 0, 0, 0])      # into L4e, L4i
        weights = nn.ModuleDict()
        for key, area in self.areas.items():
            fan_in = self.nr_input_units if key == "0" else self.areas[str(int(key) - 1)].num_columns
            base = area.feedforward_weights.clone().detach()
            target = self.ff_target_mask.repeat(area.num_columns)
            weights[key] = nn.ParameterList(
                [nn.Parameter(torch.normal(mean=base, std=0.1).abs() * target, requires_grad=True) for _ in range(fan_in)])
        self.feedforward_target_weights = weights

    def partition_firing_rates(self, firing_rate):
        out, at = {}, 0
        for key, area in self.areas.items():
            out[key] = firing_rate[at:at + area.num_populations].reshape(area.num_columns, POPS)
            at += area.num_populations
        return out

    def _weights(self):
        sizes = [a.num_populations for a in self.areas.values()]
        n, n0 = sum(sizes), sizes[0]
        dev = self.feedforward_target_weights["0"][0].device
        blocks, bias, at = [], [], 0
        for key, area in self.areas.items():
            rows = torch.zeros(area.num_populations, n, device=dev)
            rows[:, at:at + area.num_populations] = area.recurrent_weights.to(dev)
            if key != "0":
                prev_at = at - sizes[int(key) - 1]
                cols = []
                for i, w in enumerate(self.feedforward_target_weights[key]):
                    cols.append((prev_at + POPS * i, 10.0 * w))                            # x10 "pump" (:394)
                onehots = torch.zeros(len(cols), n, device=dev)
                for q, (c, _) in enumerate(cols):
                    onehots[q, c] = 1.0
                rows = rows + torch.stack([w for _, w in cols], dim=1) @ onehots
            blocks.append(rows)
            bias.append(area.background_weights * area.background_drive)
            at += area.num_populations
        W = torch.cat(blocks, dim=0)
        # area 0: sum_i u[i] (.) w_i  ->  U[p, i*n0 + p] = w_i[p]
        U = torch.zeros(n, self.nr_input_units * n0, device=dev)
        top = torch.cat([torch.diag(w) for w in self.feedforward_target_weights["0"]], dim=1)
        U = torch.cat((top, U[n0:]), dim=0)
        sigma = torch.zeros(3 * n, device=dev)
        sigma[:n] = 10.0                                                                    # V only (:449-452)
        return W, U, torch.cat(bias).to(dev), self.network_as_area.adaptation_strength, sigma, self.network_as_area.scalars()

    def _channels(self, stim):
        # (T, n_units, n0) or (B, T, n_units, n0) -> (B|1, T, n_units*n0)
        if stim.dim() == 3:
            stim = stim.unsqueeze(0)
        return stim.reshape(stim.shape[0], stim.shape[1], -1)


class ColumnNetwork(_LinearFormNetwork):
    """Areas of columns with trainable lateral / feedforward / input / output weights (reference :458-800)."""

    def __init__(self, model_parameters: dict, network_dict: dict, device):
        super().__init__()
        nd, mp = network_dict, model_parameters
        self.device = device
        self.areas = nn.ModuleDict({
            str(a): ColumnArea(mp, nd["areas"][a], nd["nr_columns_per_area"][a]).to(device) for a in range(nd["nr_areas"])})
        self.network_as_area = ColumnArea(mp, "mt", sum(nd["nr_columns_per_area"]))
        self.nr_input_units = nd["nr_input_units"]
        self.nr_columns_per_area = nd["nr_columns_per_area"]
        self.nr_areas = nd["nr_areas"]
        masks, inits = mp["connection_masks"], mp["connection_inits"]
        self.input_mask = torch.tensor(masks["input"])
        self.output_mask = torch.tensor(masks["output"])
        self.feedforward_mask = torch.tensor(masks["feedforward"])
        self.lateral_mask = torch.tensor(masks["lateral"])
        self.lateral_scale = self.feedforward_scale = self.output_scale = 1.0

        # lateral (drawn first, like the reference's constructor order :481-485)
        for area in self.areas.values():
            c = area.num_columns
            area.inner_weights = (area.recurrent_weights * area.internal_mask).to(device)
            area.lateral_mask = self.lateral_mask.repeat(c, c) * area.external_mask
            init = torch.tensor(inits["lateral"]).repeat(c, c)
            w = torch.normal(mean=init, std=0.01) * self.lateral_scale * 0.01
            w = (w * area.lateral_mask * area.external_mask).to(device)
            area.lateral_weights = nn.Parameter(w, requires_grad=c > 1)
        # feedforward between consecutive areas, stored on the target area
        for key, area in self.areas.items():
            if key == "0":
                continue
            src, tgt = self.nr_columns_per_area[int(key) - 1], self.nr_columns_per_area[int(key)]
            init = torch.tensor(inits["feedforward"]).repeat(tgt, src)
            w = torch.normal(mean=init, std=1.0).abs() * self.feedforward_scale * 4.0
            mask = self.feedforward_mask.repeat(tgt, src)
            if tgt > 1:
                mask = self.make_mask_fan_in(mask, 2, 2)
            area.feedforward_mask = mask
            area.feedforward_weights = nn.Parameter(w * mask, requires_grad=True)
        # external input into area 0
        first = self.areas["0"]
        init = torch.tensor(inits["input"]).repeat(first.num_columns, self.nr_input_units)
        w = torch.normal(mean=init, std=3.0).abs() * self.feedforward_scale * 0.8
        mask = self.make_mask_fan_in(self.input_mask.repeat(first.num_columns, self.nr_input_units), 2, 2)
        mask[0:16, :] = mask[32:48, :]
        mask[32:48, :] = mask[16:32, :]
        first.input_mask = mask
        first.input_weights = nn.Parameter(w * mask, requires_grad=True)
        # read-out weights of the last area (applied outside the ODE, scripts/parity_ode.py:243)
        last = self.areas[str(self.nr_areas - 1)]
        init = torch.tensor(inits["output"]).repeat(last.num_columns)
        w = torch.normal(mean=init, std=0.001).abs()
        w = w * (w * self.output_mask.repeat(last.num_columns)) * self.output_scale
        self.output_weights = nn.Parameter(w, requires_grad=True)

    @staticmethod
    def make_mask_fan_in(mask, num_target_blocks, num_source_blocks):
        rows, cols = mask.shape
        fr, fc = rows // num_target_blocks, cols // num_source_blocks
        keep = torch.zeros_like(mask)
        for q in range(min(num_target_blocks, num_source_blocks)):
            keep[q * fr:(q + 1) * fr, q * fc:(q + 1) * fc] = 1.0
        return mask * keep

    def partition_firing_rates(self, firing_rate):
        out, at = {}, 0
        for key, area in self.areas.items():
            out[key] = firing_rate[at:at + area.num_populations]
            at += area.num_populations
        return out

    def _weights(self):
        sizes = [a.num_populations for a in self.areas.values()]
        n = sum(sizes)
        dev = self.areas["0"].input_weights.device
        rows, bias, at = [], [], 0
        for key, area in self.areas.items():
            m = area.num_populations
            local = area.inner_weights.to(dev) + area.lateral_weights / self.lateral_scale
            parts = []
            if key == "0":
                parts = [local, torch.zeros(m, n - m, device=dev)]
            else:
                prev = sizes[int(key) - 1]
                parts = [torch.zeros(m, at - prev, device=dev), area.feedforward_weights / self.feedforward_scale, local,
                         torch.zeros(m, n - at - m, device=dev)]
            rows.append(torch.cat(parts, dim=1))
            bias.append(area.background_weights * area.background_drive)
            at += m
        W = torch.cat(rows, dim=0)
        U = torch.cat((self.areas["0"].input_weights / self.feedforward_scale,
                       torch.zeros(n - sizes[0], self.nr_input_units, device=dev)), dim=0)
        sigma = torch.full((3 * n,), 10.0, device=dev)                                      # (:795-798) every component
        return W, U, torch.cat(bias).to(dev), self.network_as_area.adaptation_strength.to(dev), sigma, self.network_as_area.scalars()

    def _channels(self, stim):
        return stim.reshape(-1, stim.shape[-2], stim.shape[-1])                              # (T,4) -> (1,T,4)


def move_to(network, device):
    """``network.to(device)`` plus the plain-tensor attributes: like the reference, the drop-in modules keep masks,
    constants, ``stim`` and ``time_vec`` as ordinary attributes (not buffers), which ``Module.to`` does not move."""
    network = network.to(device)
    for m in [network] + list(network.modules()):
        for k, v in list(vars(m).items()):
            if torch.is_tensor(v) and not isinstance(v, torch.nn.Parameter):
                setattr(m, k, v.to(device))
    return network
