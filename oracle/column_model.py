"""Oracle restatement of the reference's model construction (TEST INFRASTRUCTURE ONLY).

Restates, in numpy with the reference's dtypes, how ``config/model.toml`` becomes the
constant tensors of a column area and how the three networks' ``forward`` methods
collapse into ONE linear form

    I(t) = W . r  +  U . s(t)  +  bias            

Follows
  * /root/reference/src/coupled_columns.py:23-37   (scalars: tau_s, tau_m, tau_a, R)
  * /root/reference/src/coupled_columns.py:39-52   (population sizes, kappa tiling)
  * /root/reference/src/coupled_columns.py:54-63   (block-diagonal connection probabilities)
  * /root/reference/src/coupled_columns.py:65-123  (synapse counts K, strengths J, W = K*J, bias, ff gains)
  * /root/reference/src/coupled_columns.py:125-140 (internal / external masks)
  * /root/reference/src/coupled_columns.py:204-249, 371-454, 717-800 (the three forwards / diffusions)

Checked against the imported reference by ``oracle/make_golden.py`` and
``tests/test_oracle_vs_reference.py`` (the latter runs only where /root/reference exists).
"""
from __future__ import annotations

import dataclasses
import tomllib
from typing import Dict, List, Sequence

import numpy as np

POPS_PER_COLUMN = 8  # L2/3e, L2/3i, L4e, L4i, L5e, L5i, L6e, L6i  (coupled_columns.py:104-111)


def load_config(path: str) -> dict:
    """src/utils.py:5-10."""
    with open(path, "rb") as fh:
        return tomllib.load(fh)


@dataclasses.dataclass
class AreaConstants:
    """Constant tensors of one ``ColumnArea`` (coupled_columns.py:8-141), all float32."""

    num_columns: int
    num_populations: int
    recurrent_weights: np.ndarray   # (n, n)  K*J incl. internal mask
    background_weights: np.ndarray  # (n,)
    feedforward_weights: np.ndarray  # (n,)
    adaptation_strength: np.ndarray  # (n,)
    internal_mask: np.ndarray       # (n, n)
    external_mask: np.ndarray       # (n, n)
    background_drive: np.float32
    tau_s: np.float32
    tau_m: np.float32
    tau_a: np.float32
    resistance: np.float32


def area_constants(cfg: dict, area: str, num_columns: int, small_network: bool = False) -> AreaConstants:
    f32 = np.float32
    n = POPS_PER_COLUMN * num_columns

    # coupled_columns.py:43-47 -- population sizes, tiled per column; halved etc. for small nets
    sizes = np.tile(np.asarray(cfg["population_size"][area.lower()], dtype=np.float64), num_columns)
    if small_network:
        sizes = sizes / num_columns

    # coupled_columns.py:125-140
    col_of = np.arange(n) // POPS_PER_COLUMN
    internal = (col_of[:, None] == col_of[None, :]).astype(f32)
    external = (1 - internal).astype(f32)

    # coupled_columns.py:58-63 -- the probabilities go through a float32 tensor first
    p8 = np.asarray(cfg["connection_probabilities"]["internal"], dtype=f32)
    prob = np.kron(np.eye(num_columns, dtype=f32), p8).astype(f32)

    # coupled_columns.py:94-98 -- log(1-p) is evaluated in float32, the rest in float64
    log_num = np.log(f32(1) - prob)                       # float32
    log_den = np.log(1 - 1 / np.outer(sizes, sizes))      # float64
    counts = (log_num / log_den / sizes[:, None]).astype(f32)

    # coupled_columns.py:104-114 -- +J for excitatory sources, -(N_E/N_I) J for inhibitory ones
    base = cfg["synaptic_strength"]["baseline"]
    strength_col = (np.ones(n, dtype=f32) * f32(base)).astype(f32)
    inh_scale = -sizes[0::2] / sizes[1::2]                # float64
    strength_col[1::2] = (inh_scale * base).astype(f32)
    strength = np.tile(strength_col, (n, 1)) * internal

    recurrent = (counts * strength).astype(f32)

    # coupled_columns.py:69-81, 121-123
    if small_network:
        bg_counts = np.full(POPS_PER_COLUMN, 2510, dtype=np.int64)
    else:
        bg_counts = np.asarray(cfg["synapse_counts"]["background"], dtype=np.int64)
    ff_counts = np.asarray(cfg["synapse_counts"]["feedforward"], dtype=np.int64)
    bg_w = (np.tile(bg_counts, num_columns).astype(f32) * f32(base)).astype(f32)
    ff_w = (np.tile(ff_counts, num_columns).astype(f32) * f32(base)).astype(f32)

    tc = cfg["time_constants"]
    return AreaConstants(
        num_columns=num_columns,
        num_populations=n,
        recurrent_weights=recurrent,
        background_weights=bg_w,
        feedforward_weights=ff_w,
        adaptation_strength=np.tile(np.asarray(cfg["adaptation_strength"], dtype=f32), num_columns),
        internal_mask=internal,
        external_mask=external,
        background_drive=f32(cfg["background_drive"]),
        tau_s=f32(tc["synapse"]),
        tau_m=f32(tc["membrane"]),
        tau_a=f32(tc["adaptation"]),
        resistance=f32(tc["membrane"] / cfg["capacitance"]),   # coupled_columns.py:36
    )


@dataclasses.dataclass
class LinearForm:
    """The unified problem a solver integrates.  All arrays float32 (or float64 on request).

    dV = (-V + (W r + U s(t) + bias) * tau_s * R) / tau_m
    dA = (-A + kappa * r) / tau_a
    dF = (-F + r) / tau_s,       r = phi(V - A)
    diffusion g = sigma (constant per state component, one scalar Brownian channel)
    """

    W: np.ndarray       # (N, N)   row = target, col = source
    U: np.ndarray       # (N, n_in)
    bias: np.ndarray    # (N,)
    kappa: np.ndarray   # (N,)
    sigma: np.ndarray   # (3N,)
    tau_s: float
    tau_m: float
    tau_a: float
    resistance: float

    @property
    def n(self) -> int:
        return self.W.shape[0]

    @property
    def n_in(self) -> int:
        return self.U.shape[1]


def wta_linear_form(cfg: dict, recurrent_weights: np.ndarray, area: str = "mt") -> LinearForm:
    """ColumnAreaWTA (coupled_columns.py:143-249): N=16, elementwise input gains, sigma=100 on ALL 48
    components (the ``g[:split_mem]`` slice at :247 indexes dim 0 of a (1,48) tensor)."""
    c = area_constants(cfg, area, 2, small_network=True)
    n = c.num_populations
    return LinearForm(
        W=np.asarray(recurrent_weights, dtype=np.float32),
        U=np.diag(c.feedforward_weights).astype(np.float32),          # :221 elementwise gain
        bias=(c.background_weights * c.background_drive).astype(np.float32),  # :222
        kappa=c.adaptation_strength,
        sigma=np.full(3 * n, 100.0, dtype=np.float32),
        tau_s=float(c.tau_s), tau_m=float(c.tau_m), tau_a=float(c.tau_a), resistance=float(c.resistance),
    )


def xor_linear_form(cfg: dict, ff_target_weights: Sequence[Sequence[np.ndarray]],
                    nr_columns_per_area: Sequence[int] = (2, 1), areas: Sequence[str] = ("mt", "mt"),
                    nr_input_units: int = 2) -> LinearForm:
    """ColumnNetworkXOR (coupled_columns.py:254-454).

    ``ff_target_weights[a][i]`` mirrors ``feedforward_target_weights.{a}.{i}``.  Area 0 receives
    ``sum_i u[i] (.) w0_i`` (elementwise, :387-388) -> U is (N, nr_input_units * n0) with the gain of
    channel (i, p) on row p.  Area a>0 receives ``10 * r[L2/3e of column i of area a-1] * w_i`` (:390-395).
    sigma = 10 on the V block only (:449-452).
    """
    consts = [area_constants(cfg, areas[a], nr_columns_per_area[a], small_network=True)
              for a in range(len(areas))]
    sizes = [c.num_populations for c in consts]
    n = int(sum(sizes))
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
    n0 = sizes[0]
    W = np.zeros((n, n), dtype=np.float32)
    U = np.zeros((n, nr_input_units * n0), dtype=np.float32)
    bias = np.zeros(n, dtype=np.float32)
    for a, c in enumerate(consts):
        sl = slice(offs[a], offs[a + 1])
        W[sl, sl] = c.recurrent_weights * c.internal_mask                      # :299
        bias[sl] = c.background_weights * c.background_drive                     # :398
        if a == 0:
            for i in range(nr_input_units):
                w = np.asarray(ff_target_weights[0][i], dtype=np.float32)
                U[np.arange(n0), i * n0 + np.arange(n0)] = w
        else:
            for i, w in enumerate(ff_target_weights[a]):
                src = offs[a - 1] + POPS_PER_COLUMN * i                          # L2/3e of column i (:307,392)
                W[sl, src] += np.float32(10.0) * np.asarray(w, dtype=np.float32)  # :394
    # network_as_area (full-size, :271) supplies kappa and the scalars
    whole = area_constants(cfg, "mt", int(sum(nr_columns_per_area)))
    sigma = np.zeros(3 * n, dtype=np.float32)
    sigma[:n] = 10.0
    return LinearForm(W=W, U=U, bias=bias, kappa=whole.adaptation_strength, sigma=sigma,
                      tau_s=float(whole.tau_s), tau_m=float(whole.tau_m), tau_a=float(whole.tau_a),
                      resistance=float(whole.resistance))


def parity_linear_form(cfg: dict, lateral_weights: Sequence[np.ndarray], feedforward_weights: Dict[int, np.ndarray],
                       input_weights: np.ndarray, nr_columns_per_area: Sequence[int] = (8, 4, 1),
                       areas: Sequence[str] = ("mt", "mt", "mt")) -> LinearForm:
    """ColumnNetwork (coupled_columns.py:458-800): block-diagonal (inner + lateral) plus sub-diagonal
    feedforward blocks (:731-749); input = input_weights @ u(t) into area 0; sigma = 10 on ALL components
    (``g[:split, :]`` at :798 slices dim 0 of a (1, 3N) tensor)."""
    consts = [area_constants(cfg, areas[a], nr_columns_per_area[a]) for a in range(len(areas))]
    sizes = [c.num_populations for c in consts]
    n = int(sum(sizes))
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
    n_in = np.asarray(input_weights).shape[1]
    W = np.zeros((n, n), dtype=np.float32)
    U = np.zeros((n, n_in), dtype=np.float32)
    bias = np.zeros(n, dtype=np.float32)
    for a, c in enumerate(consts):
        sl = slice(offs[a], offs[a + 1])
        W[sl, sl] = c.recurrent_weights * c.internal_mask + np.asarray(lateral_weights[a], dtype=np.float32)  # :739-740
        bias[sl] = c.background_weights * c.background_drive                                                   # :743
        if a == 0:
            U[sl, :] = np.asarray(input_weights, dtype=np.float32)                                             # :731
        else:
            W[sl, offs[a - 1]:offs[a]] = np.asarray(feedforward_weights[a], dtype=np.float32)                  # :736
    whole = area_constants(cfg, "mt", int(sum(nr_columns_per_area)))
    return LinearForm(W=W, U=U, bias=bias, kappa=whole.adaptation_strength,
                      sigma=np.full(3 * n, 10.0, dtype=np.float32),
                      tau_s=float(whole.tau_s), tau_m=float(whole.tau_m), tau_a=float(whole.tau_a),
                      resistance=float(whole.resistance))


def synthetic_linear_form(cfg: dict, num_columns: int, seed: int = 0, area: str = "mt",
                          lateral_mean: float = 0.1, lateral_std: float = 0.01,
                          sigma_v: float = 10.0) -> LinearForm:
    """block-diagonal of the full-size ``mt`` 8x8 internal block plus
    dense lateral inhibition ``-|N(mean, std)|`` at the TOML lateral-mask positions between ALL column pairs;
    one input channel per column feeding L4e/L4i with the TOML input gains.  Built directly (no O(N^2) Python
    loops) so that N = 8192 is cheap."""
    one = area_constants(cfg, area, 1)
    n = POPS_PER_COLUMN * num_columns
    rng = np.random.default_rng(seed)
    W = np.kron(np.eye(num_columns, dtype=np.float32), one.recurrent_weights).astype(np.float32)
    lat_mask8 = np.asarray(cfg["connection_masks"]["lateral"], dtype=np.float32)
    lat_mask = np.kron(1 - np.eye(num_columns, dtype=np.float32), lat_mask8).astype(np.float32)
    lat = -np.abs(rng.normal(lateral_mean, lateral_std, size=(n, n))).astype(np.float32)
    W = (W + lat * lat_mask).astype(np.float32)
    gains = np.asarray(cfg["connection_inits"]["input"], dtype=np.float32)[:, 0]   # [0,0,25.9,16.3,0,...]
    U = np.kron(np.eye(num_columns, dtype=np.float32), gains[:, None]).astype(np.float32)  # (N, num_columns)
    bias = np.tile(one.background_weights * one.background_drive, num_columns).astype(np.float32)
    sigma = np.zeros(3 * n, dtype=np.float32)
    sigma[:n] = sigma_v
    return LinearForm(W=W, U=U, bias=bias, kappa=np.tile(one.adaptation_strength, num_columns),
                      sigma=sigma, tau_s=float(one.tau_s), tau_m=float(one.tau_m), tau_a=float(one.tau_a),
                      resistance=float(one.resistance))
