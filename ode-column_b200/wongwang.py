"""Batched Wong-Wang training-target generator.

Drop-in for the dataset builders of the reference's WTA script:

    make_ds_wwp(ds_file, nr_samples, time_steps)        reference scripts/wta_ode.py:56-93
    get_data(nr_samples, batch_size, time_steps, fn)    reference scripts/wta_ode.py:95-107

The reference runs ``DM.run_sim`` (src/ww_model.py:113-127) once per sample in numpy: 15,003 sequential float64 updates
each, 3,010 samples.  Here every sample is one GPU thread of ``odecol_ww_generate`` (include/odecol.h); the host side only
draws the stimuli -- from numpy's global generator, consuming it exactly like the reference loop does, so that
``np.random.seed(s)`` gives the reference's dataset -- and keeps the reference's pickle cache format.
There is no CPU implementation: a CUDA device is required.
"""
from __future__ import annotations

import os
import pickle
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native

_DT, _TAU_AMPA, _I_0 = 1e-3, 0.002, 0.3255           # reference src/ww_model.py:58-71
_PHASE_SECONDS = 5.0                                 # DM.run_sim: three phases of simulate(5.)


def steps_per_phase() -> int:
    return int(_PHASE_SECONDS / _DT) + 1             # ww_model.py:106


def sample_stimuli(nr_samples: int) -> np.ndarray:
    """(nr_samples, 2) float64 (muA, muB) pairs, drawn as make_ds_wwp draws them (wta_ode.py:70-81).  The reference
    also draws randn(2) when DM() is built and in every update (ww_model.py:82,98; multiplied by sigma_noise = 0); those
    draws are consumed here too so the global stream stays aligned with the reference's."""
    np.random.randn(2)
    out = np.zeros((nr_samples, 2))
    per_sample = 2 * 3 * steps_per_phase()
    for i in range(nr_samples):
        muA = np.random.uniform(15.0, 25.0)
        muB = muA + np.random.uniform(10., 20.)
        mu_vals = [muA, muB]
        np.random.shuffle(mu_vals)
        out[i] = mu_vals
        np.random.randn(per_sample)
    return out


def initial_noise_currents(nr_samples: int) -> np.ndarray:
    """DM.reset() (ww_model.py:135-143) leaves I_noise alone: sample 0 starts from 0, every later sample from the fixed
    point the current reached during the previous one (plain float64 recursion, no transcendental involved)."""
    i = 0.0
    for _ in range(3 * steps_per_phase()):
        i += _DT * (_I_0 - i) / _TAU_AMPA
    out = np.full((nr_samples, 2), i)
    out[0] = 0.0
    return out


def generate_states(mu, time_steps: int, device="cuda", i_noise0=None, sigma_noise: float = 0.0, seed: int = 0,
                    trial_offset: int = 0, every: int = 10) -> torch.Tensor:
    """(B, time_steps, 2) float32 firing rates for stimuli mu (B, 2): one kernel launch, one thread per sample."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("odecol: the Wong-Wang generator runs on a CUDA device (no CPU path)")
    mu_t = torch.as_tensor(np.asarray(mu, dtype=np.float64)).reshape(-1, 2).to(dev).contiguous()
    if i_noise0 is None:
        i_noise0 = initial_noise_currents(mu_t.shape[0])
    i0 = torch.as_tensor(np.asarray(i_noise0, dtype=np.float64)).reshape(-1, 2).to(dev).contiguous()
    return _native.ext().ww_generate(mu_t, i0, steps_per_phase(), int(every), int(time_steps), float(sigma_noise),
                                     int(seed), int(trial_offset))


def make_ds_wwp(ds_file: Optional[str], nr_samples: int, time_steps: int, device="cuda") -> Tuple[torch.Tensor, torch.Tensor]:
    """Dataset of Wong-Wang samples, loaded from ``ds_file`` when it exists (same pickle layout as the reference:
    {'states': (n, time_steps, 2), 'stims': (n, 2)} float32 CPU tensors), generated on the GPU otherwise."""
    if ds_file is not None and os.path.exists(ds_file):
        with open(ds_file, "rb") as f:
            ds = pickle.load(f)
        return ds["states"], ds["stims"]
    mu = sample_stimuli(nr_samples)
    states = generate_states(mu, time_steps, device=device).cpu()
    ds = {"states": states, "stims": torch.tensor(mu, dtype=torch.float32)}
    if ds_file is not None:
        os.makedirs(os.path.dirname(os.path.abspath(ds_file)), exist_ok=True)
        with open(ds_file, "wb") as f:
            pickle.dump(ds, f)
    return ds["states"], ds["stims"]


def get_data(nr_samples: int, batch_size: int, time_steps: int, fn: Optional[str], device="cuda"):
    """DataLoader over (states / 20, stims), shuffled, like the reference (wta_ode.py:95-107; +10 samples, /20 to match
    the L2/3 firing-rate scale)."""
    from torch.utils.data import DataLoader, TensorDataset
    states, stims = make_ds_wwp(fn, nr_samples + 10, time_steps, device=device)
    states = states / 20.
    return DataLoader(TensorDataset(states, stims), batch_size=batch_size, shuffle=True)
