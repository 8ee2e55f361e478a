"""Oracle restatement of the Wong-Wang target generator (TEST INFRASTRUCTURE ONLY).

Follows, in float64 numpy and vectorised over samples,
  * /root/reference/src/ww_model.py:56-87    DM.__init__   constants, s = 0.1, I_noise = randn(2) * sigma_noise
  * /root/reference/src/ww_model.py:89-90    DM.f          (270 x - 108) / (1 - exp(-0.154 (270 x - 108)))
  * /root/reference/src/ww_model.py:92-103   DM.update     one exponential-Euler-free explicit step of dt = 1e-3
  * /root/reference/src/ww_model.py:105-111  DM.simulate   int(time / dt) + 1 updates, r recorded after each update
  * /root/reference/src/ww_model.py:113-127  DM.run_sim    pre-stimulus / stimulus / post-stimulus phases of 5 s
  * /root/reference/src/ww_model.py:135-143  DM.reset      resets s, x, r and mu -- NOT I_noise: every sample after the
                                                           first starts from the converged noise current
  * /root/reference/scripts/wta_ode.py:56-93 make_ds_wwp   mu sampling, R[:, ::10][:, :time_steps], float32 cast

PARITY PINNED: ``oracle/make_golden.py --ww-only`` runs the unmodified reference class and stores its output in
tests/golden/ww.npz; ``tests/test_oracle_golden.py`` checks this restatement against it.
"""
from __future__ import annotations

import numpy as np

GAMMA, TAU_S, TAU_AMPA = 0.641, 0.1, 0.002
J_WITHIN, J_BETWEEN, J_EXT, I_0, DT = 0.2609, 0.0497, 5.2e-4, 0.3255, 1e-3
PHASE_SECONDS = 5.0


def steps_per_phase(seconds: float = PHASE_SECONDS) -> int:
    return int(seconds / DT) + 1                                   # ww_model.py:106


def f(x):
    return (270.0 * x - 108) / (1.0 - np.exp(-0.154 * (270.0 * x - 108.0)))


def converged_noise_current(sigma_noise: float = 0.0) -> float:
    """I_noise after a full sample when sigma_noise = 0: the fixed point of I += dt (I_0 - I) / tau_ampa from 0."""
    assert sigma_noise == 0.0
    i = 0.0
    for _ in range(3 * steps_per_phase()):
        i += DT * (I_0 - i) / TAU_AMPA
    return i


def initial_noise_currents(n_samples: int) -> np.ndarray:
    """(n, 2): sample 0 of a fresh DM() starts at I_noise = 0, the following ones inherit the converged current."""
    out = np.full((n_samples, 2), converged_noise_current())
    out[0] = 0.0
    return out


def run_sim(mu: np.ndarray, i_noise0: np.ndarray, sigma_noise: float = 0.0, noise: np.ndarray | None = None) -> np.ndarray:
    """DM.run_sim for a batch: mu (n, 2) = (muA, muB) of the stimulus phase, i_noise0 (n, 2).  Returns R (n, 2, 3 * 5001).
    ``noise`` (steps, n, 2) standard normals, needed only when sigma_noise != 0."""
    mu = np.asarray(mu, dtype=np.float64)
    n = mu.shape[0]
    per = steps_per_phase()
    s = np.ones((n, 2)) * 0.1
    i_noise = np.array(i_noise0, dtype=np.float64).copy()
    dsig = np.sqrt(DT / TAU_AMPA) * sigma_noise
    R = np.zeros((n, 2, 3 * per))
    k = 0
    for phase in range(3):
        m = mu if phase == 1 else np.zeros_like(mu)
        i_ext = J_EXT * m
        for _ in range(per):
            i_rec = np.stack((J_WITHIN * s[:, 0] + -J_BETWEEN * s[:, 1], -J_BETWEEN * s[:, 0] + J_WITHIN * s[:, 1]), axis=1)
            i_noise += DT * (I_0 - i_noise) / TAU_AMPA + (dsig * noise[k] if noise is not None else 0.0)
            x = i_rec + i_ext + i_noise
            r = f(x)
            s += DT * (-s / TAU_S + (1.0 - s) * GAMMA * r)
            R[:, :, k] = r
            k += 1
    return R


def dataset_states(mu: np.ndarray, time_steps: int, i_noise0: np.ndarray | None = None) -> np.ndarray:
    """states of make_ds_wwp: (n, time_steps, 2) float32 = R[:, ::10][:, :time_steps] transposed (wta_ode.py:82-87)."""
    mu = np.asarray(mu, dtype=np.float64)
    if i_noise0 is None:
        i_noise0 = initial_noise_currents(mu.shape[0])
    R = run_sim(mu, i_noise0)[:, :, ::10][:, :, :time_steps]
    return np.ascontiguousarray(R.transpose(0, 2, 1)).astype(np.float32)


def sample_stimuli(n_samples: int) -> np.ndarray:
    """The (muA, muB) pairs make_ds_wwp draws from numpy's GLOBAL legacy generator, consuming it exactly like the
    reference loop does (DM() draws randn(2) once; every update draws randn(2) even though sigma_noise = 0), so that
    np.random.seed(s) gives the reference's stimuli (wta_ode.py:70-81, ww_model.py:82,98)."""
    np.random.randn(2)
    out = np.zeros((n_samples, 2))
    draws = 2 * 3 * steps_per_phase()
    for i in range(n_samples):
        muA = np.random.uniform(15.0, 25.0)
        muB = muA + np.random.uniform(10., 20.)
        mu_vals = [muA, muB]
        np.random.shuffle(mu_vals)
        out[i] = mu_vals
        np.random.randn(draws)
    return out
