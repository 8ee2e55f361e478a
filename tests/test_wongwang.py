"""Wong-Wang target generator: oracle vs the reference's own output (golden, CPU) and the
CUDA generator vs both (GPU).  tests/golden/ww.npz was produced by the UNMODIFIED reference class src/ww_model.py::DM
inside the loop of scripts/wta_ode.py::make_ds_wwp (oracle/make_golden.py::add_ww), numpy seed 7."""
import os
import pickle

import numpy as np
import pytest
import torch

import odecol
from oracle import ww

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ww.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def test_oracle_reproduces_the_reference_dataset_bit_for_bit(gold):
    np.random.seed(int(gold["seed"]))
    mu = ww.sample_stimuli(5)
    assert np.array_equal(mu.astype(np.float32), gold["stims"])
    states = ww.dataset_states(mu, 1500)
    assert states.dtype == np.float32 and np.array_equal(states, gold["states"])
    # the first sample starts from I_noise = 0, the others from the converged current: dropping that quirk is visible
    wrong = ww.dataset_states(mu[1:2], 1500, i_noise0=np.zeros((1, 2)))
    assert not np.array_equal(wrong[0], gold["states"][1])
    full = ww.run_sim(mu[:1], np.zeros((1, 2)))[0]
    # in float64 the restatement is within rounding of the reference (np.dot may contract to FMA): < 1e-12 relative
    assert np.abs(full[:, ::25] - gold["first_full"]).max() < 1e-12 * np.abs(gold["first_full"]).max()


def test_host_side_sampling_matches_oracle_and_reference(gold):
    np.random.seed(int(gold["seed"]))
    mu = odecol.wongwang.sample_stimuli(5)
    assert np.array_equal(mu.astype(np.float32), gold["stims"])
    after = np.random.uniform()
    np.random.seed(int(gold["seed"]))
    ww.sample_stimuli(5)
    assert after == np.random.uniform()                       # both leave the global stream in the same place
    assert np.array_equal(odecol.wongwang.initial_noise_currents(4), ww.initial_noise_currents(4))
    assert odecol.wongwang.steps_per_phase() == ww.steps_per_phase() == 5001
    mu_all = odecol.wongwang.sample_stimuli(200)
    lo, hi = mu_all.min(axis=1), mu_all.max(axis=1)
    assert (lo >= 15).all() and (lo <= 25).all() and (hi - lo >= 10).all() and (hi - lo <= 20).all()
    assert 0.3 < float((mu_all[:, 0] > mu_all[:, 1]).mean()) < 0.7   # shuffled


def test_generator_has_no_cpu_path():
    with pytest.raises(RuntimeError, match="CUDA"):
        odecol.wongwang.generate_states(np.array([[20.0, 35.0]]), 100, device="cpu")


@pytest.mark.gpu
def test_cuda_generator_matches_the_reference_dataset(gold, tmp_path):
    np.random.seed(int(gold["seed"]))
    fn = str(tmp_path / "data" / "ds.pkl")
    states, stims = odecol.make_ds_wwp(fn, 5, 1500)
    assert states.shape == (5, 1500, 2) and states.dtype == torch.float32 and not states.is_cuda
    assert np.array_equal(stims.numpy(), gold["stims"])
    ref = torch.tensor(gold["states"])
    err = float((states - ref).abs().max() / ref.abs().max())
    same = float((states == ref).float().mean())
    print(f"\nWong-Wang generator vs reference: max rel err {err:.2e}, bit-identical float32 entries {same:.4f}")
    assert err < 1e-6 and same > 0.99                          # float64 arithmetic in the reference's order; exp() last ulp
    # pickle cache: same layout as the reference's file, second call loads instead of generating
    with open(fn, "rb") as f:
        ds = pickle.load(f)
    assert set(ds) == {"states", "stims"} and torch.equal(ds["states"], states)
    s2, m2 = odecol.make_ds_wwp(fn, 999, 7)
    assert torch.equal(s2, states) and torch.equal(m2, stims)


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 33, 257])
def test_cuda_generator_matches_oracle_on_seeded_stimuli(B):
    rng = np.random.default_rng(B)
    mu = np.stack((rng.uniform(0, 45, B), rng.uniform(0, 45, B)), axis=1)
    mu[0] = (30.0, 30.0)                                       # symmetric input: no winner
    if B > 2:
        mu[1] = (0.0, 0.0)
        mu[2] = (40.0, 15.0)
    i0 = ww.initial_noise_currents(B)
    got = odecol.wongwang.generate_states(mu, 1501, i_noise0=i0).cpu().numpy()
    want = ww.dataset_states(mu, 1501, i_noise0=i0)
    err = np.abs(got - want).max() / np.abs(want).max()
    print(f"\nWong-Wang generator vs oracle, B={B}: max rel err {err:.2e}")
    assert got.shape == (B, 1501, 2) and err < 1e-6
    short = odecol.wongwang.generate_states(mu, 40, i_noise0=i0, every=3).cpu().numpy()
    R = ww.run_sim(mu, i0)[:, :, ::3][:, :, :40].transpose(0, 2, 1)
    assert np.abs(short - R).max() / np.abs(R).max() < 1e-6


@pytest.mark.gpu
def test_cuda_generator_noise_mode_is_reproducible_and_sharding_invariant():
    mu = np.tile(np.array([[18.0, 32.0]]), (64, 1))
    a = odecol.wongwang.generate_states(mu, 1500, sigma_noise=0.02, seed=5)
    b = odecol.wongwang.generate_states(mu, 1500, sigma_noise=0.02, seed=5)
    c = odecol.wongwang.generate_states(mu[:32], 1500, sigma_noise=0.02, seed=5, trial_offset=32,
                                        i_noise0=odecol.wongwang.initial_noise_currents(64)[32:])
    clean = odecol.wongwang.generate_states(mu, 1500)
    assert torch.equal(a, b) and torch.equal(c, a[32:])
    assert not torch.equal(a[3], a[4])                         # trials draw different noise
    # the stronger input (pool B) still wins the competition in (almost) every noisy trial, as in the clean solution
    assert float(clean[1, 990, 1]) > 10 * float(clean[1, 990, 0])
    assert float((a[:, 990, 1] > a[:, 990, 0]).float().mean()) > 0.9
    assert torch.isfinite(a).all()
