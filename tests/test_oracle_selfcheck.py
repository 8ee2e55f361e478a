"""Self-checks that anchor the restated third-party solvers (oracle/solvers.py).  torchdiffeq / torchsde are not
installable offline, so these stand in for their own test-suites: tableau order conditions, empirical convergence
order, analytic solutions, controller behaviour."""
import math

import numpy as np
import pytest
import torch

from oracle import solvers as S


def test_dopri5_tableau_order_conditions():
    b = np.array(S.DP_C_SOL)
    c = np.array([0.0] + S.DP_ALPHA)
    A = np.zeros((7, 7))
    for i, row in enumerate(S.DP_BETA):
        A[i + 1, :len(row)] = row
    assert abs(b.sum() - 1) < 1e-15
    assert np.allclose(A.sum(1), c, atol=1e-15)                 # row sums
    assert abs(b @ c - 1 / 2) < 1e-15
    assert abs(b @ c ** 2 - 1 / 3) < 1e-15
    assert abs(b @ (A @ c) - 1 / 6) < 1e-15
    assert abs(b @ c ** 3 - 1 / 4) < 1e-15
    assert abs(b @ c ** 4 - 1 / 5) < 1e-14
    e = np.array(S.DP_C_ERR)
    assert abs(e.sum()) < 1e-15                                 # the survey's 50085-vs-50001 typo fails this
    assert abs(e @ c) < 1e-15 and abs(e @ c ** 2) < 1e-15 and abs(e @ c ** 3) < 1e-15
    assert np.allclose(S.DP_BETA[-1] + [0.0], S.DP_C_SOL)       # FSAL
    m = np.array(S.DP_C_MID)
    assert abs(m.sum() - 0.5) < 1e-12                           # y_mid is the solution at t0 + h/2


def _lin(lmbda):
    return lambda t, y: lmbda * y


def test_rk4_38_is_fourth_order():
    y0 = torch.ones(1, 1, dtype=torch.float64)
    errs = []
    for n in (20, 40, 80):
        t = torch.linspace(0, 1, n + 1, dtype=torch.float64)
        y = S.odeint_rk4(_lin(-2.0), y0, t)
        errs.append(abs(float(y[-1]) - math.exp(-2.0)))
    orders = [math.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert all(3.8 < o < 4.3 for o in orders), orders


def test_rk4_outputs_by_linear_interpolation_between_grid_points():
    y0 = torch.ones(1, 1, dtype=torch.float64)
    t = torch.tensor([0.0, 0.25, 1.0], dtype=torch.float64)
    coarse = S.odeint_rk4(_lin(-1.0), y0, t, step_size=0.5)
    grid = S.odeint_rk4(_lin(-1.0), y0, torch.tensor([0.0, 0.5, 1.0], dtype=torch.float64))
    assert torch.allclose(coarse[1], 0.5 * (grid[0] + grid[1]))
    assert torch.allclose(coarse[2], grid[2])


def test_dopri5_meets_tolerance_and_dense_output_is_accurate():
    y0 = torch.tensor([[1.0, 0.0]], dtype=torch.float64)

    def osc(t, y):
        return torch.stack((y[..., 1], -y[..., 0]), dim=-1)

    t = torch.linspace(0, 6.0, 61, dtype=torch.float64)
    st = {}
    y = S.odeint_dopri5(osc, y0, t, rtol=1e-8, atol=1e-10, stats=st)
    exact = torch.stack((torch.cos(t), -torch.sin(t)), dim=-1)
    assert float((y[:, 0] - exact).abs().max()) < 2e-7
    assert st["n_accept"] < 200 and st["n_reject"] < st["n_accept"]


def test_dopri5_step_count_scales_with_tolerance():
    y0 = torch.ones(1, 1, dtype=torch.float64)
    t = torch.tensor([0.0, 5.0], dtype=torch.float64)
    counts = []
    for tol in (1e-4, 1e-6, 1e-8):
        st = {}
        S.odeint_dopri5(_lin(-1.0), y0, t, rtol=tol, atol=tol * 1e-2, stats=st)
        counts.append(st["n_accept"])
    assert counts[0] < counts[1] < counts[2]
    # a 5th-order controller: 100x tighter tolerance -> about 100^(1/5) = 2.5x the steps
    assert 1.5 < counts[2] / counts[1] < 4.0


class _OU:
    noise_type, sde_type = "scalar", "ito"

    def __init__(self, theta, sigma):
        self.theta, self.sigma = theta, sigma

    def forward(self, t, y):
        return -self.theta * y

    def diffusion(self, t, y):
        return torch.full_like(y, self.sigma).unsqueeze(-1)


def test_em_schedule_and_exact_recursion():
    ts = torch.linspace(0, 0.1, 11)
    sched = S.em_step_schedule(ts, 0.01)
    assert len(sched) in (10, 11) and abs(sched[-1][1] - 0.1) < 1e-7
    g = torch.Generator().manual_seed(0)
    dW = torch.randn(len(sched), 3, 1, generator=g) * 0.1
    sde = _OU(2.0, 0.5)
    y0 = torch.ones(3, 1)
    ys = S.sdeint_euler(sde, y0, ts, S.TabulatedBrownian(dW), dt=0.01)
    y = y0.clone()
    for k, (a, b) in enumerate(sched):                          # hand recursion
        y = y + (-2.0 * y) * (b - a) + 0.5 * dW[k]
    assert torch.allclose(ys[-1], y, atol=1e-6)


def test_em_ou_moments():
    torch.manual_seed(0)
    theta, sigma, T = 1.5, 0.8, 1.0
    ts = torch.linspace(0, T, 11)
    sched = S.em_step_schedule(ts, 0.005)
    B = 4000
    dW = torch.randn(len(sched), B, 1) * math.sqrt(0.005)
    ys = S.sdeint_euler(_OU(theta, sigma), torch.ones(B, 1), ts, S.TabulatedBrownian(dW), dt=0.005)
    mean, var = float(ys[-1].mean()), float(ys[-1].var())
    assert abs(mean - math.exp(-theta * T)) < 0.03
    assert abs(var - sigma ** 2 / (2 * theta) * (1 - math.exp(-2 * theta * T))) < 0.02


def test_adaptive_controller_limits():
    # error above 1 shrinks the step by at most 5x, error below 1 never shrinks and grows by at most 1.4x
    s, _ = S.adaptive_update(1e6, 1.0, None)
    assert abs(s - 0.2) < 1e-12
    s, r = S.adaptive_update(1e-9, 1.0, None)
    assert abs(s - 1.4) < 1e-12 and r == 0.9 / 1e-9
    s, _ = S.adaptive_update(0.95, 1.0, None)
    assert s == 1.0
