"""Read-outs and losses the reference scripts apply to solver output (plain torch on the trajectory).

Mirrors reference src/utils.py:51-88 (min_max, fr_to_binary, huber_loss_wta) and the script-level read-outs of
scripts/xor_ode.py:120-130 and scripts/parity_ode.py:239-249, batched as (T, B, 3N) trajectories.
"""
from __future__ import annotations

import torch

from .model import compute_firing_rate


def min_max(firing_rates: torch.Tensor) -> torch.Tensor:
    lo, hi = firing_rates.min(), firing_rates.max()
    return (firing_rates - lo) / (hi - lo)


def fr_to_binary(firing_rates: torch.Tensor, scaling_factor: float = 1.0) -> torch.Tensor:
    z = (firing_rates - firing_rates.mean()) / (firing_rates.std() / scaling_factor)
    return torch.sigmoid(z)


def huber_loss_wta(pred_states: torch.Tensor, true: torch.Tensor, network) -> torch.Tensor:
    """pred_states (S, T, 1, 48) as the reference stacks them, true (S, T, 2)."""
    rate = compute_firing_rate(pred_states[:, :, 0, :16] - pred_states[:, :, 0, 16:32])
    w = network.output_weights.to(rate.device)
    both = torch.stack(((rate[:, :, :8] * w).sum(2), (rate[:, :, 8:] * w).sum(2)), dim=2)
    return torch.nn.functional.smooth_l1_loss(both, true, beta=1.0)


class _HuberRateLoss(torch.autograd.Function):
    """loss and d loss / d trajectory in one fused pass (include/odecol.h: odecol_huber_rate_loss)."""

    @staticmethod
    def forward(ctx, y_sel, target, P, w, beta):
        from . import _native
        loss, grad = _native.ext().huber_rate_loss(y_sel.detach().contiguous(), target.detach(), int(P),
                                                   None if w is None else w.detach().to(torch.float32).contiguous(), float(beta))
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout, None, None, None, None


def huber_rate_loss(y_sel: torch.Tensor, target: torch.Tensor, pops_per_group: int = 1, weights=None,
                    beta: float = 1.0) -> torch.Tensor:
    """Fused ``smooth_l1_loss(sum_k w_k phi(V - A), target)`` on a trajectory restricted to read-out populations.

    ``y_sel`` (T, B, 2*G*P) as returned by ``odeint(..., components=cat(pops, N + pops))``: V of the G*P read-out
    populations (G groups of P consecutive ones) followed by their A; ``target`` broadcastable to (T, B, G).  Equals
    the reference's huber_loss_wta reduction (src/utils.py:74-88) -- one kernel computes the loss and its gradient
    w.r.t. ``y_sel`` (no gradient flows to ``target`` / ``weights``).  CUDA only."""
    if not y_sel.is_cuda:
        raise RuntimeError("odecol: huber_rate_loss is a fused CUDA read-out (no CPU path)")
    if target.dim() != 3:
        raise ValueError("odecol: target must be 3-d and broadcastable to (T, B, G)")
    return _HuberRateLoss.apply(y_sel, target.to(y_sel.device, torch.float32), pops_per_group, weights, beta)


class _WindowRateL1(torch.autograd.Function):
    """loss, prediction, d loss / d trajectory and d loss / d weights in one fused pass (include/odecol.h:
    odecol_window_rate_l1_loss)."""

    @staticmethod
    def forward(ctx, y_sel, target, last, w):
        from . import _native
        loss, pred, grad, grad_w = _native.ext().window_rate_l1_loss(y_sel.detach().contiguous(), target.detach().contiguous(),
                                                                     int(last), w.detach().to(torch.float32).contiguous())
        ctx.save_for_backward(grad, grad_w)
        ctx.mark_non_differentiable(pred)
        return loss, pred

    @staticmethod
    def backward(ctx, gout, _gpred):
        grad, grad_w = ctx.saved_tensors
        return grad * gout, None, None, grad_w * gout


def window_rate_l1_loss(y_sel: torch.Tensor, target: torch.Tensor, weights=None, last: int = 1):
    """Fused ``mean_b | sum_k w_k mean_{last points} phi(V_k - A_k) - target_b |`` on a trajectory restricted to the P read-out
    populations: ``y_sel`` (T, B, 2*P) as returned by ``odeint(..., components=readout_components(pops, N))``, ``target`` (B,).
    Returns ``(loss, pred)`` with pred (B,) the read-out itself (detached).  The XOR task reads the final point of column C
    (reference scripts/xor_ode.py:120-130: ``last=1``, ``weights=network.ff_source_mask``), the parity task the mean of the
    last 100 points of the output column (scripts/parity_ode.py:239-249: ``last=100``,
    ``weights=network.output_weights / network.output_scale`` -- trained through this read-out, so the gradient flows to
    ``weights`` as well as to ``y_sel``).  One kernel computes all of it.  CUDA only."""
    if not y_sel.is_cuda:
        raise RuntimeError("odecol: window_rate_l1_loss is a fused CUDA read-out (no CPU path)")
    P = y_sel.shape[2] // 2
    w = torch.ones(P, device=y_sel.device) if weights is None else torch.as_tensor(weights, dtype=torch.float32).to(y_sel.device)
    return _WindowRateL1.apply(y_sel, target.to(y_sel.device, torch.float32).reshape(-1), last, w)


def readout_components(pops, N: int) -> torch.Tensor:
    """State components of the read-out populations `pops`: their V, then their A (what the fused read-outs consume)."""
    pops = torch.as_tensor(pops, dtype=torch.int64).flatten()
    return torch.cat((pops, pops + N))


def xor_readout(traj: torch.Tensor, network) -> torch.Tensor:
    """traj (T, B, 72) -> final L2/3e rate of column C per trial (scripts/xor_ode.py:120-125)."""
    rate = compute_firing_rate(traj[-1, :, 16:24] - traj[-1, :, 40:48])
    return (rate * network.ff_source_mask.to(rate.device)).sum(dim=1)


def parity_readout(traj: torch.Tensor, network, last: int = 100) -> torch.Tensor:
    """traj (T, B, 3N) -> weighted mean rate of the output column over the last `last` points
    (scripts/parity_ode.py:239-243)."""
    N = traj.shape[2] // 3
    rate = compute_firing_rate(traj[-last:, :, N - 8:N] - traj[-last:, :, 2 * N - 8:2 * N])
    return (rate.mean(dim=0) * network.output_weights / network.output_scale).sum(dim=-1)
