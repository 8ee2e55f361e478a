"""Gradient of the adaptive dopri5 solve (discretise-then-optimise through the accepted steps)."""
from __future__ import annotations


def dopri5_with_grad(setup, y0, rtol, atol, options, sel_long, sel_i32, stats):
    raise NotImplementedError(
        "odecol: gradients through adaptive dopri5 are not fused yet; train with method='rk4' "
        "(exact discrete adjoint) or call odeint under torch.no_grad() for dopri5 inference")
