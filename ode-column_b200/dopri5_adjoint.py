"""Gradient of the adaptive dopri5 solve: discretise-then-optimise through the ACCEPTED steps, which is what
``loss.backward()`` through torchdiffeq's unrolled adaptive solver computes (reference scripts/xor_ode.py:114,177;
rejected attempts and the step-size controller run under no_grad there and carry no gradient here)."""
from __future__ import annotations

import torch

_ST_MAXSTEPS = 2


class _Dopri5Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, W_aug, setup, rtol, atol, max_steps, cap, sel_long, sel_i32, stats):
        ext = setup.ext
        prob = setup.problem(W_aug)
        y0c = y0.detach().to(torch.float32).contiguous()
        while True:
            y, na, nr, st, rec_y, rec_t0, rec_dt, out_step, out_x = ext.dopri5_fwd_record(prob, setup.t, y0c, rtol, atol,
                                                                                          max_steps, cap)
            overflow = bool(((st == _ST_MAXSTEPS) & (na >= cap)).any())
            if not overflow or cap >= max_steps:
                break
            cap = min(2 * cap, max_steps)          # a trial needed more accepted steps than the record holds
        if stats is not None:
            stats.update(n_accept=na, n_reject=nr, status=st)
        ctx.setup, ctx.prob, ctx.sel_i32, ctx.T = setup, prob, sel_i32, setup.t.numel()
        ctx.save_for_backward(rec_y, rec_t0, rec_dt, out_step, out_x, na)
        return y if sel_long is None else y.index_select(2, sel_long)

    @staticmethod
    def backward(ctx, grad):
        rec_y, rec_t0, rec_dt, out_step, out_x, na = ctx.saved_tensors
        gy0, gW = ctx.setup.ext.dopri5_bwd(ctx.prob, ctx.T, rec_y, rec_t0, rec_dt, out_step, out_x, na,
                                           grad.to(torch.float32).contiguous(), ctx.sel_i32)
        return (gy0, gW) + (None,) * 8


def dopri5_with_grad(setup, y0, rtol, atol, options, sel_long, sel_i32, stats):
    max_steps = int(options.get("max_num_steps", 4_000_000))
    cap = int(options.get("record_capacity", 4096))
    return _Dopri5Function.apply(y0, setup.lf.W_aug, setup, float(rtol), float(atol), max_steps, cap, sel_long, sel_i32, stats)
