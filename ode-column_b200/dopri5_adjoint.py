"""Gradient of the adaptive dopri5 solve: discretise-then-optimise through the ACCEPTED steps, which is what
``loss.backward()`` through torchdiffeq's unrolled adaptive solver computes (reference scripts/xor_ode.py:114,177;
rejected attempts and the step-size controller run under no_grad there and carry no gradient here)."""
from __future__ import annotations

import torch

_ST_MAXSTEPS = 2


class _Dopri5Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, W_aug, setup, rtol, atol, max_steps, cap, sel_long, sel_i32, stats, check_status=True):
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
        if check_status and bool((st != 0).any()):
            from .solvers import _check_status
            _check_status(st, {}, "dopri5")
        ctx.setup, ctx.prob, ctx.sel_i32, ctx.T = setup, prob, sel_i32, setup.t.numel()
        ctx.save_for_backward(rec_y, rec_t0, rec_dt, out_step, out_x, na)
        return y if sel_long is None else y.index_select(2, sel_long)

    @staticmethod
    def backward(ctx, grad):
        rec_y, rec_t0, rec_dt, out_step, out_x, na = ctx.saved_tensors
        gy0, gW = ctx.setup.ext.dopri5_bwd(ctx.prob, ctx.T, rec_y, rec_t0, rec_dt, out_step, out_x, na,
                                           grad.to(torch.float32).contiguous(), ctx.sel_i32)
        return (gy0, gW) + (None,) * 9


_SMALL_RECORD_BYTES = 256 << 20


def _record_capacity(setup, y0, rtol, atol, max_steps, options):
    """Accepted steps the record must hold.  ``options['record_capacity']`` fixes it (a trial that needs more makes the
    forward pass rerun with twice the capacity).  Otherwise 4096 steps when that is a small buffer; for large batches the
    capacity comes from a forward-only pre-pass (same controller, accepted-step counts within a few per thousand), so
    the record is about (1.02 max n_accept + 16) x B x 3N floats instead of 4096 x B x 3N -- 77 GB at B = 65,536, N = 24."""
    if options.get("record_capacity") is not None:
        return int(options["record_capacity"])
    row = 4 * y0.shape[0] * y0.shape[1]
    if 4096 * row <= _SMALL_RECORD_BYTES:
        return 4096
    _, na, _, _ = setup.ext.dopri5_fwd(setup.problem(setup.lf.W_aug), setup.t, y0.detach().to(torch.float32).contiguous(),
                                       rtol, atol, max_steps)
    # the pre-pass reads FP16 operand pairs, the recording pass TF32 pairs (stage_em.cu: stage_dopri5_fwd): their accepted-step
    # counts agree to a few steps in a thousand, hence the margin (a trial that still overflows reruns with twice the capacity)
    cap = (int(int(na.max()) * 1.02) + 16) // 8 * 8
    free, _ = torch.cuda.mem_get_info(y0.device)
    cached = torch.cuda.memory_reserved(y0.device) - torch.cuda.memory_allocated(y0.device)
    if cap * row > 0.9 * (free + cached):
        raise RuntimeError(f"odecol: recording {cap} accepted dopri5 steps of {y0.shape[0]} trials needs {cap * row / 2**30:.1f} GiB; "
                           "split the batch or use method='rk4' (checkpointed)")
    return cap


def dopri5_with_grad(setup, y0, rtol, atol, options, sel_long, sel_i32, stats):
    max_steps = int(options.get("max_num_steps", 4_000_000))
    cap = _record_capacity(setup, y0, float(rtol), float(atol), max_steps, options)
    return _Dopri5Function.apply(y0, setup.lf.W_aug, setup, float(rtol), float(atol), max_steps, cap, sel_long, sel_i32, stats,
                                 bool(options.get("check_status", True)))
