"""CPU oracle for the ODE-Column hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm the reference executes for the
coupled cortical-column ODE/SDE (``/root/reference/src/coupled_columns.py`` driven by
torchdiffeq ``odeint`` / torchsde ``sdeint``).  It exists so that the CUDA product
in ``ode-column_b200/`` can be checked against something independent.

Rules (enforced by ``tests/test_layout.py``):

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
  ``--impl reference`` legs may import it;
* nothing under ``ode-column_b200/`` may import it -- the product has no CPU fallback.

Parity status
-------------
* RHS / weight construction: PINNED.  ``oracle/make_golden.py`` imports the
  reference's own ``src/`` package unchanged (it runs in the build container) and
  the restatements in ``oracle/column_model.py`` / ``oracle/rhs.py`` are compared
  with it to fp32 rounding; the resulting vectors are committed under
  ``tests/golden/``.
* Solvers (torchdiffeq rk4 / dopri5 / odeint_adjoint, torchsde euler):
  **PARITY UNPINNED**.  Neither third-party package is vendored by the reference,
  neither is pinned (no requirements file) and neither is installed here; the
  reference holds no golden trajectory.  ``oracle/solvers.py`` restates their
  published algorithms (torchdiffeq 0.2.x ``rk4_alt_step_func``,
  ``RKAdaptiveStepsizeODESolver`` with the Dormand-Prince tableau; torchsde 0.2.x
  ``Euler`` + ``adaptive_stepping``) and anchors them with self-checks (order
  conditions, convergence order, FSAL, analytic linear ODE / OU moments).
"""
