"""The C-ABI shared library loads without a GPU and exports every symbol include/odecol.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ode-column_b200", "lib", "libodecol.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import importlib.util
        spec = importlib.util.spec_from_file_location("odecol_build", os.path.join(ROOT, "ode-column_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return ctypes.CDLL(LIB)


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "odecol.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(odecol_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/odecol.h but not exported"


def test_version_and_strerror(lib):
    assert lib.odecol_abi_version() == 2
    lib.odecol_strerror.restype = ctypes.c_char_p
    assert lib.odecol_strerror(0) == b"ok"
    assert b"workspace" in lib.odecol_strerror(-4)


def test_null_problem_is_rejected_without_touching_the_gpu(lib):
    lib.odecol_rk4_fwd.restype = ctypes.c_int
    rc = lib.odecol_rk4_fwd(None, None, 4, None, None, 1, None, ctypes.c_size_t(0), None)
    assert rc == -1


def test_em_num_steps_matches_oracle_schedule(lib):
    from oracle import solvers, stimuli
    lib.odecol_em_num_steps.restype = ctypes.c_int64
    lib.odecol_em_num_steps.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_float]
    for T, dt_grid, dt in ((1500, 1e-4, 1e-3), (1000, 1e-3, 1e-3), (1000, 1e-3, 2.5e-4), (37, 0.01, 0.003)):
        ts = stimuli.time_vec(T, dt_grid).contiguous()
        n = lib.odecol_em_num_steps(ts.data_ptr(), T, dt)
        assert n == len(solvers.em_step_schedule(ts, dt))


def test_torch_extension_imports_without_gpu():
    import odecol
    e = odecol._native.ext()
    assert e.abi_version() == 2
    with pytest.raises(RuntimeError):
        # CPU tensors are refused: there is no CPU path
        e.Problem(torch.zeros(8, 12), torch.zeros(8), None, torch.zeros(2), torch.zeros(1, 2, 1), 1, 1, 5e-4, 0.02, 10.0, 80.0, 0)
