"""Loader of the native extension.  There is no Python or CPU implementation behind it: if the extension cannot
be imported the solvers raise, loudly, at first use."""
from __future__ import annotations

import importlib.util
import sysconfig
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_EXT_PATH = _PKG / ("_odecol_ext" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
LIB_PATH = _PKG / "lib" / "libodecol.so"
_ext = None
_err = None


def ext():
    """The pybind module over libodecol.so (imports torch first so that libtorch/libc10 are resolvable)."""
    global _ext, _err
    if _ext is not None:
        return _ext
    import torch  # noqa: F401  (must be loaded before the extension)
    if not _EXT_PATH.exists():
        raise ImportError(
            f"odecol: native extension not built ({_EXT_PATH.name} missing). Run `python ode-column_b200/build.py` "
            "(needs nvcc); there is no CPU fallback.")
    try:
        spec = importlib.util.spec_from_file_location("_odecol_ext", str(_EXT_PATH))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception as e:  # pragma: no cover - depends on the machine
        _err = e
        raise ImportError(f"odecol: cannot load {_EXT_PATH.name}: {e}. There is no CPU fallback.") from e
    _ext = mod
    return mod


def available() -> bool:
    try:
        ext()
        return True
    except ImportError:
        return False
