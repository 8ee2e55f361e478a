"""Repo-layout rules of the task: the product never touches the oracle, and nothing run on the GPU box reads the
reference tree."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _files(sub, exts):
    for dp, _, fns in os.walk(os.path.join(ROOT, sub)):
        if "build" in dp.split(os.sep) or "__pycache__" in dp:
            continue
        for fn in fns:
            if fn.endswith(exts):
                yield os.path.join(dp, fn)


def test_product_does_not_import_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for path in _files("ode-column_b200", (".py", ".cpp", ".cu", ".cuh", ".h")):
        src = open(path).read()
        assert not pat.search(src), f"{path} imports the oracle"
        assert "/root/reference" not in src, f"{path} names the reference tree"


def test_gpu_paths_do_not_read_reference():
    for path in ["bench.py", "__graft_entry__.py"]:
        p = os.path.join(ROOT, path)
        if os.path.exists(p):
            assert "/root/reference" not in open(p).read(), path


def test_required_layout():
    for rel in ["include/odecol.h", "oracle/__init__.py", "tests/golden/wta.npz", "tests/golden/xor.npz",
                "tests/golden/parity.npz", "config/model.toml", "ode-column_b200/csrc/abi.cu"]:
        assert os.path.exists(os.path.join(ROOT, rel)), rel
