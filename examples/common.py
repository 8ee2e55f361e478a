"""Shared helpers of the example training loops: repository import path and `.to(device)` for the drop-in modules
(the reference keeps masks / constants as plain tensor attributes, so Module.to() alone does not move them)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
CONFIG = os.path.join(ROOT, "config", "model.toml")


def to_device(net, device):
    import odecol
    return odecol.move_to(net, device)
