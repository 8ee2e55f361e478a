"""odecol -- B200-native fused integrator for the ODE-Column hot path (package directory ``ode-column_b200``).

Drop-in surface (reference src/coupled_columns.py, src/utils.py; torchdiffeq / torchsde call shapes):

    ColumnArea, ColumnAreaWTA, ColumnNetworkXOR, ColumnNetwork, load_config,
    compute_firing_rate, soft_clamp, torch_interp, min_max, fr_to_binary, huber_loss_wta,
    odeint, odeint_adjoint, sdeint,
    make_ds_wwp, get_data            (Wong-Wang target generator, reference scripts/wta_ode.py:56-107)

The solvers run only on CUDA through the C ABI in ``include/odecol.h``; importing the package does not need a GPU.
"""
from .model import (ColumnArea, ColumnAreaWTA, ColumnNetwork, ColumnNetworkXOR, LinearForm, compute_firing_rate,
                    load_config, move_to, pack_w_aug, soft_clamp, torch_interp)
from .losses import (fr_to_binary, huber_loss_wta, huber_rate_loss, min_max, parity_readout, readout_components,
                     window_rate_l1_loss, xor_readout)
from .solvers import odeint, odeint_adjoint, sdeint, sdeint_adjoint
from .stimulus import compress_knots, step_knots
from .synthetic import SyntheticColumnSheet
from .wongwang import get_data, make_ds_wwp
from . import distributed, wongwang
from . import _native

__all__ = [
    "ColumnArea", "ColumnAreaWTA", "ColumnNetwork", "ColumnNetworkXOR", "SyntheticColumnSheet", "LinearForm",
    "compute_firing_rate", "soft_clamp", "torch_interp", "load_config", "pack_w_aug", "move_to",
    "min_max", "fr_to_binary", "huber_loss_wta", "huber_rate_loss", "window_rate_l1_loss", "readout_components", "xor_readout",
    "parity_readout",
    "odeint", "odeint_adjoint", "sdeint", "sdeint_adjoint", "compress_knots", "step_knots", "distributed",
    "make_ds_wwp", "get_data", "wongwang",
]
__version__ = "0.1.0"
