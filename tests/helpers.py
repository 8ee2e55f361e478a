"""Shared builders for the tests: product networks loaded with the golden (reference-generated) parameters, and the
oracle's unified form of the same networks."""
import numpy as np
import torch

import odecol
from oracle import column_model as cm

XOR_DICT = {"nr_areas": 2, "areas": ["mt", "mt"], "nr_columns_per_area": [2, 1], "nr_input_units": 2}
PARITY_DICT = {"nr_areas": 3, "areas": ["mt"] * 3, "nr_columns_per_area": [8, 4, 1], "nr_input_units": 4}


def product_network(name, cfg, g, device="cpu"):
    """Product module carrying exactly the reference's seed-0 parameters (copied from the golden file)."""
    torch.manual_seed(0)
    if name == "wta":
        net = odecol.ColumnAreaWTA(cfg, "mt")
        with torch.no_grad():
            net.recurrent_weights.copy_(torch.tensor(g["recurrent_weights"]))
    elif name == "xor":
        net = odecol.ColumnNetworkXOR(cfg, XOR_DICT)
        with torch.no_grad():
            for a in "01":
                for i in range(2):
                    net.feedforward_target_weights[a][i].copy_(torch.tensor(g[f"ffw_{a}_{i}"]))
    elif name == "parity":
        net = odecol.ColumnNetwork(cfg, PARITY_DICT, torch.device("cpu"))
        with torch.no_grad():
            for k in "012":
                net.areas[k].lateral_weights.copy_(torch.tensor(g[f"lateral_{k}"]))
            for k in "12":
                net.areas[k].feedforward_weights.copy_(torch.tensor(g[f"feedforward_{k}"]))
            net.areas["0"].input_weights.copy_(torch.tensor(g["input_weights"]))
            net.output_weights.copy_(torch.tensor(g["output_weights"]))
    else:
        raise KeyError(name)
    net.time_vec = torch.tensor(g["time_vec"])
    return to_device(net, device)


def to_device(net, device):
    """Module.to() plus the plain-tensor attributes the reference keeps outside buffers."""
    net = net.to(device)
    mods = [net] + list(net.modules())
    for m in mods:
        for k, v in list(vars(m).items()):
            if torch.is_tensor(v) and not isinstance(v, torch.nn.Parameter):
                setattr(m, k, v.to(device))
    return net


def oracle_form(name, cfg, g):
    if name == "wta":
        return cm.wta_linear_form(cfg, g["recurrent_weights"])
    if name == "xor":
        return cm.xor_linear_form(cfg, [[g["ffw_0_0"], g["ffw_0_1"]], [g["ffw_1_0"], g["ffw_1_1"]]])
    if name == "parity":
        return cm.parity_linear_form(cfg, [g[f"lateral_{k}"] for k in range(3)],
                                     {1: g["feedforward_1"], 2: g["feedforward_2"]}, g["input_weights"])
    raise KeyError(name)


def stim_table(name, stim):
    """Reference-shaped stimulus -> (B, T, n_in) channel table (B = leading batch if present)."""
    s = torch.as_tensor(stim)
    if name == "xor":
        if s.dim() == 3:
            s = s[None]
        return s.reshape(s.shape[0], s.shape[1], -1)
    if s.dim() == 2:
        s = s[None]
    return s


def block_errs(a, b):
    """Norm-relative error of the V, A and F blocks (elementwise relative error is meaningless
    at zero crossings of V)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    n = a.shape[-1] // 3
    out = []
    for c in range(3):
        x, y = a[..., c * n:(c + 1) * n], b[..., c * n:(c + 1) * n]
        out.append(float((x - y).abs().max() / y.abs().max().clamp_min(1e-30)))
    return out


def rel_err(a, b):
    return max(block_errs(a, b))


def sheet_oracle_form(sheet):
    """Oracle LinearForm of a product SyntheticColumnSheet (its W_aug split back into W | U | bias)."""
    from oracle.column_model import LinearForm
    lfp = sheet.export_linear_form()
    n, n_in = lfp.N, lfp.n_in
    Wa = lfp.W_aug.detach().cpu().numpy()
    return LinearForm(W=Wa[:, :n], U=Wa[:, n:n + n_in], bias=Wa[:, n + n_in], kappa=lfp.kappa.cpu().numpy(),
                      sigma=lfp.sigma.cpu().numpy(), tau_s=lfp.tau_s, tau_m=lfp.tau_m, tau_a=lfp.tau_a,
                      resistance=lfp.resistance)
