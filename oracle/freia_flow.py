"""CPU oracle for the inverse (latent -> joint space) pass of the IKFlow conditional flow.

TEST INFRASTRUCTURE -- never imported by the product.  **Parity unpinned** for this file: FrEIA 0.2
is an un-vendored PyPI dependency of the reference (``pyproject.toml:11``, ``uv.lock:533-541``) and
the reference has no numeric golden vector for the flow (see ``oracle/__init__.py``).  Every function
below restates the published FrEIA 0.2 algorithm of the module the reference instantiates; the
citation is the reference call site that constrains it.

Everything is plain torch ops in the dtype of the state dict (fp32 for the reference,
``ikflow/config.py:8``; pass an fp64 state dict / inputs to obtain a "ground truth" for error budgets).
"""

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

# FrEIA GLOWCouplingBlock: ``f_clamp(u) = 0.636 * atan(u)`` -- the literal constant, not 2/pi.
ATAN_CLAMP_CONSTANT = 0.636
LEAKY_RELU_SLOPE = 0.01  # nn.LeakyReLU() default, ikflow/model.py:74-83


def permute_random_tables(width: int, seed: int) -> Tuple[np.ndarray, np.ndarray]:
    """FrEIA ``PermuteRandom(dims_in, seed)`` as instantiated at ``ikflow/model.py:339``.

    ``np.random.seed(seed); perm = np.random.permutation(width)``; ``perm_inv[perm[i]] = i``.
    Uses a private legacy ``RandomState`` (same MT19937 stream as the global ``np.random.seed``)
    so that the oracle has no global side effect.
    """
    rs = np.random.RandomState(seed)
    perm = rs.permutation(width)
    perm_inv = np.zeros_like(perm)
    for i, p in enumerate(perm):
        perm_inv[p] = i
    return perm.astype(np.int64), perm_inv.astype(np.int64)


def n_linear_layers(coeff_fn_config: int) -> int:
    """``subnet_constructor`` (``ikflow/model.py:51-96``): n_layers hidden activations -> n_layers+1 Linear."""
    assert coeff_fn_config in (1, 2, 3, 4)
    return coeff_fn_config + 1


def subnet_forward(sd: Dict[str, torch.Tensor], prefix: str, n_linear: int, x: torch.Tensor) -> torch.Tensor:
    """nn.Sequential(Linear, LeakyReLU, ..., Linear) of ``ikflow/model.py:51-96``.

    Linear indices in the Sequential are 0, 2, 4, ... (odd indices are the LeakyReLUs).
    """
    h = x
    for li in range(n_linear):
        w = sd[f"{prefix}.{2 * li}.weight"]
        b = sd[f"{prefix}.{2 * li}.bias"]
        h = torch.nn.functional.linear(h, w, b)
        if li != n_linear - 1:
            h = torch.nn.functional.leaky_relu(h, LEAKY_RELU_SLOPE)
    return h


def glow_coupling_reverse(
    sd: Dict[str, torch.Tensor],
    prefix: str,
    n_linear: int,
    u: torch.Tensor,
    cond: torch.Tensor,
    split_len1: int,
    clamp: float,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """FrEIA 0.2 ``GLOWCouplingBlock.forward(rev=True)`` (built at ``ikflow/model.py:342-351``).

    subnet1 reads the first half (+cond) and emits (scale, shift) for the second half; subnet2 reads
    the (already inverted) second half and emits (scale, shift) for the first half.  Returns
    (output, log-det) -- the solver discards the log-det (``ikflow/ikflow_solver.py:98``).
    """
    width = u.shape[1]
    split_len2 = width - split_len1
    x1, x2 = u[:, :split_len1], u[:, split_len1:]

    a1 = subnet_forward(sd, f"{prefix}.subnet1", n_linear, torch.cat([x1, cond], dim=1))
    s1, t1 = a1[:, :split_len2], a1[:, split_len2:]
    s1 = clamp * ATAN_CLAMP_CONSTANT * torch.atan(s1)
    y2 = (x2 - t1) * torch.exp(-s1)

    a2 = subnet_forward(sd, f"{prefix}.subnet2", n_linear, torch.cat([y2, cond], dim=1))
    s2, t2 = a2[:, :split_len1], a2[:, split_len1:]
    s2 = clamp * ATAN_CLAMP_CONSTANT * torch.atan(s2)
    y1 = (x1 - t2) * torch.exp(-s2)

    logdet = -(s1.sum(dim=1) + s2.sum(dim=1))
    return torch.cat([y1, y2], dim=1), logdet


def glow_coupling_forward(
    sd: Dict[str, torch.Tensor],
    prefix: str,
    n_linear: int,
    x: torch.Tensor,
    cond: torch.Tensor,
    split_len1: int,
    clamp: float,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """FrEIA 0.2 ``GLOWCouplingBlock.forward(rev=False)`` -- used only to check invertibility in tests."""
    x1, x2 = x[:, :split_len1], x[:, split_len1:]
    a2 = subnet_forward(sd, f"{prefix}.subnet2", n_linear, torch.cat([x2, cond], dim=1))
    s2, t2 = a2[:, :split_len1], a2[:, split_len1:]
    s2 = clamp * ATAN_CLAMP_CONSTANT * torch.atan(s2)
    y1 = torch.exp(s2) * x1 + t2
    split_len2 = x.shape[1] - split_len1
    a1 = subnet_forward(sd, f"{prefix}.subnet1", n_linear, torch.cat([y1, cond], dim=1))
    s1, t1 = a1[:, :split_len2], a1[:, split_len2:]
    s1 = clamp * ATAN_CLAMP_CONSTANT * torch.atan(s1)
    y2 = torch.exp(s1) * x2 + t1
    return torch.cat([y1, y2], dim=1), s1.sum(dim=1) + s2.sum(dim=1)


def flow_inverse(
    sd: Dict[str, torch.Tensor],
    latent: torch.Tensor,
    cond: torch.Tensor,
    nb_nodes: int,
    coeff_fn_config: int,
    rnvp_clamp: float,
    return_intermediates: bool = False,
):
    """``nn_model(latent, c=cond, rev=True)`` of ``ikflow/ikflow_solver.py:98``.

    FrEIA ``GraphINN.forward(rev=True)`` walks ``node_list[::-1]``: glow_{nb-1}^-1, perm_{nb-1}^-1, ...,
    glow_0^-1, perm_0^-1, FixedLinearTransform^-1 (graph built at ``ikflow/model.py:300-354``;
    ``module_list.0`` = FLT, ``module_list.(1+2i)`` = perm_i, ``module_list.(2+2i)`` = glow_i).
    Returns ``(out [B, W], logdet [B])``.
    """
    assert latent.shape[0] == cond.shape[0]
    width = latent.shape[1]
    split_len1 = width // 2  # ikflow/model.py:336
    n_linear = n_linear_layers(coeff_fn_config)
    u = latent
    logdet = torch.zeros(latent.shape[0], dtype=latent.dtype, device=latent.device)
    inter: List[torch.Tensor] = []
    for i in range(nb_nodes - 1, -1, -1):
        u, ld = glow_coupling_reverse(sd, f"module_list.{2 + 2 * i}", n_linear, u, cond, split_len1, rnvp_clamp)
        logdet = logdet + ld
        perm_inv = sd[f"module_list.{1 + 2 * i}.perm_inv"]
        u = u[:, perm_inv]  # PermuteRandom reverse, jac 0
        if return_intermediates:
            inter.append(u.clone())
    # FixedLinearTransform reverse: (x - b).mm(M_inv); in-repo evidence ikflow/model.py:220
    b = sd["module_list.0.b"]
    m_inv = sd["module_list.0.M_inv"]
    out = (u - b).mm(m_inv)
    logdet = logdet - sd["module_list.0.logDetM"].expand(latent.shape[0])
    if return_intermediates:
        return out, logdet, inter
    return out, logdet


def flow_forward(
    sd: Dict[str, torch.Tensor],
    x: torch.Tensor,
    cond: torch.Tensor,
    nb_nodes: int,
    coeff_fn_config: int,
    rnvp_clamp: float,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """x -> z direction (``ikflow/training/lt_model.py:156``); only used by tests as an invertibility property."""
    width = x.shape[1]
    split_len1 = width // 2
    n_linear = n_linear_layers(coeff_fn_config)
    u = x.mm(sd["module_list.0.M"]) + sd["module_list.0.b"]
    logdet = sd["module_list.0.logDetM"].expand(x.shape[0]).clone()
    for i in range(nb_nodes):
        u = u[:, sd[f"module_list.{1 + 2 * i}.perm"]]
        u, ld = glow_coupling_forward(sd, f"module_list.{2 + 2 * i}", n_linear, u, cond, split_len1, rnvp_clamp)
        logdet = logdet + ld
    return u, logdet


def build_fixed_linear_transform(
    width: int, joint_limits: List[Tuple[float, float]], dtype=torch.float32
) -> Dict[str, torch.Tensor]:
    """State-dict entries of the FixedLinearTransform node of ``ikflow/model.py:311-316``.

    ``M = diag(1/max(|lo_i|,|hi_i|))`` for the ndof joint columns, 1 for the padding columns, ``b = 0``.
    FrEIA stores ``M = M_arg.t()``, ``M_inv = M_arg.t().inverse()``, ``b.unsqueeze(0)``, ``logDetM``.
    """
    m = torch.eye(width, dtype=dtype)
    for i, (lo, hi) in enumerate(joint_limits):
        m[i, i] = 1.0 / max(abs(lo), abs(hi))
    return {
        "module_list.0.M": m.t().contiguous(),
        "module_list.0.M_inv": m.t().inverse().contiguous(),
        "module_list.0.b": torch.zeros(1, width, dtype=dtype),
        "module_list.0.logDetM": torch.slogdet(m)[1],
    }


def state_dict_to(sd: Dict[str, torch.Tensor], dtype: Optional[torch.dtype] = None, device=None):
    """Cast the floating tensors of a state dict (index tensors stay int64)."""
    out = {}
    for k, v in sd.items():
        if v.is_floating_point():
            out[k] = v.to(dtype=dtype or v.dtype, device=device or v.device)
        else:
            out[k] = v.to(device=device or v.device)
    return out
