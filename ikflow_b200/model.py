"""Model description for the B200 flow engine: hyper-parameters, FrEIA-layout state dicts, graph builder.

Mirrors ``ikflow/model.py`` of the reference for the default (non-sigmoid) branch:

* ``IkflowModelParameters`` / ``TINY_MODEL_PARAMS``  -- ``ikflow/model.py:17-48`` (same attribute names/defaults).
* ``glow_cNF_model(params, robot, dim_cond, ndim_tot)``  -- ``ikflow/model.py:291-356``; instead of a FrEIA
  ``GraphINN`` it returns a :class:`ikflow_b200.flow.FlowModel`, which is callable exactly like the reference
  uses ``nn_model`` (``nn_model(latent, c=cond, rev=True) -> (out, logdet)``, ``ikflow/ikflow_solver.py:98``) and
  loads state dicts with the FrEIA key names.
* ``make_synthetic_state_dict``  -- seeded weights in that exact key layout (the released ``.pkl`` files live on GCS,
  ``ikflow/model_loading.py:31-49``; there is no network here).
"""

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


class IkflowModelParameters:
    """Same attribute bag as ``ikflow/model.py:17-41`` (``get_ik_solver`` updates ``__dict__`` from the yaml entry)."""

    def __init__(self):
        self.coupling_layer = "glow"
        self.nb_nodes = 12
        self.dim_latent_space = 9
        self.coeff_fn_config = 3
        self.coeff_fn_internal_size = 1024
        self.permute_random_enabled = True
        self.sigmoid_on_output = False

        # ___ Loss parameters (training only; kept so that attribute access keeps working)
        self.lambd_predict = 1.0
        self.init_scale = 0.04473500291638653
        self.rnvp_clamp = 2.5
        self.y_noise_scale = 1e-7
        self.zeros_noise_scale = 1e-3

        self.softflow_noise_scale = 0.01
        self.softflow_enabled = True

    def __str__(self) -> str:
        s = "IkflowModelParameters\n"
        for k, v in self.__dict__.items():
            s += f"  {k}: \t{v}\n"
        return s


# Convenience variable for testing purposes (ikflow/model.py:45-48)
TINY_MODEL_PARAMS = IkflowModelParameters()
TINY_MODEL_PARAMS.nb_nodes = 3
TINY_MODEL_PARAMS.coeff_fn_config = 2
TINY_MODEL_PARAMS.coeff_fn_internal_size = 256


def permute_random_tables(width: int, seed: int) -> Tuple[np.ndarray, np.ndarray]:
    """Permutation tables of FrEIA ``PermuteRandom(seed=i)`` (``ikflow/model.py:339``): legacy MT19937 stream,
    ``np.random.seed(seed); np.random.permutation(width)``.  A private ``RandomState`` yields the same stream
    without touching numpy's global RNG."""
    perm = np.random.RandomState(seed).permutation(width).astype(np.int64)
    perm_inv = np.empty_like(perm)
    perm_inv[perm] = np.arange(width, dtype=np.int64)
    return perm, perm_inv


def subnet_layer_dims(ch_in: int, ch_out: int, internal_size: int, n_layers: int) -> List[Tuple[int, int]]:
    """(in, out) of every nn.Linear of ``subnet_constructor`` (``ikflow/model.py:51-96``)."""
    assert n_layers in [1, 2, 3, 4], "Number of layers `n_layers` must be in [1, ..., 4]"
    dims = [(ch_in, internal_size)]
    dims += [(internal_size, internal_size)] * (n_layers - 1)
    dims += [(internal_size, ch_out)]
    return dims


def state_dict_keys(params: IkflowModelParameters, dim_cond: int) -> Dict[str, Tuple[int, ...]]:
    """Every key of the FrEIA state dict with its shape."""
    w = params.dim_latent_space
    s1 = w // 2  # ikflow/model.py:336
    s2 = w - s1
    keys: Dict[str, Tuple[int, ...]] = {
        "module_list.0.M": (w, w),
        "module_list.0.M_inv": (w, w),
        "module_list.0.b": (1, w),
        "module_list.0.logDetM": (),
    }
    for i in range(params.nb_nodes):
        keys[f"module_list.{1 + 2 * i}.perm"] = (w,)
        keys[f"module_list.{1 + 2 * i}.perm_inv"] = (w,)
        for name, cin, cout in (("subnet1", s1 + dim_cond, 2 * s2), ("subnet2", s2 + dim_cond, 2 * s1)):
            dims = subnet_layer_dims(cin, cout, params.coeff_fn_internal_size, params.coeff_fn_config)
            for li, (fi, fo) in enumerate(dims):
                keys[f"module_list.{2 + 2 * i}.{name}.{2 * li}.weight"] = (fo, fi)
                keys[f"module_list.{2 + 2 * i}.{name}.{2 * li}.bias"] = (fo,)
    return keys


def make_synthetic_state_dict(
    params: IkflowModelParameters,
    joint_limits: Sequence[Tuple[float, float]],
    seed: int = 0,
    stress: float = 1.0,
    dim_cond: Optional[int] = None,
) -> Dict[str, torch.Tensor]:
    """Seeded random weights in the reference's state-dict layout.

    nn.Linear default init (U(+-1/sqrt(fan_in)) for weight and bias), permutations from ``PermuteRandom(seed=i)``,
    ``M = diag(1/max|limit|)`` as ``ikflow/model.py:311-316``.  ``stress=k`` multiplies the last Linear of every
    subnet by k so that the scale outputs span more of the atan/exp range (k=8 as suggested there
    makes the untrained flow expand to |q|~1e8, k=2..3 keeps |q| in a sane range while exercising the clamp).
    """
    if dim_cond is None:
        dim_cond = 8 if params.softflow_enabled else 7
    g = torch.Generator().manual_seed(seed)
    w = params.dim_latent_space
    sd: Dict[str, torch.Tensor] = {}
    m = torch.eye(w, dtype=torch.float32)
    for i, (lo, hi) in enumerate(joint_limits):
        m[i, i] = 1.0 / max(abs(lo), abs(hi))
    sd["module_list.0.M"] = m.t().contiguous()
    sd["module_list.0.M_inv"] = m.t().inverse().contiguous()
    sd["module_list.0.b"] = torch.zeros(1, w, dtype=torch.float32)
    sd["module_list.0.logDetM"] = torch.slogdet(m)[1]
    for key, shape in state_dict_keys(params, dim_cond).items():
        if key.startswith("module_list.0."):
            continue
        if key.endswith(".perm"):
            idx = (int(key.split(".")[1]) - 1) // 2
            perm, perm_inv = permute_random_tables(w, idx)
            sd[key] = torch.from_numpy(perm)
            sd[key + "_inv"] = torch.from_numpy(perm_inv)
        elif key.endswith(".perm_inv"):
            continue
        else:
            fan_in = shape[1] if key.endswith(".weight") else sd[key.replace(".bias", ".weight")].shape[1]
            bound = 1.0 / float(fan_in) ** 0.5
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound
            last = int(key.split(".")[3]) == 2 * params.coeff_fn_config
            if last and stress != 1.0:
                t = t * float(stress)
            sd[key] = t
    return sd


DEFAULT_PRECISION = "auto"


def glow_cNF_model(params: IkflowModelParameters, robot, dim_cond: int, ndim_tot: int, precision: Optional[str] = None):
    """Build the conditional flow for ``robot`` -- ``ikflow/model.py:291-356``.

    The reference wires FrEIA nodes (FixedLinearTransform, then ``nb_nodes`` x [PermuteRandom(seed=i),
    GLOWCouplingBlock]) into a ``GraphINN``; here the same hyper-parameters configure one
    :class:`ikflow_b200.flow.FlowModel`, whose reverse pass is a single sm_100a kernel.  The weights (including the
    FixedLinearTransform matrices and the permutation tables) arrive through ``load_state_dict`` exactly as in the
    reference.

    ``precision``: operand format of the hidden-layer tensor-core products (``include/ikflow_b200.h``): ``"auto"`` (default:
    ``"fp16x3"`` where the tcgen05 engine runs the model, ``"bf16x3"`` otherwise), ``"fp16x3"`` (fp32-grade: as close to the
    fp32 reference as two fp32 implementations are to each other; hidden activations must stay inside the fp16 range),
    ``"bf16x3"`` (16 operand bits, the exponent range of fp32, 3 % faster) or ``"bf16x1"`` (fast, NOT parity grade);
    ``IKFLOW_B200_PRECISION`` overrides the default.
    """
    import os

    from .flow import FlowModel

    if precision is None:
        precision = os.environ.get("IKFLOW_B200_PRECISION", DEFAULT_PRECISION)

    assert ndim_tot >= robot.ndof, f"network width {ndim_tot} is smaller than the robot's {robot.ndof} dofs"
    return FlowModel(params, robot.actuated_joints_limits, dim_cond, ndim_tot, precision=precision)
