"""Run the reference's OWN host code here, unmodified, on top of stand-ins for its two missing dependencies.

TEST INFRASTRUCTURE (same rules as the rest of ``oracle/``).  ``import ikflow`` dies in this image because FrEIA 0.2
and jrl are neither vendored in ``/root/reference`` nor installable (no network).  Everything *else* on the hot path
is the reference's own Python and can be executed as it lies under ``/root/reference`` once those two imports resolve:

* ``ikflow/model.py``           ``subnet_constructor`` (:51-96), ``IkflowModelParameters`` / ``TINY_MODEL_PARAMS``
                                (:17-48), ``IkFlowFixedLinearTransform`` (:153-238, the reference's documented copy of
                                FrEIA's ``FixedLinearTransform``), ``get_pre_sigmoid_scaling_node`` (:241-288) and the
                                graph wiring of ``glow_cNF_model`` (:291-356);
* ``ikflow/ikflow_solver.py``   ``draw_latent``, ``IKFlowSolver.__init__ / _run_inference / _calculate_pose_error /
                                _generate_exact_ik_solutions / generate_ik_solutions / generate_exact_ik_solutions``;
* ``ikflow/evaluation_utils.py`` ``evaluate_solutions`` and friends;
* ``ikflow/model_loading.py``   the registry (``MODEL_DESCRIPTIONS``).

``install()`` registers stand-in modules in ``sys.modules``:

* ``jrl.config / jrl.robots / jrl.robot / jrl.math_utils / jrl.utils`` -- ``Robot`` objects whose methods are the
  oracle's kinematics (``oracle/jrl_kinematics.py``, pinned by the reference's FK / pose-error KATs);
* ``FrEIA.framework`` (``InputNode / ConditionNode / Node / OutputNode / GraphINN`` for a chain graph) and
  ``FrEIA.modules`` (``FixedLinearTransform / PermuteRandom / GLOWCouplingBlock`` with FrEIA's parameter names, their
  arithmetic delegated to ``oracle/freia_flow.py``).  **These three classes are this repo's restatement, not FrEIA**:
  what running the reference's ``glow_cNF_model`` on top of them pins is the WIRING (node order, permutation seeds,
  ``split_len``, the ``M`` matrix, subnet sizes, state-dict key names) and -- because the stand-in coupling block calls
  the ``nn.Sequential`` that the reference's ``subnet_constructor`` returned -- the subnet arithmetic.  The GLOW
  coupling formula itself stays the one unpinned function (see ``oracle/__init__.py``).

and a package shell for ``ikflow`` that points at ``/root/reference/ikflow`` WITHOUT executing its ``__init__``
(which pulls ``visualizations`` -> klampt).  No reference source is copied or edited; nothing here is imported by the
product.  ``/root/reference`` does not exist on the GPU box: ``available()`` says whether this can run at all, and
``scripts/make_golden_reference.py`` freezes what it computes into ``tests/golden/reference_*.npz``.
"""

import importlib
import importlib.util
import os
import sys
import types
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import freia_flow, jrl_kinematics as jk

REFERENCE_ROOT = os.environ.get("IKFLOW_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ikflow", "ikflow_solver.py"))


# ----------------------------------------------------------------------------------------------------------------------
# jrl stand-ins: the reference's call sites (ikflow_solver.py:102,114,116,205,208; evaluation_utils.py:86,96,124;
# model.py:261,314) on top of the oracle's kinematics
# ----------------------------------------------------------------------------------------------------------------------


class Robot:
    """``jrl.robot.Robot`` as the reference uses it; ``chain`` is an ``oracle.jrl_kinematics.ChainRobot``."""

    def __init__(self, chain: jk.ChainRobot):
        self._chain = chain

    @property
    def name(self) -> str:
        return self._chain.name

    @property
    def ndof(self) -> int:
        return self._chain.ndof

    @property
    def actuated_joints_limits(self) -> List[Tuple[float, float]]:
        return self._chain.actuated_joints_limits

    def forward_kinematics(self, x: torch.Tensor) -> torch.Tensor:
        return jk.forward_kinematics(self._chain, x)

    def jacobian(self, x: torch.Tensor) -> torch.Tensor:
        return jk.jacobian(self._chain, x)

    def inverse_kinematics_step_levenburg_marquardt(self, target_poses, xs_current, lambd: float = 0.0001):
        return jk.lm_step(self._chain, target_poses, xs_current, lambd)

    def clamp_to_joint_limits(self, x: torch.Tensor) -> torch.Tensor:
        return jk.clamp_to_joint_limits(self._chain, x)

    def config_self_collides(self, config) -> bool:  # klampt capsule test: out of scope
        return False

    def sample_joint_angles_and_poses(self, n: int, seed: int = 0, **_):
        q, poses = jk.sample_joint_angles_and_poses(self._chain, n, seed=seed)
        return q.numpy(), poses.numpy()


class Panda(Robot):
    def __init__(self):
        super().__init__(jk.PANDA)


class FetchArm(Robot):
    def __init__(self):
        super().__init__(jk.FETCH_ARM)


def get_robot(name: str) -> Robot:
    return {"panda": Panda, "fetch_arm": FetchArm}[name]()


# ----------------------------------------------------------------------------------------------------------------------
# FrEIA stand-ins (chain graphs only).  Parameter names follow FrEIA 0.2 so that the reference's state dicts load
# (key names quoted at scripts/download_model_from_wandb_checkpoint.py:14-18).
# ----------------------------------------------------------------------------------------------------------------------


class InvertibleModule(nn.Module):
    def __init__(self, dims_in, dims_c=None):
        super().__init__()
        self.dims_in = list(dims_in)
        self.dims_c = [] if dims_c is None else list(dims_c)


class FixedLinearTransform(InvertibleModule):
    def __init__(self, dims_in, dims_c=None, M: torch.Tensor = None, b: Optional[torch.Tensor] = None):
        super().__init__(dims_in, dims_c)
        self.M = nn.Parameter(M.t(), requires_grad=False)
        self.M_inv = nn.Parameter(M.t().inverse(), requires_grad=False)
        self.b = nn.Parameter(b.unsqueeze(0), requires_grad=False)
        self.logDetM = nn.Parameter(torch.slogdet(M)[1], requires_grad=False)

    def forward(self, x, c=None, rev=False, jac=True):
        j = self.logDetM.expand(x[0].shape[0])
        if not rev:
            return (x[0].mm(self.M) + self.b,), j
        return ((x[0] - self.b).mm(self.M_inv),), -j


class PermuteRandom(InvertibleModule):
    def __init__(self, dims_in, dims_c=None, seed: Optional[int] = None):
        super().__init__(dims_in, dims_c)
        self.in_channels = dims_in[0][0]
        if seed is not None:
            np.random.seed(seed)  # global, as FrEIA does: building a model reseeds numpy
        self.perm = np.random.permutation(self.in_channels)
        self.perm_inv = np.zeros_like(self.perm)
        for i, p in enumerate(self.perm):
            self.perm_inv[p] = i
        self.perm = nn.Parameter(torch.LongTensor(self.perm), requires_grad=False)
        self.perm_inv = nn.Parameter(torch.LongTensor(self.perm_inv), requires_grad=False)

    def forward(self, x, c=None, rev=False, jac=True):
        if not rev:
            return (x[0][:, self.perm],), 0.0
        return (x[0][:, self.perm_inv],), 0.0


class GLOWCouplingBlock(InvertibleModule):
    """Stand-in with FrEIA's constructor signature.  ``subnet1`` / ``subnet2`` are whatever ``subnet_constructor``
    returns -- for the reference that is its own ``nn.Sequential`` (``ikflow/model.py:51-96``).  The coupling formula
    is ``oracle.freia_flow.glow_coupling_*`` with the subnets evaluated through those modules."""

    def __init__(self, dims_in, dims_c=None, subnet_constructor=None, clamp: float = 2.0, clamp_activation="ATAN", split_len=None):
        super().__init__(dims_in, dims_c)
        assert clamp_activation == "ATAN"
        self.channels = dims_in[0][0]
        self.condition_length = sum(d[0] for d in self.dims_c)
        self.split_len1 = self.channels // 2 if split_len is None else split_len
        self.split_len2 = self.channels - self.split_len1
        self.clamp = clamp
        self.subnet1 = subnet_constructor(self.split_len1 + self.condition_length, 2 * self.split_len2)
        self.subnet2 = subnet_constructor(self.split_len2 + self.condition_length, 2 * self.split_len1)

    def _clamped(self, s):
        return self.clamp * freia_flow.ATAN_CLAMP_CONSTANT * torch.atan(s)

    def forward(self, x, c=None, rev=False, jac=True):
        u, cond = x[0], list(c or [])
        x1, x2 = u[:, : self.split_len1], u[:, self.split_len1 :]
        if rev:
            a1 = self.subnet1(torch.cat([x1, *cond], 1))
            s1, t1 = a1[:, : self.split_len2], a1[:, self.split_len2 :]
            s1 = self._clamped(s1)
            y2 = (x2 - t1) * torch.exp(-s1)
            a2 = self.subnet2(torch.cat([y2, *cond], 1))
            s2, t2 = a2[:, : self.split_len1], a2[:, self.split_len1 :]
            s2 = self._clamped(s2)
            y1 = (x1 - t2) * torch.exp(-s2)
            return (torch.cat([y1, y2], 1),), -(s1.sum(1) + s2.sum(1))
        a2 = self.subnet2(torch.cat([x2, *cond], 1))
        s2, t2 = a2[:, : self.split_len1], a2[:, self.split_len1 :]
        s2 = self._clamped(s2)
        y1 = torch.exp(s2) * x1 + t2
        a1 = self.subnet1(torch.cat([y1, *cond], 1))
        s1, t1 = a1[:, : self.split_len2], a1[:, self.split_len2 :]
        s1 = self._clamped(s1)
        y2 = torch.exp(s1) * x2 + t1
        return (torch.cat([y1, y2], 1),), s1.sum(1) + s2.sum(1)


class _Out:
    def __init__(self, node):
        self.node = node


class _GraphNode:
    module = None

    def __init__(self, dims, name=None):
        self.output_dims = [tuple(dims)]
        self.name = name
        self.out0 = _Out(self)


class InputNode(_GraphNode):
    def __init__(self, *dims, name=None):
        super().__init__(dims, name)


class ConditionNode(_GraphNode):
    def __init__(self, *dims, name=None):
        super().__init__(dims, name)


class OutputNode(_GraphNode):
    def __init__(self, inputs, name=None):
        src = inputs[0] if isinstance(inputs, (list, tuple)) else inputs
        super().__init__(src.node.output_dims[0], name)
        self.inputs = [src]


class Node(_GraphNode):
    def __init__(self, inputs, module_type, module_args, conditions=None, name=None):
        inputs = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        self.inputs = list(inputs)
        self.conditions = [] if conditions is None else (conditions if isinstance(conditions, (list, tuple)) else [conditions])
        dims_in = [i.node.output_dims[0] for i in self.inputs]
        dims_c = [cn.output_dims[0] for cn in self.conditions]
        super().__init__(dims_in[0], name)
        self.module = module_type(dims_in, dims_c=dims_c, **module_args) if dims_c else module_type(dims_in, **module_args)


class GraphINN(InvertibleModule):
    """Chain-only stand-in: ``module_list`` holds the modules of the non-special nodes in list order (FrEIA's key
    layout), ``forward`` walks them (reversed for ``rev=True``) and sums the log-dets."""

    def __init__(self, node_list, verbose=False, force_tuple_output=False):
        in_nodes = [n for n in node_list if isinstance(n, InputNode)]
        super().__init__([in_nodes[0].output_dims[0]])
        self.node_list = list(node_list)
        prev = in_nodes[0]
        for n in self.node_list:  # a chain: every node consumes the previous one
            if isinstance(n, (Node, OutputNode)):
                assert n.inputs[0].node is prev, "the stand-in GraphINN only supports chain graphs"
                prev = n
        self.module_list = nn.ModuleList([n.module for n in self.node_list if n.module is not None])

    def forward(self, x_or_z, c=None, rev=False, jac=True, intermediate_outputs=False, x=None):
        u = (x_or_z,) if torch.is_tensor(x_or_z) else tuple(x_or_z)
        conds = [] if c is None else ([c] if torch.is_tensor(c) else list(c))
        jacobian = torch.zeros(u[0].shape[0], dtype=u[0].dtype, device=u[0].device)
        nodes = [n for n in self.node_list if isinstance(n, Node)]
        for n in nodes[::-1] if rev else nodes:
            if n.conditions:
                u, j = n.module(u, c=conds, rev=rev, jac=jac)
            else:
                u, j = n.module(u, rev=rev, jac=jac)
            jacobian = jacobian + j
        return u[0], jacobian


# ----------------------------------------------------------------------------------------------------------------------
# installation
# ----------------------------------------------------------------------------------------------------------------------


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__ikflow_b200_stub__ = True
    sys.modules[name] = m
    return m


def install(device: str = "cpu") -> None:
    """Register the stand-ins (idempotent).  ``device`` becomes ``jrl.config.DEVICE`` (``ikflow/config.py:6``)."""
    if getattr(sys.modules.get("jrl"), "__ikflow_b200_stub__", False):
        sys.modules["jrl.config"].DEVICE = device
        return
    assert available(), f"the reference tree is not present at {REFERENCE_ROOT}"
    assert "jrl" not in sys.modules and "FrEIA" not in sys.modules, "a real jrl / FrEIA is importable: use it instead"
    from typing import Union

    jrl = _module("jrl")
    jrl.__path__ = []
    jrl.config = _module("jrl.config", DEVICE=device, GPU_IDX=0, PT_NP_TYPE=Union[np.ndarray, torch.Tensor], DEFAULT_TORCH_DTYPE=torch.float32)
    jrl.robot = _module("jrl.robot", Robot=Robot)
    jrl.robots = _module("jrl.robots", Robot=Robot, Panda=Panda, FetchArm=FetchArm, get_robot=get_robot)
    jrl.math_utils = _module("jrl.math_utils", geodesic_distance_between_quaternions=jk.geodesic_distance_between_quaternions)
    jrl.utils = _module(
        "jrl.utils",
        mm_to_m=lambda x: x / 1000.0,
        make_text_green_or_red=lambda text, print_green: text,
        set_seed=lambda seed=0: torch.manual_seed(seed),
    )
    fr = _module("FrEIA")
    fr.__path__ = []
    fr.framework = _module(
        "FrEIA.framework", InputNode=InputNode, ConditionNode=ConditionNode, Node=Node, OutputNode=OutputNode, GraphINN=GraphINN
    )
    fr.modules = _module(
        "FrEIA.modules",
        InvertibleModule=InvertibleModule,
        FixedLinearTransform=FixedLinearTransform,
        PermuteRandom=PermuteRandom,
        GLOWCouplingBlock=GLOWCouplingBlock,
    )
    fr.modules.__path__ = []
    fr.modules.base = _module("FrEIA.modules.base", InvertibleModule=InvertibleModule)
    # package shell: sub-modules import from the reference tree, ikflow/__init__.py is never executed
    pkg_dir = os.path.join(REFERENCE_ROOT, "ikflow")
    spec = importlib.util.spec_from_file_location("ikflow", os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    shell = importlib.util.module_from_spec(spec)
    shell.__ikflow_b200_stub__ = True
    sys.modules["ikflow"] = shell


def load(device: str = "cpu") -> types.SimpleNamespace:
    """The reference's own modules, imported from where they lie: ``.model``, ``.ikflow_solver``,
    ``.evaluation_utils``, ``.model_loading``, ``.utils``, ``.config``."""
    install(device)
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):  # ikflow/config.py:9 prints the device at import
        mods = {n: importlib.import_module(f"ikflow.{n}") for n in ("config", "utils", "model", "evaluation_utils", "ikflow_solver", "model_loading")}
    for m in mods.values():
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(REFERENCE_ROOT)), m.__file__
    return types.SimpleNamespace(**mods)


def reference_solver(ref, hyper_parameters, robot: Robot, state_dict=None):
    """``ikflow.ikflow_solver.IKFlowSolver(hyper_parameters, robot)`` (reference code), weights from ``state_dict``."""
    hp = ref.model.IkflowModelParameters()
    hp.__dict__.update(hyper_parameters.__dict__)
    solver = ref.ikflow_solver.IKFlowSolver(hp, robot)
    if state_dict is not None:
        solver.nn_model.load_state_dict(state_dict)  # what IKFlowSolver.load_state_dict does after unpickling (:428)
        solver._model_weights_loaded = True
    return solver
