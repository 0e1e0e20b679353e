"""Kinematic chains and the ``Robot`` operator interface of the hot path.

Mirrors the part of jrl's ``Robot`` (jrl @ 2ba7c39, un-vendored dependency of the reference) that
``ikflow/ikflow_solver.py`` calls -- same method names, argument meaning, in-place semantics and return layout:

* ``forward_kinematics(x[m,ndof]) -> [m,7]``                        (``ikflow_solver.py:114``)
* ``inverse_kinematics_step_levenburg_marquardt(poses, x) -> x'``   (``ikflow_solver.py:205,208``)
* ``clamp_to_joint_limits(x) -> x`` (in place)                      (``ikflow_solver.py:102``)
* ``actuated_joints_limits``, ``ndof``, ``name``                    (``ikflow/model.py:261,314``, ``model_loading.py:83``)
* ``sample_joint_angles_and_poses(n)``                              (``scripts/benchmark_runtime.py:83-86``)

Every method runs on the GPU through ``libikflow_b200`` (``csrc/robot.cu``); tensors must be CUDA fp32.  Collision
checking, meshes and klampt are out of scope.
"""

import math
import warnings
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_KIND = {"fixed": 0, "revolute": 1, "continuous": 1, "prismatic": 2}


def _rpy_matrix(rpy: Sequence[float]) -> np.ndarray:
    """URDF fixed-axis roll/pitch/yaw: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr],
        ],
        dtype=np.float64,
    )


class Joint:
    """One URDF joint on the base -> end-effector chain."""

    def __init__(self, name, kind, xyz, rpy, axis=(0.0, 0.0, 1.0), limits=None):
        assert kind in _KIND, f"unknown joint type '{kind}'"
        self.name, self.kind = name, kind
        self.xyz, self.rpy, self.axis = tuple(xyz), tuple(rpy), tuple(axis)
        self.limits = None if limits is None else (float(limits[0]), float(limits[1]))
        if kind != "fixed":
            assert limits is not None, f"actuated joint '{name}' needs limits"


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _check(t: torch.Tensor, name: str, cols: Optional[int] = None) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor (got {type(t)})")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on '{t.device}': ikflow_b200 computes on CUDA (sm_100a) only, there is no CPU path")
    assert t.dtype == torch.float32, f"{name} must be float32 (got {t.dtype})"
    assert t.dim() == 2, f"{name} must be 2-dimensional (got shape {tuple(t.shape)})"
    if cols is not None:
        assert t.shape[1] == cols, f"{name} must be [n x {cols}] (got {tuple(t.shape)})"


class Robot:
    def __init__(self, name: str, joints: Sequence[Joint]):
        self._name = name
        self._joints = list(joints)
        self._actuated = [j for j in self._joints if j.kind != "fixed"]
        assert 1 <= len(self._actuated) <= _lib.IKF_MAX_DOF, f"{len(self._actuated)} actuated joints (max {_lib.IKF_MAX_DOF})"
        assert len(self._joints) <= _lib.IKF_MAX_LINKS
        self._handles: Dict[int, int] = {}

    # ---- description ------------------------------------------------------------------------------------------------
    @property
    def name(self) -> str:
        return self._name

    @property
    def ndof(self) -> int:
        return len(self._actuated)

    @property
    def n_dofs(self) -> int:  # older jrl spelling, used by some reference scripts
        return self.ndof

    @property
    def actuated_joints_limits(self) -> List[Tuple[float, float]]:
        return [j.limits for j in self._actuated]

    @property
    def actuated_joint_names(self) -> List[str]:
        return [j.name for j in self._actuated]

    def __str__(self) -> str:
        return f"<Robot[{self.name}] ndof={self.ndof}>"

    # ---- device handle ----------------------------------------------------------------------------------------------
    def _handle(self, device: torch.device) -> int:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._handles:
            n = len(self._joints)
            kind = np.array([_KIND[j.kind] for j in self._joints], dtype=np.int32)
            fixed = np.zeros((n, 12), dtype=np.float64)
            axis = np.zeros((n, 3), dtype=np.float64)
            for i, j in enumerate(self._joints):
                t = np.zeros((3, 4), dtype=np.float64)
                t[:, :3] = _rpy_matrix(j.rpy)
                t[:, 3] = j.xyz
                fixed[i] = t.reshape(-1)
                axis[i] = j.axis
            lims = np.array(self.actuated_joints_limits, dtype=np.float64)
            lo, hi = np.ascontiguousarray(lims[:, 0]), np.ascontiguousarray(lims[:, 1])
            out = _lib.ctypes.c_void_p()
            code = _lib.lib().ikf_robot_create(
                n, kind.ctypes.data, fixed.ctypes.data, axis.ctypes.data, lo.ctypes.data, hi.ctypes.data, idx,
                _lib.ctypes.byref(out),
            )
            _lib.check(code, "ikf_robot_create")
            self._handles[idx] = out.value
        return self._handles[idx]

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.lib().ikf_robot_destroy(h)
        except Exception:
            pass

    # ---- operators on the hot path ----------------------------------------------------------------------------------
    def forward_kinematics(self, x: torch.Tensor, out_device: Optional[str] = None) -> torch.Tensor:
        """[m x ndof] joint angles -> [m x 7] end-effector poses [x, y, z, qw, qx, qy, qz]."""
        _check(x, "x", self.ndof)
        x = x.contiguous()
        poses = torch.empty((x.shape[0], 7), dtype=torch.float32, device=x.device)
        code = _lib.lib().ikf_forward_kinematics(self._handle(x.device), x.data_ptr(), poses.data_ptr(), x.shape[0], _stream_ptr(x))
        _lib.check(code, "ikf_forward_kinematics")
        return poses if out_device is None else poses.to(out_device)

    def clamp_to_joint_limits(self, x: torch.Tensor) -> torch.Tensor:
        """Clamp every column to its joint limits IN PLACE and return ``x`` (jrl semantics)."""
        _check(x, "x", self.ndof)
        if x.is_contiguous():
            code = _lib.lib().ikf_clamp_to_joint_limits(self._handle(x.device), x.data_ptr(), x.shape[0], _stream_ptr(x))
            _lib.check(code, "ikf_clamp_to_joint_limits")
            return x
        tmp = x.contiguous()
        code = _lib.lib().ikf_clamp_to_joint_limits(self._handle(x.device), tmp.data_ptr(), tmp.shape[0], _stream_ptr(x))
        _lib.check(code, "ikf_clamp_to_joint_limits")
        x.copy_(tmp)
        return x

    def inverse_kinematics_step_levenburg_marquardt(
        self, target_poses: torch.Tensor, xs_current: torch.Tensor, lambd: float = 0.0001
    ) -> torch.Tensor:
        """One Levenberg-Marquardt step towards ``target_poses`` ([m x 7]); returns the clamped new configs."""
        _check(target_poses, "target_poses", 7)
        _check(xs_current, "xs_current", self.ndof)
        assert target_poses.shape[0] == xs_current.shape[0], f"{target_poses.shape[0]} != {xs_current.shape[0]}"
        assert target_poses.device == xs_current.device
        poses, xs = target_poses.contiguous(), xs_current.contiguous()
        out = torch.empty_like(xs)
        code = _lib.lib().ikf_lm_step(
            self._handle(xs.device), poses.data_ptr(), poses.shape[0], xs.data_ptr(), out.data_ptr(), xs.shape[0],
            float(lambd), 1, _stream_ptr(xs),
        )
        _lib.check(code, "ikf_lm_step")
        return out

    def pose_errors(self, qs: torch.Tensor, target_poses: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Fused ``IKFlowSolver._calculate_pose_error`` (``ikflow_solver.py:112-117``): FK + L2 + quaternion geodesic.
        ``target_poses`` may have fewer rows than ``qs``; row i uses pose i % rows."""
        _check(qs, "qs", self.ndof)
        _check(target_poses, "target_poses", 7)
        qs, poses = qs.contiguous(), target_poses.contiguous()
        pos = torch.empty(qs.shape[0], dtype=torch.float32, device=qs.device)
        rot = torch.empty_like(pos)
        code = _lib.lib().ikf_pose_error(
            self._handle(qs.device), qs.data_ptr(), poses.data_ptr(), poses.shape[0], pos.data_ptr(), rot.data_ptr(),
            qs.shape[0], _stream_ptr(qs),
        )
        _lib.check(code, "ikf_pose_error")
        return pos, rot

    def evaluate_solutions(self, qs: torch.Tensor, target_poses: torch.Tensor):
        """pose errors + joint-limit flag in one launch (``ikflow/evaluation_utils.py:65-112``)."""
        _check(qs, "qs", self.ndof)
        _check(target_poses, "target_poses", 7)
        qs, poses = qs.contiguous(), target_poses.contiguous()
        pos = torch.empty(qs.shape[0], dtype=torch.float32, device=qs.device)
        rot = torch.empty_like(pos)
        exceeded = torch.empty(qs.shape[0], dtype=torch.uint8, device=qs.device)
        code = _lib.lib().ikf_evaluate_solutions(
            self._handle(qs.device), qs.data_ptr(), poses.data_ptr(), poses.shape[0], pos.data_ptr(), rot.data_ptr(),
            exceeded.data_ptr(), qs.shape[0], _stream_ptr(qs),
        )
        _lib.check(code, "ikf_evaluate_solutions")
        return pos, rot, exceeded.bool()

    def lm_refine(
        self, target_poses: torch.Tensor, q_seeds: torch.Tensor, repeat_count: int, n_steps: int, pos_thr: float,
        rot_thr: float, lambd: float = 0.0001,
    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Device-side LM / select loop of ``_generate_exact_ik_solutions`` (``ikflow_solver.py:197-233``).
        ``q_seeds`` is [repeat_count*n x ndof], repeat-major, and is updated in place.  Returns
        (final_q [n x ndof], final_valid [n] bool, n_valid int32[1] on the device)."""
        _check(target_poses, "target_poses", 7)
        _check(q_seeds, "q_seeds", self.ndof)
        n = target_poses.shape[0]
        assert q_seeds.shape[0] == n * repeat_count and q_seeds.is_contiguous()
        poses = target_poses.contiguous()
        final_q = torch.empty((n, self.ndof), dtype=torch.float32, device=poses.device)
        valid = torch.empty(n, dtype=torch.uint8, device=poses.device)
        n_valid = torch.empty(1, dtype=torch.int32, device=poses.device)
        code = _lib.lib().ikf_lm_refine(
            self._handle(poses.device), poses.data_ptr(), q_seeds.data_ptr(), n, int(repeat_count), int(n_steps),
            float(pos_thr), float(rot_thr), float(lambd), final_q.data_ptr(), valid.data_ptr(), n_valid.data_ptr(),
            _stream_ptr(poses),
        )
        _lib.check(code, "ikf_lm_refine")
        return final_q, valid.bool(), n_valid

    # ---- sampling (the step before the hot path) --------------------------------------------------------------------
    def sample_joint_angles_and_poses(
        self, n: int, joint_limit_eps: float = 1e-6, only_non_self_colliding: bool = False, tqdm_enabled: bool = False,
        return_torch: bool = False, seed: Optional[int] = None, first_index: int = 0, generator=None, device=None,
    ):
        """Uniform joint samples inside the limits and their FK poses -- jrl ``Robot.sample_joint_angles_and_poses``
        (``scripts/benchmark_runtime.py:83-86``, ``scripts/evaluate.py:137-139``) in ONE kernel launch: a counter-based
        Philox4x32-10 draw per (sample, joint) and the forward kinematics of the sample (``ikf_sample_joint_angles_and_poses``).
        Nothing is drawn on the host and nothing crosses PCIe unless numpy arrays are asked for (jrl's return type,
        ``return_torch=False``).

        ``seed``: 64-bit key of the stream; ``None`` draws one from ``generator`` (default: torch's global CPU generator,
        so ``torch.manual_seed`` / the reference's ``set_seed`` make the samples reproducible).  Sample ``i`` depends
        only on ``(seed, first_index + i)``: shards of one index range are the same whatever the split.
        The self-collision rejection of jrl needs klampt capsule geometry, which is outside this package."""
        if only_non_self_colliding:
            warnings.warn("ikflow_b200 has no self-collision checker: samples are NOT filtered for self-collisions")
        assert isinstance(n, int) and n >= 0
        device = torch.device(device or _default_device())
        if seed is None:
            seed = int(torch.randint(0, 2**62, (1,), generator=generator).item())
        q = torch.empty((n, self.ndof), dtype=torch.float32, device=device)
        poses = torch.empty((n, 7), dtype=torch.float32, device=device)
        code = _lib.lib().ikf_sample_joint_angles_and_poses(
            self._handle(device), int(seed) & (2**64 - 1), int(first_index), float(joint_limit_eps), q.data_ptr(),
            poses.data_ptr(), n, _stream_ptr(q),
        )
        _lib.check(code, "ikf_sample_joint_angles_and_poses")
        if return_torch:
            return q, poses
        return q.cpu().numpy(), poses.cpu().numpy()

    def sample_joint_angles(self, n: int, joint_limit_eps: float = 1e-6, seed: Optional[int] = None, generator=None, device=None) -> torch.Tensor:
        """jrl ``Robot.sample_joint_angles``: the joint half of :meth:`sample_joint_angles_and_poses` (torch tensor)."""
        return self.sample_joint_angles_and_poses(n, joint_limit_eps, return_torch=True, seed=seed, generator=generator, device=device)[0]


def _default_device() -> str:
    from .config import DEVICE

    if "cuda" not in str(DEVICE):
        raise RuntimeError("ikflow_b200 needs a CUDA (sm_100a) device; none is visible and there is no CPU path")
    return DEVICE


_HP = math.pi / 2


class Panda(Robot):
    """Franka Panda, base -> panda_hand.  Joint limits as pinned by the reference's ``tests/model_test.py:18-25``; the
    chain reproduces the golden FK vector of ``tests/evaluation_utils_test.py:20-24``."""

    name_ = "panda"

    def __init__(self):
        super().__init__(
            "panda",
            [
                Joint("panda_joint1", "revolute", (0, 0, 0.333), (0, 0, 0), limits=(-2.8973, 2.8973)),
                Joint("panda_joint2", "revolute", (0, 0, 0), (-_HP, 0, 0), limits=(-1.7628, 1.7628)),
                Joint("panda_joint3", "revolute", (0, -0.316, 0), (_HP, 0, 0), limits=(-2.8973, 2.8973)),
                Joint("panda_joint4", "revolute", (0.0825, 0, 0), (_HP, 0, 0), limits=(-3.0718, -0.0698)),
                Joint("panda_joint5", "revolute", (-0.0825, 0.384, 0), (-_HP, 0, 0), limits=(-2.8973, 2.8973)),
                Joint("panda_joint6", "revolute", (0, 0, 0), (_HP, 0, 0), limits=(-0.0175, 3.7525)),
                Joint("panda_joint7", "revolute", (0.088, 0, 0), (_HP, 0, 0), limits=(-2.8973, 2.8973)),
                Joint("panda_joint8", "fixed", (0, 0, 0.107), (0, 0, 0)),
                Joint("panda_hand_joint", "fixed", (0, 0, 0), (0, 0, -math.pi / 4)),
            ],
        )


_FETCH_ARM_JOINTS = [
    ("shoulder_pan_joint", "revolute", (0.119525, 0, 0.34858), (0, 0, 1), (-1.6056, 1.6056)),
    ("shoulder_lift_joint", "revolute", (0.117, 0, 0.06), (0, 1, 0), (-1.221, 1.518)),
    ("upperarm_roll_joint", "revolute", (0.219, 0, 0), (1, 0, 0), (-math.pi, math.pi)),
    ("elbow_flex_joint", "revolute", (0.133, 0, 0), (0, 1, 0), (-2.251, 2.251)),
    ("forearm_roll_joint", "revolute", (0.197, 0, 0), (1, 0, 0), (-math.pi, math.pi)),
    ("wrist_flex_joint", "revolute", (0.1245, 0, 0), (0, 1, 0), (-2.16, 2.16)),
    ("wrist_roll_joint", "revolute", (0.1385, 0, 0), (1, 0, 0), (-math.pi, math.pi)),
]


class FetchArm(Robot):
    """Fetch arm with the torso fixed, base_link -> gripper_link.  NOTE: the reference tree holds no URDF or test
    constant for this robot; the constants are the public Fetch URDF with continuous joints limited
    to +-pi.  Only the 7 limit pairs influence the flow (M_inv and the clamp)."""

    def __init__(self):
        joints = [Joint("torso_lift_joint", "fixed", (-0.086875, 0, 0.37743), (0, 0, 0))]
        joints += [Joint(n, k, xyz, (0, 0, 0), ax, lim) for n, k, xyz, ax, lim in _FETCH_ARM_JOINTS]
        joints += [Joint("gripper_axis", "fixed", (0.16645, 0, 0), (0, 0, 0))]
        super().__init__("fetch_arm", joints)


class Fetch(Robot):
    """Fetch with the prismatic torso lift (8 dof).  Same provenance caveat as :class:`FetchArm`."""

    def __init__(self):
        joints = [Joint("torso_lift_joint", "prismatic", (-0.086875, 0, 0.37743), (0, 0, 0), (0, 0, 1), (0.0, 0.38615))]
        joints += [Joint(n, k, xyz, (0, 0, 0), ax, lim) for n, k, xyz, ax, lim in _FETCH_ARM_JOINTS]
        joints += [Joint("gripper_axis", "fixed", (0.16645, 0, 0), (0, 0, 0))]
        super().__init__("fetch", joints)


ALL_CLCS = [Panda, Fetch, FetchArm]
_ROBOTS = {"panda": Panda, "fetch": Fetch, "fetch_arm": FetchArm}


def get_robot(robot_name: str) -> Robot:
    """jrl ``get_robot(name)`` (``ikflow/model_loading.py:81-83``)."""
    if robot_name == "rizon4":
        raise ValueError(
            "robot 'rizon4' (model 'rizon4__snowy-brook-208__global_step=2.75M', ikflow/model_descriptions.yaml:90-97) is "
            "registered but its kinematic chain is not available in ikflow_b200: the reference tree ships no URDF or test "
            "constant for the Flexiv Rizon 4 (they live in the un-vendored jrl package), so its joint limits and link "
            "transforms cannot be reproduced offline.  Pass your own chain: ikflow_b200.Robot('rizon4', [Joint(...), ...]) "
            "as the `robot` argument of get_ik_solver()."
        )
    if robot_name not in _ROBOTS:
        raise ValueError(f"Unable to find robot '{robot_name}' (available: {sorted(_ROBOTS)})")
    return _ROBOTS[robot_name]()
