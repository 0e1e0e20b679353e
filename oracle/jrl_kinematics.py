"""CPU oracle for the jrl kinematics calls on the IKFlow hot path.

TEST INFRASTRUCTURE -- never imported by the product.  jrl (git pin ``2ba7c399`` at
``pyproject.toml:22``) is not vendored in ``/root/reference``; this file restates its batched
forward kinematics, geometric Jacobian, Levenberg-Marquardt step, joint-limit clamp and quaternion
geodesic distance as they are *called* by the reference (``ikflow/ikflow_solver.py:102,114,116,
205,208``; ``ikflow/evaluation_utils.py:86,96``).  PINNED by the reference's known-answer tests
(``tests/evaluation_utils_test.py:20-32``, ``tests/model_test.py:18-44``) -- see
``tests/test_oracle_kats.py``.

The op sequence deliberately mirrors jrl's (a Python loop over chain links issuing small batched
ops, ``torch.linalg.solve`` for the 7x7 systems) so that timing this file is a fair stand-in for the
reference's own CPU/GPU cost.  The robot constants here are an independent copy of the ones the
product uses.
"""

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch


@dataclass(frozen=True)
class ChainJoint:
    name: str
    kind: str  # "revolute" | "prismatic" | "fixed"
    xyz: Tuple[float, float, float]
    rpy: Tuple[float, float, float]
    axis: Tuple[float, float, float] = (0.0, 0.0, 1.0)
    limits: Optional[Tuple[float, float]] = None


@dataclass(frozen=True)
class ChainRobot:
    name: str
    joints: Tuple[ChainJoint, ...]

    @property
    def actuated(self) -> List[ChainJoint]:
        return [j for j in self.joints if j.kind != "fixed"]

    @property
    def ndof(self) -> int:
        return len(self.actuated)

    @property
    def actuated_joints_limits(self) -> List[Tuple[float, float]]:
        return [j.limits for j in self.actuated]


_HP = math.pi / 2

# Franka Panda, base -> panda_hand.  Limits: reference tests/model_test.py:18-25.  Chain verified against
# the golden FK vector of reference tests/evaluation_utils_test.py:20-24.
PANDA = ChainRobot(
    "panda",
    (
        ChainJoint("panda_joint1", "revolute", (0, 0, 0.333), (0, 0, 0), limits=(-2.8973, 2.8973)),
        ChainJoint("panda_joint2", "revolute", (0, 0, 0), (-_HP, 0, 0), limits=(-1.7628, 1.7628)),
        ChainJoint("panda_joint3", "revolute", (0, -0.316, 0), (_HP, 0, 0), limits=(-2.8973, 2.8973)),
        ChainJoint("panda_joint4", "revolute", (0.0825, 0, 0), (_HP, 0, 0), limits=(-3.0718, -0.0698)),
        ChainJoint("panda_joint5", "revolute", (-0.0825, 0.384, 0), (-_HP, 0, 0), limits=(-2.8973, 2.8973)),
        ChainJoint("panda_joint6", "revolute", (0, 0, 0), (_HP, 0, 0), limits=(-0.0175, 3.7525)),
        ChainJoint("panda_joint7", "revolute", (0.088, 0, 0), (_HP, 0, 0), limits=(-2.8973, 2.8973)),
        ChainJoint("panda_joint8", "fixed", (0, 0, 0.107), (0, 0, 0)),
        ChainJoint("panda_hand_joint", "fixed", (0, 0, 0), (0, 0, -math.pi / 4)),
    ),
)

# Fetch arm (torso fixed), base_link -> gripper_link.  SYNTHETIC CONSTANTS: the reference tree holds no URDF or
# test constant for this robot; values are the public Fetch URDF as recalled, continuous
# joints limited to +-pi as jrl does.  Only the 7 limit pairs matter for the flow (M_inv + clamp).
FETCH_ARM = ChainRobot(
    "fetch_arm",
    (
        ChainJoint("torso_lift_joint", "fixed", (-0.086875, 0, 0.37743), (0, 0, 0)),
        ChainJoint("shoulder_pan_joint", "revolute", (0.119525, 0, 0.34858), (0, 0, 0), (0, 0, 1), (-1.6056, 1.6056)),
        ChainJoint("shoulder_lift_joint", "revolute", (0.117, 0, 0.06), (0, 0, 0), (0, 1, 0), (-1.221, 1.518)),
        ChainJoint("upperarm_roll_joint", "revolute", (0.219, 0, 0), (0, 0, 0), (1, 0, 0), (-math.pi, math.pi)),
        ChainJoint("elbow_flex_joint", "revolute", (0.133, 0, 0), (0, 0, 0), (0, 1, 0), (-2.251, 2.251)),
        ChainJoint("forearm_roll_joint", "revolute", (0.197, 0, 0), (0, 0, 0), (1, 0, 0), (-math.pi, math.pi)),
        ChainJoint("wrist_flex_joint", "revolute", (0.1245, 0, 0), (0, 0, 0), (0, 1, 0), (-2.16, 2.16)),
        ChainJoint("wrist_roll_joint", "revolute", (0.1385, 0, 0), (0, 0, 0), (1, 0, 0), (-math.pi, math.pi)),
        ChainJoint("gripper_axis", "fixed", (0.16645, 0, 0), (0, 0, 0)),
    ),
)

ROBOTS = {"panda": PANDA, "fetch_arm": FETCH_ARM}


# ----------------------------------------------------------------------------------------------------------------------
# small math helpers


def _rpy_matrix(rpy, dtype) -> torch.Tensor:
    """URDF fixed-axis rpy: R = Rz(yaw) @ Ry(pitch) @ Rx(roll)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    rx = torch.tensor([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=torch.float64)
    ry = torch.tensor([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=torch.float64)
    rz = torch.tensor([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], dtype=torch.float64)
    return (rz @ ry @ rx).to(dtype)


def _fixed_transform(joint: ChainJoint, dtype, device) -> torch.Tensor:
    t = torch.eye(4, dtype=dtype)
    t[:3, :3] = _rpy_matrix(joint.rpy, dtype)
    t[:3, 3] = torch.tensor(joint.xyz, dtype=dtype)
    return t.to(device)


def _axis_angle_transform(axis, angle: torch.Tensor) -> torch.Tensor:
    """[m] angles about a fixed unit axis -> [m,4,4] homogeneous rotation (Rodrigues)."""
    m = angle.shape[0]
    ax = torch.tensor(axis, dtype=angle.dtype, device=angle.device)
    ax = ax / ax.norm()
    kx, ky, kz = ax[0], ax[1], ax[2]
    c, s = torch.cos(angle), torch.sin(angle)
    v = 1.0 - c
    t = torch.zeros(m, 4, 4, dtype=angle.dtype, device=angle.device)
    t[:, 0, 0] = kx * kx * v + c
    t[:, 0, 1] = kx * ky * v - kz * s
    t[:, 0, 2] = kx * kz * v + ky * s
    t[:, 1, 0] = ky * kx * v + kz * s
    t[:, 1, 1] = ky * ky * v + c
    t[:, 1, 2] = ky * kz * v - kx * s
    t[:, 2, 0] = kz * kx * v - ky * s
    t[:, 2, 1] = kz * ky * v + kx * s
    t[:, 2, 2] = kz * kz * v + c
    t[:, 3, 3] = 1.0
    return t


def _translation_transform(axis, dist: torch.Tensor) -> torch.Tensor:
    m = dist.shape[0]
    t = torch.eye(4, dtype=dist.dtype, device=dist.device).repeat(m, 1, 1)
    ax = torch.tensor(axis, dtype=dist.dtype, device=dist.device)
    t[:, :3, 3] = ax[None, :] * dist[:, None]
    return t


def rotation_matrix_to_quaternion(rot: torch.Tensor) -> torch.Tensor:
    """[m,3,3] -> [m,4] wxyz.  Branch on the best-conditioned of the four candidates (as jrl / pytorch3d)."""
    m00, m01, m02 = rot[:, 0, 0], rot[:, 0, 1], rot[:, 0, 2]
    m10, m11, m12 = rot[:, 1, 0], rot[:, 1, 1], rot[:, 1, 2]
    m20, m21, m22 = rot[:, 2, 0], rot[:, 2, 1], rot[:, 2, 2]
    q_abs = torch.sqrt(
        torch.clamp(
            torch.stack(
                [1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=1
            ),
            min=0.0,
        )
    )
    cands = torch.stack(
        [
            torch.stack([q_abs[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=1),
            torch.stack([m21 - m12, q_abs[:, 1] ** 2, m10 + m01, m02 + m20], dim=1),
            torch.stack([m02 - m20, m10 + m01, q_abs[:, 2] ** 2, m12 + m21], dim=1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[:, 3] ** 2], dim=1),
        ],
        dim=1,
    )  # [m, 4(candidate), 4]
    cands = cands / (2.0 * torch.clamp(q_abs, min=0.1))[:, :, None]
    best = torch.argmax(q_abs, dim=1)
    return cands[torch.arange(rot.shape[0], device=rot.device), best]


def quaternion_inverse(q: torch.Tensor) -> torch.Tensor:
    """Conjugate / |q|^2 (wxyz)."""
    conj = torch.cat([q[:, 0:1], -q[:, 1:4]], dim=1)
    return conj / (q * q).sum(dim=1, keepdim=True)


def quaternion_product(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Hamilton product a (x) b, wxyz."""
    w1, x1, y1, z1 = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    w2, x2, y2, z2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    return torch.stack(
        [
            w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
            w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
            w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
            w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
        ],
        dim=1,
    )


def quaternion_to_rpy(q: torch.Tensor) -> torch.Tensor:
    """wxyz -> [roll, pitch, yaw]."""
    q0, q1, q2, q3 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    roll = torch.atan2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 * q1 + q2 * q2))
    pitch = torch.asin(torch.clamp(2 * (q0 * q2 - q3 * q1), -1.0, 1.0))
    yaw = torch.atan2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 * q2 + q3 * q3))
    return torch.stack([roll, pitch, yaw], dim=1)


def geodesic_distance_between_quaternions(
    q1: torch.Tensor, q2: torch.Tensor, acos_epsilon: Optional[float] = None
) -> torch.Tensor:
    """jrl ``math_utils.geodesic_distance_between_quaternions`` (called ``ikflow/ikflow_solver.py:116``).

    ``d = 2*acos(clamp(<q1,q2>, -1+eps, 1-eps))`` wrapped into [0, pi].  Pinned at one point by
    ``tests/evaluation_utils_test.py:28-32``.
    """
    eps = 1e-7 if acos_epsilon is None else acos_epsilon
    dot = torch.clamp((q1 * q2).sum(dim=1), -1.0 + eps, 1.0 - eps)
    d = 2.0 * torch.acos(dot)
    return torch.abs(torch.remainder(d + math.pi, 2.0 * math.pi) - math.pi)


# ----------------------------------------------------------------------------------------------------------------------
# robot functions


def clamp_to_joint_limits(robot: ChainRobot, q: torch.Tensor) -> torch.Tensor:
    """jrl ``Robot.clamp_to_joint_limits``: per-column clamp, in place (``ikflow/ikflow_solver.py:102``)."""
    for i, (lo, hi) in enumerate(robot.actuated_joints_limits):
        q[:, i] = torch.clamp(q[:, i], lo, hi)
    return q


def _chain_transforms(robot: ChainRobot, q: torch.Tensor):
    """Walk the chain; returns (T_ee [m,4,4], list of (axis_world [m,3], origin_world [m,3], kind))."""
    m = q.shape[0]
    t = torch.eye(4, dtype=q.dtype, device=q.device).repeat(m, 1, 1)
    frames = []
    k = 0
    for joint in robot.joints:
        t = torch.matmul(t, _fixed_transform(joint, q.dtype, q.device))
        if joint.kind == "fixed":
            continue
        ax = torch.tensor(joint.axis, dtype=q.dtype, device=q.device)
        ax = ax / ax.norm()
        frames.append((torch.matmul(t[:, :3, :3], ax), t[:, :3, 3].clone(), joint.kind))
        if joint.kind == "revolute":
            t = torch.bmm(t, _axis_angle_transform(joint.axis, q[:, k]))
        else:
            t = torch.bmm(t, _translation_transform(joint.axis, q[:, k]))
        k += 1
    return t, frames


def forward_kinematics(robot: ChainRobot, q: torch.Tensor) -> torch.Tensor:
    """jrl ``Robot.forward_kinematics(q[m,ndof]) -> [m,7]`` = [x, y, z, qw, qx, qy, qz] (``ikflow_solver.py:114``)."""
    assert q.shape[1] == robot.ndof
    t, _ = _chain_transforms(robot, q)
    return torch.cat([t[:, :3, 3], rotation_matrix_to_quaternion(t[:, :3, :3])], dim=1)


def jacobian(robot: ChainRobot, q: torch.Tensor) -> torch.Tensor:
    """jrl ``Robot.jacobian(q) -> [m,6,ndof]``: geometric Jacobian in the base frame, rows 0-2 angular, 3-5 linear."""
    t, frames = _chain_transforms(robot, q)
    p_ee = t[:, :3, 3]
    jac = torch.zeros(q.shape[0], 6, robot.ndof, dtype=q.dtype, device=q.device)
    for k, (axis_w, origin_w, kind) in enumerate(frames):
        if kind == "revolute":
            jac[:, 0:3, k] = axis_w
            jac[:, 3:6, k] = torch.cross(axis_w, p_ee - origin_w, dim=1)
        else:
            jac[:, 3:6, k] = axis_w
    return jac


def lm_step(
    robot: ChainRobot, target_poses: torch.Tensor, q: torch.Tensor, lambd: float = 1e-4, clamp: bool = True
) -> torch.Tensor:
    """jrl ``Robot.inverse_kinematics_step_levenburg_marquardt`` (``ikflow/ikflow_solver.py:205,208``).

    ``e = [rpy(q_target (x) q_cur^-1); p_target - p_cur]``, ``dq = solve(J^T J + lambd I, J^T e)``,
    ``q <- clamp(q + dq)``.
    """
    m = q.shape[0]
    jac = jacobian(robot, q)
    jac_t = jac.transpose(1, 2)
    cur = forward_kinematics(robot, q)
    err = torch.zeros(m, 6, 1, dtype=q.dtype, device=q.device)
    err[:, 3:6, 0] = target_poses[:, 0:3] - cur[:, 0:3]
    rot_err = quaternion_product(target_poses[:, 3:7], quaternion_inverse(cur[:, 3:7]))
    err[:, 0:3, 0] = quaternion_to_rpy(rot_err)
    eye = torch.eye(robot.ndof, dtype=q.dtype, device=q.device)[None]
    lhs = torch.bmm(jac_t, jac) + lambd * eye
    rhs = torch.bmm(jac_t, err)
    dq = torch.linalg.solve(lhs, rhs)
    out = q + dq[:, :, 0]
    if clamp:
        out = clamp_to_joint_limits(robot, out)
    return out


def pose_error(robot: ChainRobot, q: torch.Tensor, target_poses: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``IKFlowSolver._calculate_pose_error`` (``ikflow/ikflow_solver.py:112-117``)."""
    realized = forward_kinematics(robot, q)
    pos = torch.norm(realized[:, 0:3] - target_poses[:, 0:3], dim=1)
    rot = geodesic_distance_between_quaternions(target_poses[:, 3:], realized[:, 3:])
    return pos, rot


def calculate_joint_limits_exceeded(configs: torch.Tensor, joint_limits) -> torch.Tensor:
    """``ikflow/evaluation_utils.py:100-112``."""
    hi = torch.tensor([x[1] for x in joint_limits], dtype=torch.float32, device=configs.device)
    lo = torch.tensor([x[0] for x in joint_limits], dtype=torch.float32, device=configs.device)
    return torch.logical_or(configs > hi, configs < lo).any(dim=1)


def sample_joint_angles_and_poses(robot: ChainRobot, n: int, seed: int, dtype=torch.float32):
    """Uniform joint samples inside the limits + FK (jrl ``sample_joint_angles_and_poses`` minus the
    self-collision filter, which needs klampt geometry)."""
    g = torch.Generator().manual_seed(seed)
    lims = torch.tensor(robot.actuated_joints_limits, dtype=torch.float64)
    u = torch.rand(n, robot.ndof, generator=g, dtype=torch.float64)
    q = (lims[:, 0] + u * (lims[:, 1] - lims[:, 0])).to(dtype)
    return q, forward_kinematics(robot, q)


def joint_angles_from_uniforms(robot: ChainRobot, u: torch.Tensor, joint_limit_eps: float = 1e-6) -> torch.Tensor:
    """The affine map of jrl's sampler applied to given uniforms ``u`` [n, ndof] in (0, 1), fp64:
    ``q = lo' + u * (hi' - lo')`` with ``lo' = lo + eps``, ``hi' = hi - eps``, rounded to fp32 once.  The product's
    device-side sampler (``ikf_sample_joint_angles_and_poses``) is checked against this with the uniforms of
    ``oracle/philox.py``."""
    lims = torch.tensor(robot.actuated_joints_limits, dtype=torch.float64)
    lo, hi = lims[:, 0] + joint_limit_eps, lims[:, 1] - joint_limit_eps
    return (lo + u.double() * (hi - lo)).to(torch.float32)
