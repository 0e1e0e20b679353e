"""Solution evaluation -- ``ikflow/evaluation_utils.py`` of the reference, minus the klampt self-collision check.

``evaluate_solutions`` keeps the reference's 4-tuple ``(l2_errors, angular_errors, joint_limits_exceeded,
self_collisions)``; the last entry is all-False with a one-time warning (capsule geometry is outside this package).
Pose errors and the joint-limit flag come from one fused launch (``ikf_evaluate_solutions``).
"""

import warnings
from typing import List, Optional, Tuple

import torch

SOLUTION_EVALUATION_RESULT_TYPE = Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, float]

_warned_self_collision = False


def _get_target_pose_batch(target_pose: torch.Tensor, n_solutions: int) -> torch.Tensor:
    """[7] -> [1 x 7] (the kernels broadcast a single row), [n x 7] unchanged (``evaluation_utils.py:21-34``)."""
    if target_pose.shape[0] == 7 and target_pose.dim() == 1:
        return target_pose.reshape(1, 7)
    return target_pose


def geodesic_distance_between_quaternions(q1: torch.Tensor, q2: torch.Tensor, acos_epsilon: Optional[float] = None):
    """jrl ``math_utils.geodesic_distance_between_quaternions``: 2*acos(<q1,q2>) wrapped into [0, pi] (host-side torch
    ops; the hot path uses the fused kernel instead)."""
    eps = 1e-7 if acos_epsilon is None else acos_epsilon
    dot = torch.clamp(torch.sum(q1 * q2, dim=1), -1.0 + eps, 1.0 - eps)
    d = 2.0 * torch.acos(dot)
    return torch.abs(torch.remainder(d + torch.pi, 2 * torch.pi) - torch.pi)


def pose_errors(poses_1: torch.Tensor, poses_2: torch.Tensor, acos_epsilon: Optional[float] = None):
    """Positional and rotational error between two batches of poses (``evaluation_utils.py:37-52``)."""
    assert poses_1.shape == poses_2.shape, f"Poses are of different shape: {poses_1.shape} != {poses_2.shape}"
    l2_errors = torch.norm(poses_1[:, 0:3] - poses_2[:, 0:3], dim=1)
    angular_errors = geodesic_distance_between_quaternions(poses_1[:, 3:7], poses_2[:, 3:7], acos_epsilon=acos_epsilon)
    return l2_errors, angular_errors


def pose_errors_cm_deg(poses_1: torch.Tensor, poses_2: torch.Tensor, acos_epsilon: Optional[float] = None):
    l2_errors, angular_errors = pose_errors(poses_1, poses_2, acos_epsilon=acos_epsilon)
    return 100 * l2_errors, torch.rad2deg(angular_errors)


def solution_pose_errors(robot, solutions: torch.Tensor, target_poses: torch.Tensor):
    """L2 and angular errors of IK solutions w.r.t. their target pose(s) (``evaluation_utils.py:65-97``)."""
    assert isinstance(target_poses, torch.Tensor), f"target_poses must be a torch.Tensor (got {type(target_poses)})"
    assert isinstance(solutions, torch.Tensor), f"solutions must be a torch.Tensor (got {type(solutions)})"
    target_poses = _get_target_pose_batch(target_poses, solutions.shape[0])
    return robot.pose_errors(solutions[:, 0 : robot.ndof], target_poses)


def calculate_joint_limits_exceeded(configs: torch.Tensor, joint_limits: List[Tuple[float, float]]) -> torch.Tensor:
    """``evaluation_utils.py:100-112`` (plain torch; any device)."""
    toolarge = configs > torch.tensor([x[1] for x in joint_limits], dtype=torch.float32, device=configs.device)
    toosmall = configs < torch.tensor([x[0] for x in joint_limits], dtype=torch.float32, device=configs.device)
    return torch.logical_or(toolarge, toosmall).any(dim=1)


def evaluate_solutions(robot, target_poses: torch.Tensor, solutions: torch.Tensor):
    """``evaluation_utils.py:130-147``."""
    global _warned_self_collision
    assert isinstance(target_poses, torch.Tensor), f"target_poses must be a torch.Tensor (got {type(target_poses)})"
    assert isinstance(solutions, torch.Tensor), f"solutions must be a torch.Tensor (got {type(solutions)})"
    target_poses = _get_target_pose_batch(target_poses, solutions.shape[0])
    l2_errors, angular_errors, joint_limits_exceeded = robot.evaluate_solutions(solutions, target_poses)
    if not _warned_self_collision:
        warnings.warn("ikflow_b200 does not check self-collisions: the self_collisions entry is all False")
        _warned_self_collision = True
    # (the reference builds this one tensor on the CPU, evaluation_utils.py:124-127; here all four live with the solutions)
    self_collisions = torch.zeros(solutions.shape[0], dtype=torch.bool, device=solutions.device)
    return l2_errors, angular_errors, joint_limits_exceeded, self_collisions
