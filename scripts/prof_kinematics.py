"""Tiny driver for ncu / timing of the kinematics kernels at BASELINE config 3 sizes (m = n * r = 2048 * 10 rows)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ikflow_b200

m = int(sys.argv[1]) if len(sys.argv) > 1 else 20480
robot = ikflow_b200.get_robot("panda")
q_true, poses = robot.sample_joint_angles_and_poses(m, seed=3, return_torch=True, device="cuda")
q = robot.clamp_to_joint_limits((q_true + 0.05 * torch.randn(m, 7, generator=torch.Generator().manual_seed(1)).cuda()).contiguous())


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


print(f"m = {m}: lm_step {timed(lambda: robot.inverse_kinematics_step_levenburg_marquardt(poses, q)):.1f} us, "
      f"pose_errors {timed(lambda: robot.pose_errors(q, poses)):.1f} us, forward_kinematics {timed(lambda: robot.forward_kinematics(q)):.1f} us, "
      f"lm_refine (n = {m // 10}, r = 10, 3 steps) {timed(lambda: robot.lm_refine(poses[: m // 10], q, 10, 3, 1e-3, 1e-2)):.1f} us")
