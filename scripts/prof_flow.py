"""Tiny driver for ncu: a few inverse-flow calls at one batch size (synthetic panda weights)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ikflow_b200
from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hp = IkflowModelParameters()
hp.dim_latent_space = 7
robot = ikflow_b200.get_robot("panda")
solver = ikflow_b200.IKFlowSolver(hp, robot)
solver.load_state_dict_from_dict(make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=0))
g = torch.Generator().manual_seed(0)
latent = torch.randn(batch, 7, generator=g).cuda()
poses = robot.forward_kinematics(robot.sample_joint_angles(batch, generator=g))
for _ in range(calls):
    q = solver.generate_ik_solutions(poses, latent=latent)
torch.cuda.synchronize()
print("ok", q.shape, solver.nn_model.status())
