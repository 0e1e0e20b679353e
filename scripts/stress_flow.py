"""Stress: the flow engine is deterministic, so N repeated calls on the same inputs must be bitwise identical; any
difference is a synchronisation bug (stale exchange data)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ikflow_b200
from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
hp = IkflowModelParameters(); hp.dim_latent_space = 7
robot = ikflow_b200.get_robot("panda")
solver = ikflow_b200.IKFlowSolver(hp, robot)
solver.load_state_dict_from_dict(make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=0))
g = torch.Generator().manual_seed(0)
latent = torch.randn(batch, 7, generator=g).cuda()
poses = robot.forward_kinematics(robot.sample_joint_angles(batch, generator=g))
flush = torch.empty(300 << 20, dtype=torch.uint8, device="cuda")
ref = solver.generate_ik_solutions(poses, latent=latent).clone()
bad = 0; worst = 0.0
for i in range(iters):
    if i % 3 == 0: flush.fill_(i & 255)
    out = solver.generate_ik_solutions(poses, latent=latent)
    if not torch.equal(out, ref):
        bad += 1; worst = max(worst, (out - ref).abs().max().item())
torch.cuda.synchronize()
print(f"batch {batch} {solver.nn_model.last_kernel().split('kernel')[-1]}: {bad}/{iters} calls differ from the first (worst {worst:.3e}); status {solver.nn_model.status()}")
