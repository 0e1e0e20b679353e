"""Small driver for compute-sanitizer: tiny model (hidden 256, 3 blocks), 1024-wide models at small batches (k-split
kernel) and at 2400 rows (ping-pong kernel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ikflow_b200
from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict
robot = ikflow_b200.get_robot("panda")
for nb, w, cfg, hid, batch in ((3, 9, 2, 256, 70), (2, 7, 3, 1024, 40), (1, 7, 3, 1024, 200), (1, 7, 3, 1024, 1000), (1, 7, 3, 1024, 2400)):
    hp = IkflowModelParameters()
    hp.nb_nodes, hp.dim_latent_space, hp.coeff_fn_config, hp.coeff_fn_internal_size = nb, w, cfg, hid
    s = ikflow_b200.IKFlowSolver(hp, robot)
    s.load_state_dict_from_dict(make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=0))
    g = torch.Generator().manual_seed(0)
    poses = robot.forward_kinematics(robot.sample_joint_angles(batch, generator=g))
    q = s.generate_ik_solutions(poses, latent=torch.randn(batch, w, generator=g).cuda())
    kernel = s.nn_model.last_kernel()
    sol, valid = s.generate_exact_ik_solutions(poses[:16], repeat_counts=(1, 3))
    torch.cuda.synchronize()
    print("ok", nb, hid, batch, float(q.abs().max()), s.nn_model.status(), kernel)
