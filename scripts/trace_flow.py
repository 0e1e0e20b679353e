"""Developer aid: per-layer timeline of CTA 0 of the flow kernel (globaltimer stamps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ikflow_b200
from ikflow_b200 import _lib
from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
hp = IkflowModelParameters()
hp.dim_latent_space = 7
robot = ikflow_b200.get_robot("panda")
solver = ikflow_b200.IKFlowSolver(hp, robot)
solver.load_state_dict_from_dict(make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=0))
g = torch.Generator().manual_seed(0)
latent = torch.randn(batch, 7, generator=g).cuda()
poses = robot.forward_kinematics(robot.sample_joint_angles(batch, generator=g))
for _ in range(3):
    solver.generate_ik_solutions(poses, latent=latent)
torch.cuda.synchronize()
nl = 24
NT = 16
stamps = torch.zeros(NT * nl * 16, dtype=torch.int64, device="cuda")
h = solver.nn_model._handle(torch.device("cuda", 0))
_lib.check(_lib.lib().ikf_flow_debug_trace(h, stamps.data_ptr(), nl), "trace")
solver.generate_ik_solutions(poses, latent=latent)
torch.cuda.synchronize()
_lib.lib().ikf_flow_debug_trace(h, None, 0)
sa = stamps.cpu().view(NT, nl, 16)
t0 = int(sa[sa > 0].min())
s = sa[0]
names = ["W0 issue", "A first", "A last", "st:sync", "st:done", "st:flag", "c:start", "c:full0", "c:mma end", "c:v ready", "c:staged", "p:start", "p:flags", "p:coupled", "p:dot", "p:myflag"]
print("layer " + " ".join(f"{n:>10s}" for n in names))
for i in range(8):
    row = [(int(v) - t0) / 1000.0 if v > 0 else float("nan") for v in s[i, :16]]
    print(f"{i:5d} " + " ".join(f"{v:10.2f}" for v in row))

print("per-CTA stamps of selected events (us):")
for layer, ev, nm in [(1, 7, "full0"), (1, 8, "mma end"), (1, 5, "flag"), (3, 11, "p:start"), (3, 12, "p:flags"), (4, 10, "L0 staged")]:
    print(f"layer {layer} {nm:>10s}: " + " ".join(f"{(int(v) - t0) / 1000.0:7.2f}" for v in sa[:, layer, ev]))



print("dot phase detail (row 4g+2): vt stored, bar passed, dot loop done, shuffles+ptile done, fence done")
for i in (2, 6, 10):
    print(i, " ".join(f"{(int(v) - t0) / 1000.0:8.2f}" for v in sa[0, i, :5]))
