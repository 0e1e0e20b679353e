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
nl = 120
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


print("per-chunk stamps, layer 1 of CTA 0 (us): loader [reach, stage free, issued(+8=W only)]  mma [reach, full, committed]")
for i in range(16):
    r = sa[0, 100 + i]
    f = lambda v: (int(v) - t0) / 1000.0 if v > 0 else float("nan")
    print(f"chunk {i:2d}: L {f(r[0]):7.2f} {f(r[1]):7.2f} {f(r[2]):7.2f} {f(r[10]):7.2f} | M {f(r[3]):7.2f} {f(r[4]):7.2f} {f(r[5]):7.2f}")
