"""Developer aid: per-layer timeline of team 0 of the flow kernel (globaltimer stamps), after a long warm-up.

usage: python scripts/trace_flow.py [batch] [subnets traced] [nb_nodes]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ikflow_b200
from ikflow_b200 import _lib
from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 6
hp = IkflowModelParameters()
hp.dim_latent_space = 7
if len(sys.argv) > 3:
    hp.nb_nodes = int(sys.argv[3])
robot = ikflow_b200.get_robot("panda")
solver = ikflow_b200.IKFlowSolver(hp, robot)
if os.environ.get("TRACE_PRECISION"):
    solver.nn_model.precision = os.environ["TRACE_PRECISION"]
solver.load_state_dict_from_dict(make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=0))
g = torch.Generator().manual_seed(0)
latent = torch.randn(batch, 7, generator=g).cuda()
poses = robot.forward_kinematics(robot.sample_joint_angles(batch, generator=g))
for _ in range(300):
    solver.generate_ik_solutions(poses, latent=latent)
torch.cuda.synchronize()
EV = 96
nl = 4 * nsub
NT = 16
stamps = torch.zeros(NT * nl * EV, dtype=torch.int64, device="cuda")
h = solver.nn_model._handle(torch.device("cuda", 0))
_lib.check(_lib.lib().ikf_flow_debug_trace(h, stamps.data_ptr(), nl), "trace")
solver.generate_ik_solutions(poses, latent=latent)
torch.cuda.synchronize()
_lib.lib().ikf_flow_debug_trace(h, None, 0)
sa = stamps.cpu().view(NT, nl, EV)
t0 = int(sa[:, :, :16][sa[:, :, :16] > 0].min())
T = 16 if os.environ.get('TRACE_PAIRS') else 8


def us(v):
    return (int(v) - t0) / 1000.0 if v > 0 else float("nan")


names = ["W0 issue", "A first", "A last", "flag seen", "st:done", "st:flag", "c:start", "c:full0", "c:mma end", "c:v ready", "c:staged", "p:start", "p:flags", "p:coupled", "p:dot", "p:myflag"]
print("CTA 0:")
print("layer " + " ".join(f"{n:>9s}" for n in names))
for i in range(nl):
    print(f"{i:5d} " + " ".join(f"{us(v):9.2f}" for v in sa[0, i, :16]))

if os.environ.get("TRACE_PAIRS"):
    print("k-split hand-over, CTA 0, us after the layer's accumulators were complete (c:full0): peer rows loaded / sent / own rows loaded / peer's sums arrived / v ready")
    for i in range(4, nl):
        if i % 4 < 2:
            b0 = int(sa[0, i, 7])
            print(f"  layer {i}: " + " ".join(f"{(int(sa[0, i, e]) - b0) / 1000.0:6.2f}" for e in (11, 12, 13, 14, 9)))
print("first layer, CTA 0, us after the previous subnet's coupling: input staged / tiles computed+stored / barrier passed / flag released")
for i in range(4, nl, 4):
    b0 = int(sa[0, i - 1, 13])
    print(f"  layer {i}: " + " ".join(f"{(int(sa[0, i, e]) - b0) / 1000.0:6.2f}" for e in (2, 3, 4, 5)))
print("per-subnet duration (coupled -> coupled, CTA 0, us): " + " ".join(f"{(int(sa[0, 4 * s_ + 3, 13]) - int(sa[0, 4 * s_ - 1, 13])) / 1000.0:5.1f}" for s_ in range(1, nsub)))
print("\nper-CTA stamps (us), subnets 1..:")
for sub in range(1, min(nsub, 6)):
    b = 4 * sub
    for layer, ev, nm in [(b, 10, "L0 staged"), (b, 5, "L0 flag"), (b, 1, "H0 A first"), (b, 2, "H0 A last"), (b, 8, "H0 mma end"), (b + 1, 10, "H0 staged"),
                          (b + 1, 1, "H1 A first"), (b + 1, 2, "H1 A last"), (b + 1, 8, "H1 mma end"), (b + 3, 11, "p:start"), (b + 3, 14, "p:dot"), (b + 3, 15, "p:myflag"),
                          (b + 3, 12, "p:flags"), (b + 3, 13, "p:coupled")]:
        print(f"sub {sub} {nm:>10s}: " + " ".join(f"{us(v):7.2f}" for v in sa[:T, layer, ev]))
    print()

print("per k-chunk SM-clock stamps (cycles, relative to the layer's first landed chunk): MMA warp saw the stage full / loader saw the stage free / loader issued the copies")
for cta in (0, 3):
    for layer in (4, 5, 8, 9) if os.environ.get('IKFLOW_B200_DEBUG') == '4' else ():
        if layer < nl:
            base = int(sa[cta, layer, 16])
            print(f"cta {cta} layer {layer}:\n   landed " + " ".join(f"{int(v) - base:6d}" for v in sa[cta, layer, 16:32]))
            print("   free   " + " ".join(f"{int(v) - base:6d}" for v in sa[cta, layer, 32:48]))
            print("   expect " + " ".join(f"{int(v) - base:6d}" for v in sa[cta, layer, 64:80]))
            print("   W sent " + " ".join(f"{int(v) - base:6d}" for v in sa[cta, layer, 80:96]))
            print("   issued " + " ".join(f"{int(v) - base:6d}" for v in sa[cta, layer, 48:64]))

if os.environ.get('IKFLOW_B200_DEBUG') == '4':
    print("just-in-time first layer, CTA 0, SM-clock cycles relative to the MMA warp's first chunk of the layer; per group (0: epilogue warps -> even chunks, 1: helpers -> odd chunks): iteration start / waits passed / math+stores issued / fence+barrier passed")
    for sub in (1, 2):
        base = int(sa[0, 4 * sub, 16])  # MMA warp: chunk 0 landed
        print(f"  subnet {sub}: MMA saw chunk i full at " + " ".join(f"{int(v) - base:6d}" for v in sa[0, 4 * sub, 16:32]))
        for grp in (0, 1):
            for nm, lo in (("start", 16), ("ready", 32), ("stored", 48), ("bar", 64)):
                print(f"    group {grp} {nm:>7s} " + " ".join(f"{int(v) - base:6d}" for v in sa[0, 4 * sub + 2, lo + 8 * grp:lo + 8 * grp + 8]))

# phase summary, averaged over subnets 1.. and the CTAs of the team
import statistics as st
rows = []
for sub in range(1, nsub - 1):
    b = 4 * sub
    for c in range(T):
        e = lambda l, v: int(sa[c, l, v])
        prev_coupled = int(sa[c, b - 1, 13])
        nxt_coupled = e(b + 3, 13)
        rows.append(dict(total=nxt_coupled - prev_coupled, first=e(b, 10) - prev_coupled, x1=e(b, 1) - e(b, 10), h0=e(b, 8) - e(b, 1), pub=e(b + 1, 10) - e(b, 8),
                         x2=e(b + 1, 1) - e(b + 1, 10), h1=e(b + 1, 8) - e(b + 1, 1), epi=e(b + 3, 11) - e(b + 1, 8), dot=e(b + 3, 14) - e(b + 3, 11),
                         myflag=e(b + 3, 15) - e(b + 3, 14), rel0=e(b, 5) - e(b, 4), rel1=e(b + 1, 5) - e(b + 1, 4), wait=e(b + 3, 12) - e(b + 3, 15), couple=e(b + 3, 13) - e(b + 3, 12)))
print("\nphase means over subnets 1..%d x %d CTAs (us):" % (nsub - 2, T))
for k in rows[0]:
    print(f"  {k:>7s} {st.mean(r[k] for r in rows) / 1000.0:6.2f}")

# kernel time with CUDA events: back to back, and with an L2 flush (256 MB fill) before every call as bench.py does
def timed(flush_buf, n=100):
    tot = 0.0
    for _ in range(n):
        if flush_buf is not None:
            flush_buf.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        solver.generate_ik_solutions(poses, latent=latent)
        b.record()
        b.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1000.0


fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
small_fl = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
print(f"\ncall time (us): back to back {timed(None):.1f}, after a 1 MB fill {timed(small_fl):.1f}, after a 256 MB fill {timed(fl):.1f}")
