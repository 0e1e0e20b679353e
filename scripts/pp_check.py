"""Developer aid: the ping-pong kernel (two 128-row groups per CTA) against the single-group kernels and the oracle.

usage: python scripts/pp_check.py [model] [batch ...]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ikflow_b200
from ikflow_b200.model import make_synthetic_state_dict

model = sys.argv[1] if len(sys.argv) > 1 else "panda__full__lp191_5.25m"
batches = [int(a) for a in sys.argv[2:]] or [2305, 4096, 8192]


def make(pp):
    os.environ["IKFLOW_B200_PP"] = "1" if pp else "0"
    solver, hp = ikflow_b200.get_ik_solver(model, synthetic_seed=0)
    solver.nn_model._handle(torch.device("cuda", 0))  # the switches are read when the handle is created
    return solver, hp


for prec in ("bf16x3", "fp16x3"):
    os.environ["IKFLOW_B200_PRECISION"] = prec
    a, hp = make(True)
    b, _ = make(False)
    for batch in batches:
        g = torch.Generator().manual_seed(batch)
        latent = torch.randn(batch, a.network_width, generator=g).cuda()
        q = a.robot.sample_joint_angles(batch, generator=g)
        poses = a.robot.forward_kinematics(q)
        ya = a.generate_ik_solutions(poses, latent=latent)
        torch.cuda.synchronize()
        ka = a.nn_model.last_kernel()
        yb = b.generate_ik_solutions(poses, latent=latent)
        torch.cuda.synchronize()
        kb = b.nn_model.last_kernel()
        # repeatability + timing
        ts = {}
        for name, s in (("pp", a), ("ref", b)):
            for _ in range(5):
                s.generate_ik_solutions(poses, latent=latent)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(20):
                y2 = s.generate_ik_solutions(poses, latent=latent)
            ev[1].record()
            torch.cuda.synchronize()
            ts[name] = ev[0].elapsed_time(ev[1]) / 20
            assert torch.equal(y2, ya if name == "pp" else yb), "not repeatable: " + name
        print(f"{prec} B={batch}: {ka[-24:]} vs {kb[-18:]}  max|diff| {(ya - yb).abs().max().item():.3e}  finite {bool(torch.isfinite(ya).all())}  status {a.nn_model.status()} "
              f" {ts['pp']:.3f} ms vs {ts['ref']:.3f} ms ({batch / ts['pp'] / 1e3:.2f} vs {batch / ts['ref'] / 1e3:.2f} M/s)", flush=True)
print("done")
