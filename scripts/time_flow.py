"""Same-box A/B timing of the flow kernel: p50 / mean CUDA-event time of `generate_ik_solutions` per batch size, with an
L2 flush before every call.  Configuration through the developer switches (IKFLOW_B200_CLUSTER, _RT, _JIT, _PRECISION,
_LIB, ...), one process per configuration:

    IKFLOW_B200_CLUSTER=2 python scripts/time_flow.py panda__full__lp191_5.25m 512 8192
"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("IKFLOW_B200_OFFLINE", "1")
import ikflow_b200  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "panda__full__lp191_5.25m"
batches = [int(b) for b in sys.argv[2:]] or [512]
solver, hp = ikflow_b200.get_ik_solver(model, synthetic_seed=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tag = {k: v for k, v in os.environ.items() if k.startswith("IKFLOW_B200_") and k != "IKFLOW_B200_OFFLINE"}
ref = {}
for b in batches:
    q, poses = solver.robot.sample_joint_angles_and_poses(b, seed=1, return_torch=True, device="cuda")
    latent = torch.randn(b, solver.network_width, generator=torch.Generator().manual_seed(2)).cuda()
    for _ in range(10):
        out = solver.generate_ik_solutions(poses, latent=latent)
    torch.cuda.synchronize()
    evs = []
    for _ in range(200 if b <= 2048 else 60):
        flush.fill_(1)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = solver.generate_ik_solutions(poses, latent=latent)
        e.record()
        evs.append((a, e))
    torch.cuda.synchronize()
    ts = [a.elapsed_time(e) for a, e in evs]
    print(json.dumps({"cfg": tag, "model": model, "batch": b, "p50_ms": round(statistics.median(ts), 4), "mean_ms": round(statistics.fmean(ts), 4),
                      "min_ms": round(min(ts), 4), "solutions_per_s": round(b / (statistics.fmean(ts) * 1e-3)), "kernel": solver.nn_model.last_kernel().split("::")[-1],
                      "cluster": solver.nn_model.last_cluster(), "grid": solver.nn_model.info()["grid_ctas_last"], "status": solver.nn_model.status(),
                      "checksum": float(out.double().sum())}))
