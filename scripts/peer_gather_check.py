"""Multi-GPU check of the fused gather (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/peer_gather_check.py

Compares PeerGather.generate_ik_solutions with generate_ik_solutions + NCCL all-gather bit for bit over many steps (even and
ragged shards), then times both (CUDA events, max over ranks)."""
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("IKFLOW_B200_OFFLINE", "1")
import ikflow_b200  # noqa: E402
from ikflow_b200.distributed import PeerGather, all_gather_rows, shard_bounds  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
solver, hp = ikflow_b200.get_ik_solver("panda__full__lp191_5.25m", synthetic_seed=0)
report = {"world": world}
for n_total in (512 * world, 100 * world + 1, 2400 * world + 3):  # the last one: ping-pong kernel (> 2304 rows per GPU), ragged
    lo, hi = shard_bounds(n_total, rank, world)
    q, poses = solver.robot.sample_joint_angles_and_poses(n_total, seed=5, return_torch=True, device=dev)  # same on every rank
    pg = PeerGather(solver, n_total)
    bad = 0
    for step in range(60 if n_total <= 512 * world else 12):
        latent = torch.randn(n_total, 7, generator=torch.Generator().manual_seed(step)).to(dev)
        fused = pg.generate_ik_solutions(poses[lo:hi], latent[lo:hi]).clone()
        ref = all_gather_rows(solver.generate_ik_solutions(poses[lo:hi], hi - lo, latent=latent[lo:hi]), n_total)
        bad += int(not torch.equal(fused, ref))
    torch.cuda.synchronize()
    t = torch.tensor([bad], device=dev)
    dist.all_reduce(t)
    report[f"mismatching_steps_n{n_total}"] = int(t.item())
    if n_total == 512 * world:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        latent = torch.randn(n_total, 7, generator=torch.Generator().manual_seed(0)).to(dev)

        def timed(fn, n=200):
            for _ in range(10):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            evs = []
            for _ in range(n):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                evs.append((a, b))
            dist.barrier()
            torch.cuda.synchronize()
            ts = torch.tensor([a.elapsed_time(b) for a, b in evs], dtype=torch.float64, device=dev)
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            return float(ts.median()), float(ts.mean())

        report["fused_ms_p50_mean"] = timed(lambda: pg.generate_ik_solutions(poses[lo:hi], latent[lo:hi]))
        report["nccl_ms_p50_mean"] = timed(lambda: all_gather_rows(solver.generate_ik_solutions(poses[lo:hi], hi - lo, latent=latent[lo:hi]), n_total))
        report["local_only_ms_p50_mean"] = timed(lambda: solver.generate_ik_solutions(poses[lo:hi], hi - lo, latent=latent[lo:hi]))
    pg.close()
report["status"] = solver.nn_model.status()
if rank == 0:
    print(json.dumps(report))
dist.destroy_process_group()
