#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json): IK solutions/s, batched approximate solve.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--model NAME] [--impl ours|reference]

A step = one pass of the hot path (IKFlowSolver.generate_ik_solutions) over one batch of synthetic target poses.
N = 1 runs BASELINE.json configs[1]: panda__full__lp191_5.25m, batch 512, approximate solve, 1 x B200.  Under torchrun
(N > 1) every rank solves its own 512 poses (weak scaling, weights replicated) and the ranks exchange one all-gather of
the joint angles per step; the timed region is bracketed by a barrier + synchronize, per-step times are the MAX over
ranks.  Rank 0 prints ONE JSON line.

Numbers:
  value        solutions/s with poses and latents already in HBM: K steps, each timed with CUDA events on the launching
               stream, an L2 flush (256 MB write) between steps outside the timed window.
  e2e          the same metric through the public API with HOST (pinned) poses: H2D copy of the poses, latent draw,
               kernel, D2H copy of the joint angles inside the timed region, host wall clock.
  roofline     algorithmic FLOPs of the flow (101,572,608 per solution for the 12-block Panda model)
               divided by the CUDA-event duration of the one kernel a step launches, against the measured bf16 peak.
  cpu_baseline the oracle restatement of the reference's torch path on the host cores, bounded sample (N = 1 only).
  --impl reference   times that CPU path alone (the reference itself cannot be installed here: FrEIA 0.2 and jrl are
               un-vendored third-party dependencies, absent from the wheelhouse -- DESIGN.md).

Weights are synthetic (seeded, reference state-dict layout): the released .pkl files live on GCS and there is no
network; timing does not depend on the weight values.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("IKFLOW_B200_OFFLINE", "1")

import torch  # noqa: E402

METRIC = "ik_solutions_per_sec"
UNIT = "solutions/s"
FLOW_FLOPS = {  # algorithmic FLOPs per solution = 2 * MACs of every nn.Linear on the reverse pass
    "panda__full__lp191_5.25m": 101_572_608,
    "fetch_arm__large__mh186_9.25m": 135_725_056,
    "panda__nb16__synthetic": 135_430_144,
}


def flops_per_solution(hp, ndim_tot: int, dim_cond: int = 8) -> int:
    s1 = ndim_tot // 2
    s2 = ndim_tot - s1
    h, nl = hp.coeff_fn_internal_size, hp.coeff_fn_config
    macs = 0
    for cin, cout in ((s1 + dim_cond, 2 * s2), (s2 + dim_cond, 2 * s1)):
        macs += cin * h + (nl - 1) * h * h + h * cout
    return 2 * macs * hp.nb_nodes


def weight_bytes(hp, ndim_tot: int, dim_cond: int = 8) -> int:
    s1 = ndim_tot // 2
    s2 = ndim_tot - s1
    h, nl = hp.coeff_fn_internal_size, hp.coeff_fn_config
    n = 0
    for cin, cout in ((s1 + dim_cond, 2 * s2), (s2 + dim_cond, 2 * s1)):
        n += cin * h + h + (nl - 1) * (h * h + h) + h * cout + cout
    return 4 * n * hp.nb_nodes


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_tflops": d["bf16_tflops"], "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


def cpu_reference_arm(model_name: str, batch: int, steps: int, warmup: int, budget_s: float):
    """The reference's own torch path restated (oracle/), on the host cores."""
    from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict
    from ikflow_b200.model_loading import MODEL_DESCRIPTIONS
    from oracle import jrl_kinematics as jk
    from oracle.solver import OracleSolver

    hp = IkflowModelParameters()
    hp.__dict__.update(MODEL_DESCRIPTIONS[model_name])
    robot = jk.ROBOTS.get(hp.robot_name, jk.PANDA)
    sd = make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=0)
    solver = OracleSolver(robot, sd, hp.nb_nodes, hp.dim_latent_space, hp.coeff_fn_config, hp.rnvp_clamp, device="cpu")
    _, poses = jk.sample_joint_angles_and_poses(jk.PANDA, batch, seed=1234)
    latent = torch.randn(batch, hp.dim_latent_space, generator=torch.Generator().manual_seed(4321))
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm is entitled to every host core
    torch.set_num_threads(max(torch.get_num_threads(), os.cpu_count() or 1))
    cores = torch.get_num_threads()
    for _ in range(warmup):
        solver.generate_ik_solutions(poses, latent=latent)
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        solver.generate_ik_solutions(poses, latent=latent)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    total = sum(times)
    return {
        "value": batch * len(times) / total,
        "ms_per_step": 1e3 * total / len(times),
        "p50_ms": 1e3 * statistics.median(times),
        "steps": len(times),
        "cores": cores,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--batch", type=int, default=512, help="target poses per GPU per step")
    ap.add_argument("--model", default="panda__full__lp191_5.25m")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16x1"])
    ap.add_argument("--mode", default="approx", choices=["approx", "exact"],
                    help="approx = generate_ik_solutions (headline); exact = generate_exact_ik_solutions (flow + LM refinement)")
    args = ap.parse_args()
    warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {
        "workload": f"{args.model}, batch={args.batch} per GPU, "
        + ("approximate solve (generate_ik_solutions)" if args.mode == "approx" else "generate_exact_ik_solutions (flow + LM refine, repeat_counts (1,3,10), 1 mm / 0.01 rad)")
        + ", synthetic seeded weights",
        "batch_per_gpu": args.batch,
        "global_batch": args.batch * world,
        "parallelism": f"batch-sharded x{world}, weights replicated, one all-gather of the joint angles per step" if world > 1 else "single GPU",
    }

    # ------------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        steps = min(args.steps, 200)
        r = cpu_reference_arm(args.model, args.batch, steps, min(warmup, 3), budget_s=120.0)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": min(warmup, 3), "ms_per_step": r["ms_per_step"], "p50_latency_ms": r["p50_ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": f"{r['steps']} calls of the oracle restatement of the reference torch path at batch {args.batch} on CPU"},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = oracle/ port of FrEIA-0.2 + jrl ops on torch-CPU; the upstream package cannot be installed offline",
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------------------------------
    assert torch.cuda.is_available(), "bench.py (impl ours) needs a GPU: ikflow_b200 has no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import ikflow_b200
    from ikflow_b200 import _lib
    from ikflow_b200.distributed import all_gather_rows

    solver, hp = ikflow_b200.get_ik_solver(args.model, synthetic_seed=0)
    if args.precision != "bf16x3":
        solver.nn_model.precision = args.precision
    robot = solver.robot
    width = solver.network_width
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    q_true = robot.sample_joint_angles(B, generator=g, device=dev)
    poses = robot.forward_kinematics(q_true)
    latent = torch.randn(B, width, generator=g).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB of L2

    def step_resident():
        if args.mode == "exact":  # BASELINE.json configs[2]: thresholds of scripts/benchmark_generate_exact_solutions.py:18-19
            local, _valid = solver.generate_exact_ik_solutions(poses, pos_error_threshold=1e-3, rot_error_threshold=1e-2)
        else:
            local = solver.generate_ik_solutions(poses, latent=latent)
        return all_gather_rows(local, B * world) if world > 1 else local

    for _ in range(warmup):
        out = step_resident()
    torch.cuda.synchronize()
    status = solver.nn_model.status()
    assert status == 0, f"flow engine status {status}"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: per-step CUDA events, L2 flush between steps (outside the timed window) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = _lib.launch_count()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = step_resident()
        b.record()
        evs.append((a, b))
    barrier()
    launches = _lib.launch_count() - launches0
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in evs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
    step_ms = step_ms.cpu()
    total_ms = float(step_ms.sum())
    value = B * world * args.steps / (total_ms * 1e-3)
    p50 = float(step_ms.median())

    # ---- end to end through the public API with host buffers ----
    poses_host = poses.cpu().pin_memory()
    out_host = torch.empty((B, solver.ndof), dtype=torch.float32).pin_memory()

    def step_e2e():
        y = poses_host.to(dev, non_blocking=True)
        if args.mode == "exact":
            sol, _valid = solver.generate_exact_ik_solutions(y, pos_error_threshold=1e-3, rot_error_threshold=1e-2)
        else:
            sol = solver.generate_ik_solutions(y)  # draws its own latent on the device, like the reference
        if world > 1:
            sol = all_gather_rows(sol, B * world)[rank * B : (rank + 1) * B]
        out_host.copy_(sol, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_steps = min(args.steps, 500)
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = B * world * e2e_steps / float(e2e_s.item())
    clocks = sampler.stop() if rank == 0 else None
    status = solver.nn_model.status()

    if rank == 0:
        peaks = measured_peaks()
        fl = FLOW_FLOPS.get(args.model, flops_per_solution(hp, width))
        kernel_ms = total_ms / args.steps if world == 1 else None
        # the dominant (only) kernel of a step: flow_inverse_kernel; at N > 1 the step also holds the all-gather, so the
        # roofline is quoted from rank-local events of the kernel alone
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                solver.generate_ik_solutions(poses, latent=latent)
            e1.record()
            torch.cuda.synchronize()
            kernel_ms = e0.elapsed_time(e1) / 50
        achieved = fl * B / (kernel_ms * 1e-3) / 1e12
        wbytes = weight_bytes(hp, width)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": total_ms / args.steps, "p50_latency_ms": p50, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (bf16x3 split-operand tensor-core products, fp32 accumulate)" if args.precision == "bf16x3" else "bf16",
            "data": "synthetic", "config": dict(config, l2="256 MB flush between timed steps; weights (203 MB) exceed L2 anyway"),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(poses_host.numel() * 4),
                    "d2h_bytes_per_step": int(out_host.numel() * 4), "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": 205_474_816 + 6_067_456 if args.model == "panda__full__lp191_5.25m" and B == 512 else None,
                "peak_source": peaks["source"] + ", burst bf16",
                "kernel": "ikf::umma::flow_inverse_umma_kernel<%s>" % (("32, true" if width - width // 2 + 8 <= 12 else "32") if B <= 576 else "64" if B <= 1152 else "128"), "kernel_ms": kernel_ms,
                "algorithmic_flops_per_launch": fl * B,
                "hbm": {"algorithmic_bytes_per_launch": wbytes + B * 84, "achieved_gbs": (wbytes + B * 84) / (kernel_ms * 1e-3) / 1e9,
                        "peak_gbs": peaks["hbm_gbs"], "frac": (wbytes + B * 84) / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                "traffic_note": "dram__bytes_read+write of one launch, profiles/r1b_flow_umma_b512_ncu_summary.txt (ncu --set full)",
            },
            "status_word": status,
        }
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_arm(args.model, B, steps=200, warmup=2, budget_s=20.0)
            line["cpu_baseline"] = {
                "value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                "sample": f"{r['steps']} calls of the oracle (torch-CPU restatement of the reference path) at batch {B}, p50 {r['p50_ms']:.1f} ms",
            }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
