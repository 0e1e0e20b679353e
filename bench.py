#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json): IK solutions/s, batched approximate solve.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--model NAME] [--impl ours|reference] [--no-extra]

A step = one pass of the hot path (IKFlowSolver.generate_ik_solutions) over one batch of synthetic target poses.
N = 1 runs BASELINE.json configs[1]: panda__full__lp191_5.25m, batch 512, approximate solve, 1 x B200.  Under torchrun
(N > 1) every rank solves its own 512 poses (weak scaling, weights replicated) and the ranks exchange the joint angles
once per step; the timed region is bracketed by a barrier + synchronize, per-step times are the MAX over ranks.  Rank 0
prints ONE JSON line.

Keys of the line:
  value        solutions/s with poses and latents already in HBM: K steps, each timed with CUDA events on the launching
               stream, an L2 flush (256 MB write) between steps outside the timed window (`timing` says so).
  e2e          the same metric through the public API with HOST (pinned) poses: H2D copy of the poses, latent draw,
               kernel, D2H copy of the joint angles inside the timed region, host wall clock.
  roofline     algorithmic FLOPs of the flow (101,572,608 per solution for the 12-block Panda model)
               divided by the CUDA-event duration of the one kernel a step launches, against the measured bf16 peak;
               `traffic` is read from the committed ncu summary of that kernel under profiles/ (null if there is none).
  cpu_baseline the oracle restatement of the reference's torch path on the host cores, bounded sample (N = 1 only).
  gpu_baseline the SAME oracle (the reference's op sequence: ~480 torch launches per call) on this B200 through
               torch-CUDA -- the comparator of the north-star target "10x the reference PyTorch-CUDA throughput at panda
               B = 512" -- eager (what the reference does) and replayed from a CUDA graph (launch overhead removed).
  extra        the other BASELINE configs, timed in the same run: config 1 (B = 16, CPU), config 3 (exact IK, B = 2048,
               flow / LM split, oracle with run_lma_on_cpu True / False), config 4 (fetch_arm, B = 4096) and config 5
               (16 blocks, B = 8192) -- at N > 1 the last two strong-scaled (B / N rows per GPU).
  --impl reference   times the CPU path alone (the reference itself cannot be installed here: FrEIA 0.2 and jrl are
               un-vendored third-party dependencies, absent from the wheelhouse -- DESIGN.md).

Weights are synthetic (seeded, reference state-dict layout): the released .pkl files live on GCS and there is no
network; timing does not depend on the weight values.
"""

import argparse
import glob
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
HEADLINE_MODEL = "panda__full__lp191_5.25m"
FLOW_FLOPS = {  # algorithmic FLOPs per solution = 2 * MACs of every nn.Linear on the reverse pass
    "panda__full__lp191_5.25m": 101_572_608,
    "fetch_arm__large__mh186_9.25m": 135_725_056,
    "panda__nb16__synthetic": 135_430_144,
}
POS_THR, ROT_THR = 1e-3, 1e-2  # reference scripts/benchmark_generate_exact_solutions.py:18-19


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


def _kernel_template_args(name: str):
    """('ikf::umma::flow_inverse_umma_kernel<32,false,true,ksplit>' | ncu's '...<128, 0, 1, 0, 1>(...)') -> (engine, [rt, jit, f16, ks, pp])."""
    tail = name.replace(" ", "").replace("(int)", "").replace("(bool)", "").split("flow_inverse")[-1].split("(")[0]
    engine, _, rest = tail.partition("<")
    vals = []
    for a in rest.rstrip(">").split(","):
        if a == "ksplit":
            vals = (vals + [0, 0, 0])[:3] + [1, 0]
        elif a == "pingpong":
            vals = (vals + [0, 0, 0])[:3] + [0, 1]
        else:
            vals.append({"true": 1, "false": 0}.get(a, int(a) if a.lstrip("-").isdigit() else -1))
    return engine, (vals + [0] * 5)[:5]  # (older profiles predate the later template arguments)


def traffic_from_profiles(kernel: str, batch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the newest committed ncu summary under
    profiles/ whose kernel (template arguments) and batch match (written by scripts/summarize_profiles.py from an
    `ncu --set full` capture).  Returns (bytes or None, file name or None)."""
    want = _kernel_template_args(kernel)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_b{batch}_ncu_summary.txt")), reverse=True):
        name, rd, wr = None, None, None
        for line in open(path):
            parts = [x.strip() for x in line.split("|")]
            if len(parts) != 3:
                continue
            if parts[0] == "Kernel Name":
                name = parts[2]
            elif parts[0] == "dram__bytes_read.sum":
                rd = float(parts[2]) * scale.get(parts[1], 1.0)
            elif parts[0] == "dram__bytes_write.sum":
                wr = float(parts[2]) * scale.get(parts[1], 1.0)
        if name and rd is not None and wr is not None and _kernel_template_args(name) == want:
            return int(rd + wr), os.path.relpath(path, ROOT)
    return None, None


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


# ----------------------------------------------------------------------------------------------------------------------
# baselines: the oracle restatement of the reference's torch path (the only code under oracle/ this file executes)


def oracle_solver(model_name: str, device: str):
    from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict
    from ikflow_b200.model_loading import MODEL_DESCRIPTIONS
    from oracle import jrl_kinematics as jk
    from oracle.solver import OracleSolver

    hp = IkflowModelParameters()
    hp.__dict__.update(MODEL_DESCRIPTIONS[model_name])
    robot = jk.ROBOTS.get(hp.robot_name, jk.PANDA)
    sd = make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=0)
    return OracleSolver(robot, sd, hp.nb_nodes, hp.dim_latent_space, hp.coeff_fn_config, hp.rnvp_clamp, device=device), hp


def oracle_inputs(batch: int, width: int, device: str):
    from oracle import jrl_kinematics as jk

    _, poses = jk.sample_joint_angles_and_poses(jk.PANDA, batch, seed=1234)
    latent = torch.randn(batch, width, generator=torch.Generator().manual_seed(4321))
    return poses.to(device), latent.to(device)


def cpu_reference_arm(model_name: str, batch: int, steps: int, warmup: int, budget_s: float):
    """The reference's own torch path restated (oracle/), on the host cores."""
    solver, hp = oracle_solver(model_name, "cpu")
    poses, latent = oracle_inputs(batch, hp.dim_latent_space, "cpu")
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


def gpu_baseline_arm(model_name: str, batch: int, dev, calls: int = 100, warmup: int = 20):
    """The oracle on torch-CUDA (fp32, TF32 off as in the reference): eager, and the same op sequence replayed from a
    CUDA graph.  CUDA events around every call, stream synchronised; p50 over `calls` calls after `warmup`."""
    from oracle import freia_flow, jrl_kinematics as jk

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    solver, hp = oracle_solver(model_name, str(dev))
    poses, latent = oracle_inputs(batch, hp.dim_latent_space, str(dev))

    def timed(fn, n):
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return ts

    eager = lambda: solver.generate_ik_solutions(poses, latent=latent)  # noqa: E731
    with torch.inference_mode():
        timed(eager, warmup)
        t_eager = timed(eager, calls)
        out_eager = eager().clone()
        # CUDA graph of the flow + clamp (the part of the call that launches kernels)
        cond = torch.cat([poses, torch.zeros(batch, 1, device=dev)], dim=1)
        graph, t_graph, graph_err = torch.cuda.CUDAGraph(), None, None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    solver._run_inference(latent, cond, True)
            torch.cuda.current_stream().wait_stream(side)
            with torch.cuda.graph(graph):
                static_out = solver._run_inference(latent, cond, True)
            timed(graph.replay, warmup)
            t_graph = timed(graph.replay, calls)
            graph_err = float((static_out - out_eager).abs().max())
        except Exception as e:  # a torch build that cannot capture this op sequence: report it, keep the eager number
            graph_err = f"capture failed: {type(e).__name__}: {e}"[:200]
    res = {
        "what": "oracle/ (plain-torch restatement of the reference's FrEIA + jrl op sequence) on this GPU, fp32, TF32 off",
        "batch": batch, "calls": calls, "warmup": warmup,
        "eager": {"p50_ms": statistics.median(t_eager), "mean_ms": statistics.fmean(t_eager), "value": batch / (statistics.median(t_eager) * 1e-3), "unit": UNIT},
    }
    if t_graph:
        res["cuda_graph"] = {"p50_ms": statistics.median(t_graph), "mean_ms": statistics.fmean(t_graph), "value": batch / (statistics.median(t_graph) * 1e-3), "unit": UNIT,
                             "max_abs_diff_vs_eager": graph_err}
    else:
        res["cuda_graph"] = {"error": graph_err}
    return res, out_eager, poses, latent


def time_calls(fn, n, warm, flush=None):
    """p50 / mean of CUDA-event times of n calls (optionally an L2 flush before each, outside the timed window)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        if flush is not None:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    return statistics.median(ts), statistics.fmean(ts)


def extra_block(args, rank, world, dev, flush, headline_solver):
    """BASELINE configs 1, 3, 4, 5 in the same run (bounded: a few seconds each).  Every rank takes part in the
    multi-GPU legs; the dict is built on rank 0."""
    import torch.distributed as dist

    import ikflow_b200
    from ikflow_b200 import ikflow_solver as solver_mod
    from ikflow_b200.distributed import PeerGather, all_gather_rows, shard_bounds

    extra = {}
    peaks = measured_peaks()

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sharded_throughput(solver, n_total, width, tag):
        """Strong scaling: n_total rows split over the ranks (B / N each), one gather per step."""
        lo, hi = shard_bounds(n_total, rank, world)
        q, poses_all = solver.robot.sample_joint_angles_and_poses(n_total, seed=77, return_torch=True, device=dev)
        latent_all = torch.randn(n_total, width, generator=torch.Generator().manual_seed(5)).to(dev)
        poses, latent = poses_all[lo:hi].contiguous(), latent_all[lo:hi].contiguous()

        spg = None
        if world > 1 and os.environ.get("IKFLOW_B200_GATHER", "fused") == "fused":
            try:
                spg = PeerGather(solver, n_total)
            except Exception:
                spg = None
            ok = torch.tensor([1 if spg is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0 and spg is not None:
                spg.close()
                spg = None

        def step():
            if spg is not None:
                return spg.generate_ik_solutions(poses, latent)
            local = solver.generate_ik_solutions(poses, latent=latent)
            return all_gather_rows(local, n_total) if world > 1 else local

        p50, mean = time_calls(step, 20, 5, flush)
        p50, mean = reduce_max(p50), reduce_max(mean)
        if spg is not None:
            spg.close()
        hp = solver.nn_model.params
        fl = flops_per_solution(hp, width)
        return {
            "workload": tag, "rows_total": n_total, "rows_per_gpu": hi - lo, "n_gpus": world, "scaling": "strong" if world > 1 else "single GPU",
            "p50_ms": p50, "mean_ms": mean, "value": n_total / (mean * 1e-3), "unit": UNIT,
            "kernel": solver.nn_model.last_kernel(), "tensor_frac_of_measured_bf16": fl * (hi - lo) / (mean * 1e-3) / 1e12 / peaks["bf16_tflops"],
        }

    # ---- config 4: fetch_arm, 16 blocks, B = 4096 --------------------------------------------------------------------
    fa, _ = ikflow_b200.get_ik_solver("fetch_arm__large__mh186_9.25m", synthetic_seed=0)
    extra["config4_fetch_arm_b4096"] = sharded_throughput(fa, 4096, fa.network_width, "fetch_arm__large__mh186_9.25m, batch=4096 total, approximate solve")
    del fa
    # ---- config 5: panda geometry, 16 blocks, B = 8192 ---------------------------------------------------------------
    p16, _ = ikflow_b200.get_ik_solver("panda__nb16__synthetic", synthetic_seed=0)
    extra["config5_panda_nb16_b8192"] = sharded_throughput(p16, 8192, p16.network_width, "panda__nb16__synthetic (nb_nodes=16), batch=8192 total, approximate solve")
    del p16
    torch.cuda.empty_cache()
    if world > 1:
        # the latency floor that bounds strong scaling of the headline batch: 512 / N rows per GPU
        lo, hi = shard_bounds(512, rank, world)
        q, poses_all = headline_solver.robot.sample_joint_angles_and_poses(512, seed=78, return_torch=True, device=dev)
        latent_all = torch.randn(512, 7, generator=torch.Generator().manual_seed(6)).to(dev)
        poses, latent = poses_all[lo:hi].contiguous(), latent_all[lo:hi].contiguous()
        spg = None
        try:
            spg = PeerGather(headline_solver, 512) if os.environ.get("IKFLOW_B200_GATHER", "fused") == "fused" else None
        except Exception:
            spg = None
        ok = torch.tensor([1 if spg is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0 and spg is not None:
            spg.close()
            spg = None
        strong_step = (lambda: spg.generate_ik_solutions(poses, latent)) if spg is not None else (lambda: all_gather_rows(headline_solver.generate_ik_solutions(poses, hi - lo, latent=latent), 512))
        p50, mean = time_calls(strong_step, 20, 5, flush)
        p50, mean = reduce_max(p50), reduce_max(mean)
        if spg is not None:
            spg.close()
        extra["config2_strong_b512"] = {"workload": f"{HEADLINE_MODEL}, batch=512 total", "rows_per_gpu": hi - lo, "n_gpus": world, "scaling": "strong",
                                        "p50_ms": p50, "mean_ms": mean, "value": 512 / (mean * 1e-3), "unit": UNIT,
                                        "note": "a launch streams all weights whatever the row count (B=64: ~0.5 ms): the headline batch does not strong-scale"}
        return extra if rank == 0 else None

    # ---- single-GPU only from here -------------------------------------------------------------------------------------
    solver = headline_solver
    # large-batch throughput of the headline model, default precision and bf16x3
    lb = {}
    for prec in (None, "bf16x3"):
        s2 = solver
        if prec is not None:
            os.environ["IKFLOW_B200_PRECISION"] = prec
            s2, _ = ikflow_b200.get_ik_solver(HEADLINE_MODEL, synthetic_seed=0)
            os.environ.pop("IKFLOW_B200_PRECISION", None)
        for b in (2048, 8192, 9216):
            q, poses = s2.robot.sample_joint_angles_and_poses(b, seed=80, return_torch=True, device=dev)
            latent = torch.randn(b, 7, generator=torch.Generator().manual_seed(8)).to(dev)
            p50, mean = time_calls(lambda: s2.generate_ik_solutions(poses, latent=latent), 20, 5, flush)
            lb[f"b{b}_{s2.nn_model.effective_precision()}"] = {"p50_ms": p50, "value": b / (mean * 1e-3), "unit": UNIT, "kernel": s2.nn_model.last_kernel()}
        if prec is not None:
            del s2
    extra["large_batch_" + HEADLINE_MODEL.split("__")[0]] = lb
    # latency floor
    for b in (16, 64):
        q, poses = solver.robot.sample_joint_angles_and_poses(b, seed=79, return_torch=True, device=dev)
        latent = torch.randn(b, 7, generator=torch.Generator().manual_seed(7)).to(dev)
        p50, mean = time_calls(lambda: solver.generate_ik_solutions(poses, latent=latent), 50, 5, flush)
        extra[f"latency_b{b}"] = {"p50_ms": p50, "mean_ms": mean, "value": b / (mean * 1e-3), "unit": UNIT}
    # ---- config 1: B = 16 on torch-CPU (the reference's own CPU-runnable case) ---------------------------------------
    r = cpu_reference_arm(HEADLINE_MODEL, 16, steps=30, warmup=2, budget_s=8.0)
    extra["config1_cpu_b16"] = {"workload": f"{HEADLINE_MODEL}, batch=16, oracle on torch-CPU", "p50_ms": r["p50_ms"], "value": r["value"], "unit": UNIT,
                                "cores": r["cores"], "calls": r["steps"], "ours_same_batch": extra["latency_b16"]}
    # ---- config 3: generate_exact_ik_solutions, B = 2048, (1, 3, 10), 1 mm / 0.01 rad --------------------------------
    n = 2048
    q_true, poses = solver.robot.sample_joint_angles_and_poses(n, seed=2048, return_torch=True, device=dev)
    kw = dict(repeat_counts=(1, 3, 10), pos_error_threshold=POS_THR, rot_error_threshold=ROT_THR)
    cfg3 = {"workload": f"{HEADLINE_MODEL}, batch=2048, generate_exact_ik_solutions (flow + LM refine), repeat_counts (1,3,10), 1 mm / 0.01 rad"}

    class Phases:  # CUDA-event split of a call into its flow launches and its LM launches (bench-side wrappers only)
        def __init__(self):
            self.flow, self.lm, self.rows = [], [], 0

        def wrap(self, fn, bucket, count_rows=False):
            def inner(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = fn(*a, **k)
                e1.record()
                bucket.append((e0, e1))
                if count_rows:
                    self.rows += a[0].shape[0]
                return out
            return inner

        def ms(self, bucket):
            return sum(a.elapsed_time(b) for a, b in bucket)

    def run_ours(seed_fn, tag, calls=10):
        real_inverse, real_refine = solver.nn_model.inverse, solver.robot.lm_refine
        ph = Phases()
        solver.nn_model.inverse = ph.wrap(seed_fn or real_inverse, ph.flow, count_rows=True)
        solver.robot.lm_refine = ph.wrap(real_refine, ph.lm)
        try:
            torch.manual_seed(0)
            for _ in range(2):
                solver.generate_exact_ik_solutions(poses, **kw)
            torch.cuda.synchronize()
            ph.flow.clear(), ph.lm.clear()
            ph.rows = 0
            ts, valid = [], None
            for _ in range(calls):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                sol, valid = solver.generate_exact_ik_solutions(poses, **kw)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
        finally:
            solver.nn_model.inverse, solver.robot.lm_refine = real_inverse, real_refine
        cfg3[tag] = {"p50_ms": statistics.median(ts), "mean_ms": statistics.fmean(ts), "value": n / (statistics.fmean(ts) * 1e-3), "unit": "poses/s",
                     "flow_ms_per_call": ph.ms(ph.flow) / calls, "lm_ms_per_call": ph.ms(ph.lm) / calls, "flow_rows_per_call": ph.rows / calls,
                     "valid_fraction_last_call": float(valid.float().mean())}

    run_ours(None, "ours_untrained_flow")

    def trained_like(latent, cond, out_cols=None, clamp=False):  # what a trained flow delivers: seeds near a solution
        r = latent.shape[0] // cond.shape[0]
        idx = (cond[:, None, :3] == poses[None, :, :3]).all(-1).float().argmax(1) if cond.shape[0] != n else torch.arange(n, device=dev)
        return solver.robot.clamp_to_joint_limits((q_true[idx].repeat(r, 1) + 0.3 * latent[:, :7]).contiguous())

    run_ours(trained_like, "ours_trained_like_seeds")
    cfg3["ours_trained_like_seeds"]["note"] = "flow replaced by q_true + 0.3 z (torch ops, counted in flow_ms): LM / select / retry logic with ~50 % converging per pass"
    # the oracle on this GPU, reference defaults: run_lma_on_cpu=True (LM on the host, ikflow_solver.py:352) and False
    osolver, _ = oracle_solver(HEADLINE_MODEL, str(dev))
    for flag in (True, False):
        torch.manual_seed(0)
        ts = []
        for i in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            osolver.generate_exact_ik_solutions(poses, (1, 3, 10), POS_THR, ROT_THR, run_lma_on_cpu=flag)
            torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
        cfg3[f"oracle_cuda_run_lma_on_cpu_{flag}"] = {"ms_per_call": min(ts), "value": n / (min(ts) * 1e-3), "unit": "poses/s", "calls": 2,
                                                      "note": "untrained flow: all three passes run (28,672 flow rows)"}
    extra["config3_exact_b2048"] = cfg3
    return extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--batch", type=int, default=512, help="target poses per GPU per step")
    ap.add_argument("--model", default=HEADLINE_MODEL)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra block (BASELINE configs 1, 3, 4, 5)")
    ap.add_argument("--precision", default=None, choices=["auto", "bf16x3", "fp16x3", "bf16x1"])
    ap.add_argument("--mode", default="approx", choices=["approx", "exact"],
                    help="approx = generate_ik_solutions (headline); exact = generate_exact_ik_solutions (flow + LM refinement)")
    args = ap.parse_args()
    warmup = max(args.warmup, 3)
    if args.precision:
        os.environ["IKFLOW_B200_PRECISION"] = args.precision

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {  # identical in both arms (the driver compares it)
        "workload": f"{args.model}, batch={args.batch} per GPU, "
        + ("approximate solve (generate_ik_solutions)" if args.mode == "approx" else "generate_exact_ik_solutions (flow + LM refine, repeat_counts (1,3,10), 1 mm / 0.01 rad)")
        + ", synthetic seeded weights",
        "batch_per_gpu": args.batch,
        "global_batch": args.batch * world,
        "parallelism": f"batch-sharded x{world}, weights replicated, one exchange of the joint angles per step" if world > 1 else "single GPU",
    }

    # ------------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_arm(args.model, args.batch, args.steps, warmup, budget_s=240.0)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": warmup, "ms_per_step": r["ms_per_step"], "p50_latency_ms": r["p50_ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": f"{r['steps']} calls of the oracle restatement of the reference torch path at batch {args.batch} on CPU"},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = oracle/ port of FrEIA-0.2 + jrl ops on torch-CPU (bit-equal to the reference's own host code run on stand-in modules, tests/test_reference_host_logic.py); the upstream package cannot be installed offline",
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
    from ikflow_b200.distributed import PeerGather, all_gather_rows

    solver, hp = ikflow_b200.get_ik_solver(args.model, synthetic_seed=0)
    precision = solver.nn_model.effective_precision(dev)
    robot = solver.robot
    width = solver.network_width
    B = args.batch
    q_true, poses = robot.sample_joint_angles_and_poses(B, seed=1234 + rank, return_torch=True, device=dev)  # one launch, on the device
    latent = torch.randn(B, width, generator=torch.Generator().manual_seed(4321 + rank)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB of L2

    # N > 1: the gather is fused into the kernel's epilogue (peer stores over NVLink + flags, ikflow_b200.distributed.PeerGather);
    # IKFLOW_B200_GATHER=nccl (or a box without peer access) falls back to one NCCL all-gather per step
    pg, gather_kind = None, "none (single GPU)"
    if world > 1 and args.mode == "approx":
        gather_kind = "nccl all_gather_into_tensor"
        if os.environ.get("IKFLOW_B200_GATHER", "fused") == "fused":
            try:
                pg = PeerGather(solver, B * world)
                gather_kind = "fused into the flow kernel: peer stores (NVLink P2P) + per-rank flags, no collective"
            except Exception as e:  # no symmetric memory / no peer access: keep the collective
                gather_kind += f" (fused gather unavailable: {type(e).__name__}: {str(e)[:120]})"
        ok = torch.tensor([1 if pg is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks or none
        if int(ok.item()) == 0 and pg is not None:
            pg.close()
            pg = None

    def step_resident():
        if args.mode == "exact":  # BASELINE.json configs[2]: thresholds of scripts/benchmark_generate_exact_solutions.py:18-19
            local, _valid = solver.generate_exact_ik_solutions(poses, pos_error_threshold=POS_THR, rot_error_threshold=ROT_THR)
        elif pg is not None:
            return pg.generate_ik_solutions(poses, latent)
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
            sol, _valid = solver.generate_exact_ik_solutions(y, pos_error_threshold=POS_THR, rot_error_threshold=ROT_THR)
        else:
            sol = pg.generate_ik_solutions(y) if pg is not None else solver.generate_ik_solutions(y)  # draws its own latent on the device, like the reference
        if world > 1:
            sol = (sol if pg is not None else all_gather_rows(sol, B * world))[rank * B : (rank + 1) * B]
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

    # the dominant (only) kernel of a step is the flow kernel; at N > 1 the step also holds the exchange, so the roofline
    # is quoted from rank-local events of the kernel alone
    kernel_ms = total_ms / args.steps
    if world > 1 or args.mode != "approx":
        kernel_ms, _ = time_calls(lambda: solver.generate_ik_solutions(poses, latent=latent), 50, 3, flush)
    kernel_name = solver.nn_model.last_kernel()

    if pg is not None:
        pg.close()
    extra = None
    if not args.no_extra and args.model == HEADLINE_MODEL and args.mode == "approx":
        extra = extra_block(args, rank, world, dev, flush, solver)

    if rank == 0:
        peaks = measured_peaks()
        fl = FLOW_FLOPS.get(args.model, flops_per_solution(hp, width))
        achieved = fl * B / (kernel_ms * 1e-3) / 1e12
        wbytes = weight_bytes(hp, width)
        traffic, traffic_src = traffic_from_profiles(kernel_name, B)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": total_ms / args.steps, "p50_latency_ms": p50, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": {"bf16x3": "f32 (bf16x3 split-operand tensor-core products, fp32 accumulate)",
                      "fp16x3": "f32 (fp16x3 split-operand tensor-core products with scaled tails, fp32 accumulate)", "bf16x1": "bf16"}[precision],
            "data": "synthetic", "config": config,
            "timing": "CUDA events per step on the launching stream, max over ranks; 256 MB L2 flush before every timed step, outside the timed window (the 203 MB of weights exceed L2 anyway)",
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(poses_host.numel() * 4),
                    "d2h_bytes_per_step": int(out_host.numel() * 4), "steps": e2e_steps},
            "gpu_launches": int(launches), "gather": gather_kind,
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": traffic,
                "peak_source": peaks["source"] + ", burst bf16",
                "kernel": kernel_name, "kernel_ms": kernel_ms,
                "algorithmic_flops_per_launch": fl * B,
                "issued_flops_per_launch": 3 * fl * B if precision != "bf16x1" else fl * B,
                "hbm": {"algorithmic_bytes_per_launch": wbytes + B * 84, "achieved_gbs": (wbytes + B * 84) / (kernel_ms * 1e-3) / 1e9,
                        "peak_gbs": peaks["hbm_gbs"], "frac": (wbytes + B * 84) / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                "traffic_source": traffic_src if traffic is not None else "none: no committed ncu capture of this kernel variant at this batch size",
            },
            "status_word": status,
        }
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_arm(args.model, B, steps=200, warmup=2, budget_s=20.0)
            line["cpu_baseline"] = {
                "value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                "sample": f"{r['steps']} calls of the oracle (torch-CPU restatement of the reference path) at batch {B}, p50 {r['p50_ms']:.1f} ms",
            }
        if world == 1 and not args.no_gpu_baseline and args.mode == "approx":
            gb, ref_out, gposes, glatent = gpu_baseline_arm(args.model, B, dev)
            ours = solver.generate_ik_solutions(gposes, latent=glatent)
            p50_ours, _ = time_calls(lambda: solver.generate_ik_solutions(gposes, latent=glatent), 100, 20, None)
            gb["ours_same_inputs"] = {"p50_ms": p50_ours, "max_abs_diff_vs_oracle_cuda": float((ours - ref_out).abs().max()),
                                      "speedup_vs_eager_p50": gb["eager"]["p50_ms"] / p50_ours,
                                      "speedup_vs_cuda_graph_p50": (gb["cuda_graph"]["p50_ms"] / p50_ours) if "p50_ms" in gb["cuda_graph"] else None}
            line["gpu_baseline"] = gb
        if extra is not None:
            line["extra"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
