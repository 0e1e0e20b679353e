"""Developer check on a B200 (run under gpurun): compares every kernel with the oracle and prints timings.
Test infrastructure -- imports oracle/."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ikflow_b200
from ikflow_b200 import _lib
from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict
from oracle import freia_flow, jrl_kinematics as jk
from oracle.solver import OracleSolver

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
print(torch.cuda.get_device_name(0), _lib.version())


def kin_checks():
    robot = ikflow_b200.get_robot("panda")
    q, poses = jk.sample_joint_angles_and_poses(jk.PANDA, 1000, seed=1)
    fk = robot.forward_kinematics(q.to(dev)).cpu()
    # quaternion sign is not unique; compare up to sign
    sign = torch.sign((fk[:, 3:] * poses[:, 3:]).sum(1, keepdim=True))
    print("FK max err pos", (fk[:, :3] - poses[:, :3]).abs().max().item(), "quat", (fk[:, 3:] * sign - poses[:, 3:]).abs().max().item(),
          "sign flips", int((sign < 0).sum()))
    fk0 = robot.forward_kinematics(torch.zeros(1, 7, device=dev)).cpu()
    print("FK(0)", fk0)
    noise = 0.05 * torch.randn(1000, 7, generator=torch.Generator().manual_seed(2))
    q0 = jk.clamp_to_joint_limits(jk.PANDA, (q + noise).clone())
    ref = jk.lm_step(jk.PANDA, poses, q0.clone())
    got = robot.inverse_kinematics_step_levenburg_marquardt(poses.to(dev), q0.to(dev)).cpu()
    d = (ref - got).abs().max(dim=1).values
    print("LM step max err", d.max().item(), "median", d.median().item(), "n>1e-4", int((d > 1e-4).sum()))
    pe_ref, re_ref = jk.pose_error(jk.PANDA, ref, poses)
    pe, re = robot.pose_errors(got.to(dev), poses.to(dev))
    print("pose err diff", (pe.cpu() - pe_ref).abs().max().item(), (re.cpu() - re_ref).abs().max().item())
    t = torch.tensor([[1.0, 1, 1, 1, 0, 0, 0]], device=dev)
    pe, re = robot.pose_errors(torch.zeros(1, 7, device=dev), t)
    print("KAT-2 pos/rot", pe.item(), re.item())


def flow_checks(name, hp, robot_name, batch, stress=1.0, blockwise=False):
    robot = ikflow_b200.get_robot(robot_name)
    orobot = jk.ROBOTS[robot_name] if robot_name in jk.ROBOTS else jk.PANDA
    sd = make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=0, stress=stress)
    solver = ikflow_b200.IKFlowSolver(hp, robot)
    solver.load_state_dict_from_dict(sd)
    W = hp.dim_latent_space
    g = torch.Generator().manual_seed(4321)
    latent = torch.randn(batch, W, generator=g)
    _, poses = jk.sample_joint_angles_and_poses(jk.PANDA, batch, seed=1234)
    cond = torch.cat([poses, torch.zeros(batch, 1)], 1)
    t0 = time.time()
    ref, _, inter = freia_flow.flow_inverse(sd, latent, cond, hp.nb_nodes, hp.coeff_fn_config, hp.rnvp_clamp, return_intermediates=True)
    t_cpu = time.time() - t0
    out, _ = solver.nn_model(latent.to(dev), c=cond.to(dev), rev=True)
    torch.cuda.synchronize()
    st = solver.nn_model.status()
    err = (out.cpu() - ref).abs()
    print(f"[{name}] B={batch} max|q|={ref.abs().max():.3f} max err {err.max().item():.3e} mean {err.mean().item():.3e} status {st} (cpu oracle {t_cpu*1e3:.1f} ms)")
    if blockwise or err.max() > 1e-3:
        state = latent.to(dev)
        for i in range(hp.nb_nodes - 1, -1, -1):
            state = solver.nn_model.inverse_blocks(state, cond.to(dev), i, i)
            e = (state.cpu() - inter[hp.nb_nodes - 1 - i]).abs().max().item()
            print(f"   after block {i}: max err {e:.3e}")
    # solver API: single launch incl. slice + clamp
    sol = solver.generate_ik_solutions(poses.to(dev), latent=latent.to(dev))
    ref_sol = jk.clamp_to_joint_limits(orobot, ref[:, : robot.ndof].clone())
    print(f"   solutions err {(sol.cpu() - ref_sol).abs().max().item():.3e}")
    return solver, latent, poses


def timing(solver, batch, iters=50):
    W = solver.network_width
    latent = torch.randn(batch, W, device=dev)
    _, poses = jk.sample_joint_angles_and_poses(jk.PANDA, batch, seed=5)
    poses = poses.to(dev)
    for _ in range(5):
        solver.generate_ik_solutions(poses, latent=latent)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        solver.generate_ik_solutions(poses, latent=latent)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    p50 = ts[len(ts) // 2]
    print(f"   timing B={batch}: p50 {p50*1e3:.1f} us  min {ts[0]*1e3:.1f} us -> {batch/p50*1e3:.0f} solutions/s; status {solver.nn_model.status()}")


def oracle_cuda_timing(hp, batch, iters=20):
    sd = make_synthetic_state_dict(hp, jk.PANDA.actuated_joints_limits, seed=0)
    sdc = freia_flow.state_dict_to(sd, device=dev)
    latent = torch.randn(batch, hp.dim_latent_space, device=dev)
    cond = torch.randn(batch, 8, device=dev)
    with torch.inference_mode():
        for _ in range(3):
            freia_flow.flow_inverse(sdc, latent, cond, hp.nb_nodes, hp.coeff_fn_config, hp.rnvp_clamp)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(iters):
            freia_flow.flow_inverse(sdc, latent, cond, hp.nb_nodes, hp.coeff_fn_config, hp.rnvp_clamp)
        torch.cuda.synchronize()
    dt = (time.time() - t0) / iters
    print(f"   oracle torch-CUDA B={batch}: {dt*1e3:.3f} ms -> {batch/dt:.0f} solutions/s")


def exact_checks(solver, n):
    robot = solver.robot
    q_true, poses = jk.sample_joint_angles_and_poses(jk.PANDA, n, seed=77)
    # seeds near the truth exercise convergence
    for r in (1, 3):
        seeds = (q_true.repeat(r, 1) + 0.05 * torch.randn(n * r, 7, generator=torch.Generator().manual_seed(5))).clamp(-2.8, 2.8)
        seeds = jk.clamp_to_joint_limits(jk.PANDA, seeds)
        fq, fv, nv = robot.lm_refine(poses.to(dev), seeds.to(dev).clone(), r, 3, 1e-3, 1e-2)
        # oracle loop
        qq = seeds.clone()
        final = torch.zeros(n, 7)
        fvalid = torch.zeros(n, dtype=torch.bool)
        for step in range(3):
            act = ~fvalid
            for k in range(r):
                rows = torch.arange(n)[act] + k * n
                qq[rows] = jk.lm_step(jk.PANDA, poses[act], qq[rows])
            newly = torch.zeros(n, dtype=torch.bool)
            for k in range(r):
                rows = torch.arange(n) + k * n
                pe, re = jk.pose_error(jk.PANDA, qq[rows], poses)
                ok = (pe < 1e-3) & (re < 1e-2) & act
                final[ok] = qq[rows][ok]
                newly |= ok
            fvalid |= newly
        agree = (fv.cpu() == fvalid).float().mean().item()
        both = fv.cpu() & fvalid
        print(f"   lm_refine r={r}: valid {int(fv.sum())}/{n} (oracle {int(fvalid.sum())}), mask agreement {agree:.4f}, n_valid_dev {int(nv.item())}, "
              f"max |dq| on common {(fq.cpu()[both] - final[both]).abs().max().item():.3e}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["kin", "tiny", "panda", "time", "exact"]
    if "kin" in which:
        kin_checks()
    if "tiny" in which:
        hp = IkflowModelParameters()
        hp.nb_nodes, hp.coeff_fn_config, hp.coeff_fn_internal_size, hp.dim_latent_space = 3, 2, 256, 9
        for b in (5, 64, 200):
            flow_checks("tiny", hp, "panda", b, blockwise=(b == 5))
    solver = None
    if "panda" in which or "time" in which or "exact" in which:
        hp = IkflowModelParameters()
        hp.dim_latent_space = 7
        solver, _, _ = flow_checks("panda", hp, "panda", 512, blockwise=True)
        flow_checks("panda", hp, "panda", 1000)
        flow_checks("panda", hp, "panda", 2500)
    if "time" in which:
        for b in (64, 512, 2048, 8192):
            timing(solver, b, iters=30)
        oracle_cuda_timing(hp, 512)
    if "exact" in which:
        exact_checks(solver, 512)
        poses = jk.sample_joint_angles_and_poses(jk.PANDA, 256, seed=9)[1].to(dev)
        sol, valid = solver.generate_exact_ik_solutions(poses, pos_error_threshold=1e-3, rot_error_threshold=1e-2)
        print("   generate_exact_ik_solutions:", sol.shape, int(valid.sum()), "valid of", len(valid))
