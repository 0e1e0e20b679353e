"""Developer aid: host-side enqueue cost of the public API calls (what the e2e number adds to the kernel time)."""
import time, torch, sys, os
sys.path.insert(0, os.getcwd())
import ikflow_b200
solver, hp = ikflow_b200.get_ik_solver("panda__full__lp191_5.25m", synthetic_seed=0)
robot = solver.robot
g = torch.Generator().manual_seed(0)
poses = robot.forward_kinematics(robot.sample_joint_angles(512, generator=g, device="cuda"))
latent = torch.randn(512, 7, device="cuda")
for _ in range(20): solver.generate_ik_solutions(poses, latent=latent)
torch.cuda.synchronize()
def enqueue_cost(fn, n=300):
    tot = 0.0
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); tot += time.perf_counter() - t0
    return tot / n * 1e6
print("generate_ik_solutions(latent given) enqueue us:", round(enqueue_cost(lambda: solver.generate_ik_solutions(poses, latent=latent)), 1))
print("generate_ik_solutions(draws latent) enqueue us:", round(enqueue_cost(lambda: solver.generate_ik_solutions(poses)), 1))
print("nn_model.inverse enqueue us:", round(enqueue_cost(lambda: solver.nn_model.inverse(latent, poses, out_cols=7, clamp=True)), 1))
print("torch.randn enqueue us:", round(enqueue_cost(lambda: torch.randn(512, 7, device="cuda")), 1))
ph = poses.cpu().pin_memory(); oh = torch.empty(512, 7).pin_memory()
print("h2d enqueue us:", round(enqueue_cost(lambda: ph.to("cuda", non_blocking=True)), 1))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(300): solver.generate_ik_solutions(poses)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
