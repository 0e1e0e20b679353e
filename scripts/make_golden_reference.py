"""Freezes what the REFERENCE'S OWN host code computes (executed here, unmodified, from /root/reference) into
``tests/golden/reference_*.npz``.  TEST INFRASTRUCTURE; run in the build container (the GPU box has no /root/reference):

    python scripts/make_golden_reference.py

How the reference runs without FrEIA / jrl: ``oracle/ref_stub.py`` (stand-in modules; read its header for what that
does and does not pin).  Two fixtures:

``reference_panda_approx.npz``  ``ikflow.ikflow_solver.IKFlowSolver.generate_ik_solutions`` (:254-343) on the graph wired
    by ``ikflow.model.glow_cNF_model`` (:291-356) with the synthetic seed-0 weights of panda__full__lp191_5.25m:
    512 poses (BASELINE config 2), latent scales 1.0 / 0.75, the single-pose ``n=`` form, and 64 rows of the
    fetch_arm 16-block model (config 4 geometry).
``reference_panda_exact_n2048.npz``  ``IKFlowSolver.generate_exact_ik_solutions`` (:345-411, thresholds of
    scripts/benchmark_generate_exact_solutions.py:18-19, repeat_counts (1,3,10), BASELINE config 3) in two scenarios:
    A  the real (untrained) flow -> every pass runs with almost every pose, nothing converges;
    B  a trained-like stand-in for ``nn_model`` (``q_true(pose) + 0.3 * latent``): pass 1 solves ~45 %, the r=3 and
       r=10 retries run on the rest, ~4 % are never solved -- every branch of :197-233 and :387-408 is taken.
    The latent draws are ``torch.randn(shape, generator=manual_seed(1000 + k))`` for the k-th draw (the reference's
    ``draw_latent`` is re-bound to that, its source is not touched); the fixture stores their sha256.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict  # noqa: E402
from oracle import jrl_kinematics as jk, ref_stub  # noqa: E402
from oracle.scenarios import DRAW_SEED0, PseudoFlow, seeded_draws  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
POS_THR, ROT_THR = 1e-3, 1e-2  # scripts/benchmark_generate_exact_solutions.py:18-19
REPEATS = (1, 3, 10)
SIGMA_B = 0.3


def pass1_fp64(poses: torch.Tensor, seeds: torch.Tensor):
    """The r=1 pass of ``ikflow_solver.py:197-233`` evaluated in fp64 from the same fp32 seeds: the yardstick for
    "well-conditioned" rows.  (``cond(J^T J + 1e-4 I)`` is ~1e5 for EVERY pose of a 7-dof arm -- the null-space
    eigenvalue is lambda itself -- so SURVEY 8d's ``cond < 1e4`` selects nothing; rows on which the reference's own fp32
    arithmetic reproduces the fp64 result to 1e-5 are the ones where a second fp32 implementation can be held to 1e-4.)"""
    q, p = seeds.double().clone(), poses.double()
    sol = torch.zeros_like(q)
    valid = torch.zeros(q.shape[0], dtype=torch.bool)
    for _ in range(3):
        act = ~valid
        q[act] = jk.lm_step(jk.PANDA, p[act], q[act])
        pe, re = jk.pose_error(jk.PANDA, q, p)
        ok = act & (pe < POS_THR) & (re < ROT_THR)
        sol[ok] = q[ok]
        valid |= ok
    return sol, valid


def panda_full():
    hp = IkflowModelParameters()
    hp.nb_nodes, hp.dim_latent_space, hp.coeff_fn_config, hp.coeff_fn_internal_size = 12, 7, 3, 1024
    return hp, make_synthetic_state_dict(hp, jk.PANDA.actuated_joints_limits, seed=0)


def approx_fixture(ref):
    hp, sd = panda_full()
    solver = ref_stub.reference_solver(ref, hp, ref_stub.Panda(), sd)
    _, poses = jk.sample_joint_angles_and_poses(jk.PANDA, 512, seed=10)
    out = {"poses": poses.numpy()}
    for tag, scale in (("s100", 1.0), ("s075", 0.75)):
        latent = scale * torch.randn(512, 7, generator=torch.Generator().manual_seed(11))
        out[f"latent_{tag}"] = latent.numpy()
        out[f"q_{tag}"] = solver.generate_ik_solutions(poses, latent=latent).numpy()
        out[f"q_unclamped_{tag}"] = solver.generate_ik_solutions(poses, latent=latent, clamp_to_joint_limits=False).numpy()
    lat1 = torch.randn(33, 7, generator=torch.Generator().manual_seed(12))
    out["single_pose_latent"] = lat1.numpy()
    out["single_pose_q"] = solver.generate_ik_solutions(poses[5], 33, latent=lat1).numpy()
    # fetch_arm, 16 blocks, width 10 (config 4 geometry)
    hp2 = IkflowModelParameters()
    hp2.nb_nodes, hp2.dim_latent_space, hp2.coeff_fn_config, hp2.coeff_fn_internal_size = 16, 10, 3, 1024
    sd2 = make_synthetic_state_dict(hp2, jk.FETCH_ARM.actuated_joints_limits, seed=0)
    solver2 = ref_stub.reference_solver(ref, hp2, ref_stub.FetchArm(), sd2)
    _, poses2 = jk.sample_joint_angles_and_poses(jk.FETCH_ARM, 64, seed=20)
    lat2 = torch.randn(64, 10, generator=torch.Generator().manual_seed(21))
    out.update(fetch_arm_poses=poses2.numpy(), fetch_arm_latent=lat2.numpy(), fetch_arm_q=solver2.generate_ik_solutions(poses2, latent=lat2).numpy())
    np.savez_compressed(os.path.join(OUT, "reference_panda_approx.npz"), **out)
    print("approx: |q| max", float(np.abs(out["q_s100"]).max()))


def exact_fixture(ref):
    n = 2048
    q_true, poses = jk.sample_joint_angles_and_poses(jk.PANDA, n, seed=2048)
    out = {"poses": poses.numpy(), "q_true": q_true.numpy(), "sigma_b": SIGMA_B, "pos_thr": POS_THR, "rot_thr": ROT_THR, "repeat_counts": np.array(REPEATS), "draw_seed0": DRAW_SEED0}
    hp, sd = panda_full()
    solver = ref_stub.reference_solver(ref, hp, ref_stub.Panda(), sd)

    # scenario A: the real flow
    draw, log = seeded_draws()
    ref.ikflow_solver.draw_latent = draw
    sols, valids = solver.generate_exact_ik_solutions(poses, repeat_counts=REPEATS, pos_error_threshold=POS_THR, rot_error_threshold=ROT_THR, run_lma_on_cpu=False)
    out.update(a_solutions=sols.numpy(), a_valids=valids.numpy(), a_draw_shapes=np.array([s for s, _ in log]), a_draw_sha256=np.array([h for _, h in log]))
    print("A: draws", [s for s, _ in log], "valid", int(valids.sum()))

    # scenario B: trained-like seeds
    flow = solver.nn_model
    solver.nn_model = PseudoFlow(poses, q_true, SIGMA_B)
    draw, log = seeded_draws()
    ref.ikflow_solver.draw_latent = draw
    sols, valids = solver.generate_exact_ik_solutions(poses, repeat_counts=REPEATS, pos_error_threshold=POS_THR, rot_error_threshold=ROT_THR, run_lma_on_cpu=False)
    # the default of the reference (run_lma_on_cpu=True, n >= 750: LM on the CPU, the rest where the poses live) is the
    # same arithmetic on a CPU box -- must be bit-identical
    draw2, _ = seeded_draws()
    ref.ikflow_solver.draw_latent = draw2
    sols2, valids2 = solver.generate_exact_ik_solutions(poses, repeat_counts=REPEATS, pos_error_threshold=POS_THR, rot_error_threshold=ROT_THR)
    assert torch.equal(sols, sols2) and torch.equal(valids, valids2)
    solver.nn_model = flow
    # which pass solved a pose: re-run the passes one by one through the reference's own per-pass function
    draw3, _ = seeded_draws()
    ref.ikflow_solver.draw_latent = draw3
    solver.nn_model = PseudoFlow(poses, q_true, SIGMA_B)
    solved_in = torch.zeros(n, dtype=torch.int64)
    s1, v1 = solver._generate_exact_ik_solutions(poses, REPEATS[0], 3, POS_THR, ROT_THR, lambda *a, **k: None, False)
    solved_in[v1] = 1
    assert torch.equal(s1[v1], sols[v1])
    seed_pass1 = jk.clamp_to_joint_limits(jk.PANDA, q_true + SIGMA_B * torch.randn((n, 7), generator=torch.Generator().manual_seed(DRAW_SEED0)))
    sol64, valid64 = pass1_fp64(poses, seed_pass1)
    well = v1 & valid64 & ((s1.double() - sol64).abs().max(dim=1).values < 1e-5)
    out.update(
        b_solutions=sols.numpy(), b_valids=valids.numpy(), b_solved_in_pass1=v1.numpy(), b_pass1_well_conditioned=well.numpy(), b_pass1_fp64=sol64.numpy(),
        b_draw_shapes=np.array([s for s, _ in log]), b_draw_sha256=np.array([h for _, h in log]),
    )
    print("B: draws", [s for s, _ in log], "valid", int(valids.sum()), "pass 1", int(v1.sum()), "well-conditioned", int(well.sum()))
    np.savez_compressed(os.path.join(OUT, "reference_panda_exact_n2048.npz"), **out)


if __name__ == "__main__":
    assert ref_stub.available(), "needs the reference tree (build container only)"
    torch.set_num_threads(8)
    ref = ref_stub.load()
    approx_fixture(ref)
    exact_fixture(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
