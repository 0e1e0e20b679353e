"""The oracle against every known-answer vector the reference's own tests hold for the hot path, against the committed
golden fixtures, and against the algebraic properties of the flow (CPU only)."""
import math
import os

import numpy as np
import pytest
import torch

from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict, permute_random_tables
from oracle import freia_flow, jrl_kinematics as jk
from oracle.solver import OracleSolver

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# reference tests/model_test.py:18-25
PANDA_LIMITS = [(-2.8973, 2.8973), (-1.7628, 1.7628), (-2.8973, 2.8973), (-3.0718, -0.0698), (-2.8973, 2.8973), (-0.0175, 3.7525), (-2.8973, 2.8973)]


def test_panda_joint_limits_match_reference_table():
    # reference tests/model_test.py:33-34,41-42 (5 decimal places)
    for (lo, hi), (rlo, rhi) in zip(jk.PANDA.actuated_joints_limits, PANDA_LIMITS):
        assert round(lo - rlo, 5) == 0 and round(hi - rhi, 5) == 0


def test_fk_golden_vector_kat1():
    # reference tests/evaluation_utils_test.py:20-24
    pose = jk.forward_kinematics(jk.PANDA, torch.zeros(1, 7))[0]
    expected = torch.tensor([0.088, 0.0, 0.926, 0.0, 0.92387953, 0.38268343, 0.0])
    torch.testing.assert_close(pose, expected, atol=1e-5, rtol=0)


def test_pose_error_kat2():
    # reference tests/evaluation_utils_test.py:26-32: target [1,1,1, 1,0,0,0] vs FK(0)
    target = torch.tensor([[1.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0]])
    pos, rot = jk.pose_error(jk.PANDA, torch.zeros(1, 7), target)
    assert abs(pos.item() - 1.355440887681938) < 1e-6
    assert abs(rot.item() - 3.1415927) < 5e-4


def test_joint_limits_exceeded_truth_table_kat3():
    # reference tests/evaluation_utils_test.py:37-55
    limits = [(0, 1), (0, 1), (0, 1)]
    configs = torch.tensor([[0.5, 0.5, 0.5], [0.0, 0.5, 1.0], [-0.1, 0.5, 0.5], [0.5, 1.1, 0.5], [0.5, 0.5, 1.0001]])
    got = jk.calculate_joint_limits_exceeded(configs, limits)
    assert got.tolist() == [False, False, True, True, True]


def test_permute_random_tables_match_survey_appendix_f():
    # np.random.seed(i); np.random.permutation(W) -- legacy MT19937 stream (ikflow/model.py:339 -> FrEIA PermuteRandom)
    w7 = [[6, 2, 1, 3, 0, 5, 4], [6, 2, 1, 0, 4, 3, 5], [4, 1, 3, 2, 6, 5, 0], [4, 6, 5, 3, 1, 0, 2]]
    w10 = [[2, 8, 4, 9, 1, 6, 7, 3, 0, 5], [2, 9, 6, 4, 0, 3, 1, 7, 8, 5]]
    for seed, exp in enumerate(w7):
        perm, inv = freia_flow.permute_random_tables(7, seed)
        assert perm.tolist() == exp
        assert perm[inv].tolist() == list(range(7))
        np.random.seed(seed)
        assert np.random.permutation(7).tolist() == exp
        assert permute_random_tables(7, seed)[0].tolist() == exp
    for seed, exp in enumerate(w10):
        assert freia_flow.permute_random_tables(10, seed)[0].tolist() == exp


def _model(nb, w, cfg, hidden, robot=jk.PANDA, seed=0):
    hp = IkflowModelParameters()
    hp.nb_nodes, hp.dim_latent_space, hp.coeff_fn_config, hp.coeff_fn_internal_size = nb, w, cfg, hidden
    return hp, make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=seed)


def test_flow_is_invertible_forward_then_reverse():
    hp, sd = _model(3, 9, 2, 256)
    sd64 = freia_flow.state_dict_to(sd, torch.float64)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(32, 9, generator=g, dtype=torch.float64)
    cond = torch.randn(32, 8, generator=g, dtype=torch.float64)
    x, ld_rev = freia_flow.flow_inverse(sd64, z, cond, 3, 2, 2.5)
    z2, ld_fwd = freia_flow.flow_forward(sd64, x, cond, 3, 2, 2.5)
    assert (z - z2).abs().max() < 1e-6  # M and M_inv are stored in fp32: M @ M_inv = I only to 1e-7
    assert (ld_rev + ld_fwd).abs().max() < 1e-6  # log-determinants are opposite


def test_fixed_linear_transform_scales_by_joint_limit_magnitude():
    # ikflow/model.py:311-316: M = diag(1/max|limit|) on the joint columns, 1 on padding, b = 0
    hp, sd = _model(1, 9, 1, 64)
    m_inv = sd["module_list.0.M_inv"]
    expected = [2.8973, 1.7628, 2.8973, 3.0718, 2.8973, 3.7525, 2.8973, 1.0, 1.0]
    torch.testing.assert_close(torch.diagonal(m_inv), torch.tensor(expected), atol=1e-6, rtol=1e-6)
    assert sd["module_list.0.b"].abs().max() == 0


def test_relational_kat4_same_rows_same_outputs_different_poses_differ():
    # reference tests/ikflow_solver_test.py:94-117 with TINY_MODEL_PARAMS
    hp, sd = _model(3, 9, 2, 256)
    solver = OracleSolver(jk.PANDA, sd, 3, 9, 2)
    pose = torch.tensor([0.5, 0.1, 0.4, 1.0, 0.0, 0.0, 0.0])
    latent = torch.randn(1, 9, generator=torch.Generator().manual_seed(0)).repeat(5, 1)
    out = solver.generate_ik_solutions(pose.repeat(5, 1), latent=latent, clamp_to_joint_limits=False)
    assert (out - out[0:1]).abs().max() < 1e-8
    poses = pose.repeat(5, 1)
    poses[:, 0] += torch.arange(5) * 0.05
    out2 = solver.generate_ik_solutions(poses, latent=latent, clamp_to_joint_limits=False)
    for i in range(5):
        for j in range(i + 1, 5):
            assert (out2[i] - out2[j]).abs().max() > 1e-8


@pytest.mark.parametrize("name", ["tiny_w9", "panda_nb12", "fetch_arm_nb16"])
def test_oracle_reproduces_golden_flow_fixture(name):
    d = np.load(os.path.join(GOLD, f"flow_{name}.npz"))
    robot = jk.ROBOTS[str(d["robot"])]
    hp, sd = _model(int(d["nb_nodes"]), int(d["width"]), int(d["coeff_fn_config"]), int(d["hidden"]), robot)
    latent, poses = torch.from_numpy(d["latent"]), torch.from_numpy(d["poses"])
    n = 16  # a slice keeps the CPU suite fast; rows are independent
    cond = torch.cat([poses, torch.zeros(len(poses), 1)], dim=1)
    out, _ = freia_flow.flow_inverse(sd, latent[:n], cond[:n], hp.nb_nodes, hp.coeff_fn_config, float(d["rnvp_clamp"]))
    assert (out - torch.from_numpy(d["out_fp32"])[:n]).abs().max() < 2e-5  # BLAS blocking differs with the batch size
    assert (out.double() - torch.from_numpy(d["out_fp64"])[:n]).abs().max() < 5e-5
    # forward direction (x -> z, log-det) of the same graph
    x = torch.from_numpy(d["fwd_x"])
    z, ld = freia_flow.flow_forward(sd, x[:n], cond[:n], hp.nb_nodes, hp.coeff_fn_config, float(d["rnvp_clamp"]))
    assert (z - torch.from_numpy(d["fwd_z_fp32"])[:n]).abs().max() < 2e-5
    assert (z.double() - torch.from_numpy(d["fwd_z_fp64"])[:n]).abs().max() < 5e-5
    assert (ld.double() - torch.from_numpy(d["fwd_logdet_fp64"])[:n]).abs().max() < 1e-4
    # ... and the reverse pass undoes it (the flow is a bijection)
    back, ld_rev = freia_flow.flow_inverse(sd, z, cond[:n], hp.nb_nodes, hp.coeff_fn_config, float(d["rnvp_clamp"]))
    assert (back - x[:n]).abs().max() < 1e-4
    assert (ld_rev + ld).abs().max() < 1e-3


def test_oracle_reproduces_golden_kinematics_fixture():
    d = np.load(os.path.join(GOLD, "kinematics_panda.npz"))
    q = torch.from_numpy(d["q"])
    fk = jk.forward_kinematics(jk.PANDA, q)
    assert (fk - torch.from_numpy(d["fk_fp32"])).abs().max() < 1e-6
    assert (fk.double() - torch.from_numpy(d["poses_fp64"])).abs()[:, :3].max() < 1e-6
    step = jk.lm_step(jk.PANDA, torch.from_numpy(d["poses_fp64"]).float(), torch.from_numpy(d["lm_seeds"]).clone())
    err = (step.double() - torch.from_numpy(d["lm_step_fp64"])).abs().max(dim=1).values
    assert err.median() < 1e-4  # fp32 solve of an ill-conditioned 7x7 system; the fp64 column is the truth
    assert torch.from_numpy(d["pos_err_after_fp64"]).median() < 5e-3  # one LM step from 0.05 rad noise nearly closes


def test_lm_step_converges_and_respects_limits_kat5():
    # the closure properties the reference asserts after refinement (tests/ikflow_solver_test.py:82-87), exercised
    # with seeds = truth + noise (no trained weights offline)
    q_true, poses = jk.sample_joint_angles_and_poses(jk.PANDA, 200, seed=3)
    q = jk.clamp_to_joint_limits(jk.PANDA, q_true + 0.02 * torch.randn(200, 7, generator=torch.Generator().manual_seed(1)))
    for _ in range(3):
        q = jk.lm_step(jk.PANDA, poses, q)
    pos, rot = jk.pose_error(jk.PANDA, q, poses)
    assert (pos < 1e-3).float().mean() > 0.9 and (rot < 1e-2).float().mean() > 0.9
    assert torch.equal(q, jk.clamp_to_joint_limits(jk.PANDA, q.clone()))


def test_oracle_exact_solver_last_valid_repeat_wins():
    # ikflow_solver.py:217-222: rows are repeat-major, later rows overwrite earlier ones
    hp, sd = _model(2, 7, 1, 64)
    q_true, poses = jk.sample_joint_angles_and_poses(jk.PANDA, 6, seed=5)
    solver = OracleSolver(jk.PANDA, sd, 2, 7, 1)
    # replace the flow seeds by truth + noise with a different noise per repeat
    noise = 0.01 * torch.randn(3 * 6, 7, generator=torch.Generator().manual_seed(2))
    solver._run_inference = lambda latent, cond, clamp: jk.clamp_to_joint_limits(jk.PANDA, q_true.repeat(3, 1) + noise)
    sol, valid = solver._generate_exact_ik_solutions(poses, 3, 3, 1e-3, 1e-2)
    assert valid.all()
    # recompute: at the first step where any repeat is valid, the largest valid repeat index supplies the solution
    q = jk.clamp_to_joint_limits(jk.PANDA, q_true.repeat(3, 1) + noise)
    q1 = jk.lm_step(jk.PANDA, poses.repeat(3, 1), q)
    pos, rot = jk.pose_error(jk.PANDA, q1, poses.repeat(3, 1))
    ok = ((pos < 1e-3) & (rot < 1e-2)).view(3, 6)
    for p in range(6):
        ks = [k for k in range(3) if ok[k, p]]
        if ks:
            assert torch.equal(sol[p], q1[ks[-1] * 6 + p])


def test_oracle_logdet_is_the_log_jacobian_determinant_of_its_own_forward_map():
    """Internal consistency of the FrEIA restatement (it cannot be pinned against FrEIA itself, which is not installed):
    the log-det returned by the forward pass equals log|det dz/dx| of that very map (autograd), and the reverse pass
    returns its negative -- i.e. the coupling / clamp / permutation / fixed-linear formulas form a proper flow."""
    hp, sd = _model(3, 7, 2, 64, jk.PANDA)
    sd64 = freia_flow.state_dict_to(sd, torch.float64)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(4, 7, generator=g, dtype=torch.float64) * 2 - 1) * 2.0
    _, poses = jk.sample_joint_angles_and_poses(jk.PANDA, 4, seed=3, dtype=torch.float64)
    cond = torch.cat([poses, torch.zeros(4, 1, dtype=torch.float64)], dim=1)
    z, logdet = freia_flow.flow_forward(sd64, x, cond, hp.nb_nodes, hp.coeff_fn_config, hp.rnvp_clamp)
    for i in range(4):
        f = lambda xi: freia_flow.flow_forward(sd64, xi[None], cond[i : i + 1], hp.nb_nodes, hp.coeff_fn_config, hp.rnvp_clamp)[0][0]
        jac = torch.autograd.functional.jacobian(f, x[i])
        assert abs(float(torch.linalg.slogdet(jac)[1]) - float(logdet[i])) < 1e-5  # logDetM of the state dict is an fp32 number
    x_back, logdet_rev = freia_flow.flow_inverse(sd64, z, cond, hp.nb_nodes, hp.coeff_fn_config, hp.rnvp_clamp)
    assert (x_back - x).abs().max() < 1e-6  # M and M_inv of the state dict are fp32 inverses of each other
    assert (logdet_rev + logdet).abs().max() < 1e-6


# --- fixtures frozen from the reference's own code (scripts/make_golden_reference.py; oracle/ref_stub.py) -------------
def test_oracle_reproduces_reference_approx_fixture():
    """``reference_panda_approx.npz`` is the output of ``ikflow.ikflow_solver.IKFlowSolver.generate_ik_solutions`` run
    from /root/reference; the oracle must give the same numbers wherever this suite runs."""
    d = np.load(os.path.join(GOLD, "reference_panda_approx.npz"))
    hp, sd = _model(12, 7, 3, 1024)
    solver = OracleSolver(jk.PANDA, sd, 12, 7, 3)
    poses = torch.from_numpy(d["poses"])
    n = 64
    for tag in ("s100", "s075"):
        latent = torch.from_numpy(d[f"latent_{tag}"])
        got = solver.generate_ik_solutions(poses[:n], latent=latent[:n])
        assert (got - torch.from_numpy(d[f"q_{tag}"])[:n]).abs().max() < 2e-5  # BLAS blocking differs with the batch size
        raw = solver.generate_ik_solutions(poses[:n], latent=latent[:n], clamp_to_joint_limits=False)
        assert (raw - torch.from_numpy(d[f"q_unclamped_{tag}"])[:n]).abs().max() < 2e-5
    lat1 = torch.from_numpy(d["single_pose_latent"])
    assert (solver.generate_ik_solutions(poses[5], 33, latent=lat1) - torch.from_numpy(d["single_pose_q"])).abs().max() < 2e-5


def test_oracle_reproduces_reference_exact_fixture_scenario_b():
    """BASELINE config 3 sizes (n = 2048, repeat_counts (1, 3, 10)) with trained-like seeds: the oracle's restatement of
    ikflow_solver.py:119-247 / :345-411 must reproduce the reference's own output bit for bit."""
    from oracle.scenarios import PseudoFlow, seeded_draws

    d = np.load(os.path.join(GOLD, "reference_panda_exact_n2048.npz"))
    poses, q_true = torch.from_numpy(d["poses"]), torch.from_numpy(d["q_true"])
    hp, sd = _model(1, 7, 1, 32)
    draw, log = seeded_draws(int(d["draw_seed0"]))
    solver = OracleSolver(jk.PANDA, sd, 1, 7, 1, latent_source=lambda shape, dev: draw("gaussian", 1.0, shape, dev))
    flow = PseudoFlow(poses, q_true, float(d["sigma_b"]))
    solver._run_inference = lambda latent, cond, clamp: jk.clamp_to_joint_limits(jk.PANDA, flow(latent, c=cond)[0])
    sols, valids = solver.generate_exact_ik_solutions(
        poses, tuple(int(r) for r in d["repeat_counts"]), float(d["pos_thr"]), float(d["rot_thr"]), run_lma_on_cpu=False
    )
    assert [list(s) for s, _ in log] == d["b_draw_shapes"].tolist()
    assert [h for _, h in log] == d["b_draw_sha256"].tolist()  # torch's CPU generator still produces the frozen stream
    assert torch.equal(valids, torch.from_numpy(d["b_valids"]))
    assert torch.equal(sols, torch.from_numpy(d["b_solutions"]))
    assert int(valids.sum()) == 1966 and len(log) == 3 and log[2][0][0] % 10 == 0  # the r = 10 pass ran


def test_philox4x32_10_known_answer_vectors():
    """Random123 kat_vectors (philox4x32, 10 rounds): zero, all-ones and the pi-digit counter/key."""
    from oracle.philox import philox4x32_10, sample_uniforms

    u32 = lambda *a: np.array(a, dtype=np.uint32)
    assert philox4x32_10(u32(0, 0, 0, 0), u32(0, 0)).tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    ones = 0xFFFFFFFF
    assert philox4x32_10(u32(ones, ones, ones, ones), u32(ones, ones)).tolist() == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox4x32_10(u32(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), u32(0xA4093822, 0x299F31D0)).tolist() == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u = sample_uniforms(seed=7, first_index=0, n=20000, ndof=7)
    assert u.shape == (20000, 7) and 0.0 < u.min() and u.max() < 1.0
    assert np.abs(u.mean(0) - 0.5).max() < 0.01 and np.abs(np.corrcoef(u.T) - np.eye(7)).max() < 0.03
    # a sample depends on (seed, index) only: any split of the index range gives the same stream
    assert np.array_equal(sample_uniforms(7, 1000, 50, 7), u[1000:1050])
    q = jk.joint_angles_from_uniforms(jk.PANDA, torch.from_numpy(u))
    lims = torch.tensor(jk.PANDA.actuated_joints_limits)
    assert (q >= lims[:, 0]).all() and (q <= lims[:, 1]).all()
