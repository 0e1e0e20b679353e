"""The oracle pinned against the reference's OWN code, executed here unmodified from /root/reference on top of the
stand-ins of ``oracle/ref_stub.py`` (FrEIA 0.2 and jrl are not installable in this image; read that file's header for
what the stand-ins are).  CPU only; skipped where the reference tree is absent (the GPU box) -- there the frozen outputs
``tests/golden/reference_*.npz`` (``scripts/make_golden_reference.py``) take over, see ``test_gpu_reference_fixtures.py``.

What is pinned here, bit for bit unless stated:
  (a) ``oracle.freia_flow.subnet_forward``          vs the module built by ``ikflow.model.subnet_constructor`` (:51-96)
  (b) the oracle's FixedLinearTransform             vs ``ikflow.model.IkFlowFixedLinearTransform`` (:153-238) + the
                                                       reference's scaling KATs (tests/model_test.py:50-106)
  (c) state-dict layout, permutation seeds, wiring  vs ``ikflow.model.glow_cNF_model`` (:291-356)
  (d) ``oracle.solver.OracleSolver`` host logic     vs ``ikflow.ikflow_solver.IKFlowSolver`` (:85-117, 119-247, 254-411)
  (e) the product's host-side mirror (argument checks, hyper-parameters, registry) vs the same reference objects
  (f) the committed fixtures are what the reference computes today
The GLOW coupling formula inside the stand-in ``GLOWCouplingBlock`` is this repo's restatement: it stays unpinned.
"""
import hashlib
import os

import numpy as np
import pytest
import torch

from ikflow_b200 import model as product_model
from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict, state_dict_keys
from oracle import freia_flow, jrl_kinematics as jk, ref_stub
from oracle.scenarios import DRAW_SEED0, PseudoFlow, seeded_draws
from oracle.solver import OracleSolver

pytestmark = pytest.mark.skipif(not ref_stub.available(), reason="needs /root/reference (build container only)")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    return ref_stub.load()


def _hp(nb_nodes, width, cfg, hidden, softflow=True):
    hp = IkflowModelParameters()
    hp.nb_nodes, hp.dim_latent_space, hp.coeff_fn_config, hp.coeff_fn_internal_size, hp.softflow_enabled = nb_nodes, width, cfg, hidden, softflow
    return hp


def _pair(ref, hp, chain=jk.PANDA, seed=0):
    sd = make_synthetic_state_dict(hp, chain.actuated_joints_limits, seed=seed)
    robot = ref_stub.Robot(chain)
    solver = ref_stub.reference_solver(ref, hp, robot, sd)
    oracle = OracleSolver(chain, sd, hp.nb_nodes, hp.dim_latent_space, hp.coeff_fn_config, hp.rnvp_clamp, hp.softflow_enabled)
    return solver, oracle, sd


def test_reference_modules_come_from_the_reference_tree(ref):
    for m in (ref.model, ref.ikflow_solver, ref.evaluation_utils, ref.model_loading):
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(ref_stub.REFERENCE_ROOT))


# ---------------------------------------------------------------------------------------------------------------- (a)
@pytest.mark.parametrize("n_layers", [1, 2, 3, 4])
def test_subnet_forward_equals_reference_subnet_constructor(ref, n_layers):
    g = torch.Generator().manual_seed(n_layers)
    net = ref.model.subnet_constructor(96, n_layers, 11, 8)  # ikflow/model.py:51-96
    sd = {f"p.{k}": torch.randn(v.shape, generator=g) * 0.3 for k, v in net.state_dict().items()}
    net.load_state_dict({k[2:]: v for k, v in sd.items()})
    x = torch.randn(37, 11, generator=g)
    with torch.inference_mode():
        want = net(x)
    got = freia_flow.subnet_forward(sd, "p", freia_flow.n_linear_layers(n_layers), x)
    assert torch.equal(got, want)
    assert [type(m).__name__ for m in net][1::2] == ["LeakyReLU"] * n_layers and net[1].negative_slope == freia_flow.LEAKY_RELU_SLOPE


# ---------------------------------------------------------------------------------------------------------------- (b)
def test_reference_fixed_linear_transform_scaling_kats(ref):
    """tests/model_test.py:50-106 (the reference runs them on "cuda"; same assertions on the CPU)."""
    panda = ref_stub.Panda()
    upper = torch.tensor([[2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973]])
    lower = torch.tensor([[-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973]])
    mid = torch.tensor([[0.0, 0.0, 0.0, -1.5708, 0.0, 1.8675, 0.0]])
    import FrEIA.framework as Ff  # the stand-in

    _, node = ref.model.get_pre_sigmoid_scaling_node(7, panda, [Ff.InputNode(7, name="input")])
    for x, y in ((upper - 1e-8, torch.ones(1, 7)), (lower + 1e-8, torch.zeros(1, 7)), (mid, 0.5 * torch.ones(1, 7))):
        torch.testing.assert_close(node.forward([x], rev=False)[0][0], y)
    for y, x in ((torch.ones(1, 7), upper), (torch.zeros(1, 7), lower), (0.5 * torch.ones(1, 7), mid)):
        torch.testing.assert_close(node.forward([y], rev=True)[0][0], x)


def test_oracle_fixed_linear_transform_equals_reference_copy_of_freia(ref):
    """``IkFlowFixedLinearTransform`` is the reference's documented copy of FrEIA's FixedLinearTransform
    (ikflow/model.py:151-153): parameters, forward, reverse and log-det of the oracle's version must be identical."""
    lim = jk.PANDA.actuated_joints_limits
    width = 9
    m = torch.eye(width)
    for i, (lo, hi) in enumerate(lim):
        m[i, i] = 1.0 / max(abs(lo), abs(hi))  # ikflow/model.py:311-314
    wide = [(-10.0, 10.0)] * width
    node = ref.model.IkFlowFixedLinearTransform([(width,)], M=m, b=torch.zeros(width), joint_limits=wide)
    mine = freia_flow.build_fixed_linear_transform(width, lim)
    for k in ("M", "M_inv", "b", "logDetM"):
        assert torch.equal(mine[f"module_list.0.{k}"], getattr(node, k).data), k
    u = torch.rand(64, width, generator=torch.Generator().manual_seed(0))  # the copy asserts its reverse input in [0, 1]
    (want,), jac = node.forward([u], rev=True)
    got = (u - mine["module_list.0.b"]).mm(mine["module_list.0.M_inv"])  # oracle/freia_flow.py flow_inverse, last step
    assert torch.equal(got, want)
    assert torch.equal(jac, -mine["module_list.0.logDetM"].expand(64))
    x = want * 0.1  # forward asserts its OUTPUT in [0, 1]
    (fw,), jf = node.forward([x], rev=False)
    assert torch.equal(x.mm(mine["module_list.0.M"]) + mine["module_list.0.b"], fw) and torch.equal(jf, -jac)
    # and the stand-in FixedLinearTransform that glow_cNF_model instantiates is that same arithmetic
    stand_in = ref_stub.FixedLinearTransform([(width,)], M=m, b=torch.zeros(width))
    assert torch.equal(stand_in([u], rev=True)[0][0], want)


# ---------------------------------------------------------------------------------------------------------------- (c)
@pytest.mark.parametrize("nb_nodes,width,cfg,hidden,chain,softflow", [
    (3, 9, 2, 256, jk.PANDA, True),      # TINY_MODEL_PARAMS, odd width (split 4/5, the note at model.py:320-336)
    (12, 7, 3, 64, jk.PANDA, True),      # panda__full geometry (narrow hidden)
    (16, 10, 3, 64, jk.FETCH_ARM, True), # fetch_arm__large geometry
    (2, 8, 1, 32, jk.PANDA, False),      # no softflow column: dim_cond 7
    (2, 7, 4, 32, jk.PANDA, True),
])
def test_state_dict_layout_and_wiring_equal_reference_glow_cnf_model(ref, nb_nodes, width, cfg, hidden, chain, softflow):
    hp = _hp(nb_nodes, width, cfg, hidden, softflow)
    np.random.seed(12345)
    solver, oracle, sd = _pair(ref, hp, chain)
    # building the model reseeds numpy's global RNG to nb_nodes - 1 (PermuteRandom(seed=i))
    np.random.seed(nb_nodes - 1)
    np.random.permutation(width)
    probe = np.random.rand()
    np.random.seed(12345)
    ref_stub.reference_solver(ref, hp, ref_stub.Robot(chain))
    assert np.random.rand() == probe
    built = solver.nn_model.state_dict()
    dim_cond = 8 if softflow else 7
    assert solver.dim_cond == dim_cond
    want_keys = state_dict_keys(hp, dim_cond)
    assert list(built.keys()) == list(sd.keys())
    assert {k: tuple(v.shape) for k, v in built.items()} == {k: tuple(s) for k, s in want_keys.items()}
    assert all(torch.equal(built[k], sd[k]) for k in built)  # incl. M, M_inv, perm tables built by the reference wiring
    fresh = ref_stub.reference_solver(ref, hp, ref_stub.Robot(chain)).nn_model.state_dict()  # before loading weights
    for k in fresh:
        if k.startswith("module_list.0.") or ".perm" in k:
            assert torch.equal(fresh[k], sd[k]), k
    g = torch.Generator().manual_seed(1)
    latent = torch.randn(50, width, generator=g)
    cond = torch.randn(50, dim_cond, generator=g)
    with torch.inference_mode():
        want, want_jac = solver.nn_model(latent, c=cond, rev=True)
        got, got_jac = freia_flow.flow_inverse(sd, latent, cond, nb_nodes, cfg, hp.rnvp_clamp)
        assert torch.equal(got, want) and torch.allclose(got_jac, want_jac, rtol=0, atol=1e-5)
        x = 0.5 * torch.randn(50, width, generator=g)
        want, want_jac = solver.nn_model(x, c=cond, rev=False)
        got, got_jac = freia_flow.flow_forward(sd, x, cond, nb_nodes, cfg, hp.rnvp_clamp)
        assert torch.equal(got, want) and torch.allclose(got_jac, want_jac, rtol=0, atol=1e-5)


def test_hyper_parameters_and_registry_equal_reference(ref):
    mine, theirs = IkflowModelParameters(), ref.model.IkflowModelParameters()
    assert mine.__dict__ == theirs.__dict__
    assert product_model.TINY_MODEL_PARAMS.__dict__ == ref.model.TINY_MODEL_PARAMS.__dict__
    from ikflow_b200.model_loading import MODEL_DESCRIPTIONS, model_filename

    theirs = ref.model_loading.MODEL_DESCRIPTIONS  # yaml.safe_load of ikflow/model_descriptions.yaml
    assert {k: v for k, v in MODEL_DESCRIPTIONS.items() if k in theirs} == theirs
    extra = set(MODEL_DESCRIPTIONS) - set(theirs)  # this repo's synthetic benchmark entries (BASELINE config 5)
    assert all(MODEL_DESCRIPTIONS[k]["model_weights_url"].startswith("synthetic://") for k in extra)
    for d in theirs.values():
        assert model_filename(d["model_weights_url"]) == ref.model_loading.model_filename(d["model_weights_url"])


# ---------------------------------------------------------------------------------------------------------------- (d)
def test_generate_ik_solutions_equals_reference(ref):
    solver, oracle, _ = _pair(ref, _hp(3, 9, 2, 256))
    _, poses = jk.sample_joint_angles_and_poses(jk.PANDA, 40, seed=3)
    latent = torch.randn(40, 9, generator=torch.Generator().manual_seed(4))
    assert torch.equal(solver.generate_ik_solutions(poses, latent=latent), oracle.generate_ik_solutions(poses, latent=latent))
    assert torch.equal(
        solver.generate_ik_solutions(poses, latent=latent, clamp_to_joint_limits=False),
        oracle.generate_ik_solutions(poses, latent=latent, clamp_to_joint_limits=False),
    )
    assert torch.equal(solver.generate_ik_solutions(poses[0], 40, latent=latent), oracle.generate_ik_solutions(poses[0], 40, latent=latent))
    for dist, scale in (("gaussian", 1.0), ("gaussian", 0.75), ("uniform", 0.5)):  # the draws themselves
        torch.manual_seed(9)
        a = solver.generate_ik_solutions(poses, latent_distribution=dist, latent_scale=scale)
        torch.manual_seed(9)
        b = oracle.generate_ik_solutions(poses, latent_distribution=dist, latent_scale=scale)
        assert torch.equal(a, b)
    # return_detailed goes through the reference's evaluate_solutions (evaluation_utils.py:130-147)
    sol, pe, re, exceeded, collides, runtime = solver.generate_ik_solutions(poses, latent=latent, return_detailed=True)
    pe_o, re_o = jk.pose_error(jk.PANDA, sol, poses)
    assert torch.equal(pe, pe_o) and torch.equal(re, re_o)
    assert torch.equal(exceeded, jk.calculate_joint_limits_exceeded(sol, jk.PANDA.actuated_joints_limits))
    assert collides.dtype == torch.bool and isinstance(runtime, float)


def test_reference_relational_flow_tests_hold(ref):
    """tests/ikflow_solver_test.py:94-117 (test_solve_multiple_poses), run against the reference solver itself."""
    hp = IkflowModelParameters()
    hp.__dict__.update(ref.model.TINY_MODEL_PARAMS.__dict__)
    solver = ref_stub.reference_solver(ref, hp, ref_stub.Panda())
    latent = torch.zeros(2, 9)
    ys = torch.zeros(2, 7)
    sols = solver.generate_ik_solutions(ys, None, latent=latent, refine_solutions=False, allow_uninitialized=True)
    torch.testing.assert_close(sols[0], sols[1])
    ys[1, 0] = 1.0
    sols = solver.generate_ik_solutions(ys, None, latent=latent, refine_solutions=False, allow_uninitialized=True)
    # _assert_different (:17-24) as written: element [i, j] of the first row vs every element of the second, for
    # i, j < shape[0] (= 1 for these [1 x 9] arguments)
    assert ((sols[1][None, :] - sols[0][None, :][0, 0]).abs() < 1e-8).sum().item() == 0


@pytest.mark.parametrize("n,sigma,run_lma_on_cpu", [(300, 0.3, False), (300, 0.3, True), (800, 0.3, True), (200, 0.05, False), (120, 1.0, False)])
def test_generate_exact_ik_solutions_equals_reference(ref, n, sigma, run_lma_on_cpu):
    """Every branch of ikflow_solver.py:119-247 / :345-411: early return when all poses converge (sigma 0.05), retries
    with r = 3 and r = 10, poses that never converge, the CPU detour for n < 750 and for n >= 750."""
    solver, oracle, _ = _pair(ref, _hp(1, 7, 1, 32))
    q_true, poses = jk.sample_joint_angles_and_poses(jk.PANDA, n, seed=n)
    flow = PseudoFlow(poses, q_true, sigma)
    solver.nn_model = flow
    oracle._run_inference = lambda latent, cond, clamp: jk.clamp_to_joint_limits(jk.PANDA, flow(latent, c=cond)[0])
    kw = dict(repeat_counts=(1, 3, 10), pos_error_threshold=1e-3, rot_error_threshold=1e-2, run_lma_on_cpu=run_lma_on_cpu)
    torch.manual_seed(n)
    want_s, want_v = solver.generate_exact_ik_solutions(poses, **kw)
    torch.manual_seed(n)
    got_s, got_v = oracle.generate_exact_ik_solutions(poses, **kw)
    assert torch.equal(got_v, want_v) and torch.equal(got_s, want_s)
    assert got_v.dtype == torch.bool and got_s.shape == (n, 7)
    if sigma == 0.05:
        assert want_v.all()
    if sigma == 1.0:
        assert not want_v.all() and want_v.any()
    # the reference's closure assertions (tests/ikflow_solver_test.py:82-86) on what it marked valid
    pe, re = jk.pose_error(jk.PANDA, want_s[want_v], poses[want_v])
    assert (pe < 1e-3).all() and (re < 1e-2).all()
    assert torch.equal(want_s[want_v], jk.clamp_to_joint_limits(jk.PANDA, want_s[want_v].clone()))
    assert (want_s[~want_v] == 0).all()  # unsolved rows keep the zeros of :195


def test_generate_exact_with_the_real_flow_equals_reference(ref):
    solver, oracle, _ = _pair(ref, _hp(3, 7, 2, 64))
    _, poses = jk.sample_joint_angles_and_poses(jk.PANDA, 96, seed=5)
    kw = dict(repeat_counts=(1, 3, 10), pos_error_threshold=1e-3, rot_error_threshold=1e-2, run_lma_on_cpu=False)
    torch.manual_seed(0)
    want_s, want_v = solver.generate_exact_ik_solutions(poses, **kw)
    torch.manual_seed(0)
    got_s, got_v = oracle.generate_exact_ik_solutions(poses, **kw)
    assert torch.equal(got_v, want_v) and torch.equal(got_s, want_s)


# ---------------------------------------------------------------------------------------------------------------- (e)
def test_product_argument_checks_fail_where_the_reference_asserts(ref):
    """Same AssertionErrors for the same bad arguments (ikflow_solver.py:311-326, 359-362) -- evaluated before any
    compute, so this runs without a GPU."""
    import ikflow_b200

    hp = IkflowModelParameters()
    hp.__dict__.update(product_model.TINY_MODEL_PARAMS.__dict__)
    theirs = ref_stub.reference_solver(ref, hp, ref_stub.Panda())
    mine = ikflow_b200.IKFlowSolver(hp, ikflow_b200.Panda())
    y = torch.zeros(7)
    bad_calls = [
        lambda s: s.generate_ik_solutions(y, 5),                                      # weights not loaded
        lambda s: s.generate_ik_solutions([0.0] * 7, 5, allow_uninitialized=True),    # not a tensor
        lambda s: s.generate_ik_solutions(y, None, allow_uninitialized=True),         # single pose needs n
        lambda s: s.generate_ik_solutions(y, 0, allow_uninitialized=True),
        lambda s: s.generate_ik_solutions(torch.zeros(4, 6), allow_uninitialized=True),
        lambda s: s.generate_ik_solutions(y, 5, latent_scale=1, allow_uninitialized=True),
        lambda s: s.generate_ik_solutions(y, 5, latent="z", allow_uninitialized=True),
        lambda s: s.generate_ik_solutions(y, 5, refine_solutions=True, allow_uninitialized=True),
        lambda s: s.generate_exact_ik_solutions(torch.zeros(4, 6)),
        lambda s: s.generate_exact_ik_solutions(torch.zeros(4, 7), repeat_counts=[1, 3]),
        lambda s: s.generate_exact_ik_solutions(torch.zeros(4, 7), return_detailed=True),
        lambda s: s.generate_exact_ik_solutions(torch.zeros(4, 7)),                   # weights not loaded
    ]
    for i, call in enumerate(bad_calls):
        with pytest.raises(AssertionError) as e_ref:
            call(theirs)
        with pytest.raises(AssertionError) as e_mine:
            call(mine)
        assert str(e_mine.value).split("\n")[0] == str(e_ref.value).split("\n")[0], i
    for solver in (theirs, mine):
        assert (solver.ndof, solver.dim_cond, solver.network_width, solver.conditional_size) == (7, 8, 9, 8)
        assert solver.robot.name == "panda" and solver._model_weights_loaded is False
    with pytest.raises(AssertionError):
        ref.ikflow_solver.draw_latent("cauchy", 1.0, (2, 2), "cpu")
    with pytest.raises(AssertionError):
        ikflow_b200.draw_latent("cauchy", 1.0, (2, 2), "cpu")
    torch.manual_seed(1)
    a = ref.ikflow_solver.draw_latent("uniform", 0.5, (5, 7), "cpu"), ref.ikflow_solver.draw_latent("gaussian", 0.75, (5, 7), "cpu"), ref.ikflow_solver.draw_latent("gaussian", 1.0, (5, 7), "cpu")
    torch.manual_seed(1)
    b = ikflow_b200.draw_latent("uniform", 0.5, (5, 7), "cpu"), ikflow_b200.draw_latent("gaussian", 0.75, (5, 7), "cpu"), ikflow_b200.draw_latent("gaussian", 1.0, (5, 7), "cpu")
    assert all(torch.equal(x, y_) for x, y_ in zip(a, b))


def test_evaluation_utils_equal_reference(ref):
    """calculate_joint_limits_exceeded truth table (tests/evaluation_utils_test.py:37-55) through the reference's own
    function and the oracle's."""
    limits = [(0, 1), (0, 1), (0, 1)]
    configs = torch.tensor([[0.5, 0.5, 0.5], [0.0, 0.5, 1.0], [-0.1, 0.5, 0.5], [0.5, 1.1, 0.5], [0.5, 0.5, 1.0001]])
    want = ref.evaluation_utils.calculate_joint_limits_exceeded(configs, limits)
    assert want.tolist() == [False, False, True, True, True]
    assert torch.equal(jk.calculate_joint_limits_exceeded(configs, limits), want)
    q, poses = jk.sample_joint_angles_and_poses(jk.PANDA, 30, seed=1)
    l2, ang = ref.evaluation_utils.solution_pose_errors(ref_stub.Panda(), q + 0.01, poses)
    pe, re = jk.pose_error(jk.PANDA, q + 0.01, poses)
    assert torch.equal(l2, pe) and torch.equal(ang, re)
    l2b, angb = ref.evaluation_utils.solution_pose_errors(ref_stub.Panda(), q + 0.01, poses[0])  # one pose for all
    pe_b, re_b = jk.pose_error(jk.PANDA, q + 0.01, poses[:1].repeat(30, 1))
    assert torch.equal(l2b, pe_b) and torch.equal(angb, re_b)


# ---------------------------------------------------------------------------------------------------------------- (f)
def test_committed_approx_fixture_is_what_the_reference_computes(ref):
    d = np.load(os.path.join(GOLD, "reference_panda_approx.npz"))
    solver, oracle, sd = _pair(ref, _hp(12, 7, 3, 1024))
    poses = torch.from_numpy(d["poses"])
    # tolerance, not equality: MKL's blocking (and with it the fp32 summation order) depends on the thread count
    for tag in ("s100", "s075"):
        latent = torch.from_numpy(d[f"latent_{tag}"])
        got = solver.generate_ik_solutions(poses, latent=latent)
        assert (got - torch.from_numpy(d[f"q_{tag}"])).abs().max() < 2e-5
        assert torch.equal(got, oracle.generate_ik_solutions(poses, latent=latent))  # same process, same shapes: bit-equal


def test_committed_exact_fixture_is_reproducible(ref):
    """Scenario B of scripts/make_golden_reference.py through the reference solver again (0.3 s)."""
    d = np.load(os.path.join(GOLD, "reference_panda_exact_n2048.npz"))
    poses, q_true = torch.from_numpy(d["poses"]), torch.from_numpy(d["q_true"])
    solver, _, _ = _pair(ref, _hp(1, 7, 1, 32))
    solver.nn_model = PseudoFlow(poses, q_true, float(d["sigma_b"]))
    draw, log = seeded_draws()
    original = ref.ikflow_solver.draw_latent
    ref.ikflow_solver.draw_latent = draw
    try:
        sols, valids = solver.generate_exact_ik_solutions(poses, repeat_counts=tuple(int(r) for r in d["repeat_counts"]), pos_error_threshold=float(d["pos_thr"]), rot_error_threshold=float(d["rot_thr"]), run_lma_on_cpu=False)
    finally:
        ref.ikflow_solver.draw_latent = original
    assert torch.equal(sols, torch.from_numpy(d["b_solutions"])) and torch.equal(valids, torch.from_numpy(d["b_valids"]))
    assert [list(s) for s, _ in log] == d["b_draw_shapes"].tolist() and [h for _, h in log] == d["b_draw_sha256"].tolist()
    z = torch.randn(2048, 7, generator=torch.Generator().manual_seed(DRAW_SEED0))
    assert hashlib.sha256(z.numpy().tobytes()).hexdigest() == str(d["b_draw_sha256"][0])
