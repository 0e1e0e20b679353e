"""Host-side logic of the product (no GPU): hyper-parameters, state-dict layout, registry, argument checks, and the
rule that nothing computes on the CPU."""
import os
import pickle

import numpy as np
import pytest
import torch

import ikflow_b200
from ikflow_b200 import evaluation_utils
from ikflow_b200.flow import FlowModel
from ikflow_b200.model import (
    TINY_MODEL_PARAMS,
    IkflowModelParameters,
    glow_cNF_model,
    make_synthetic_state_dict,
    state_dict_keys,
)
from ikflow_b200.model_loading import MODEL_DESCRIPTIONS, get_all_model_names, get_ik_solver, model_filename


def _tiny():
    hp = IkflowModelParameters()
    hp.__dict__.update(TINY_MODEL_PARAMS.__dict__)
    hp.dim_latent_space = 9
    return hp


def test_model_filename_reference_test():
    # reference tests/model_loading_test.py:11-14
    url = "https://storage.googleapis.com/ikflow_models/atlas_desert-sweep-6.pkl"
    assert model_filename(url) == "atlas_desert-sweep-6.pkl"


def test_registry_has_the_reference_models_with_the_reference_hyper_parameters():
    names = get_all_model_names()
    for n in ("panda__full__lp191_5.25m", "panda_lite_tpm", "fetch_full_temp_nsc_tpm", "fetch__large__ns183_9.75m",
              "fetch_arm__large__mh186_9.25m", "rizon4__snowy-brook-208__global_step=2.75M"):
        assert n in names
    d = MODEL_DESCRIPTIONS["panda__full__lp191_5.25m"]  # ikflow/model_descriptions.yaml:10-17
    assert (d["nb_nodes"], d["dim_latent_space"], d["coeff_fn_config"], d["coeff_fn_internal_size"], d["rnvp_clamp"], d["robot_name"]) == (12, 7, 3, 1024, 2.5, "panda")
    d = MODEL_DESCRIPTIONS["fetch_arm__large__mh186_9.25m"]  # :56-63
    assert (d["nb_nodes"], d["dim_latent_space"], d["robot_name"]) == (16, 10, "fetch_arm")
    assert d["model_weights_url"].endswith("fetch_arm__major-hill-186__global_step%3D9.25M.pkl")


def test_default_hyper_parameters_match_reference():
    hp = IkflowModelParameters()  # ikflow/model.py:17-41
    assert (hp.coupling_layer, hp.nb_nodes, hp.dim_latent_space, hp.coeff_fn_config, hp.coeff_fn_internal_size) == ("glow", 12, 9, 3, 1024)
    assert hp.rnvp_clamp == 2.5 and hp.softflow_enabled and not hp.sigmoid_on_output and hp.permute_random_enabled
    assert (TINY_MODEL_PARAMS.nb_nodes, TINY_MODEL_PARAMS.coeff_fn_config, TINY_MODEL_PARAMS.coeff_fn_internal_size) == (3, 2, 256)


def test_state_dict_layout_and_parameter_counts():
    # 4 + nb*18 tensors for coeff_fn_config=3
    hp = IkflowModelParameters()
    hp.dim_latent_space = 7
    keys = state_dict_keys(hp, 8)
    assert len(keys) == 4 + 12 * 18
    assert keys["module_list.2.subnet1.0.weight"] == (1024, 11) and keys["module_list.2.subnet1.6.weight"] == (8, 1024)
    assert keys["module_list.2.subnet2.0.weight"] == (1024, 12) and keys["module_list.2.subnet2.6.bias"] == (6,)
    assert keys["module_list.1.perm"] == (7,) and keys["module_list.0.M_inv"] == (7, 7)
    n_linear = sum(int(np.prod(s)) for k, s in keys.items() if "subnet" in k)
    assert n_linear == 50_860_200


def test_synthetic_state_dict_is_seeded_and_in_layout():
    hp = _tiny()
    lim = ikflow_b200.Panda().actuated_joints_limits
    a, b = make_synthetic_state_dict(hp, lim, seed=0), make_synthetic_state_dict(hp, lim, seed=0)
    c = make_synthetic_state_dict(hp, lim, seed=1)
    assert set(a) == set(state_dict_keys(hp, 8))
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["module_list.2.subnet1.0.weight"], c["module_list.2.subnet1.0.weight"])
    assert a["module_list.1.perm"].dtype == torch.int64
    w = a["module_list.2.subnet1.2.weight"]
    assert w.abs().max() <= 1 / 16 + 1e-7  # U(+-1/sqrt(256))


def test_flow_model_load_state_dict_prefixes_and_errors():
    hp = _tiny()
    robot = ikflow_b200.Panda()
    sd = make_synthetic_state_dict(hp, robot.actuated_joints_limits)
    model = glow_cNF_model(hp, robot, 8, 9)
    assert isinstance(model, FlowModel)
    model.load_state_dict({"_orig_mod." + k: v for k, v in sd.items()})  # ikflow_solver.py:420-426
    model.load_state_dict({"nn_model." + k: v for k, v in sd.items()})   # scripts/download_model_from_wandb_checkpoint.py:13-28
    w = model.flat_weights()
    assert w.dtype == np.float32 and w.size == 429_366
    first = sd["module_list.2.subnet1.0.weight"].reshape(-1).numpy()
    assert np.array_equal(w[: first.size], first)  # block 0 / subnet1 / Linear 0 comes first
    bad = dict(sd)
    del bad["module_list.3.perm_inv"]
    with pytest.raises(RuntimeError, match="missing keys"):
        model.load_state_dict(bad)
    bad = dict(sd)
    bad["module_list.2.subnet1.0.weight"] = torch.zeros(256, 11)
    with pytest.raises(RuntimeError, match="size mismatch"):
        model.load_state_dict(bad)


def test_solver_argument_checks_raise_assertion_errors_like_the_reference(tmp_path):
    hp = _tiny()
    robot = ikflow_b200.Panda()
    solver = ikflow_b200.IKFlowSolver(hp, robot)
    assert solver.ndof == 7 and solver.dim_cond == 8 and solver.network_width == 9 and solver.conditional_size == 8
    y = torch.zeros(4, 7)
    with pytest.raises(AssertionError, match="Model weights have not been loaded"):  # ikflow_solver.py:311
        solver.generate_ik_solutions(y)
    solver.load_state_dict_from_dict(make_synthetic_state_dict(hp, robot.actuated_joints_limits))
    with pytest.raises(AssertionError):  # :313-317
        solver.generate_ik_solutions(torch.zeros(4, 6))
    with pytest.raises(AssertionError):  # single pose needs n
        solver.generate_ik_solutions(torch.zeros(7))
    with pytest.raises(AssertionError, match="refine_solutions is deprecated"):  # :324
        solver.generate_ik_solutions(y, refine_solutions=True)
    with pytest.raises(AssertionError):  # :359
        solver.generate_exact_ik_solutions(torch.zeros(4, 6))
    with pytest.raises(AssertionError, match="repeat_counts must be a tuple"):  # :360
        solver.generate_exact_ik_solutions(y, repeat_counts=[1, 3])
    with pytest.raises(AssertionError):
        ikflow_b200.IKFlowSolver({"nb_nodes": 3}, robot)
    # the pickled state-dict file format of the reference (ikflow_solver.py:413-441)
    path = tmp_path / "w.pkl"
    with open(path, "wb") as f:
        pickle.dump(make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=3), f)
    solver.load_state_dict(str(path))
    with open(path, "wb") as f:
        f.write(b"not a pickle")
    with pytest.raises(pickle.UnpicklingError):
        solver.load_state_dict(str(path))


def test_nothing_computes_on_the_cpu():
    hp = _tiny()
    robot = ikflow_b200.Panda()
    solver = ikflow_b200.IKFlowSolver(hp, robot)
    solver.load_state_dict_from_dict(make_synthetic_state_dict(hp, robot.actuated_joints_limits))
    if torch.cuda.is_available():
        pytest.skip("this check is about the CPU-only container")
    with pytest.raises(RuntimeError, match="no CPU path"):
        solver.generate_ik_solutions(torch.zeros(4, 7))
    with pytest.raises(RuntimeError, match="no CPU path"):
        solver.nn_model(torch.zeros(4, 9), c=torch.zeros(4, 8), rev=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        robot.forward_kinematics(torch.zeros(1, 7))
    with pytest.raises(RuntimeError, match="no CPU path"):
        robot.inverse_kinematics_step_levenburg_marquardt(torch.zeros(1, 7), torch.zeros(1, 7))
    with pytest.raises(RuntimeError, match="no CPU path"):  # the forward (x -> z) direction has no CPU path either
        solver.nn_model(torch.zeros(4, 9), c=torch.zeros(4, 8), rev=False)


def test_product_does_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "ikflow_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text, f"{fn} reaches into oracle/"


def test_get_ik_solver_synthetic_and_robot_checks():
    solver, hp = get_ik_solver("panda__tiny__synthetic")
    assert isinstance(solver, ikflow_b200.IKFlowSolver) and solver._model_weights_loaded
    assert hp.nb_nodes == 3 and hp.robot_name == "panda" and solver.robot.name == "panda"
    with pytest.raises(AssertionError, match="not found in model descriptions"):
        get_ik_solver("no_such_model")
    with pytest.raises(AssertionError):
        get_ik_solver("panda__tiny__synthetic", robot=ikflow_b200.FetchArm())
    with pytest.raises(ValueError):
        ikflow_b200.get_robot("rizon4")


def test_robot_descriptions():
    p = ikflow_b200.Panda()
    assert p.name == "panda" and p.ndof == 7 and len(p.actuated_joints_limits) == 7
    assert p.actuated_joints_limits[3] == (-3.0718, -0.0698)  # reference tests/model_test.py:18-25
    assert ikflow_b200.FetchArm().ndof == 7 and ikflow_b200.Fetch().ndof == 8


def test_joint_limits_exceeded_truth_table():
    # reference tests/evaluation_utils_test.py:37-55
    limits = [(0, 1), (0, 1), (0, 1)]
    configs = torch.tensor([[0.5, 0.5, 0.5], [0.0, 0.5, 1.0], [-0.1, 0.5, 0.5], [0.5, 1.1, 0.5], [0.5, 0.5, 1.0001]])
    assert evaluation_utils.calculate_joint_limits_exceeded(configs, limits).tolist() == [False, False, True, True, True]


def test_draw_latent_matches_reference_semantics():
    torch.manual_seed(0)
    a = ikflow_b200.draw_latent("gaussian", 0.75, (5, 7), "cpu")
    torch.manual_seed(0)
    assert torch.equal(a, 0.75 * torch.randn((5, 7)))
    u = ikflow_b200.draw_latent("uniform", 0.5, (1000, 3), "cpu")
    assert u.min() >= -0.5 and u.max() <= 0.5
    with pytest.raises(AssertionError):
        ikflow_b200.draw_latent("laplace", 1.0, (2, 2), "cpu")
