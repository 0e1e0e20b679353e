"""Weight files either side of the path: the reference's pickled FrEIA state dicts
(`ikflow_solver.py:413-441`, `scripts/download_model_from_wandb_checkpoint.py:13-28`) and the .ikfw container."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

from ikflow_b200 import weight_files as wf
from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict, state_dict_keys
from ikflow_b200.robots import get_robot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tiny(nb=3, w=9, cfg=2, hidden=64):
    hp = IkflowModelParameters()
    hp.nb_nodes, hp.dim_latent_space, hp.coeff_fn_config, hp.coeff_fn_internal_size = nb, w, cfg, hidden
    robot = get_robot("panda")
    return hp, robot, make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=3)


def test_format_state_dict_strips_the_lightning_prefix_like_the_reference():
    sd = {"nn_model.module_list.0.M": 1, "nn_model.module_list.0.M_inv": 2}
    assert wf.format_state_dict(sd) == {"module_list.0.M": 1, "module_list.0.M_inv": 2}
    with pytest.raises(AssertionError):  # the reference asserts that the dict is malformatted first (:24-25)
        wf.format_state_dict({"module_list.0.M": 1})


def test_ikfw_round_trip_is_bit_exact():
    hp, robot, sd = _tiny()
    blob = wf.pack_state_dict(sd, hp, 8, robot.ndof)
    sd2, hp2, dim_cond, ndof = wf.unpack_state_dict(blob)
    assert (hp2.nb_nodes, hp2.dim_latent_space, hp2.coeff_fn_config, hp2.coeff_fn_internal_size) == (3, 9, 2, 64)
    assert (dim_cond, ndof, hp2.rnvp_clamp) == (8, 7, 2.5)
    assert list(sd2) == list(state_dict_keys(hp, 8))
    for k, v in sd2.items():
        assert v.dtype == sd[k].dtype or k.endswith("logDetM"), k
        assert torch.equal(v.to(sd[k].dtype), sd[k]), k
    assert wf.pack_state_dict(sd2, hp2, dim_cond, ndof) == blob  # and back: identical bytes


def test_ikfw_sections_are_what_ikf_flow_create_takes():
    """The Linear section is FlowModel.flat_weights() (the array handed to ikf_flow_create), byte for byte."""
    import ikflow_b200

    hp, robot, sd = _tiny()
    model = ikflow_b200.glow_cNF_model(hp, robot, 8, 9)
    model.load_state_dict(sd)
    flat = model.flat_weights()
    blob = wf.pack_state_dict(sd, hp, 8, robot.ndof)
    assert np.array_equal(np.frombuffer(blob, dtype="<f4", count=flat.size, offset=wf.HEADER_BYTES), flat)


def test_ikfw_detects_corruption_truncation_and_foreign_files():
    hp, robot, sd = _tiny()
    blob = bytearray(wf.pack_state_dict(sd, hp, 8, robot.ndof))
    bad = bytearray(blob)
    bad[wf.HEADER_BYTES + 1000] ^= 0x01
    with pytest.raises(wf.WeightFileError, match="checksum"):
        wf.unpack_state_dict(bytes(bad))
    with pytest.raises(wf.WeightFileError):
        wf.unpack_state_dict(bytes(blob[:-5]))
    with pytest.raises(wf.WeightFileError, match="magic"):
        wf.unpack_state_dict(b"PK\x03\x04" + bytes(blob[4:]))
    newer = bytearray(blob)
    newer[8] = 2  # version 2
    with pytest.raises(wf.WeightFileError, match="version"):
        wf.unpack_state_dict(bytes(newer))
    with pytest.raises(wf.WeightFileError):
        wf.unpack_state_dict(b"")


def test_pack_rejects_wrong_shapes_and_broken_permutations():
    hp, robot, sd = _tiny()
    wrong = dict(sd)
    wrong["module_list.2.subnet1.0.weight"] = torch.zeros(3, 3)
    with pytest.raises(wf.WeightFileError, match="shape"):
        wf.pack_state_dict(wrong, hp, 8, robot.ndof)
    wrong = dict(sd)
    wrong["module_list.1.perm_inv"] = sd["module_list.1.perm"].clone()
    if not torch.equal(sd["module_list.1.perm"][sd["module_list.1.perm"]], torch.arange(9)):
        with pytest.raises(wf.WeightFileError, match="inverse permutations"):
            wf.pack_state_dict(wrong, hp, 8, robot.ndof)
    missing = {k: v for k, v in sd.items() if k != "module_list.0.b"}
    with pytest.raises(wf.WeightFileError, match="no 'module_list.0.b'"):
        wf.pack_state_dict(missing, hp, 8, robot.ndof)


def test_solver_loads_both_containers(tmp_path):
    import ikflow_b200

    hp, robot, sd = _tiny(w=9)
    solver = ikflow_b200.IKFlowSolver(hp, robot)
    pkl, ikfw = str(tmp_path / "m.pkl"), str(tmp_path / "m.ikfw")
    wf.save_pickled_state_dict(pkl, {"nn_model." + k: v for k, v in sd.items()})
    wf.save_pickled_state_dict(pkl, wf.format_state_dict(wf.load_pickled_state_dict(pkl)))
    wf.save_ikfw(ikfw, sd, hp, solver.dim_cond, robot.ndof)
    for path in (pkl, ikfw):
        s = ikflow_b200.IKFlowSolver(hp, robot)
        assert not s._model_weights_loaded
        s.load_state_dict(path)
        assert s._model_weights_loaded
        assert all(torch.equal(a, sd[k]) for k, a in s.nn_model.state_dict().items() if not k.endswith("logDetM"))
    other = IkflowModelParameters()
    other.__dict__.update(hp.__dict__)
    other.nb_nodes = 4
    with pytest.raises(AssertionError, match="describes"):
        ikflow_b200.IKFlowSolver(other, robot).load_state_dict(ikfw)
    with open(str(tmp_path / "junk.pkl"), "wb") as f:
        f.write(b"not a pickle")
    with pytest.raises(pickle.UnpicklingError):  # re-raised as in ikflow_solver.py:439-441
        solver.load_state_dict(str(tmp_path / "junk.pkl"))


def test_convert_script_both_directions(tmp_path):
    hp = IkflowModelParameters()
    hp.nb_nodes, hp.dim_latent_space, hp.coeff_fn_config, hp.coeff_fn_internal_size = 6, 7, 3, 1024  # panda_lite_tpm
    robot = get_robot("panda")
    sd = make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=1)
    pkl, ikfw, back = str(tmp_path / "a.pkl"), str(tmp_path / "a.ikfw"), str(tmp_path / "b.pkl")
    wf.save_pickled_state_dict(pkl, sd)
    script = os.path.join(ROOT, "scripts", "convert_weights.py")
    subprocess.run([sys.executable, script, pkl, ikfw, "--model_name", "panda_lite_tpm"], check=True, capture_output=True)
    subprocess.run([sys.executable, script, ikfw, back], check=True, capture_output=True)
    sd2 = wf.load_pickled_state_dict(back)
    assert all(torch.equal(sd2[k].to(sd[k].dtype), sd[k]) for k in sd)
