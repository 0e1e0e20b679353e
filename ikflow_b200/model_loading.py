"""Model registry and loader -- ``ikflow/model_loading.py`` of the reference (jstmn/ikflow @ 2f4636e).

``get_ik_solver(model_name)`` keeps the reference signature and return value.  Weight files are the reference's own
pickled state dicts, cached under ``~/.cache/ikflow/models/`` (``ikflow/config.py:12-18``); they are downloaded on
first use like the reference does.  Two additions, both opt-in:

* ``synthetic_seed``: if the weight file is not cached and cannot be downloaded (no network), build seeded weights in
  the same state-dict layout instead of failing -- benchmarks and tests on machines without the released files;
* registry entries whose ``model_weights_url`` starts with ``synthetic://`` never touch the network.
"""

import os
import sys
import urllib.error
from typing import Dict, Optional, Tuple
from urllib.request import urlretrieve

import yaml

from .config import MODELS_DIR
from .ikflow_solver import IKFlowSolver
from .model import IkflowModelParameters, make_synthetic_state_dict
from .robots import Robot, get_robot

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_descriptions.yaml"), "r") as _f:
    MODEL_DESCRIPTIONS = yaml.safe_load(_f)


def _assert_model_downloaded_correctly(filepath: str):
    filesize_mb = os.path.getsize(filepath) * 0.000001
    assert filesize_mb > 10, (
        f"Model weights saved at '{filepath}' has only {filesize_mb} MB - was it saved correctly? Tip: Check that the"
        " file is publically available on GCP."
    )


def get_all_model_names() -> Tuple[str]:
    """Return a tuple of the model names"""
    return tuple(MODEL_DESCRIPTIONS.keys())


def model_filename(url: str) -> str:
    """https://storage.googleapis.com/ikflow_models/atlas_desert-sweep-6.pkl -> atlas_desert-sweep-6.pkl"""
    return url.split("/")[-1]


def download_model(url: str, download_dir: Optional[str] = None) -> str:
    """Return the cached path of the weight file at ``url``, downloading it into ``download_dir`` (default
    ``MODELS_DIR``) if it is not there yet."""
    if download_dir is None:
        download_dir = MODELS_DIR
    os.makedirs(download_dir, exist_ok=True)
    assert os.path.isdir(download_dir), f"Download directory {download_dir} does not exist"
    save_filepath = os.path.join(download_dir, model_filename(url))
    if os.path.isfile(save_filepath):
        _assert_model_downloaded_correctly(save_filepath)
        return save_filepath
    urlretrieve(url, filename=save_filepath)
    _assert_model_downloaded_correctly(save_filepath)
    return save_filepath


def _synthetic_seed_from_url(url: str) -> Optional[int]:
    if not url.startswith("synthetic://"):
        return None
    for kv in url[len("synthetic://"):].split("&"):
        k, _, v = kv.partition("=")
        if k == "seed":
            return int(v)
    return 0


def get_ik_solver(
    model_name: str,
    robot: Optional[Robot] = None,
    compile_model: Optional[Dict] = None,
    synthetic_seed: Optional[int] = None,
) -> Tuple[IKFlowSolver, IkflowModelParameters]:
    """Build and return the `IKFlowSolver` for the given model. The input `model_name` should match an index in
    `model_descriptions.yaml`.

    Returns:
        Tuple[IKFlowSolver, IkflowModelParameters]: the solver and its hyper-parameters
    """
    assert model_name in MODEL_DESCRIPTIONS, f"Model name '{model_name}' not found in model descriptions"
    hparams = MODEL_DESCRIPTIONS[model_name]
    model_weights_url = hparams["model_weights_url"]
    robot_name = hparams["robot_name"]
    assert isinstance(robot_name, str), f"robot_name must be a string, got {type(robot_name)}"
    assert isinstance(hparams, dict), f"model_hyperparameters must be a Dict, got {type(hparams)}"

    if robot is None:
        robot = get_robot(robot_name)
    assert robot.name == robot_name

    hyper_parameters = IkflowModelParameters()
    hyper_parameters.__dict__.update(hparams)
    ik_solver = IKFlowSolver(hyper_parameters, robot, compile_model=compile_model)

    url_seed = _synthetic_seed_from_url(model_weights_url)
    if url_seed is not None:
        seed = url_seed if synthetic_seed is None else synthetic_seed
        ik_solver.load_state_dict_from_dict(make_synthetic_state_dict(hyper_parameters, robot.actuated_joints_limits, seed=seed))
        return ik_solver, hyper_parameters

    try:
        cached = os.path.join(MODELS_DIR, model_filename(model_weights_url))
        if synthetic_seed is not None and os.environ.get("IKFLOW_B200_OFFLINE") == "1" and not os.path.isfile(cached):
            raise OSError("offline (IKFLOW_B200_OFFLINE=1) and the weight file is not cached")
        model_weights_filepath = download_model(model_weights_url)
    except (urllib.error.URLError, OSError) as e:
        if synthetic_seed is None:
            raise
        print(f"get_ik_solver(): '{model_name}' weights unavailable ({e}); using synthetic weights, seed {synthetic_seed}", file=sys.stderr)
        ik_solver.load_state_dict_from_dict(
            make_synthetic_state_dict(hyper_parameters, robot.actuated_joints_limits, seed=synthetic_seed)
        )
        return ik_solver, hyper_parameters
    assert os.path.isfile(
        model_weights_filepath
    ), f"File '{model_weights_filepath}' was not found. Unable to load model weights"
    ik_solver.load_state_dict(model_weights_filepath)
    return ik_solver, hyper_parameters
