"""ikflow_b200 -- B200-native (sm_100a) engine for the IKFlow hot path: batched inverse-flow sampling and the
Levenberg-Marquardt refinement loop, behind the reference's ``get_ik_solver()`` / ``IKFlowSolver`` API.

The compute lives in ``lib/libikflow_b200.so`` (C ABI: ``include/ikflow_b200.h``), built from ``csrc/``; Python is the
host layer only.  There is no CPU or PyTorch fallback.
"""

from .config import DEVICE  # noqa: F401
from .ikflow_solver import IKFlowSolver, draw_latent  # noqa: F401
from .model import IkflowModelParameters, TINY_MODEL_PARAMS, glow_cNF_model, make_synthetic_state_dict  # noqa: F401
from .model_loading import get_all_model_names, get_ik_solver  # noqa: F401
from .robots import Fetch, FetchArm, Panda, Robot, get_robot  # noqa: F401

__version__ = "0.1.0"
