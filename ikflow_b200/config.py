"""Global configuration -- mirrors ``ikflow/config.py`` (reference: jstmn/ikflow @ 2f4636e).

``DEVICE`` is what ``jrl.config.DEVICE`` is in the reference (``ikflow/config.py:6``): the CUDA device of this process.
One process drives one GPU (``LOCAL_RANK`` selects it under ``torchrun``); without a GPU the package still imports
(host logic, weight tooling, tests) but every compute call raises -- there is no CPU execution path.
"""

import os

import torch


def _pick_device() -> str:
    if torch.cuda.is_available():
        idx = int(os.environ.get("LOCAL_RANK", "0")) % max(torch.cuda.device_count(), 1)
        return f"cuda:{idx}"
    return "cpu"


DEVICE = _pick_device()
DEFAULT_TORCH_DTYPE = torch.float32  # ikflow/config.py:8

# ~/.cache/ikflow/  (ikflow/config.py:12-18) -- released weight files are looked up in MODELS_DIR
DEFAULT_DATA_DIR = os.path.join(os.path.expanduser("~"), ".cache/ikflow/")
MODELS_DIR = os.path.join(DEFAULT_DATA_DIR, "models/")
