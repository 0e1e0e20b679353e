"""Weight files on either side of the hot path.

The reference ships its models as pickled FrEIA state dicts (``ikflow_solver.py:413-441``), written by
``scripts/download_model_from_wandb_checkpoint.py`` after ``format_state_dict`` has stripped the Lightning prefix
(``:13-28``).  This module keeps that path working (:func:`format_state_dict`, :func:`load_pickled_state_dict`,
:func:`save_pickled_state_dict`) and adds a second, pickle-free container for the same numbers:

``.ikfw`` -- little endian, versioned, checksummed, exactly the arrays ``ikf_flow_create`` takes (include/ikflow_b200.h)::

    0   8   magic  b"IKFLOWB2"
    8   4   u32    format version (1)
    12  4   u32    header bytes (72)
    16  24  6 x i32  ndim_tot, dim_cond, nb_nodes, coeff_fn_config, hidden (coeff_fn_internal_size), ndof
    40  4   f32    rnvp_clamp
    44  4   i32    reserved (0)
    48  8   u64    number of fp32 Linear parameters W
    56  16  reserved
    72      f32[W]                Linear weights/biases in state-dict order (block, subnet1/2, layer: weight then bias)
            i64[nb_nodes][ndim]   PermuteRandom.perm, then i64[nb_nodes][ndim] perm_inv
            f32[ndim][ndim] x 2   FixedLinearTransform M, M_inv;  f32[ndim] b;  f32 logDetM
    end-32  sha256 of every byte before it

A C or C++ caller can mmap the file and hand the sections to ``ikf_flow_create`` without Python or pickle; the Python
side converts both ways without loss (the round trip is bit exact, ``tests/test_weight_files.py``).
"""
from __future__ import annotations

import hashlib
import pickle
import struct
from typing import Dict, Tuple

import numpy as np
import torch

from .model import IkflowModelParameters, state_dict_keys

MAGIC = b"IKFLOWB2"
VERSION = 1
HEADER_BYTES = 72
_HEADER = struct.Struct("<8sII6ifiQ16x")
assert _HEADER.size == HEADER_BYTES


class WeightFileError(RuntimeError):
    """Malformed, truncated, corrupted or unsupported weight file."""


def format_state_dict(state_dict: Dict) -> Dict:
    """``scripts/download_model_from_wandb_checkpoint.py:13-28``: Lightning checkpoints prefix every key with
    ``nn_model.``; strip it.  Like the reference, asserts that EVERY key carries the prefix."""
    bad_prefix = "nn_model."
    updated = {}
    for k, v in state_dict.items():
        assert k[: len(bad_prefix)] == bad_prefix, f"key '{k}' does not start with '{bad_prefix}'"
        updated[k[len(bad_prefix):]] = v
    return updated


def load_pickled_state_dict(filename: str) -> Dict[str, torch.Tensor]:
    """The reference's container (``ikflow_solver.py:416-418``); ``pickle.UnpicklingError`` propagates as there."""
    with open(filename, "rb") as f:
        return pickle.load(f)


def save_pickled_state_dict(filename: str, state_dict: Dict[str, torch.Tensor]) -> None:
    """What ``download_model_from_wandb_checkpoint.py`` writes: a plain pickle of the (CPU) state dict."""
    with open(filename, "wb") as f:
        pickle.dump({k: (v.detach().cpu() if isinstance(v, torch.Tensor) else v) for k, v in state_dict.items()}, f)


def _linear_keys(params: IkflowModelParameters):
    for i in range(params.nb_nodes):
        for sub in ("subnet1", "subnet2"):
            for li in range(params.coeff_fn_config + 1):
                p = f"module_list.{2 + 2 * i}.{sub}.{2 * li}"
                yield p + ".weight"
                yield p + ".bias"


def pack_state_dict(state_dict: Dict[str, torch.Tensor], params: IkflowModelParameters, dim_cond: int, ndof: int) -> bytes:
    """FrEIA state dict -> ``.ikfw`` bytes.  Shapes are checked against the hyper-parameters first."""
    expected = state_dict_keys(params, dim_cond)
    for k, shape in expected.items():
        if k not in state_dict:
            raise WeightFileError(f"state dict has no '{k}'")
        if tuple(state_dict[k].shape) != tuple(shape):
            raise WeightFileError(f"'{k}' has shape {tuple(state_dict[k].shape)}, the hyper-parameters need {tuple(shape)}")
    w = params.dim_latent_space
    f32 = lambda k: state_dict[k].detach().cpu().to(torch.float32).contiguous().numpy().reshape(-1)
    linear = np.concatenate([f32(k) for k in _linear_keys(params)]).astype("<f4")
    perm = np.stack([state_dict[f"module_list.{1 + 2 * i}.perm"].numpy().astype("<i8") for i in range(params.nb_nodes)])
    perm_inv = np.stack([state_dict[f"module_list.{1 + 2 * i}.perm_inv"].numpy().astype("<i8") for i in range(params.nb_nodes)])
    for i in range(params.nb_nodes):
        if sorted(perm[i].tolist()) != list(range(w)) or not np.array_equal(perm[i][perm_inv[i]], np.arange(w)):
            raise WeightFileError(f"block {i}: perm / perm_inv are not inverse permutations of 0..{w - 1}")
    header = _HEADER.pack(MAGIC, VERSION, HEADER_BYTES, w, dim_cond, params.nb_nodes, params.coeff_fn_config,
                          params.coeff_fn_internal_size, ndof, float(params.rnvp_clamp), 0, linear.size)
    body = b"".join([
        header, linear.tobytes(), perm.tobytes(), perm_inv.tobytes(),
        f32("module_list.0.M").astype("<f4").tobytes(), f32("module_list.0.M_inv").astype("<f4").tobytes(),
        f32("module_list.0.b").astype("<f4").tobytes(), f32("module_list.0.logDetM").astype("<f4").tobytes(),
    ])
    return body + hashlib.sha256(body).digest()


def unpack_state_dict(blob: bytes) -> Tuple[Dict[str, torch.Tensor], IkflowModelParameters, int, int]:
    """``.ikfw`` bytes -> (state dict, hyper-parameters, dim_cond, ndof).  Raises :class:`WeightFileError` on a bad
    magic, an unknown version, a wrong size or a checksum mismatch."""
    if len(blob) < HEADER_BYTES + 32:
        raise WeightFileError(f"{len(blob)} bytes is shorter than a header and a checksum")
    magic, version, header_bytes, w, dim_cond, nb, cfg, hidden, ndof, clamp, _res, n_lin = _HEADER.unpack_from(blob, 0)
    if magic != MAGIC:
        raise WeightFileError(f"bad magic {magic!r} (expected {MAGIC!r})")
    if version != VERSION or header_bytes != HEADER_BYTES:
        raise WeightFileError(f"unsupported format version {version} / header of {header_bytes} bytes (this reader: {VERSION} / {HEADER_BYTES})")
    if hashlib.sha256(blob[:-32]).digest() != blob[-32:]:
        raise WeightFileError("checksum mismatch: the file is corrupted or truncated")
    params = IkflowModelParameters()
    params.dim_latent_space, params.nb_nodes, params.coeff_fn_config, params.coeff_fn_internal_size = w, nb, cfg, hidden
    params.rnvp_clamp = float(np.float32(clamp))
    params.softflow_enabled = dim_cond == 8
    expected = state_dict_keys(params, dim_cond)
    want_lin = sum(int(np.prod(expected[k])) for k in _linear_keys(params))
    want = HEADER_BYTES + 4 * want_lin + 2 * 8 * nb * w + 4 * (2 * w * w + w + 1) + 32
    if n_lin != want_lin or len(blob) != want:
        raise WeightFileError(f"size mismatch: header says {n_lin} Linear parameters in {len(blob)} bytes, the description needs {want_lin} in {want}")
    off = HEADER_BYTES
    sd: Dict[str, torch.Tensor] = {}

    def take(dtype, shape):
        nonlocal off
        n = int(np.prod(shape)) if len(shape) else 1
        a = np.frombuffer(blob, dtype=dtype, count=n, offset=off).reshape(shape).copy()
        off += a.nbytes
        return torch.from_numpy(a) if len(shape) else torch.tensor(a.reshape(()).item(), dtype=torch.float32)

    for k in _linear_keys(params):
        sd[k] = take("<f4", expected[k])
    perm = take("<i8", (nb, w))
    perm_inv = take("<i8", (nb, w))
    for i in range(nb):
        sd[f"module_list.{1 + 2 * i}.perm"] = perm[i].clone()
        sd[f"module_list.{1 + 2 * i}.perm_inv"] = perm_inv[i].clone()
    sd["module_list.0.M"] = take("<f4", (w, w))
    sd["module_list.0.M_inv"] = take("<f4", (w, w))
    sd["module_list.0.b"] = take("<f4", (1, w))
    sd["module_list.0.logDetM"] = take("<f4", ())
    return {k: sd[k] for k in expected}, params, dim_cond, ndof


def save_ikfw(filename: str, state_dict: Dict[str, torch.Tensor], params: IkflowModelParameters, dim_cond: int, ndof: int) -> None:
    with open(filename, "wb") as f:
        f.write(pack_state_dict(state_dict, params, dim_cond, ndof))


def load_ikfw(filename: str) -> Tuple[Dict[str, torch.Tensor], IkflowModelParameters, int, int]:
    with open(filename, "rb") as f:
        return unpack_state_dict(f.read())
