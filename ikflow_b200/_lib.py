"""ctypes binding of ``libikflow_b200.so`` (the C ABI declared in ``include/ikflow_b200.h``).

There is no CPU fallback: if the shared library is missing or fails to load, every entry point raises.  The library is
built in-tree by ``python -m ikflow_b200.csrc.build`` (``__graft_entry__.build()``).
"""

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p

# IKFLOW_B200_LIB: A/B comparisons of two builds on the same GPU box (developer aid)
LIB_PATH = os.environ.get("IKFLOW_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libikflow_b200.so")

IKF_OK = 0
IKF_ESTATUS = -5
IKF_STATUS_NONFINITE = 1
IKF_STATUS_SYNC_TIMEOUT = 2
IKF_STATUS_RANGE = 4
IKF_PRECISION_BF16X3 = 0
IKF_PRECISION_BF16X1 = 1
IKF_PRECISION_FP16X3 = 2
IKF_PRECISION_AUTO = 3
IKF_MAX_WIDTH = 16
IKF_MAX_LINKS = 16
IKF_MAX_DOF = 8


class IkfFlowDesc(ctypes.Structure):
    _fields_ = [
        ("ndim_tot", c_int32),
        ("dim_cond", c_int32),
        ("nb_nodes", c_int32),
        ("coeff_fn_config", c_int32),
        ("hidden", c_int32),
        ("ndof", c_int32),
        ("rnvp_clamp", c_float),
        ("precision", c_int32),
    ]


# name -> (restype, argtypes); every symbol include/ikflow_b200.h declares
PROTOTYPES = {
    "ikf_flow_create": (c_int, [POINTER(IkfFlowDesc), c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, POINTER(c_void_p)]),
    "ikf_flow_destroy": (None, [c_void_p]),
    "ikf_flow_weight_count": (c_size_t, [POINTER(IkfFlowDesc)]),
    "ikf_flow_inverse": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ikf_flow_inverse_blocks": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ikf_flow_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "ikf_flow_set_peers": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ikf_flow_inverse_gather": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_size_t, c_int, c_int, c_void_p]),
    "ikf_flow_set_forward_tables": (c_int, [c_void_p, c_void_p, c_float]),
    "ikf_flow_status": (c_int, [c_void_p, c_void_p, POINTER(c_uint32)]),
    "ikf_flow_poll_status": (c_int, [c_void_p, POINTER(c_uint32)]),
    "ikf_flow_precision": (c_int, [c_void_p]),
    "ikf_flow_last_kernel": (c_char_p, [c_void_p]),
    "ikf_flow_last_cluster": (c_int, [c_void_p]),
    "ikf_flow_debug_trace": (c_int, [c_void_p, c_void_p, c_int]),
    "ikf_flow_info": (c_int, [c_void_p, POINTER(c_size_t), POINTER(c_int), POINTER(c_int)]),
    "ikf_robot_create": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, POINTER(c_void_p)]),
    "ikf_robot_destroy": (None, [c_void_p]),
    "ikf_robot_ndof": (c_int, [c_void_p]),
    "ikf_forward_kinematics": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ikf_sample_joint_angles_and_poses": (c_int, [c_void_p, c_uint64, c_uint64, c_double, c_void_p, c_void_p, c_int, c_void_p]),
    "ikf_clamp_to_joint_limits": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "ikf_lm_step": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_int, c_void_p]),
    "ikf_pose_error": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "ikf_lm_refine": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ikf_evaluate_solutions": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ikf_last_error": (c_char_p, []),
    "ikf_version": (c_char_p, []),
    "ikf_launch_count": (c_uint64, []),
}

_lib = None


class IkflowB200Error(RuntimeError):
    """A call into libikflow_b200 returned a non-zero code."""


def lib() -> ctypes.CDLL:
    """The loaded shared library.  Raises if it is missing -- the product has no other execution path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IkflowB200Error(
                f"{LIB_PATH} is missing: build it with `python -m ikflow_b200.csrc.build` "
                "(ikflow_b200 has no CPU or PyTorch fallback)"
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI and the binding ever diverge
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    if code != IKF_OK:
        msg = lib().ikf_last_error().decode("utf-8", "replace")
        raise IkflowB200Error(f"{what} failed with code {code}: {msg}")


def launch_count() -> int:
    return int(lib().ikf_launch_count())


def version() -> str:
    return lib().ikf_version().decode()
