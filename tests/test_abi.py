"""The C-ABI library loads and exports every symbol include/ikflow_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from ikflow_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ikflow_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ikf_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol(built_lib):
    assert os.path.exists(built_lib)
    handle = ctypes.CDLL(built_lib)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), f"{name} is declared in include/ikflow_b200.h but not exported"
    # and the Python binding covers exactly the declared ABI
    assert sorted(_lib.PROTOTYPES) == declared


def test_version_and_error_string(built_lib):
    assert "sm_100a" in _lib.version()
    assert isinstance(_lib.lib().ikf_last_error(), bytes)
    assert _lib.launch_count() >= 0


def test_weight_count_matches_reference_parameter_counts(built_lib):
    # SURVEY.md App. E (nn.Linear parameters only)
    def count(nb, w, cfg, hidden, ndof=7):
        d = _lib.IkfFlowDesc(w, 8, nb, cfg, hidden, ndof, 2.5, 0)
        return _lib.lib().ikf_flow_weight_count(ctypes.byref(d))

    def expected(nb, w, cfg, hidden):
        s1, s2 = w // 2, w - w // 2
        n = 0
        for cin, cout in ((s1 + 8, 2 * s2), (s2 + 8, 2 * s1)):
            n += cin * hidden + hidden + (cfg - 1) * (hidden * hidden + hidden) + hidden * cout + cout
        return n * nb

    assert count(12, 7, 3, 1024) == expected(12, 7, 3, 1024) == 50_860_200
    assert count(16, 10, 3, 1024) == expected(16, 10, 3, 1024)
    assert count(3, 9, 2, 256) == expected(3, 9, 2, 256)
    # invalid descriptions are rejected, not mis-sized
    assert count(12, 7, 3, 1000) == 0  # hidden not a multiple of 64
    assert count(12, 40, 3, 1024) == 0  # width > 16
    assert count(12, 7, 5, 1024) == 0  # coeff_fn_config out of range


def test_calls_fail_cleanly_without_valid_arguments(built_lib):
    lib = _lib.lib()
    assert lib.ikf_flow_inverse(None, None, 0, None, 0, 0, 0, None, 0, 0, 0, 0, None) == -1  # IKF_EINVAL
    assert b"NULL" in lib.ikf_last_error()
    assert lib.ikf_lm_step(None, None, 0, None, None, 4, 1e-4, 1, None) == -1
    out = ctypes.c_void_p()
    assert lib.ikf_robot_create(0, None, None, None, None, None, 0, ctypes.byref(out)) == -1
    assert out.value is None
