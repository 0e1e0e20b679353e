"""bench.py's host-side helpers: kernel-name matching for the committed ncu summaries, work / byte counts of SURVEY 8(d)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_kernel_template_arguments_match_library_and_ncu_spellings():
    b = _bench()
    lib = b._kernel_template_args("ikf::umma::flow_inverse_umma_kernel<32,false,true,ksplit>")
    ncu = b._kernel_template_args("void flow_inverse_umma_kernel<32, 0, 1, 1, 0>(FlowParams)")
    assert lib == ncu == ("_umma_kernel", [32, 0, 1, 1, 0])
    assert b._kernel_template_args("ikf::umma::flow_inverse_umma_kernel<128,false,false,pingpong>")[1] == [128, 0, 0, 0, 1]
    # summaries of round 1 predate the later template arguments
    assert b._kernel_template_args("void ikf::umma::flow_inverse_umma_kernel<(int)32, (bool)1>(ikf::FlowParams)")[1] == [32, 1, 0, 0, 0]
    assert b._kernel_template_args("ikf::flow_inverse_kernel<64>") == ("_kernel", [64, 0, 0, 0, 0])


def test_traffic_lookup_finds_the_committed_captures():
    b = _bench()
    ks, src = b.traffic_from_profiles("ikf::umma::flow_inverse_umma_kernel<32,false,true,ksplit>", 512)
    assert src == "profiles/r2b_flow_umma_ks_b512_ncu_summary.txt" and 203e6 < ks < 215e6  # the weights once + scratch
    pp, src = b.traffic_from_profiles("ikf::umma::flow_inverse_umma_kernel<128,false,true,pingpong>", 8192)
    assert src == "profiles/r2b_flow_umma_pp_b8192_ncu_summary.txt" and pp > 2 * 203e6  # one pass over the weights per round
    assert b.traffic_from_profiles("ikf::umma::flow_inverse_umma_kernel<32,false,false,ksplit>", 512) == (None, None)  # no capture of that variant


def test_algorithmic_work_is_the_survey_figure():
    b = _bench()
    assert b.FLOW_FLOPS["panda__full__lp191_5.25m"] == 101_572_608  # SURVEY 8(d), FLOPs per solution
