"""Multi-GPU: the gather fused into the flow kernel's epilogue (peer stores over NVLink + per-rank flags) against
generate_ik_solutions + NCCL all-gather, bit for bit.  Needs two GPUs on the box (skipped otherwise); the host-side
sharding logic is covered on CPU by tests/test_distributed_gloo.py."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_gather_equals_nccl_all_gather_two_gpus():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "peer_gather_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    report = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert report["world"] == 2 and report["status"] == 0
    assert report["mismatching_steps_n1024"] == 0 and report["mismatching_steps_n201"] == 0  # even and ragged shards, 60 steps each
    assert report["mismatching_steps_n4803"] == 0  # 2401 / 2402 rows per GPU: the ping-pong kernel writes the peers' buffers


def test_set_peers_argument_checks():
    import ctypes

    import ikflow_b200
    from ikflow_b200 import _lib

    solver, _ = ikflow_b200.get_ik_solver("panda__full__lp191_5.25m", synthetic_seed=0)
    dev = torch.device("cuda", 0)
    h = solver.nn_model._handle(dev)
    lib = _lib.lib()
    assert lib.ikf_flow_set_peers(h, 9, 0, None, None) == -1  # more ranks than a node has GPUs
    assert lib.ikf_flow_set_peers(h, 2, 2, None, None) == -1  # rank out of range / NULL tables
    assert lib.ikf_flow_set_peers(h, 0, 0, None, None) == 0   # off
    lat = torch.zeros(4, 7, device=dev)
    code = lib.ikf_flow_inverse_gather(h, lat.data_ptr(), 7, lat.data_ptr(), 7, 4, 7, 7, 4, 1, 0, 7, 0, None)
    assert code == -1 and b"ikf_flow_set_peers first" in lib.ikf_last_error()
    # a one-rank "node": the gathered tensor is this rank's own buffer, the kernel waits for its own flag only
    buf = torch.zeros(2 * 64 * 7 + 64, device=dev)
    bufs = (ctypes.c_void_p * 1)(buf.data_ptr())
    flags = (ctypes.c_void_p * 1)(buf.data_ptr() + 4 * 2 * 64 * 7)
    assert lib.ikf_flow_set_peers(h, 1, 0, bufs, flags) == 0
    q, poses = solver.robot.sample_joint_angles_and_poses(64, seed=1, return_torch=True, device=dev)
    latent = torch.randn(64, 7, generator=torch.Generator().manual_seed(2)).to(dev)
    ref = solver.generate_ik_solutions(poses, latent=latent)
    for call in range(3):  # ping-pong halves, increasing sequence numbers
        solver.nn_model.inverse_gather(latent, poses, 7, True, (call & 1) * 64 * 7, 7, 0)
        got = buf[(call & 1) * 64 * 7 : ((call & 1) + 1) * 64 * 7].view(64, 7)
        assert torch.equal(got, ref)
    assert int(buf[2 * 64 * 7 :].view(torch.int32)[0]) == 3
    assert lib.ikf_flow_set_peers(h, 0, 0, None, None) == 0
    assert solver.nn_model.status() == 0
