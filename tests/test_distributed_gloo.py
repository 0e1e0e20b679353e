"""The N > 1 path on CPU: world_size 2 over gloo.  The sharding / all-gather logic is what is under test, so the solver
is a stand-in that computes a deterministic function of (pose, latent) row by row."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ikflow_b200.distributed import all_gather_rows, generate_exact_ik_solutions_sharded, generate_ik_solutions_sharded, shard_bounds


def test_shard_bounds_cover_every_row_once():
    for n in (0, 1, 7, 512, 1000, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeSolver:
    """Computes a row-wise function, behind the REAL argument checks of the solver (ikflow_solver.py:311-326, 359-362):
    a [1 x 7] shard is "a single pose" there and needs n, an empty shard must never reach it."""

    ndof = 7

    def generate_ik_solutions(self, y, n=None, latent=None, **kw):
        assert isinstance(y, torch.Tensor)
        if y.numel() == 7:
            assert isinstance(n, int)
            assert n > 0
        else:
            assert y.shape[1] == 7
        n = y.shape[0] if n is None else n
        assert latent is None or latent.shape[0] == n
        assert n % y.reshape(-1, 7).shape[0] == 0
        return y.reshape(-1, 7).expand(n, 7) * 2.0 + (0.0 if latent is None else latent[:, :7])

    def generate_exact_ik_solutions(self, y, **kw):
        assert y.shape[1] == 7 and y.shape[0] > 0
        return y + 1.0, y[:, 0] > 0


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        poses = torch.randn(n, 7, generator=g)
        latent = torch.randn(n, 7, generator=g)
        out = generate_ik_solutions_sharded(_FakeSolver(), poses, latent)
        sol, valid = generate_exact_ik_solutions_sharded(_FakeSolver(), poses)
        lo, hi = shard_bounds(n, rank, world)
        rows = all_gather_rows(torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1), n)
        ok = (
            torch.equal(out, poses * 2.0 + latent)
            and torch.equal(sol, poses + 1.0)
            and torch.equal(valid, poses[:, 0] > 0)
            and torch.equal(rows[:, 0], torch.arange(n, dtype=torch.float32))
        )
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [512, 7, 3, 1, 0])  # even and ragged splits, a one-row shard (n = world + 1), empty shards (n < world)
def test_sharded_solve_equals_unsharded_world_size_2(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
