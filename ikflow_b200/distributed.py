"""Batch-axis sharding of the hot path over the GPUs of one node (one process per GPU, ``torch.distributed``).

Every row (target pose, latent draw) of ``generate_ik_solutions`` is independent, and in ``generate_exact_ik_solutions``
all repeats of a pose stay with the pose, so the work shards over contiguous blocks of rows with NO data-path
collective; the weights are replicated (203-272 MB per GPU).  The only exchange is one all-gather of the final
``[n, ndof]`` joint angles (+ the ``[n]`` valid mask for the exact path) -- SURVEY.md section 8(e).
"""

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Rows [lo, hi) of rank ``rank``: contiguous blocks, the first ``n % world_size`` ranks get one extra row."""
    assert 0 <= rank < world_size
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate the per-rank row blocks (sizes given by :func:`shard_bounds`) on every rank: ONE collective."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    if n_total == 0:
        return local
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    if len(set(sizes)) == 1:
        out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # ragged tail: pad to the largest block so that it is still a single all-gather
    mx = max(sizes)
    padded = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * mx : r * mx + sizes[r]] for r in range(world)], dim=0)


def generate_ik_solutions_sharded(
    solver, target_poses: torch.Tensor, latent: Optional[torch.Tensor] = None, gather: bool = True, group=None, **kwargs
) -> torch.Tensor:
    """``solver.generate_ik_solutions`` for the rows of this rank (``target_poses`` / ``latent`` hold ALL n rows on
    every rank, as after a broadcast), followed by the single all-gather of the joint angles."""
    assert target_poses.dim() == 2 and target_poses.shape[1] == 7, f"target_poses must be [n x 7], got {tuple(target_poses.shape)}"
    n = target_poses.shape[0]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(n, rank, world)
    if hi == lo:
        # n < world_size: this rank has no row.  It must still take part in the collective (a rank that raised here
        # would leave the others blocked in the all-gather).
        local = torch.empty((0, solver.ndof), dtype=torch.float32, device=target_poses.device)
    else:
        local_latent = None if latent is None else latent[lo:hi]
        # n=hi-lo: a one-row shard is a [1 x 7] tensor, which generate_ik_solutions reads as "a single pose, n solutions"
        local = solver.generate_ik_solutions(target_poses[lo:hi], hi - lo, latent=local_latent, **kwargs)
    return all_gather_rows(local, n, group) if gather else local


def generate_exact_ik_solutions_sharded(solver, target_poses: torch.Tensor, gather: bool = True, group=None, **kwargs):
    """Pose-sharded ``generate_exact_ik_solutions`` (all repeats of a pose stay on one GPU)."""
    assert target_poses.dim() == 2 and target_poses.shape[1] == 7, f"target_poses must be [n x 7], got {tuple(target_poses.shape)}"
    n = target_poses.shape[0]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(n, rank, world)
    if hi == lo:  # no pose for this rank (n < world_size): empty shard, but stay in the collective
        sol = torch.empty((0, solver.ndof), dtype=torch.float32, device=target_poses.device)
        valid = torch.empty((0,), dtype=torch.bool, device=target_poses.device)
    else:
        sol, valid = solver.generate_exact_ik_solutions(target_poses[lo:hi], **kwargs)
    if not gather:
        return sol, valid
    packed = torch.cat([sol, valid.to(sol.dtype).unsqueeze(1)], dim=1)  # one collective for both
    packed = all_gather_rows(packed, n, group)
    return packed[:, :-1].contiguous(), packed[:, -1] > 0.5
