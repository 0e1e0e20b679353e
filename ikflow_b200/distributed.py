"""Batch-axis sharding of the hot path over the GPUs of one node (one process per GPU, ``torch.distributed``).

Every row (target pose, latent draw) of ``generate_ik_solutions`` is independent, and in ``generate_exact_ik_solutions``
all repeats of a pose stay with the pose, so the work shards over contiguous blocks of rows with NO data-path
collective; the weights are replicated (203-272 MB per GPU).  The only exchange is one all-gather of the final
``[n, ndof]`` joint angles (+ the ``[n]`` valid mask for the exact path) -- SURVEY.md section 8(e).
"""

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .ikflow_solver import draw_latent


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


class PeerGather:
    """The gather fused into the flow kernel (the alternative to the collective).

    Every rank allocates the SAME symmetric buffer -- two gathered tensors ``[n_total, ndof]`` (ping-pong) -- plus a flag
    array, and maps its peers' copies (``torch.distributed._symmetric_memory``: CUDA VMM handles exchanged over the
    process group, NVLink P2P).  ``generate_ik_solutions`` then is ONE flow launch: its final epilogue stores this rank's
    joint angles into the gathered tensor of every rank, its last CTA raises this rank's flag everywhere and waits for the
    other ranks' flags -- when the kernel ends the gathered tensor is complete; no NCCL call, no second launch.  All ranks must call it the same number of times (SPMD), each with
    the rows ``shard_bounds(n_total, rank, world)`` of the batch; the returned tensor is valid until the call after the
    next one.
    """

    def __init__(self, solver, n_total: int, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.solver, self.group = solver, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n_total, self.ndof = int(n_total), solver.ndof
        assert self.n_total >= self.world, "the fused gather needs at least one row per rank (use all_gather_rows otherwise)"
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.words = 2 * self.n_total * self.ndof
        self.flag_off = (self.words + 31) // 32 * 32  # flags behind the two gathered tensors, 128-byte aligned
        self.buf = symm_mem.empty(self.flag_off + 32, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, dist.group.WORLD if group is None else group)
        torch.cuda.synchronize()
        dist.barrier(group)  # every rank's flags are zero before anybody launches
        base = [int(x) for x in self.handle.buffer_ptrs]
        solver.nn_model.set_peers(self.device, self.world, self.rank, base, [b + 4 * self.flag_off for b in base])
        self.calls = 0
        self.lo, self.hi = shard_bounds(self.n_total, self.rank, self.world)

    def generate_ik_solutions(self, target_poses_local: torch.Tensor, latent_local: Optional[torch.Tensor] = None,
                              latent_distribution: str = "gaussian", latent_scale: float = 1.0, clamp_to_joint_limits: bool = True) -> torch.Tensor:
        """``generate_ik_solutions`` for this rank's rows + the fused gather; returns the gathered ``[n_total, ndof]``."""
        n = self.hi - self.lo
        assert target_poses_local.shape == (n, 7), f"expected this rank's {n} poses, got {tuple(target_poses_local.shape)}"
        assert self.solver._model_weights_loaded, "Model weights have not been loaded. Call load_state_dict(...)"
        if latent_local is None:
            latent_local = draw_latent(latent_distribution, latent_scale, (n, self.solver.network_width), target_poses_local.device)
        half = self.calls & 1
        self.calls += 1
        self.solver.nn_model.inverse_gather(latent_local, target_poses_local, self.ndof, clamp_to_joint_limits, half * self.n_total * self.ndof, self.ndof, self.lo)
        return self.buf[half * self.n_total * self.ndof : (half + 1) * self.n_total * self.ndof].view(self.n_total, self.ndof)

    def close(self):
        self.solver.nn_model.set_peers(self.device, 0, 0, [], [])
