"""Shared pieces of the exact-IK test scenarios (TEST INFRASTRUCTURE): the seeded latent stream and the trained-like
stand-in for ``nn_model`` used by ``scripts/make_golden_reference.py`` (through the reference's own solver),
``tests/test_reference_host_logic.py`` (reference vs oracle) and ``tests/test_gpu_reference_fixtures.py`` (CUDA path vs
the frozen reference outputs)."""
import hashlib

import torch

DRAW_SEED0 = 1000


def seeded_draws(seed0: int = DRAW_SEED0):
    """A ``draw_latent`` replacement (same signature, ``ikflow/ikflow_solver.py:16-29``): the k-th call returns
    ``torch.randn(shape, generator=manual_seed(seed0 + k))`` drawn on the CPU.  Returns (function, log of
    (shape, sha256))."""
    log = []

    def draw(latent_distribution, latent_scale, shape, device):
        assert latent_distribution == "gaussian" and latent_scale == 1.0
        z = torch.randn(tuple(shape), generator=torch.Generator().manual_seed(seed0 + len(log)))
        log.append((tuple(shape), hashlib.sha256(z.numpy().tobytes()).hexdigest()))
        return z.to(device)

    return draw, log


class PseudoFlow:
    """Duck-typed ``nn_model`` (``ikflow/ikflow_solver.py:98``): ``q_true(pose of the row) + sigma * latent`` -- what a
    trained flow delivers (seeds near a true solution), so that the LM / selection / retry logic sees converging,
    slowly converging and failing poses.  Evaluated on the CPU whatever the device of the inputs (bit-identical seeds
    for every implementation)."""

    def __init__(self, poses: torch.Tensor, q_true: torch.Tensor, sigma: float):
        poses = poses.cpu()
        self.table = {poses[i].numpy().tobytes(): i for i in range(poses.shape[0])}
        self.q_true, self.sigma = q_true.cpu(), sigma

    def rows(self, cond: torch.Tensor) -> torch.Tensor:
        return torch.tensor([self.table[r.numpy().tobytes()] for r in cond[:, :7].cpu().contiguous()], dtype=torch.int64)

    def __call__(self, latent, c=None, rev=True):
        assert rev
        return (self.q_true[self.rows(c)] + self.sigma * latent.cpu()).to(latent.device), None
