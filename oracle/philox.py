"""CPU restatement of the target-pose generator's random stream (TEST INFRASTRUCTURE).

The product draws ``q ~ U(lo + eps, hi - eps)`` on the GPU with the counter-based generator Philox4x32-10 (Salmon, Moraes,
Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; the generator behind curand / torch.cuda's Philox
engine).  The reference's own sampler (jrl ``Robot.sample_joint_angles_and_poses``, called at
``scripts/benchmark_runtime.py:83-86``) uses numpy's global generator on the host; a device-side generator cannot
reproduce that stream, so parity here is defined on the uniforms: this file restates the published algorithm in numpy,
is pinned by the known-answer vectors of the Random123 distribution (``tests/test_oracle_kats.py``), and the CUDA kernel
must reproduce its words bit for bit; the joint angles then follow from the oracle's affine map fed those uniforms.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """``counter`` [..., 4] uint32, ``key`` [..., 2] uint32 (broadcastable) -> [..., 4] uint32."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k = np.broadcast_to(np.array(key, dtype=np.uint32), c.shape[:-1] + (2,)).copy()
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0, k1 = k[..., 0].copy(), k[..., 1].copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1)


def sample_uniforms(seed: int, first_index: int, n: int, ndof: int) -> np.ndarray:
    """The uniforms of ``ikf_sample_joint_angles_and_poses``: sample s, joint j <- word j % 4 of the block with
    counter (index lo, index hi, j // 4, 0) and key (seed lo, seed hi); u = (word + 0.5) * 2^-32 in (0, 1), fp64."""
    assert 1 <= ndof <= 8
    idx = np.arange(n, dtype=np.uint64) + np.uint64(first_index)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    words = []
    for block in range((ndof + 3) // 4):
        ctr = np.stack(
            [(idx & MASK).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32), np.full(n, block, np.uint32), np.zeros(n, np.uint32)], axis=-1
        )
        words.append(philox4x32_10(ctr, key))
    w = np.concatenate(words, axis=-1)[:, :ndof]
    return (w.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
