"""Multi-GPU plan: one process per GPU, MSM sharded by contiguous point range — exactly the
reference's thread split (scalar_multiplication.cpp:703-728, table stride 2) — NTT batches sharded by polynomial.
The only exchange is a gather of one 128-byte XYZZ partial per rank, then the host-side fold
(scalar_multiplication.cpp:750-765 is the CPU analogue).  Works over NCCL (GPU tensors) or gloo (CPU tensors)."""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous point range [lo, hi) of rank `rank`; scalars[lo:hi] pair with table[2*lo:2*hi]."""
    return rank * n // world, (rank + 1) * n // world


def shard_batch(batch, rank, world):
    """Indices of the polynomials rank `rank` transforms (round-robin; empty when world > batch: replicas only)."""
    return [i for i in range(batch) if i % world == rank]


def gather_partials(partial16, world, device=None):
    """All-gather one XYZZ partial (16 uint64) per rank -> (world, 16) uint64 on the host."""
    if world == 1:
        return np.asarray(partial16, dtype=np.uint64).reshape(1, 16)
    import torch
    import torch.distributed as dist

    mine = torch.from_numpy(np.ascontiguousarray(partial16, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        mine = mine.to(device)
    out = torch.empty((world * 16,), dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    return out.cpu().numpy().view(np.uint64).reshape(world, 16)


def normalized_to_partial(jac12):
    """A normalised Jacobian point (x, y, one) as an XYZZ partial (x, y, one, one); infinity -> all zero."""
    part = np.zeros(16, dtype=np.uint64)
    if not (int(jac12[7]) >> 63):
        part[:8] = jac12[:8]
        part[8:12] = jac12[8:12]
        part[12:16] = jac12[8:12]
    return part


def sharded_msm(lib, scalars, table, rank, world, device=None):
    """sum_i scalars[i] P_i with this rank computing only its point range; every rank returns the full result."""
    n = scalars.shape[0]
    lo, hi = shard_range(n, rank, world)
    jac = lib.msm(np.ascontiguousarray(scalars[lo:hi]), np.ascontiguousarray(table[2 * lo:2 * hi]), hi - lo)
    return lib.fold_partials(gather_partials(normalized_to_partial(jac), world, device))
