"""Host-side helpers for the one-process-per-GPU launch (torchrun) used by bench.py.

The reference's only parallelism is data parallelism over read batches
(classify.cpp:211-219 round-robins 1024-read buffers over worker threads and
:226-229,57-63 sums the per-thread maps).  The same scheme across GPUs: batches are
dealt round-robin to ranks, the k-mer table is replicated, and the int32
per-barcode partial counts are summed to rank 0.  On GPUs the sum is the
ncclReduce inside hast_finish; `reduce_counts` is the torch.distributed form of
the same step, used on CPU (gloo) to test the sharding logic without a GPU.
"""
from __future__ import annotations

import numpy as np


def shard_batches(n_batches: int, rank: int, world: int) -> list[int]:
    """Indices of the batches rank `rank` classifies (round-robin, classify.cpp:214-218)."""
    return list(range(rank, n_batches, world))


def batch_bounds(n_reads: int, reads_per_batch: int) -> list[tuple[int, int]]:
    return [(lo, min(lo + reads_per_batch, n_reads)) for lo in range(0, n_reads, reads_per_batch)]


def strong_slices(n_pairs: int, world: int, fit_per_rank: int | None = None) -> tuple[int, int]:
    """Strong scaling (bench.py cfg3 leg): the pair index space [0, n_pairs) is cut into `world` contiguous slices of
    `per_rank` pairs, rank r taking [r * per_rank, min((r + 1) * per_rank, used)).  `fit_per_rank` caps a slice at what
    one GPU's HBM holds; every rank must call this with the SAME cap (bench.py takes the minimum over ranks).
    -> (per_rank, used): used <= n_pairs is what the job classifies in total."""
    per_rank = (n_pairs + world - 1) // world
    if fit_per_rank is not None:
        per_rank = max(1, min(per_rank, int(fit_per_rank)))
    return per_rank, min(n_pairs, per_rank * world)


def slice_of(rank: int, per_rank: int, used: int) -> tuple[int, int]:
    """(first pair, number of pairs) of rank `rank` under strong_slices"""
    lo = rank * per_rank
    return lo, max(0, min(per_rank, used - lo))


def reduce_counts(counts: np.ndarray, dst: int = 0):
    """Sum int32 counts[n_barcodes][2] over ranks to `dst` (BarcodeCache::Add, classify.cpp:57-63)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(counts, np.int32))
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    return t.numpy() if dist.get_rank() == dst else None


def max_over_ranks(x: float) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_bytes(b: bytes | None, src: int = 0) -> bytes:
    """Ship the 128-byte ncclUniqueId made by hast_comm_unique_id on rank `src` to every rank."""
    import torch.distributed as dist
    box = [b]
    dist.broadcast_object_list(box, src=src)
    return box[0]
