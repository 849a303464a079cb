"""CPU, world_size 2 over gloo: the N>1 host logic -- round-robin batch sharding, replicated
table, integer sum of the per-rank partial counts to rank 0 -- gives exactly the single-rank
result (the per-rank classification itself is done by the CPU oracle here; on GPUs the same
shards go through hast_submit_batch and the sum is the ncclReduce inside hast_finish)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    import oracle as orc
    from hast_b200 import dist as hd
    from hast_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = synth.make_trio(synth.config("tiny"))          # same seeds on every rank: replicated table
    o = orc.Oracle()
    o.load_kmers(t.kmer_text(0), 0)
    o.load_kmers(t.kmer_text(1), 1)
    o.init_adaptor()
    bases, off, bc = t.batch()
    n = bc.size
    L = t.spec.read_len
    bounds = hd.batch_bounds(n, 333)
    mine = hd.shard_batches(len(bounds), rank, world)
    part = np.zeros((t.n_barcodes, 2), np.int32)
    for i in mine:
        lo, hi = bounds[i]
        c, _ = o.classify_batch(bases[lo * L:hi * L], (off[lo:hi + 1] - off[lo]).astype(np.uint64), bc[lo:hi],
                                t.n_barcodes, nthreads=1)
        part += c
    uid = hd.broadcast_bytes(bytes(range(128)) if rank == 0 else None)
    assert uid == bytes(range(128))
    total = hd.reduce_counts(part, dst=0)
    slowest = hd.max_over_ranks(float(rank + 1))
    assert slowest == float(world)
    if rank == 0:
        full, _ = o.classify_batch(bases, off.astype(np.uint64), bc, t.n_barcodes, nthreads=1)
        np.save(Path(out_dir) / "ok.npy", np.array([int((total == full).all()), int(full.sum() > 0),
                                                    int(sorted(mine) == list(range(0, len(bounds), world)))]))
    else:
        assert total is None
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_sharding_and_reduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok.tolist() == [1, 1, 1]


def test_shard_batches_partition():
    from hast_b200 import dist as hd
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in hd.shard_batches(37, r, world))
        assert seen == list(range(37))
    assert hd.batch_bounds(10, 4) == [(0, 4), (4, 8), (8, 10)]
