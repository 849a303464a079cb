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


def _worker_strong(rank, world, port, out_dir):
    """bench.py's cfg3 leg on CPU: every rank builds the same streamed trio, generates ITS contiguous slice of the pair
    index space from (seed, pair index), classifies it (oracle standing in for the GPU), the partial counts are summed to
    rank 0, and rank 0 checks a barcode-complete subsample of the REDUCED counts against the oracle on reads it
    regenerates itself -- pairs that lie in both ranks' slices."""
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    import oracle as orc
    from hast_b200 import dist as hd
    from hast_b200 import synth_stream as ss
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = ss.stream_config("stream_tiny")
    spec.n_pairs = 3001                                    # not a multiple of the world size
    t = ss.StreamTrio(spec, "cpu")
    o = orc.Oracle()
    o.load_kmers_packed(t.pat, spec.k, 0)
    o.load_kmers_packed(t.mat, spec.k, 1)
    o.init_adaptor()
    per_rank, used = hd.strong_slices(spec.n_pairs, world)
    lo, n = hd.slice_of(rank, per_rank, used)
    bases, bc = t.gen_pairs(lo, n)
    L = spec.read_len
    off = np.arange(2 * n + 1, dtype=np.uint64) * np.uint64(L)
    part, _ = o.classify_batch(bases.numpy().reshape(-1), off, bc.numpy().view(np.uint32), t.n_barcodes, nthreads=1)
    local_sum = np.array([int(part[:, 0].sum()), int(part[:, 1].sum())], np.int64)     # before the reduce (in place on rank 0)
    total = hd.reduce_counts(part, dst=0)
    import torch
    ls = torch.from_numpy(local_sum.copy())
    dist.all_reduce(ls)
    if rank == 0:
        ids = np.arange(3, spec.n_barcodes, 7)
        idx = t.pairs_of_barcodes(ids, 0, used)
        sb, s_bc = t.gen_pairs_idx(idx)
        s_off = np.arange(sb.shape[0] + 1, dtype=np.uint64) * np.uint64(L)
        want, _ = o.classify_batch(sb.reshape(-1), s_off, s_bc, t.n_barcodes, nthreads=1)
        by_rank = np.bincount(idx // per_rank, minlength=world)
        np.save(Path(out_dir) / "strong.npy", np.array([
            int((total[ids] == want[ids]).all()), int(want[ids].sum() > 0), int((by_rank > 0).all()),
            int(ls.tolist() == [int(total[:, 0].sum()), int(total[:, 1].sum())]), int(used == spec.n_pairs)]))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_strong_scaling_slices_and_subsample_parity(tmp_path):
    port = _free_port()
    mp.spawn(_worker_strong, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "strong.npy").tolist() == [1, 1, 1, 1, 1]


def test_strong_slices_cover_the_index_space():
    from hast_b200 import dist as hd
    for n, world in ((600_000_000, 8), (3001, 2), (7, 8), (100, 1)):
        per, used = hd.strong_slices(n, world)
        sl = [hd.slice_of(r, per, used) for r in range(world)]
        assert used == n and sum(c for _, c in sl) == n
        pos = 0
        for lo, c in sl:
            assert (lo == pos or c == 0) and c >= 0
            pos += c
    per, used = hd.strong_slices(600_000_000, 2, fit_per_rank=250_000_000)     # HBM holds less than a slice
    assert (per, used) == (250_000_000, 500_000_000)
