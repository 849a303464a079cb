"""GPU, >= 2 devices: sharding read batches over GPUs with a replicated table and ONE
ncclReduce of the int32 per-barcode counts (hast_finish) gives exactly the single-GPU
result -- through the C ABI (two contexts in one process, hast_comm_init_all +
hast_table_clone) and through the drop-in process (bin/classify --gpus 2).
Skipped on a one-GPU box; run with `gpurun --gpus 2`."""
import ctypes as C
import json
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

import cases
from hast_b200 import dist as hd
from hast_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _n_gpus():
    try:
        from hast_b200 import capi
        return capi.load_library().hast_device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")


@needs2
def test_two_contexts_reduce_equals_one_gpu():
    from hast_b200.capi import Engine
    t = synth.make_trio(synth.config("small"))
    k = t.spec.k
    bases, off, bc = t.batch()
    L = t.spec.read_len
    nb = t.n_barcodes

    def build(e):
        e.table_begin(k, t.pat.size + t.mat.size)
        e.table_add_packed(t.pat, 0)
        e.table_add_packed(t.mat, 1)
        e.table_erase_seq(cases.ADAPTOR_F)
        e.table_erase_seq(cases.ADAPTOR_R)

    with Engine(0) as single:
        build(single)
        single.reserve_barcodes(nb)
        single.submit_batch(bases, off, bc)
        want = single.finish(nb)
        want_lookups = single.stats()["lookups"]
    assert want.sum() > 0

    e0, e1 = Engine(0), Engine(1)
    try:
        build(e0)
        e1.table_clone_from(e0)                        # replica over NVLink instead of a rebuild
        i0, i1 = e0.table_info(), e1.table_info()
        assert (i0.size[0], i0.size[1], i0.n_entries) == (i1.size[0], i1.size[1], i1.n_entries)
        arr = (C.c_void_p * 2)(e0.handle, e1.handle)
        rc = e0.lib.hast_comm_init_all(arr, 2)
        assert rc == 0, e0.lib.hast_last_error(e0.handle)
        bounds = hd.batch_bounds(bc.size, 3001)
        for rank, e in enumerate((e0, e1)):
            e.reserve_barcodes(nb)
            for i in hd.shard_batches(len(bounds), rank, 2):
                lo, hi = bounds[i]
                e.submit_batch(bases[lo * L:hi * L], (off[lo:hi + 1] - off[lo]).astype(np.uint32), bc[lo:hi])
        # ncclReduce is a collective: both ranks of this process must be inside it together
        import threading
        out = {}
        th = threading.Thread(target=lambda: out.setdefault("r1", e1.finish(nb, want_counts=False)))
        th.start()
        got = e0.finish(nb)
        th.join()
        assert (got == want).all()
        assert e0.stats()["lookups"] + e1.stats()["lookups"] == want_lookups
        assert e0.stats()["lookups"] > 0 and e1.stats()["lookups"] > 0
    finally:
        e1.close()
        e0.close()


@needs2
def test_cli_two_gpus_same_bytes(tmp_path):
    t = synth.make_trio(synth.config("small"))
    pat, mat = t.write_kmer_lists(tmp_path)
    r1, r2 = t.write_fastq(tmp_path, gz=False)
    outs = []
    for g in (1, 2):
        stats = tmp_path / f"stats{g}.json"
        r = subprocess.run([str(ROOT / "bin" / "classify"), "--hap0", pat, "--hap1", mat, "--read", r1, "--read", r2,
                            "--weight0", "1.04", "--gpus", str(g), "--stats-json", str(stats)],
                           capture_output=True, env=dict(os.environ, HAST_BLOCK_MB="1"))
        assert r.returncode == 0, r.stderr[-600:].decode(errors="replace")
        outs.append(r.stdout)
        s = json.loads(stats.read_text())
        assert s["reads"] == 2 * t.spec.n_pairs
    a, b = outs[0].splitlines(), outs[1].splitlines()
    assert len(a) == len(b), (len(a), len(b))
    diff = [(x, y) for x, y in zip(a, b) if x != y]
    assert not diff, (len(diff), diff[:5])
    assert len(a) == len(set(t.pair_bc.tolist()))
