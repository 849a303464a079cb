"""configs[1] at FULL size (100 Mbp trio, k=21, 20 M read pairs, 500 k barcodes) on one B200.

The oracle cannot finish this size in seconds, so parity is carried by properties that hold for any
input (the oracle-checked cases of test_gpu_parity.py pin the per-read arithmetic itself):
  * linearity: counts of the whole job == sum of the counts of its parts, for any split into batches
  * permutation invariance: read order does not matter (per-barcode integer sums)
  * independent kernels agree: pre-filtered fused kernel == direct-probe fused kernel (different table
    access path, same table) == host-packed batches
  * a checksum of checksums: total hap0 / hap1 hits == the number of positions whose standalone K2 -> K3
    lookup carries the tag, computed on a 1 M-read slice by the unfused kernels
  * an oracle spot check on a barcode-complete subsample (every read of 64 barcodes)
The reads are generated on the device (hast_b200/synth.py, torch) and never leave it for the big legs.
"""
import numpy as np
import pytest

import oracle as orc
from hast_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def job():
    import torch
    spec = synth.config("cfg2")
    t = synth.make_trio(spec, device="cuda:0", keep_reads_on_device=True)
    L, P = spec.read_len, spec.n_pairs
    assert t.r2.data_ptr() == t.r1.data_ptr() + P * L
    d_bases = torch.as_strided(t.r1, (2 * P * L,), (1,))
    bc = np.concatenate([t.pair_bc, t.pair_bc]).astype(np.int32)
    d_bc = torch.from_numpy(bc).to("cuda:0")
    return dict(t=t, L=L, P=P, n=2 * P, d_bases=d_bases, d_bc=d_bc, bc=bc, torch=torch)


def _engine(kernel, t):
    from hast_b200.capi import Engine
    e = Engine(0)
    e.set_option("kernel", kernel)
    e.table_begin(t.spec.k, t.pat.size + t.mat.size)
    e.table_add_packed(t.pat, 0)
    e.table_add_packed(t.mat, 1)
    e.table_erase_seq(orc.ADAPTOR_F)
    e.table_erase_seq(orc.ADAPTOR_R)
    e.reserve_barcodes(t.n_barcodes)
    return e


def _run(e, job, sub, order=None):
    torch = job["torch"]
    L, n, nb = job["L"], job["n"], job["t"].n_barcodes
    d_bases, d_bc = job["d_bases"], job["d_bc"]
    d_off = (torch.arange(sub + 1, dtype=torch.int64, device="cuda:0") * L).to(torch.int32)
    e.reset_counts()
    keep = []
    for lo in range(0, n, sub):
        m = min(sub, n - lo)
        if order is None:
            e.submit_batch_device(d_bases.data_ptr() + lo * L, m * L, d_off.data_ptr(), d_bc.data_ptr() + 4 * lo, m)
        else:
            idx = order[lo:lo + m]
            b = d_bases.view(n, L)[idx].contiguous()
            c = d_bc[idx].contiguous()
            keep += [b, c]
            torch.cuda.synchronize()           # torch gathers on its own stream; the engine launches on another
            e.submit_batch_device(b.data_ptr(), m * L, d_off.data_ptr(), c.data_ptr(), m)
            e.sync()
            keep.clear()
    return e.finish(nb), e.stats()["lookups"]


def test_full_size_properties(job):
    torch = job["torch"]
    t, L, n = job["t"], job["L"], job["n"]
    e = _engine(3, t)                              # the default kernel (minimizer-addressed pre-filter), as benchmarked
    whole, lookups = _run(e, job, 4_000_000)
    assert whole.sum() > 1_000_000 and lookups > 3_000_000_000
    # linearity under a different, ragged split
    parts, lookups2 = _run(e, job, 1_234_568)      # ragged, but a multiple of 4 reads: device batches start 16-byte aligned
    assert (parts == whole).all() and lookups2 == lookups
    # permutation invariance
    g = torch.Generator(device="cuda:0")
    g.manual_seed(7)
    perm = torch.randperm(n, device="cuda:0", generator=g)
    shuffled, lookups3 = _run(e, job, 4_000_000, order=perm)
    assert (shuffled == whole).all() and lookups3 == lookups
    e.close()
    # the per-k-mer filter word (kernel 1) and a fixed small tile must agree with it
    e1 = _engine(1, t)
    e1.set_option("reads_per_tile", 100)
    k1, lookups1 = _run(e1, job, 4_000_000)
    assert (k1 == whole).all() and lookups1 == lookups
    e1.close()
    # the direct-probe kernel walks the table differently and must agree
    e0 = _engine(0, t)
    direct, lookups4 = _run(e0, job, 4_000_000)
    assert (direct == whole).all() and lookups4 == lookups

    # checksum of checksums on a slice through the UNFUSED kernels: extract (K2) then lookup (K3)
    m = 1_000_000
    lo = 3_000_000
    d_off = (torch.arange(m + 1, dtype=torch.int64, device="cuda:0") * L).to(torch.int32)
    d_km = torch.full((m * L,), -1, dtype=torch.int64, device="cuda:0")
    d_hasn = torch.zeros(m, dtype=torch.uint8, device="cuda:0")
    e0.extract_kmers_device(job["d_bases"].data_ptr() + lo * L, m * L, d_off.data_ptr(), m, d_km.data_ptr(), d_hasn.data_ptr())
    e0.sync()
    valid = (d_km != -1) & (d_hasn.repeat_interleave(L) == 0)
    km = d_km[valid].contiguous()
    d_tags = torch.zeros(km.numel(), dtype=torch.uint8, device="cuda:0")
    e0.lookup_device(km.data_ptr(), km.numel(), d_tags.data_ptr())
    e0.sync()
    want0, want1 = int((d_tags & 1).sum()), int((d_tags >> 1).sum())
    e0.reset_counts()
    e0.submit_batch_device(job["d_bases"].data_ptr() + lo * L, m * L, d_off.data_ptr(), job["d_bc"].data_ptr() + 4 * lo, m)
    sl = e0.finish(t.n_barcodes)
    assert (int(sl[:, 0].sum()), int(sl[:, 1].sum())) == (want0, want1) and want0 + want1 > 10_000
    assert e0.stats()["lookups"] == km.numel()
    e0.close()

    # oracle spot check: every read of 64 barcodes (barcode-complete subsample)
    pick = np.random.Generator(np.random.PCG64(3)).choice(t.spec.n_barcodes, 64, replace=False)
    sel = np.nonzero(np.isin(job["bc"], pick))[0]
    sub_bases = job["d_bases"].view(n, L)[torch.from_numpy(sel).to("cuda:0")].cpu().numpy()
    o = orc.Oracle()
    o.load_kmers(t.kmer_text(0), 0)
    o.load_kmers(t.kmer_text(1), 1)
    o.init_adaptor()
    off = (np.arange(sel.size + 1, dtype=np.uint64) * L)
    want, _ = o.classify_batch(sub_bases.reshape(-1), off, job["bc"][sel].astype(np.uint32), t.n_barcodes, nthreads=8)
    assert (want[pick] == whole[pick]).all() and want[pick].sum() > 0


def test_barcode_scale_20m():
    """configs[2]'s barcode count on one GPU: 20 M barcodes = 160 MB of per-barcode counters (larger than L2),
    10 M read pairs.  The default fused kernel and the direct-probe kernel agree bit for bit, a ragged
    re-split gives the same counts, barcodes no read carries stay zero, and a barcode-complete subsample
    matches the oracle."""
    import torch
    from hast_b200.capi import Engine
    spec = synth.config("cfg3b")
    t = synth.make_trio(spec, device="cuda:0", keep_reads_on_device=True)
    L, P = spec.read_len, spec.n_pairs
    n = 2 * P
    d_bases = torch.as_strided(t.r1, (n * L,), (1,))
    bc = np.concatenate([t.pair_bc, t.pair_bc]).astype(np.int32)
    d_bc = torch.from_numpy(bc).to("cuda:0")
    nb = t.n_barcodes
    assert nb >= 20_000_000
    job = dict(t=t, L=L, P=P, n=n, d_bases=d_bases, d_bc=d_bc, bc=bc, torch=torch)
    res = {}
    for kern in (3, 0):
        e = _engine(kern, t)
        res[kern], lookups = _run(e, job, 4_000_000)
        if kern == 3:
            again, lookups2 = _run(e, job, 2_345_676)
            assert (again == res[3]).all() and lookups2 == lookups
        e.close()
    assert (res[3] == res[0]).all()
    whole = res[3]
    assert whole.shape == (nb, 2) and whole.sum() > 500_000
    # every counted hit belongs to a barcode some read carries; untouched barcodes stay zero
    seen = np.zeros(nb, bool)
    seen[np.unique(bc)] = True
    assert not whole[~seen].any()
    # the reference's per-barcode sums on a barcode-complete subsample (all reads of 48 barcodes)
    pick = np.random.Generator(np.random.PCG64(5)).choice(np.unique(bc), 48, replace=False)
    sel = np.nonzero(np.isin(bc, pick))[0]
    sub_bases = d_bases.view(n, L)[torch.from_numpy(sel).to("cuda:0")].cpu().numpy()
    o = orc.Oracle()
    o.load_kmers(t.kmer_text(0), 0)
    o.load_kmers(t.kmer_text(1), 1)
    o.init_adaptor()
    off = (np.arange(sel.size + 1, dtype=np.uint64) * L)
    want, _ = o.classify_batch(sub_bases.reshape(-1), off, bc[sel].astype(np.uint32), nb, nthreads=8)
    assert (want[pick] == whole[pick]).all()
