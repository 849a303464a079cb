"""The streamed synthetic trio (hast_b200/synth_stream.py) is test/bench tooling, but the large-config parity
checks lean on it: its parent-unique sets must be the exact set difference, and the counter-based read source
must be reproducible pair by pair."""
import numpy as np
import pytest
import torch

import oracle as orc
from hast_b200 import synth, synth_stream as ss


def brute_sets(anc, haps, k):
    ks = {n: synth.canonical_kmers_np(h.numpy(), k) for n, h in haps.items()}
    return synth._unique_sets([ks["P1"], ks["P2"]], [ks["M1"], ks["M2"]], "cpu")


@pytest.mark.parametrize("G,het,k,seed", [(60_000, 0.002, 21, 1), (200_000, 0.001, 11, 2), (50_000, 0.01, 9, 3),
                                          (30_000, 0.0005, 31, 4), (4_000, 0.01, 5, 5)])
def test_parent_unique_sets_equal_brute_force(G, het, k, seed):
    anc = ss.ancestor(G, seed, "cpu")
    names = ("P1", "P2", "M1", "M2")
    sites = {n: ss.snp_sites(G, het, seed, i, "cpu") for i, n in enumerate(names)}
    haps = {}
    for n in names:
        h = anc.clone()
        h[sites[n][0]] = (h[sites[n][0]] + sites[n][1]) & 3
        haps[n] = h
    pat, mat = ss.parent_unique_sets(anc, haps, sites, k, chunk=7_001)     # odd chunk: exercises the chunk seams
    bp, bm = brute_sets(anc, haps, k)
    assert np.array_equal(pat.numpy().astype(np.uint64), bp)
    assert np.array_equal(mat.numpy().astype(np.uint64), bm)
    if k >= 21:
        assert pat.numel() > 0 and mat.numel() > 0


def test_ancestor_is_uniform_and_chunk_independent():
    a = ss.ancestor(100_003, 7, "cpu")
    assert a.dtype == torch.uint8 and int(a.max()) == 3
    frac = torch.bincount(a.to(torch.int64), minlength=4).double() / a.numel()
    assert (frac - 0.25).abs().max() < 0.01
    assert torch.equal(a[:1000], ss.ancestor(1000, 7, "cpu"))


@pytest.fixture(scope="module")
def trio():
    return ss.StreamTrio(ss.stream_config("stream_tiny"), "cpu")


def test_pairs_are_a_pure_function_of_the_index(trio):
    b0, c0 = trio.gen_pairs(100, 50)
    idx = np.arange(100, 150)[::-1].copy()
    b1, c1 = trio.gen_pairs_idx(idx)
    L = trio.spec.read_len
    assert np.array_equal(b0.numpy()[:50][::-1], b1[:50]) and np.array_equal(b0.numpy()[50:][::-1], b1[50:])
    assert np.array_equal(c0.numpy().view(np.uint32)[:50][::-1], c1[:50])
    assert np.array_equal(trio.barcodes_of(100, 50).numpy(), c0.numpy()[:50])
    assert b0.shape == (100, L)


def test_reads_come_from_the_child_haplotypes(trio):
    n = 1500
    bases, bc = trio.gen_pairs(0, n)
    bases = bases.numpy()
    L = trio.spec.read_len
    h0 = synth.LETTERS[trio.hap0.numpy()]
    h1 = synth.LETTERS[trio.hap1.numpy()]
    win0 = np.lib.stride_tricks.sliding_window_view(h0, L)
    win1 = np.lib.stride_tricks.sliding_window_view(h1, L)
    comp = np.zeros(256, np.uint8)
    comp[list(b"ACGTN")] = list(b"TGCAN")
    exact = 0
    for r in range(0, n, 37):
        fwd = bases[r]
        rc2 = comp[bases[n + r]][::-1]
        for read in (fwd, rc2):
            d = np.minimum((win0 != read).sum(1).min(), (win1 != read).sum(1).min())
            assert d <= 4                         # at most 3 substitutions and one N
            exact += d == 0
    assert exact > 0.6 * 2 * len(range(0, n, 37))
    special = (bc.numpy().view(np.uint32) == trio.spec.n_barcodes).mean()
    assert 0.01 < special < 0.06
    assert 0.001 < (bases == ord("N")).any(1).mean() < 0.02


def test_barcode_complete_subsample(trio):
    ids = [3, 17, 42, trio.spec.n_barcodes]
    idx = trio.pairs_of_barcodes(ids, chunk=777)
    allbc = trio.barcodes_of(0, trio.spec.n_pairs).numpy().view(np.uint32)
    want = np.nonzero(np.isin(allbc, ids))[0]
    assert np.array_equal(np.sort(idx), want)


def test_names_are_distinct_triples(trio):
    blob, off = trio.barcode_name_blob()
    names = blob.split(b"\0")[:-1]
    assert len(names) == trio.n_barcodes == len(set(names)) and names[-1] == b"0_0_0"
    for nm in names[:-1]:
        a = [int(x) for x in nm.split(b"_")]
        assert len(a) == 3 and all(1 <= v <= 1536 for v in a)
    assert blob[int(off[5]):].split(b"\0")[0] == names[5]


def test_zipf_cdf_is_heavy_tailed():
    spec = ss.stream_config("stream_tiny")
    spec.zipf_alpha = 1.2
    spec.n_barcodes = 5000
    t = ss.StreamTrio(spec, "cpu", sets=False)
    bc = t.barcodes_of(0, 200_000).numpy().view(np.uint32)
    bc = bc[bc != spec.n_barcodes]
    cnt = np.bincount(bc, minlength=spec.n_barcodes)
    assert cnt[0] > 20 * np.median(cnt[cnt > 0]) and cnt[0] == cnt.max()


@pytest.mark.parametrize("gz", [False, True])
def test_fastq_of_the_stream_matches_the_batch_through_the_oracle(trio, tmp_path, gz):
    n = 600
    paths = trio.write_fastq(tmp_path, lo=0, hi=n, gz=gz, chunk=250)
    o = orc.Oracle()
    o.load_kmers(trio.kmer_text(0), 0)
    o.load_kmers(trio.kmer_text(1), 1)
    o.init_adaptor()
    for p in paths:
        assert o.process_fastq(p) == 0, o.err()
    table = [ln.split("\t") for ln in o.table(tmp_path).decode().splitlines()]
    bases, bc = trio.gen_pairs(0, n)
    off = (np.arange(2 * n + 1, dtype=np.uint64) * trio.spec.read_len)
    want, _ = o.classify_batch(bases.numpy().reshape(-1), off, bc.numpy().view(np.uint32), trio.n_barcodes)
    blob, _ = trio.barcode_name_blob()
    names = blob.split(b"\0")[:-1]
    got = {row[0]: (int(row[2]), int(row[3])) for row in table}
    seen = set(bc.numpy().view(np.uint32).tolist())
    assert set(got) == {names[i].decode() for i in seen}
    for i in seen:
        assert got[names[i].decode()] == (int(want[i, 0]), int(want[i, 1]))


@pytest.mark.gpu
def test_device_generator_equals_host_generator():
    spec = ss.stream_config("stream_small")
    a = ss.StreamTrio(spec, "cpu")
    b = ss.StreamTrio(spec, "cuda:0")
    assert np.array_equal(a.pat, b.pat) and np.array_equal(a.mat, b.mat) and a.pat.size > 1000
    ba, ca = a.gen_pairs(12_345, 20_000)
    bb, cb = b.gen_pairs(12_345, 20_000)
    torch.cuda.synchronize()
    assert torch.equal(ba, bb.cpu()) and torch.equal(ca, cb.cpu())
    idx = np.array([5, 99_999, 17, 4242], dtype=np.int64)
    xa, ya = a.gen_pairs_idx(idx)
    xb, yb = b.gen_pairs_idx(idx)
    assert np.array_equal(xa, xb) and np.array_equal(ya, yb)
    assert np.array_equal(a.pairs_of_barcodes([1, 2, 3]), b.pairs_of_barcodes([1, 2, 3]))
