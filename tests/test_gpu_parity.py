"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Bar: bit-exact (all integer work).  Sizes are chosen so the oracle finishes in
seconds; the full-size checks use size-independent properties (linearity under
batch splitting, permutation invariance, GPU-vs-GPU agreement of independent
kernels K2+K3 vs the fused kernel).
"""
import numpy as np
import pytest

import cases
import oracle as orc
from hast_b200 import synth

pytestmark = pytest.mark.gpu

K_SWEEP = [5, 16, 17, 21, 25, 31, 32]


def build_table(engine, case, adaptor=True, expected=None):
    k = case["k"]
    n = (len(case["pat_text"]) + len(case["mat_text"])) // (k + 1)
    engine.table_begin(k, expected or n)
    engine.table_add_text(case["pat_text"], k, 0)
    engine.table_add_text(case["mat_text"], k, 1)
    erased = []
    if adaptor:
        erased += engine.table_erase_seq(cases.ADAPTOR_F)
        erased += engine.table_erase_seq(cases.ADAPTOR_R)
    return erased


def build_oracle(case, adaptor=True):
    o = orc.Oracle()
    assert o.load_kmers(case["pat_text"], 0) >= 0
    assert o.load_kmers(case["mat_text"], 1) >= 0
    n_erased = o.init_adaptor() if adaptor else 0
    return o, n_erased


# ---- reference known answers (TestAll, classify.cpp:341-367) on the device ----
def test_testall_constants_on_device(engine):
    engine.table_begin(5, 16)
    bases, off = cases.flatten([b"GAGCTA", b"AGCTC", b"GAGCT"])
    km, has_n = engine.extract_kmers(bases, off.astype(np.uint32))
    assert km[0] == 0xD9 and km[1] == 0xD8            # chopRead2Kmer("GAGCTA") -> {0xD9, 0xD8}
    assert km[6] == 0xD9 and km[11] == 0xD9           # str2Kmer(AGCTC) == str2Kmer(GAGCT) == 0xD9
    assert (km[2:6] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    assert not has_n.any()


@pytest.mark.parametrize("k", K_SWEEP)
def test_extract_parity(engine, k):
    case = cases.adversarial_case(k, 700, seed=100 + k, with_adaptor=False)
    engine.table_begin(k, 16)
    bases, off = cases.flatten(case["reads"])
    km, has_n = engine.extract_kmers(bases, off.astype(np.uint32))
    for i, r in enumerate(case["reads"]):
        want = orc.chop(r, k)
        got = km[off[i]:off[i] + len(r)]
        assert (got[:len(want)] == want).all(), (k, i)
        assert (got[len(want):] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
        assert bool(has_n[i]) == (b"N" in r)


@pytest.mark.parametrize("k", K_SWEEP)
def test_table_and_lookup_parity(engine, k):
    case = cases.adversarial_case(k, 600, seed=200 + k, with_adaptor=(k <= 45))
    erased = build_table(engine, case)
    o, n_erased = build_oracle(case)
    info = engine.table_info()
    assert info.k == k
    assert (info.size[0], info.size[1]) == (o.set_size(0), o.set_size(1))
    assert sum(bin(t).count("1") for _, t in erased) == n_erased
    # every k-mer of every read (present and absent) + random probes
    qs = [orc.chop(r, k) for r in case["reads"] if len(r) >= k]
    rng = np.random.Generator(np.random.PCG64(k))
    mask = np.uint64((1 << (2 * k)) - 1) if k < 32 else np.uint64(0xFFFFFFFFFFFFFFFF)
    rnd = rng.integers(0, 1 << 62, 5000, dtype=np.uint64) & mask
    rnd = np.minimum(rnd, synth.revcomp_packed(rnd, k))
    q = np.concatenate(qs + [rnd])
    got = engine.lookup(q)
    want = o.lookup_many(q)
    assert (got == want).all()
    assert want.max() == 3 or k == 5                  # both-parent tags are exercised


def run_fused(engine, case, split=None):
    bases, off = cases.flatten(case["reads"])
    nb = len(case["bc_names"])
    engine.reset_counts()
    engine.reserve_barcodes(nb)
    n = len(case["reads"])
    cuts = [0, n] if not split else [0] + sorted(split) + [n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        if a == b:
            continue
        sub_off = (off[a:b + 1] - off[a]).astype(np.uint32)
        engine.submit_batch(bases[off[a]:off[b]], sub_off, case["bc_ids"][a:b])
    return engine.finish(nb), engine.stats()


@pytest.mark.parametrize("k", K_SWEEP)
def test_fused_parity(engine, k):
    case = cases.adversarial_case(k, 3000, seed=300 + k)
    build_table(engine, case)
    o, _ = build_oracle(case)
    bases, off = cases.flatten(case["reads"])
    want, lookups = o.classify_batch(bases, off, case["bc_ids"], len(case["bc_names"]))
    got, st = run_fused(engine, case)
    assert (got == want).all()
    assert want.sum() > 0
    assert st["lookups"] == lookups
    assert st["reads_with_n"] == sum(b"N" in r for r in case["reads"])
    # linearity: the same reads in three ragged batches
    got3, _ = run_fused(engine, case, split=[1, 1501])
    assert (got3 == want).all()


@pytest.mark.parametrize("k", [16, 21, 31])
def test_tile_size_does_not_change_counts(engine, k):
    """The tile size is chosen per batch so that a tile's reads fill one pass (fused_reads_per_tile); any fixed size --
    one read per tile's worth of threads, an odd size, the largest -- and the automatic choice give the oracle's counts."""
    case = cases.adversarial_case(k, 3000, seed=900 + k)
    build_table(engine, case)
    o, _ = build_oracle(case)
    bases, off = cases.flatten(case["reads"])
    want, lookups = o.classify_batch(bases, off, case["bc_ids"], len(case["bc_names"]))
    try:
        for rpt in (1, 33, 240, 416, 0):
            engine.set_option("reads_per_tile", rpt)
            got, st = run_fused(engine, case)
            assert (got == want).all() and st["lookups"] == lookups, rpt
    finally:
        engine.set_option("reads_per_tile", 0)


@pytest.mark.parametrize("k", [17, 21, 32])
def test_host_pack_option_equals_ascii(engine, k, request):
    """option host_pack_threads: hast_submit_batch packs the ASCII batch to 2 bits on the host and runs the packed kernel"""
    if "direct" in request.node.name:
        pytest.skip("the direct-probe kernel takes ASCII batches only")
    case = cases.adversarial_case(k, 3000, seed=950 + k)
    build_table(engine, case)
    o, _ = build_oracle(case)
    bases, off = cases.flatten(case["reads"])
    want, lookups = o.classify_batch(bases, off, case["bc_ids"], len(case["bc_names"]))
    try:
        for threads in (1, 4):
            engine.set_option("host_pack_threads", threads)
            got, st = run_fused(engine, case, split=[700, 1501])
            assert (got == want).all() and st["lookups"] == lookups, threads
            assert st["reads_with_n"] == sum(b"N" in r for r in case["reads"])
    finally:
        engine.set_option("host_pack_threads", 0)


@pytest.mark.parametrize("k", [5, 17, 21, 32])
def test_packed_batches_equal_ascii_batches(engine, k, request):
    """hast_submit_batch_packed (2-bit words + containN bits made on the host) == hast_submit_batch."""
    from hast_b200.capi import HastError, E_STATE
    case = cases.adversarial_case(k, 2500, seed=400 + k)
    build_table(engine, case)
    o, _ = build_oracle(case)
    bases, off = cases.flatten(case["reads"])
    nb = len(case["bc_names"])
    want, lookups = o.classify_batch(bases, off, case["bc_ids"], nb)
    engine.reset_counts()
    engine.reserve_barcodes(nb)
    n = len(case["reads"])
    if "direct" in request.node.name and "prefilter" not in request.node.name:
        with pytest.raises(HastError) as e:            # the direct-probe kernel has no packed form
            engine.submit_batch_packed(bases, off.astype(np.uint32), case["bc_ids"])
        assert e.value.code == E_STATE
        return
    cuts = [0, 1, 777, n]                              # ragged batches: packed streams restart per batch
    for a, b in zip(cuts[:-1], cuts[1:]):
        engine.submit_batch_packed(bases[off[a]:off[b]], (off[a:b + 1] - off[a]).astype(np.uint32), case["bc_ids"][a:b])
    got = engine.finish(nb)
    st = engine.stats()
    assert (got == want).all() and want.sum() > 0
    assert st["lookups"] == lookups
    assert st["reads_with_n"] == sum(b"N" in r for r in case["reads"])


def test_fused_empty_and_minimal(engine):
    case = cases.adversarial_case(21, 50, seed=9)
    build_table(engine, case)
    engine.reset_counts()
    engine.reserve_barcodes(4)
    engine.submit_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint32), np.zeros(0, np.uint32))   # empty batch
    assert (engine.finish(4) == 0).all()
    # one read of exactly k bases, taken from the paternal list
    kmer = case["pat_text"][:21]
    engine.submit_batch(np.frombuffer(kmer, np.uint8), np.array([0, 21], np.uint32), np.array([2], np.uint32))
    o, _ = build_oracle(case)
    want, _ = o.classify_batch(np.frombuffer(kmer, np.uint8), np.array([0, 21], np.uint64), np.array([2], np.uint32), 4)
    assert (engine.finish(4) == want).all()


def test_short_read_is_an_error(engine):
    from hast_b200.capi import HastError, E_SHORT_READ
    case = cases.adversarial_case(21, 50, seed=10)
    build_table(engine, case)
    engine.reset_counts()
    engine.reserve_barcodes(2)
    engine.submit_batch(np.frombuffer(b"ACGTACGT", np.uint8), np.array([0, 8], np.uint32), np.array([0], np.uint32))
    with pytest.raises(HastError) as e:                # reference: assert(rlen >= overlap), kmer.h:171
        engine.finish(2)
    assert e.value.code == E_SHORT_READ
    engine.reset_counts()
    # ... but a short read that contains an N is silently skipped (classify.cpp:190-193)
    engine.submit_batch(np.frombuffer(b"ACGNACGT", np.uint8), np.array([0, 8], np.uint32), np.array([0], np.uint32))
    assert (engine.finish(2) == 0).all()


def test_bad_kmer_line_is_an_error(engine):
    from hast_b200.capi import HastError, E_KMER_LINE
    engine.table_begin(5, 16)
    with pytest.raises(HastError) as e:                # kmer.h:154 assert(str.size()==overlap)
        engine.table_add_text(b"ACGTA\nACG\nTTACGTA\n", 5, 0)
    assert e.value.code == E_KMER_LINE


def test_long_reads_multi_pass(engine):
    """Reads long enough that a 256-read tile needs several shared-memory passes."""
    case = cases.adversarial_case(21, 700, seed=11, min_len=150, max_len=900)
    build_table(engine, case)
    o, _ = build_oracle(case)
    bases, off = cases.flatten(case["reads"])
    want, lookups = o.classify_batch(bases, off, case["bc_ids"], len(case["bc_names"]))
    got, st = run_fused(engine, case)
    assert (got == want).all() and st["lookups"] == lookups


def test_very_long_reads_between_short_ones(engine):
    """Reads of 5-24 kb (a shared-memory pass of their own, or most of one) mixed with ordinary ones, all cut from
    one random genome so that the k-mer lists sampled from the short reads also hit inside the long ones."""
    k = 21
    rng = np.random.Generator(np.random.PCG64(77))
    genome = cases.LET[rng.integers(0, 4, 60_000)].tobytes()
    reads = []
    for i in range(1500):
        L = int(rng.integers(5_000, 24_001)) if i % 97 == 5 else int(rng.integers(k, 300))
        s = int(rng.integers(0, len(genome) - L))
        r = genome[s:s + L]
        if rng.random() < 0.3:
            r = cases.revcomp_ascii(r)
        if i % 211 == 7:
            r = r[:L // 2] + b"N" + r[L // 2 + 1:]
        reads.append(r)
    kmers = []
    for _ in range(4000):
        r = reads[int(rng.integers(0, len(reads)))]
        if b"N" in r:
            continue
        p = int(rng.integers(0, len(r) - k + 1))
        kmers.append(r[p:p + k])
    half = len(kmers) // 2
    case = dict(k=k, pat_text=b"".join(x + b"\n" for x in kmers[:half + 200]),
                mat_text=b"".join(x + b"\n" for x in kmers[half - 200:]), reads=reads,
                bc_ids=rng.integers(0, 23, len(reads)).astype(np.uint32), bc_names=[b"%d" % i for i in range(23)])
    build_table(engine, case)
    o, _ = build_oracle(case)
    bases, off = cases.flatten(reads)
    want, lookups = o.classify_batch(bases, off, case["bc_ids"], 23)
    got, st = run_fused(engine, case, split=[500])
    assert (got == want).all() and want.sum() > 1000
    assert st["lookups"] == lookups and max(len(r) for r in reads) > 15_000


def test_table_under_pressure(engine):
    """Overflowing buckets: displaced entries must still be found; a hopeless capacity is an error."""
    case = cases.adversarial_case(21, 4000, seed=12)
    k = 21
    engine.table_begin(k, 16)                          # 16 buckets... for ~10k keys -> HAST_E_TABLE_FULL
    from hast_b200.capi import HastError, E_TABLE_FULL
    with pytest.raises(HastError) as e:
        engine.table_add_text(case["pat_text"], k, 0)
    assert e.value.code == E_TABLE_FULL
    build_table(engine, case)                          # default sizing: ~1.5 keys per 4-slot bucket
    info = engine.table_info()
    assert info.n_overflow_buckets > 0 and info.n_displaced > 0
    o, _ = build_oracle(case)
    assert (info.size[0], info.size[1]) == (o.set_size(0), o.set_size(1))
    bases, off = cases.flatten(case["reads"])
    want, _ = o.classify_batch(bases, off, case["bc_ids"], len(case["bc_names"]))
    got, st = run_fused(engine, case)
    assert (got == want).all()
    assert st["extra_probes"] > 0


@pytest.mark.parametrize("k", [17, 21, 25, 31, 32])
def test_low_complexity_parity(engine, k):
    """Homopolymers and short tandem repeats: thousands of distinct k-mers share one minimizer, so the
    minimizer-addressed pre-filter saturates a few words and leaves the decision to the exact table."""
    case = cases.adversarial_case(k, 3000, seed=900 + k, low_complexity=True)
    build_table(engine, case)
    o, _ = build_oracle(case)
    bases, off = cases.flatten(case["reads"])
    want, lookups = o.classify_batch(bases, off, case["bc_ids"], len(case["bc_names"]))
    got, st = run_fused(engine, case, split=[777])
    assert (got == want).all() and want.sum() > 0
    assert st["lookups"] == lookups


def test_minimizer_sweep_fetches_fewer_filter_words(request):
    """kernel 3 and kernel 1 give identical counts on the cfg1-scale trio; the minimizer sweep fetches a new
    pre-filter word only where the minimizer changes (about 2/(k-m+2) of the positions plus read starts),
    the per-k-mer sweep one word per position."""
    from hast_b200.capi import Engine
    t = synth.make_trio(synth.config("small"))
    bases, off, bc = t.batch()
    res = {}
    for kern in (3, 1):
        with Engine(0) as e:
            e.set_option("kernel", kern)
            e.table_begin(t.spec.k, t.pat.size + t.mat.size)
            e.table_add_packed(t.pat, 0)
            e.table_add_packed(t.mat, 1)
            e.reserve_barcodes(t.n_barcodes)
            e.submit_batch(bases, off, bc)
            res[kern] = (e.finish(t.n_barcodes), e.stats())
    assert (res[3][0] == res[1][0]).all() and res[3][0].sum() > 0
    s3, s1 = res[3][1], res[1][1]
    assert s1["lookups"] == s3["lookups"] and s1["filter_loads"] == s1["lookups"]
    assert 0 < s3["filter_loads"] < 0.45 * s3["lookups"]
    # the filter stays selective: members plus a few per cent of false positives
    assert s3["filter_pass"] < 0.15 * s3["lookups"]


def test_saturated_filter_queue_drains(engine):
    """A pre-filter far too small for the key set passes (almost) every position, so the
    shared-memory queue fills and drains every sweep; the exact table still decides."""
    case = cases.adversarial_case(21, 3000, seed=13, min_len=80, max_len=400)
    o, _ = build_oracle(case)
    bases, off = cases.flatten(case["reads"])
    want, lookups = o.classify_batch(bases, off, case["bc_ids"], len(case["bc_names"]))
    try:
        engine.set_option("filter_max_bytes", 128)
        build_table(engine, case)
        assert engine.table_info().filter_bytes == 128
        got, st = run_fused(engine, case)
    finally:
        engine.set_option("filter_max_bytes", 64 << 20)
    assert (got == want).all() and st["lookups"] == lookups
    if st["filter_pass"]:                              # the pre-filtered kernel
        assert st["filter_pass"] > 0.9 * lookups


@pytest.fixture(scope="module")
def trio_small():
    return synth.make_trio(synth.config("small"))


def test_synthetic_trio_parity(engine, trio_small):
    t = trio_small
    k = t.spec.k
    engine.table_begin(k, t.pat.size + t.mat.size)
    engine.table_add_text(t.kmer_text(0), k, 0)
    engine.table_add_text(t.kmer_text(1), k, 1)
    engine.table_erase_seq(cases.ADAPTOR_F)
    engine.table_erase_seq(cases.ADAPTOR_R)
    o = orc.Oracle()
    o.load_kmers(t.kmer_text(0), 0)
    o.load_kmers(t.kmer_text(1), 1)
    o.init_adaptor()
    bases, off, bc = t.batch()
    want, lookups = o.classify_batch(bases, off.astype(np.uint64), bc, t.n_barcodes)
    engine.reset_counts()
    engine.reserve_barcodes(t.n_barcodes)
    engine.submit_batch(bases, off, bc)
    got = engine.finish(t.n_barcodes)
    assert (got == want).all()
    assert engine.stats()["lookups"] == lookups
    # packed insertion builds the same table as text insertion
    engine.table_begin(k, t.pat.size + t.mat.size)
    engine.table_add_packed(t.pat, 0)
    engine.table_add_packed(synth.revcomp_packed(t.mat, k), 1)      # non-canonical orientation on purpose
    engine.reset_counts()
    engine.reserve_barcodes(t.n_barcodes)
    engine.submit_batch(bases, off, bc)
    assert (engine.finish(t.n_barcodes) == want).all()
    # the calls make biological sense: most barcodes land on their true haplotype
    c = got[:t.spec.n_barcodes].astype(np.int64)
    call = np.where(c[:, 0] * t.mat.size > c[:, 1] * t.pat.size, 0, 1)
    decided = (c.sum(1) > 0)
    assert (call[decided] == t.bc_hap[decided]).mean() > 0.95


def test_fused_equals_k2_then_k3(engine, trio_small):
    """GPU-vs-GPU: the fused kernel agrees with standalone extract + lookup + host sum."""
    t = trio_small
    k = t.spec.k
    engine.table_begin(k, t.pat.size + t.mat.size)
    engine.table_add_packed(t.pat, 0)
    engine.table_add_packed(t.mat, 1)
    bases, off, bc = t.batch(0, 5000)
    km, has_n = engine.extract_kmers(bases, off)
    valid = km != np.uint64(0xFFFFFFFFFFFFFFFF)
    tags = np.zeros(km.size, np.uint8)
    tags[valid] = engine.lookup(km[valid])
    L = t.spec.read_len
    per_read = tags.reshape(-1, L)
    v0 = (per_read & 1).sum(1) * (has_n == 0)
    v1 = (per_read >> 1).sum(1) * (has_n == 0)
    want = np.zeros((t.n_barcodes, 2), np.int64)
    np.add.at(want[:, 0], bc, v0)
    np.add.at(want[:, 1], bc, v1)
    engine.reset_counts()
    engine.reserve_barcodes(t.n_barcodes)
    engine.submit_batch(bases, off, bc)
    assert (engine.finish(t.n_barcodes) == want).all()


def test_permutation_invariance_cfg1_scale(engine):
    """configs[0] size: oracle parity on the full input + order independence."""
    spec = synth.config("cfg1")
    t = synth.make_trio(spec)
    k = spec.k
    engine.table_begin(k, t.pat.size + t.mat.size)
    engine.table_add_packed(t.pat, 0)
    engine.table_add_packed(t.mat, 1)
    o = orc.Oracle()
    o.load_kmers(t.kmer_text(0), 0)
    o.load_kmers(t.kmer_text(1), 1)
    bases, off, bc = t.batch()
    want, lookups = o.classify_batch(bases, off.astype(np.uint64), bc, t.n_barcodes, nthreads=8)
    engine.reset_counts()
    engine.reserve_barcodes(t.n_barcodes)
    engine.submit_batch(bases, off, bc)
    got = engine.finish(t.n_barcodes)
    assert (got == want).all() and engine.stats()["lookups"] == lookups
    perm = np.random.Generator(np.random.PCG64(5)).permutation(bc.size)
    L = spec.read_len
    engine.reset_counts()
    engine.reserve_barcodes(t.n_barcodes)
    engine.submit_batch(bases.reshape(-1, L)[perm], off, bc[perm])
    assert (engine.finish(t.n_barcodes) == want).all()


def test_skewed_barcodes_parity(engine):
    """configs[4] shape at a size the oracle finishes in seconds: heavy-tailed (Zipf 1.2) reads per barcode,
    so a few barcodes take most of the per-barcode atomics while most barcodes see one or two reads."""
    spec = synth.TrioSpec(genome_len=400_000, het=0.002, n_pairs=60_000, n_barcodes=40_000, zipf_alpha=1.2)
    t = synth.make_trio(spec)
    k = spec.k
    top = np.bincount(t.pair_bc, minlength=t.n_barcodes).max()
    assert top > 0.02 * spec.n_pairs                   # the tail really is heavy
    engine.table_begin(k, t.pat.size + t.mat.size)
    engine.table_add_packed(t.pat, 0)
    engine.table_add_packed(t.mat, 1)
    o = orc.Oracle()
    o.load_kmers(t.kmer_text(0), 0)
    o.load_kmers(t.kmer_text(1), 1)
    bases, off, bc = t.batch()
    want, lookups = o.classify_batch(bases, off.astype(np.uint64), bc, t.n_barcodes, nthreads=8)
    engine.reset_counts()
    engine.reserve_barcodes(t.n_barcodes)
    # reads grouped by barcode: the worst case for atomic contention on the hot barcodes
    order = np.argsort(bc, kind="stable")
    L = spec.read_len
    engine.submit_batch(bases.reshape(-1, L)[order], off, bc[order])
    got = engine.finish(t.n_barcodes)
    assert (got == want).all() and engine.stats()["lookups"] == lookups
