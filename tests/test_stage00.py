"""HAST stage 00 (00.build_unshare_kmers_by_jellyfish/build_unshared_kmers.sh): the parent-unique
k-mer lists from parental reads.

CPU: oracle/stage00_unshared.py against the committed fixtures tests/golden/stage00_* (produced by the
reference's own script + its vendored jellyfish binary, tests/golden/make_golden_stage00.py) and, when
/root/reference is present, against a live run of that script.
GPU: the count table (hast_kc_*) through the C ABI against the oracle and the fixtures, bit-exact:
histograms, bounds, both lists; partitions; direct table construction; bin/build_unshared_kmers.
"""
import gzip
import os
import subprocess
import sys
from collections import Counter
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import stage00_unshared as s0  # noqa: E402

from hast_b200 import synth  # noqa: E402

GOLDEN = sorted((ROOT / "tests" / "golden").glob("stage00_*"))
REF_SCRIPT = Path("/root/reference/00.build_unshare_kmers_by_jellyfish/build_unshared_kmers.sh")
BIN = ROOT / "bin" / "build_unshared_kmers"


def parse_cmd(d: Path):
    toks = (d / "cmd.txt").read_text().split()
    a = {"paternal": [], "maternal": [], "k": 21, "pl": 9, "pu": 33, "ml": 9, "mu": 33, "auto": False}
    i = 0
    while i < len(toks):
        t = toks[i]
        if t == "--auto_bounds":
            a["auto"] = True
            i += 1
            continue
        v = toks[i + 1]
        if t == "--paternal":
            a["paternal"].append(str(d / v))
        elif t == "--maternal":
            a["maternal"].append(str(d / v))
        elif t == "--mer":
            a["k"] = int(v)
        else:
            a[{"--p-lower": "pl", "--p-upper": "pu", "--m-lower": "ml", "--m-upper": "mu"}[t]] = int(v)
        i += 2
    return a


def expected_files(d: Path) -> dict:
    return {p.name: p.read_bytes() for p in (d / "expected").iterdir()}


# ---------------------------------------------------------------- CPU: the oracle is pinned
def test_fixtures_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("d", GOLDEN, ids=lambda p: p.name)
def test_oracle_matches_reference_fixtures(d):
    a = parse_cmd(d)
    got = s0.run(a["paternal"], a["maternal"], a["k"], a["pl"], a["pu"], a["ml"], a["mu"], a["auto"])
    want = expected_files(d)
    assert want
    for name, data in want.items():
        assert got[name] == data, name


def test_find_bounds_restates_the_awk_program():
    # first bin that is not smaller than its predecessor only flips the state (find_bounds.awk:9-17)
    h = [(1, 900), (2, 300), (3, 80), (4, 120), (5, 400), (6, 700), (7, 650), (8, 100), (10001, 5)]
    b = s0.find_bounds(h)
    assert b == {"MIN_INDEX": 3, "MAX_INDEX": 6, "LOWER_INDEX": 4, "UPPER_INDEX": 3 * 6 - 2 * 3 - 1}
    if REF_SCRIPT.exists():
        awk = REF_SCRIPT.parent / "find_bounds.awk"
        r = subprocess.run(["awk", "-f", str(awk)], input=s0.histo_text(h), capture_output=True)
        assert r.stdout == s0.bounds_text(b)


def test_counting_rule_known_answers():
    # hand-checked against `jellyfish count -m 5 -C` (multi-line FASTA joined, lower case accepted,
    # N breaks the window, blank line skipped)
    c = s0.count_canonical([b"ACGTACGTNNACGTTTGAccgtagGATTACA", b"TTTTTTTTTTAAAAAAAAAA"], 5)
    assert c[b"AAAAA"] == 12 and c[b"TAAAA"] == 2 and c[b"TTAAA"] == 2 and c[b"ACGTA"] == 2 and c[b"GGTCA"] == 1
    assert sum(c.values()) == (8 - 4) + (21 - 4) + 16 and len(c) == 22


@pytest.mark.skipif(not REF_SCRIPT.exists(), reason="reference not present")
def test_oracle_matches_live_reference_script(tmp_path):
    spec = synth.TrioSpec(genome_len=1500, het=0.012, k=21, seed=9)
    r = synth.parent_reads(spec, 18.0, n_frac=0.03, lowercase_frac=0.05)
    pat = synth.write_reads_fastq(tmp_path / "p.fq.gz", r["paternal"], gz=True)
    mat = synth.write_reads_fastq(tmp_path / "m.fq.gz", r["maternal"], gz=True)
    run = tmp_path / "run"
    run.mkdir()
    res = subprocess.run(["bash", str(REF_SCRIPT), "--paternal", pat, "--maternal", mat, "--mer", "21", "--auto_bounds",
                          "--thread", "2", "--memory", "1"], cwd=run, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-1500:]
    got = s0.run([pat], [mat], 21, auto_bounds=True)
    for f in ("paternal.histo", "maternal.histo", "paternal.bounds.txt", "maternal.bounds.txt"):
        assert (run / f).read_bytes() == got[f], f
    for f in ("paternal.unique.filter.mer", "maternal.unique.filter.mer"):
        assert sorted((run / f).read_bytes().splitlines()) == got[f].splitlines(), f


# ---------------------------------------------------------------- GPU: the count table
def _batch(seqs, k, chunk=8192):
    """Sequences -> (bases, seq_off) with long sequences cut into chunks overlapping by k-1."""
    parts, off = [], [0]
    for s in seqs:
        i = 0
        while True:
            piece = s[i:i + chunk]
            parts.append(piece)
            off.append(off[-1] + len(piece))
            if i + chunk >= len(s):
                break
            i += chunk - (k - 1)
    bases = np.frombuffer(b"".join(parts), np.uint8)
    return bases, np.asarray(off, np.uint32)


def _text(kmers, k):
    from hast_b200.capi import kmers_to_text
    return kmers_to_text(kmers, k)


@pytest.fixture(scope="module")
def kc_engine():
    from hast_b200.capi import Engine
    e = Engine(0)
    yield e
    e.close()


def _oracle_histo_array(counts: Counter, high: int) -> np.ndarray:
    h = np.zeros(high + 2, np.uint64)
    for c, n in s0.histo(counts, high):
        h[c] = n
    h[0] = len(counts)
    return h


@pytest.mark.gpu
@pytest.mark.parametrize("k", [11, 16, 17, 21, 25, 31, 32])
def test_count_table_matches_oracle(kc_engine, k):
    e = kc_engine
    spec = synth.TrioSpec(genome_len=4000, het=0.01, k=k, seed=20 + k)
    r = synth.parent_reads(spec, 12.0, n_frac=0.05, lowercase_frac=0.05)
    seqs = [[x.tobytes() for x in r["paternal"]], [x.tobytes() for x in r["maternal"]]]
    # ragged input: one long multi-chunk sequence, sequences shorter than k, an empty one, odd bytes
    seqs[0] += [b"".join(seqs[0][:300]), b"ACGT", b"", b"ACGTRYACGT" * 9, seqs[0][0][:k], seqs[0][1][:k - 1]]
    seqs[1] += [b"NN".join(seqs[1][:50]), b"acgtn" * 40]
    cnt = [s0.count_canonical(seqs[p], k) for p in (0, 1)]
    e.kc_begin(k, sum(len(c) for c in cnt))
    for p in (0, 1):
        half = len(seqs[p]) // 2
        for part in (seqs[p][:half], seqs[p][half:]):          # two batches per parent
            bases, off = _batch(part, k)
            e.kc_add(bases, off, p)
    ki = e.kc_info()
    assert ki.table_full == 0
    assert (ki.distinct[0], ki.distinct[1]) == (len(cnt[0]), len(cnt[1]))
    assert (ki.occurrences[0], ki.occurrences[1]) == (sum(cnt[0].values()), sum(cnt[1].values()))
    assert ki.occupied == len(set(cnt[0]) | set(cnt[1])) and ki.both == len(set(cnt[0]) & set(cnt[1]))
    assert ki.windows == ki.occurrences[0] + ki.occurrences[1]
    for p in (0, 1):
        assert (e.kc_histo(p, 10000) == _oracle_histo_array(cnt[p], 10000)).all()
        assert (e.kc_histo(p, 6) == _oracle_histo_array(cnt[p], 6)).all()          # overflow bin
    pat, mat = s0.unshared(cnt[0], cnt[1], 3, 14, 2, 20)
    assert _text(e.kc_select(0, 3, 14), k) == b"".join(x + b"\n" for x in pat)
    assert _text(e.kc_select(1, 2, 20), k) == b"".join(x + b"\n" for x in mat)
    # `jellyfish dump -L -U` alone (no uniqueness)
    want = sorted(x for x, c in cnt[0].items() if 4 <= c <= 9)
    assert _text(e.kc_select(0, 4, 9, require_unique=False), k) == b"".join(x + b"\n" for x in want)
    e.kc_end()


@pytest.mark.gpu
@pytest.mark.parametrize("d", GOLDEN, ids=lambda p: p.name)
def test_count_table_matches_reference_fixtures(kc_engine, d):
    e = kc_engine
    a = parse_cmd(d)
    k = a["k"]
    e.kc_begin(k, 400_000)
    for p, files in ((0, a["paternal"]), (1, a["maternal"])):
        for f in files:
            bases, off = _batch(s0.sequences(f), k, chunk=3000)
            e.kc_add(bases, off, p)
    want = expected_files(d)
    pl, pu, ml, mu = a["pl"], a["pu"], a["ml"], a["mu"]
    if a["auto"]:
        bounds = {}
        for p, name in ((0, "paternal"), (1, "maternal")):
            h = e.kc_histo(p, 10000)
            hl = [(c, int(h[c])) for c in range(1, h.size) if h[c]]
            assert s0.histo_text(hl) == want[f"{name}.histo"]
            bounds[name] = s0.find_bounds(hl)
            assert s0.bounds_text(bounds[name]) == want[f"{name}.bounds.txt"]
        pl, pu = bounds["paternal"]["LOWER_INDEX"], bounds["paternal"]["UPPER_INDEX"]
        ml, mu = bounds["maternal"]["LOWER_INDEX"], bounds["maternal"]["UPPER_INDEX"]
    assert _text(e.kc_select(0, pl, pu), k) == want["paternal.unique.filter.mer"]
    assert _text(e.kc_select(1, ml, mu), k) == want["maternal.unique.filter.mer"]
    e.kc_end()


@pytest.mark.gpu
def test_partitions_union_equals_whole(kc_engine):
    """Streaming the reads once per partition (or to one GPU per partition) gives the same lists."""
    e = kc_engine
    k = 21
    spec = synth.TrioSpec(genome_len=20000, het=0.006, k=k, seed=77)
    r = synth.parent_reads(spec, 15.0)
    batches = [_batch([x.tobytes() for x in r[n]], k) for n in ("paternal", "maternal")]

    def run(part, n_parts, expected):
        e.kc_begin(k, expected, part, n_parts)
        for p in (0, 1):
            e.kc_add(batches[p][0], batches[p][1], p)
        out = (e.kc_select(0, 4, 40), e.kc_select(1, 4, 40), e.kc_histo(0, 100), e.kc_info())
        e.kc_end()
        return out

    whole = run(0, 1, 300_000)
    assert whole[0].size > 100 and whole[1].size > 100
    parts = [run(i, 3, 120_000) for i in range(3)]
    for j in (0, 1):
        merged = np.sort(np.concatenate([p[j] for p in parts]))
        assert (merged == whole[j]).all()
    assert (sum(p[2] for p in parts) == whole[2]).all()
    assert sum(p[3].occupied for p in parts) == whole[3].occupied
    assert all(p[3].windows == whole[3].windows for p in parts)
    assert all(0.2 < p[3].occupied / whole[3].occupied < 0.5 for p in parts)       # the split is even


@pytest.mark.gpu
def test_table_full_is_reported(kc_engine):
    e = kc_engine
    k = 21
    spec = synth.TrioSpec(genome_len=20000, het=0.006, k=k, seed=78)
    r = synth.parent_reads(spec, 4.0)
    bases, off = _batch([x.tobytes() for x in r["paternal"]], k)
    from hast_b200.capi import HastError, E_TABLE_FULL
    e.kc_begin(k, 256)                                  # 1024 slots for ~25 000 distinct k-mers
    e.kc_add(bases, off, 0)
    with pytest.raises(HastError) as ei:
        e.kc_select(0, 1, 100)
    assert ei.value.code == E_TABLE_FULL
    e.kc_end()


@pytest.mark.gpu
def test_counts_feed_the_classifier_without_text_round_trip(kc_engine):
    """hast_kc_to_table == writing the two lists as text and loading them with hast_table_add_text."""
    from hast_b200.capi import Engine, kmers_to_text
    import cases
    e = kc_engine
    k = 21
    spec = synth.TrioSpec(genome_len=30000, het=0.005, k=k, seed=5, n_pairs=3000, n_barcodes=200)
    r = synth.parent_reads(spec, 16.0)
    e.kc_begin(k, 400_000)
    for p, n in ((0, "paternal"), (1, "maternal")):
        bases, off = _batch([x.tobytes() for x in r[n]], k)
        e.kc_add(bases, off, p)
    pat, mat = e.kc_select(0, 4, 40), e.kc_select(1, 4, 40)
    trio = synth.make_trio(spec)
    bases, off, bc = trio.batch()

    def classify(eng):
        eng.table_erase_seq(cases.ADAPTOR_F)
        eng.table_erase_seq(cases.ADAPTOR_R)
        eng.reset_counts()
        eng.reserve_barcodes(trio.n_barcodes)
        eng.submit_batch(bases, off, bc)
        return eng.finish(trio.n_barcodes), eng.table_info()

    e.kc_to_table(4, 40, 4, 40)
    got, gi = classify(e)
    with Engine(0) as e2:
        e2.table_begin(k, pat.size + mat.size)
        e2.table_add_text(kmers_to_text(pat, k), k, 0)
        e2.table_add_text(kmers_to_text(mat, k), k, 1)
        want, wi = classify(e2)
    assert (gi.size[0], gi.size[1]) == (wi.size[0], wi.size[1]) and gi.size[0] > 100
    assert (got == want).all() and got.sum() > 0
    # the child's reads mostly hit the lists of the haplotypes it inherited
    e.kc_end()


# ---------------------------------------------------------------- GPU: the drop-in process
@pytest.mark.gpu
@pytest.mark.parametrize("d", GOLDEN, ids=lambda p: p.name)
def test_cli_matches_reference_fixtures(d, tmp_path):
    assert BIN.exists(), "bin/build_unshared_kmers not built"
    args = []
    for t in (d / "cmd.txt").read_text().split():
        args.append(str(d / t) if (d / t).exists() else t)
    r = subprocess.run([str(BIN)] + args + ["--thread", "3"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-800:]
    for name, data in expected_files(d).items():
        got = (tmp_path / name).read_bytes()
        if name.endswith(".mer"):
            got = b"".join(x + b"\n" for x in sorted(got.splitlines()))
        assert got == data, name
    # the lists are what stage 01 loads next: same k, one k-mer per line
    k = parse_cmd(d)["k"]
    first = (tmp_path / "paternal.unique.filter.mer").read_bytes().split(b"\n")[0]
    assert len(first) == k


@pytest.mark.gpu
def test_cli_retry_with_more_partitions_counts_every_partition(tmp_path):
    """a table-full retry that raises the partition count must sweep ALL the new partitions.  A tiny table
    cap (HAST_KC_TABLE_MB) and an undersized --expected-distinct force the retry; the lists must equal the fixtures."""
    d = GOLDEN[0]
    args = []
    for t in (d / "cmd.txt").read_text().split():
        args.append(str(d / t) if (d / t).exists() else t)
    r = subprocess.run([str(BIN)] + args + ["--thread", "2", "--expected-distinct", "300"], cwd=tmp_path,
                       capture_output=True, text=True, env=dict(os.environ, HAST_KC_TABLE_MB="0"))
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-800:]
    assert "count table too small, retrying" in r.stdout and "partition(s)" in r.stdout
    n_parts = int(r.stdout.split("distinct k-mers in ")[-1].split(" partition")[0])
    assert n_parts > 1, r.stdout[-400:]
    for name, data in expected_files(d).items():
        got = (tmp_path / name).read_bytes()
        if name.endswith(".mer"):
            got = b"".join(x + b"\n" for x in sorted(got.splitlines()))
        assert got == data, name


def test_cli_usage_and_errors(tmp_path):
    if not BIN.exists():
        pytest.skip("bin/build_unshared_kmers not built")
    r = subprocess.run([str(BIN)], capture_output=True, text=True)
    assert r.returncode == 0 and "Usage" in r.stdout                         # script :56-59
    r = subprocess.run([str(BIN), "--paternal", "nope.fq", "--maternal", "nope2.fq"], capture_output=True, text=True,
                       cwd=tmp_path)
    assert r.returncode == 1 and "is not exist" in r.stdout                  # script :153-158
    r = subprocess.run([str(BIN), "--bogus"], capture_output=True, text=True)
    assert "invalid params" in r.stdout                                      # script :118-121
    r = subprocess.run([str(BIN), "--paternal", str(GOLDEN[0] / "pat.fq"), "--maternal", str(GOLDEN[0] / "mat.fq"),
                        "--mer", "9"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "arguments invalid" in r.stdout             # script :141-152
