"""CPU: pin the oracle (oracle/hast_oracle.c) to the reference.

1. every known answer of the reference's own TestAll() (classify.cpp:341-367);
2. the committed golden fixtures = stdout of the untouched reference binary
   (tests/golden/make_golden.py);
3. when oracle/_ref exists (this container), live runs of that binary on fresh
   seeded inputs, and a check that the fixtures are not stale.
"""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

import cases
import oracle as orc

GOLDEN = Path(__file__).resolve().parent / "golden"
CLASSIFY_CASES = sorted(p.name for p in GOLDEN.iterdir() if (p / "pat.mer").exists())


def test_golden_fixtures_present():
    assert len(CLASSIFY_CASES) >= 12 and (GOLDEN / "merge" / "expected.tsv").exists()


# ---- 1. TestAll (classify.cpp:341-367) -------------------------------------------
def test_reference_testall_known_answers():
    l = orc.lib()
    assert orc.parse_name(b"VSDSDS#XXX_xxx_s/1") == b"XXX_xxx_s"                     # :342
    assert [l.ho_base2int(c) for c in b"AGCTC"] == [0, 3, 1, 2, 1]                    # :344-346
    assert [l.ho_base2int(c) for c in b"GAGCT"] == [3, 0, 3, 1, 2]                    # :347-349
    assert l.ho_str2kmer(b"AGCTC", 5) == 0xD9                                         # :351-352 (high == 0)
    assert l.ho_str2kmer(b"GAGCT", 5) == 0xD9                                         # :353-354
    km = orc.chop(b"GAGCTA", 5)
    assert list(km) == [0xD9, 0xD8]                                                   # :355-362
    assert orc.kmer2str(km[0], 5) == b"AGCTC" and orc.kmer2str(km[1], 5) == b"AGCTA"  # :363-366


def test_base_code_is_total_and_case_insensitive():
    l = orc.lib()
    for c in range(256):
        assert l.ho_base2int(c) == (c & 6) >> 1                                      # kmer.h:11
    assert [l.ho_base2int(c) for c in b"acgtNn\r"] == [0, 1, 3, 2, 3, 3, 2]


@pytest.mark.parametrize("k", [1, 5, 16, 21, 31, 32])
def test_revcomp_involution_and_string_model(k):
    rng = np.random.Generator(np.random.PCG64(k))
    l = orc.lib()
    for _ in range(200):
        s = cases.LET[rng.integers(0, 4, k)].tobytes()
        w = 0
        for c in s:
            w = (w << 2) | ((c & 6) >> 1)
        rc = l.ho_revcomp(w, k)
        assert l.ho_revcomp(rc, k) == w
        w2 = 0
        for c in cases.revcomp_ascii(s):
            w2 = (w2 << 2) | ((c & 6) >> 1)
        assert rc == w2
        assert l.ho_str2kmer(s, k) == min(w, rc)


def test_parse_name_edge_cases():
    assert orc.parse_name(b"@V300#203_1533_1069/1") == b"203_1533_1069"
    assert orc.parse_name(b"@a#b#c/d/e") == b"c/d"            # LAST '#', LAST '/'
    assert orc.parse_name(b"@noseparators") == b"@noseparators"  # s=-1,e=-1 -> substr(0,-1)
    assert orc.parse_name(b"@x/1#bc") == b"bc"                # '/' before '#': to end of string
    assert orc.parse_name(b"@x#/1") == b""
    assert orc.parse_name(b"@x#bc") == b"bc"
    assert orc.parse_name(b"") == b""


def test_get_hap_ladder():
    g = orc.lib().ho_get_hap
    assert g(b"0_0_0", 1, 9, 0, 0, 10, 10, 1.0, 1.0) == -1                           # classify.cpp:67-68
    assert g(b"0_0", 1, 9, 0, 0, 10, 10, 1.0, 1.0) == -1 and g(b"0", 1, 9, 0, 0, 10, 10, 1.0, 1.0) == -1
    assert g(b"1_2_3", 1, 5, 1, 5, 10, 10, 1.0, 1.0) == -1                           # tie
    assert g(b"1_2_3", 1, 5, 1, 5, 10, 10, 1.04, 1.0) == 0                           # weight breaks it
    assert g(b"1_2_3", 1, 5, 1, 5, 10, 11, 1.0, 1.0) == 0                            # smaller set -> larger ratio
    assert g(b"1_2_3", 1, 5, 1, 6, 10, 10, 1.0, 1.0) == 1
    assert g(b"1_2_3", 1, 5, 0, 0, 10, 10, 1.0, 1.0) == 0
    assert g(b"1_2_3", 0, 0, 1, 2, 10, 10, 1.0, 1.0) == 1
    assert g(b"1_2_3", 0, 0, 0, 0, 10, 10, 1.0, 1.0) == -1


# ---- 2. golden fixtures ---------------------------------------------------------------
def run_oracle_on_case(d: Path, tmp_path) -> bytes:
    args = json.loads((d / "cmd.txt").read_text())
    opt = {"--weight0": "1.0", "--weight1": "1.0", "--adaptor_f": orc.ADAPTOR_F.decode(),
           "--adaptor_r": orc.ADAPTOR_R.decode()}
    reads = []
    for a, b in zip(args[::2], args[1::2]):
        if a == "--read":
            reads.append(b)
        else:
            opt[a] = b
    o = orc.Oracle()
    assert o.load_kmers_file(d / opt["--hap0"], 0) >= 0, o.err()
    assert o.load_kmers_file(d / opt["--hap1"], 1) >= 0, o.err()
    assert o.init_adaptor(opt["--adaptor_f"].encode(), opt["--adaptor_r"].encode()) >= 0
    o.set_weights(float(opt["--weight0"]), float(opt["--weight1"]))
    for r in reads:
        assert o.process_fastq(d / r) == 0, o.err()
    return o.table(tmp_path)


@pytest.mark.parametrize("name", CLASSIFY_CASES)
def test_oracle_matches_golden(name, tmp_path):
    d = GOLDEN / name
    assert run_oracle_on_case(d, tmp_path) == (d / "expected.tsv").read_bytes()


def test_oracle_merge_matches_golden(tmp_path):
    d = GOLDEN / "merge"
    args = json.loads((d / "cmd.txt").read_text())
    inputs = [d / b for a, b in zip(args[::2], args[1::2]) if a == "--input"]
    out = tmp_path / "m.tsv"
    assert orc.merge_result(inputs, out, w0=2.5) == 0
    got = out.read_bytes()
    assert got == (d / "expected.tsv").read_bytes()
    assert b"1_2_3\t0\t8\t0\n" in got                       # "1_2_3 1 1 3" twice -> "1_2_3 0 8 0"


# ---- 3. live reference ------------------------------------------------------------------
needs_ref = pytest.mark.skipif(orc.ref_binary("classify") is None,
                               reason="oracle/_ref not built (no /root/reference on this box)")


@needs_ref
@pytest.mark.parametrize("name", CLASSIFY_CASES)
def test_fixture_not_stale(name):
    d = GOLDEN / name
    args = json.loads((d / "cmd.txt").read_text())
    r = subprocess.run([str(orc.ref_binary("classify_O2"))] + args, cwd=d, capture_output=True)
    assert r.returncode == 0 and r.stdout == (d / "expected.tsv").read_bytes()


@needs_ref
@pytest.mark.parametrize("k,seed", [(7, 51), (19, 52), (21, 53), (27, 54), (32, 55)])
def test_oracle_matches_live_reference(k, seed, tmp_path):
    c = cases.adversarial_case(k, 400, seed)
    (tmp_path / "p.mer").write_bytes(c["pat_text"])
    (tmp_path / "m.mer").write_bytes(c["mat_text"])
    cases.write_fastq(tmp_path / "a.fq", c["heads"][:250], c["reads"][:250])
    cases.write_fastq(tmp_path / "b.fq.gz", c["heads"][250:], c["reads"][250:])
    ref = orc.run_ref_classify(tmp_path / "p.mer", tmp_path / "m.mer", [tmp_path / "a.fq", tmp_path / "b.fq.gz"],
                               extra=["--weight0", "1.04"], binary="classify")
    o = orc.Oracle()
    o.load_kmers_file(tmp_path / "p.mer", 0)
    o.load_kmers_file(tmp_path / "m.mer", 1)
    o.init_adaptor()
    o.set_weights(1.04, 1.0)
    o.process_fastq(tmp_path / "a.fq")
    o.process_fastq(tmp_path / "b.fq.gz")
    assert o.table(tmp_path) == ref
    # the dense batch form used to check the device interface gives the same counts
    bases, off = cases.flatten(c["reads"])
    counts, _ = o.classify_batch(bases, off, c["bc_ids"], len(c["bc_names"]))
    table = {ln.split(b"\t")[0]: ln.split(b"\t") for ln in ref.splitlines()}
    for i, nm in enumerate(c["bc_names"]):
        if nm in table:
            assert (int(table[nm][2]), int(table[nm][3])) == tuple(counts[i]), nm
        else:
            assert not (c["bc_ids"] == i).any()


def test_packed_loader_equals_text_loader():
    """ho_load_kmers_packed (harness convenience for the large configurations) builds the same sets as load_kmers."""
    from hast_b200 import synth
    t = synth.make_trio(synth.config("tiny"))
    a, b = orc.Oracle(), orc.Oracle()
    for i in (0, 1):
        a.load_kmers(t.kmer_text(i), i)
        km = t.pat if i == 0 else t.mat
        flip = np.arange(km.size) % 3 == 0                      # any orientation is canonicalised on load
        b.load_kmers_packed(np.where(flip, synth.revcomp_packed(km, t.spec.k), km), t.spec.k, i)
    a.init_adaptor(); b.init_adaptor()
    assert (a.set_size(0), a.set_size(1)) == (b.set_size(0), b.set_size(1))
    bases, off, bc = t.batch()
    ca, la = a.classify_batch(bases, off.astype(np.uint64), bc, t.n_barcodes)
    cb, lb = b.classify_batch(bases, off.astype(np.uint64), bc, t.n_barcodes)
    assert la == lb and (ca == cb).all() and ca.sum() > 0
