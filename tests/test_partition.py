"""CPU: the stage-01 post-processing -- bin/quartering_fastq and the
barcode-list split -- against the reference's own awk programs (run live when /root/reference
is present) and against the Python restatement oracle/stage01_post.py.  Byte-identical files."""
import gzip
import os
import shutil
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import stage01_post as post  # noqa: E402

QUART = ROOT / "bin" / "quartering_fastq"
REF_AWK = Path("/root/reference/01.classify_stlfr_reads/quartering_fastq.awk")
AWK = shutil.which("awk")
SUFFIXES = ["nobarcode", "paternal", "maternal", "homozygous"]


def make_case(seed, n=400, crlf=False, trailing_newline=True, partial_tail=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    bcs = [b"%d_%d_%d" % tuple(rng.integers(1, 1537, 3)) for _ in range(40)]
    odd = [b"lib2_7_8_9", b"a b", b"x#y", b"p/q", b"", b"0_0_0", b"0_0", b"0", b"7_7_7\r"]
    pat, mat, hom = bcs[:12] + [b"lib2_7_8_9", b"x"], bcs[12:24] + [b"p", b"dup_1"], bcs[24:36] + [b"0_0_0", b"dup_1", b""]
    recs = []
    for i in range(n):
        r = rng.random()
        bc = bcs[rng.integers(0, 40)] if r < 0.8 else odd[rng.integers(0, len(odd))]
        style = rng.integers(0, 10)
        if style == 0:
            head = b"@noseparators%d" % i                       # NF == 1 -> no barcode
        elif style == 1:
            head = b"@r%d#" % i + bc                            # no '/'
        elif style == 2:
            head = b"@dir/r%d#" % i + bc + b"/1"                # '/' before '#': $2 is not the barcode
        else:
            head = b"@V300R%010d#" % i + bc + b"/%d" % (1 + i % 2)
        L = int(rng.integers(1, 120))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), L))
        recs.append((head, seq, b"+", b"F" * L))
    eol = b"\r\n" if crlf else b"\n"
    text = b"".join(eol.join(r) + eol for r in recs)
    if partial_tail:
        text += b"@tail#" + bcs[0] + b"/1" + eol + b"ACGT"      # record cut short, no final newline
    elif not trailing_newline:
        text = text[:-len(eol)]
    lists = tuple(b"".join(x + b"\n" for x in l) for l in (pat, mat, hom))
    return lists, text


def run_ours(tmp, lists, fastq_path, prefix, filename=None, extra=()):
    for nm, data in zip(("p.lst", "m.lst", "h.lst"), lists):
        (tmp / nm).write_bytes(data)
    cmd = [str(QUART), "--prefix", prefix] + (["--filename", filename] if filename is not None else []) + list(extra)
    cmd += ["p.lst", "m.lst", "h.lst", str(fastq_path)]
    r = subprocess.run(cmd, cwd=tmp, capture_output=True)
    assert r.returncode == 0, r.stderr[-500:]
    files = {s: (tmp / f"{prefix}.{s}.fastq").read_bytes() for s in SUFFIXES if (tmp / f"{prefix}.{s}.fastq").exists()}
    return files, (tmp / "filter_reads.log").read_bytes(), r.stderr


def run_awk(tmp, lists, fastq_name, prefix, via_stdin=False):
    for nm, data in zip(("p.lst", "m.lst", "h.lst"), lists):
        (tmp / nm).write_bytes(data)
    base = [AWK, "-v", f"prefix={prefix}", "-F", "#|/", "-f", str(REF_AWK), "p.lst", "m.lst", "h.lst"]
    if via_stdin:
        with open(tmp / fastq_name, "rb") as f:
            data = gzip.decompress(f.read()) if fastq_name.endswith(".gz") else f.read()
        r = subprocess.run(base + ["-"], cwd=tmp, input=data, capture_output=True)
    else:
        r = subprocess.run(base + [fastq_name], cwd=tmp, capture_output=True)
    assert r.returncode == 0, r.stderr[-500:]
    files = {s: (tmp / f"{prefix}.{s}.fastq").read_bytes() for s in SUFFIXES if (tmp / f"{prefix}.{s}.fastq").exists()}
    return files, (tmp / "filter_reads.log").read_bytes(), r.stderr


CASES = {"plain": dict(seed=1), "crlf": dict(seed=2, crlf=True), "no_final_newline": dict(seed=3, trailing_newline=False),
         "partial_tail": dict(seed=4, partial_tail=True), "tiny": dict(seed=5, n=3), "empty": dict(seed=6, n=0)}


@pytest.mark.parametrize("name", sorted(CASES))
def test_quartering_matches_awk_and_oracle(tmp_path, name):
    lists, text = make_case(**CASES[name])
    ours_dir, awk_dir = tmp_path / "ours", tmp_path / "awk"
    ours_dir.mkdir(); awk_dir.mkdir()
    (ours_dir / "in.fq").write_bytes(text)
    (awk_dir / "in.fq").write_bytes(text)
    got = run_ours(ours_dir, lists, "in.fq", "in.fq")
    want = post.quartering(*lists, text, b"in.fq")
    assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2]
    if REF_AWK.exists() and AWK:                                # the reference's own program, live
        ref = run_awk(awk_dir, lists, "in.fq", "in.fq")
        assert ref[0] == got[0] and ref[1] == got[1] and ref[2] == got[2]
    if name == "plain":
        assert set(got[0]) == set(SUFFIXES) and got[2].count(b"unclassify") > 0


def test_quartering_gz_input_and_stdin_and_log_append(tmp_path):
    lists, text = make_case(seed=7, n=300)
    (tmp_path / "lane1.fq.gz").write_bytes(gzip.compress(text[:len(text) // 2]) + gzip.compress(text[len(text) // 2:]))
    got = run_ours(tmp_path, lists, "lane1.fq.gz", "lane1.fq", filename="-")
    want = post.quartering(*lists, text, b"-")
    assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2]
    # standard input + a second run appends to filter_reads.log
    for nm, data in zip(("p.lst", "m.lst", "h.lst"), lists):
        (tmp_path / nm).write_bytes(data)
    r = subprocess.run([str(QUART), "--prefix", "again", "p.lst", "m.lst", "h.lst", "-"], cwd=tmp_path, input=text,
                       capture_output=True)
    assert r.returncode == 0
    assert (tmp_path / "filter_reads.log").read_bytes() == want[1] + want[1]
    assert (tmp_path / "again.paternal.fastq").read_bytes() == want[0]["paternal"]
    if REF_AWK.exists() and AWK:
        d = tmp_path / "awk"
        d.mkdir()
        shutil.copy(tmp_path / "lane1.fq.gz", d / "lane1.fq.gz")
        ref = run_awk(d, lists, "lane1.fq.gz", "lane1.fq", via_stdin=True)     # `gzip -dc $x | awk ... -`
        assert ref[0] == got[0] and ref[1] == want[1] and ref[2] == got[2]


def test_quartering_large_blocks(tmp_path):
    """More than one 16 MiB reader block; record and line boundaries fall anywhere."""
    lists, text = make_case(seed=8, n=2000)
    big = text * 120
    (tmp_path / "big.fq").write_bytes(big)
    got = run_ours(tmp_path, lists, "big.fq", "big")
    want = post.quartering(*lists, big, b"big.fq")
    assert got[0] == want[0] and got[1] == want[1]


def test_quartering_missing_input_is_an_error(tmp_path):
    for nm in ("p.lst", "m.lst", "h.lst"):
        (tmp_path / nm).write_bytes(b"")
    r = subprocess.run([str(QUART), "--prefix", "x", "p.lst", "m.lst", "h.lst", "nope.fq"], cwd=tmp_path, capture_output=True)
    assert r.returncode != 0 and b"cannot open" in r.stderr
    r = subprocess.run([str(QUART), "--prefix", "x", "p.lst"], cwd=tmp_path, capture_output=True)
    assert r.returncode == 255


TABLE = (b"0_0_0\t-1\t5\t6\n10_1_1\t0\t7\t0\n1_2_3\t1\t0\t9\n2_2_2\t-1\t0\t0\nlib2_7_8_9\t0\t3\t1\n"
         b"a b\t1\t0\t4\n\t-1\t0\t0\nx y z\t0\t2\t0\n")


@pytest.mark.skipif(AWK is None, reason="awk not installed")
def test_split_oracle_matches_the_scripts_awk_one_liners(tmp_path):
    (tmp_path / "phased.barcodes").write_bytes(TABLE)
    progs = ['{if($2 == 0) print $1;}', '{if($2 == 1) print $1;}', '{if($2 == "-1") print $1;}']   # script :157,160,163
    want = tuple(subprocess.run([AWK, p, "phased.barcodes"], cwd=tmp_path, capture_output=True).stdout for p in progs)
    assert post.split_barcodes(TABLE) == want
