"""Stage 03's per-sequence classifier: bin/classify_seq against the untouched
reference binary (oracle/_ref/classify03, compiled by oracle/Makefile from
03.mkoutput_by_fabulous2.0/src_main/classify.cpp) and the restatement oracle/stage03_classify.py."""
import gzip
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import stage03_classify as s3  # noqa: E402

BIN = ROOT / "bin" / "classify_seq"
REF = ROOT / "oracle" / "_ref" / "classify03"
ACGT = np.frombuffer(b"ACGT", np.uint8)


def rc(s: bytes) -> bytes:
    return s[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))


def make_case(seed, k=21, n_seq=60, long_seq=0, fmt="fasta", crlf=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    genome = bytes(rng.choice(ACGT, 6000))
    def kmers(n, lo, hi):
        out = []
        for _ in range(n):
            p = int(rng.integers(lo, hi - k))
            km = genome[p:p + k]
            out.append(km if rng.random() < 0.5 else rc(km))
        return out
    shared = kmers(15, 2500, 3500)
    pat = kmers(150, 0, 3000) + shared + [b"ACGT"] * 2              # other-length lines: counted, never match
    mat = kmers(120, 3000, 6000) + shared + kmers(5, 3000, 6000) * 2   # duplicates count as lines
    recs = []
    for i in range(n_seq):
        kind = i % 8
        L = int(rng.integers(1, 400))
        p = int(rng.integers(0, len(genome) - 400))
        seq = bytearray(genome[p:p + L])
        if kind == 1:
            seq = bytearray(rc(bytes(seq)))
        elif kind == 2 and L > 30:                                    # N run: only the windows over it are lost
            a = int(rng.integers(0, L - 5))
            seq[a:a + int(rng.integers(1, 30))] = b"N" * len(seq[a:a + int(rng.integers(1, 30))])
        elif kind == 3:
            seq = bytearray(bytes(seq).lower())                       # string matching: lower case never matches
        elif kind == 4 and L > 10:
            seq[int(rng.integers(0, L))] = ord("n")
        elif kind == 5:
            seq = bytearray(seq[:int(rng.integers(0, k))])            # shorter than k (also empty)
        elif kind == 6 and L > 40:
            q = int(rng.integers(0, L))
            seq[q] = ord("R")                                         # IUPAC
        recs.append((b"seq%d some description/%d" % (i, i), bytes(seq)))
    for j in range(long_seq):                                         # longer than one 8192-base chunk
        L = int(rng.integers(9000, 40000))
        seq = (genome * (L // len(genome) + 1))[:L]
        recs.append((b"long%d" % j, seq))
    eol = b"\r\n" if crlf else b"\n"
    if fmt == "fasta":
        text = b""
        for name, seq in recs:
            text += b">" + name + eol
            w = int(rng.integers(20, 90))
            for a in range(0, len(seq), w):
                text += seq[a:a + w] + eol
            if rng.random() < 0.2:
                text += b"\n"                                          # blank lines are skipped
    else:
        text = b"".join(b"@" + n + eol + s + eol + b"+" + eol + b"F" * len(s) + eol for n, s in recs)
    to_list = lambda ks: b"".join(x + b"\n" for x in ks)
    return to_list(pat), to_list(mat), text


def run(binary, tmp, pat, mat, text, fmt, name="in.fa", threads=3):
    (tmp / "p.mer").write_bytes(pat)
    (tmp / "m.mer").write_bytes(mat)
    (tmp / name).write_bytes(gzip.compress(text) if name.endswith(".gz") else text)
    r = subprocess.run([str(binary), "--hap", "p.mer", "--hap", "m.mer", "--read", name, "--thread", str(threads),
                        "--format", fmt], cwd=tmp, capture_output=True)
    return r


def test_restatement_matches_reference_binary(tmp_path):
    """CPU: pins oracle/stage03_classify.py to the untouched reference."""
    if not REF.exists():
        pytest.skip("oracle/_ref/classify03 not built (no /root/reference here)")
    for seed, fmt, crlf in ((1, "fasta", False), (2, "fastq", False), (3, "fasta", True)):
        pat, mat, text = make_case(seed, fmt=fmt, crlf=crlf, long_seq=1)
        r = run(REF, tmp_path, pat, mat, text, fmt)
        assert r.returncode == 0
        assert r.stdout == s3.classify(pat, mat, text, fmt)
        assert b"haplotype0" in r.stdout and b"haplotype1" in r.stdout and b"ambiguous" in r.stdout


def test_reference_matches_odd_kmer_lines_as_strings(tmp_path):
    """The one documented divergence of bin/classify_seq, pinned from the reference's side: its string sets accept ANY
    line (03.mkoutput_by_fabulous2.0/src_main/classify.cpp:52-72), so a k-mer holding 'N' or lower case scores the
    sequences that contain exactly that text.  The restatement reproduces it; bin/classify_seq refuses such a list
    with an error (test_classify_seq_rejects_lists_it_cannot_represent) instead of silently scoring differently."""
    if not REF.exists():
        pytest.skip("oracle/_ref/classify03 not built (no /root/reference here)")
    pat, mat, text = make_case(21)
    odd = b"ACGTNACGTACGTACGTACGT"
    low = b"acgtacgtacgtacgtacgta"
    text += b">with_odd\n" + b"TT" + odd + b"GG" + low + b"CC\n"
    r = run(REF, tmp_path, pat + odd + b"\n" + low + b"\n", mat, text, "fasta")
    assert r.returncode == 0
    assert r.stdout == s3.classify(pat + odd + b"\n" + low + b"\n", mat, text, "fasta")
    row = [ln for ln in r.stdout.splitlines() if ln.startswith(b"with_odd")]
    base = run(REF, tmp_path, pat, mat, text, "fasta").stdout
    assert row and row != [ln for ln in base.splitlines() if ln.startswith(b"with_odd")]     # the odd lines did score


@pytest.mark.parametrize("args", [[], ["--hap", "a", "--read", "r"], ["--hap", "a", "--hap", "b"], ["-h"],
                                  ["--hap", "a", "--hap", "b", "--read", "r", "--format", "bam"]])
def test_usage_exit_code(args):
    r = subprocess.run([str(BIN)] + args, capture_output=True)
    assert r.returncode == 255 and r.stdout == b""
    if REF.exists():
        assert subprocess.run([str(REF)] + args, capture_output=True).returncode == 255


@pytest.mark.gpu
@pytest.mark.parametrize("seed,k,fmt,crlf,name", [(11, 21, "fasta", False, "in.fa"), (12, 21, "fastq", False, "in.fq"),
                                                  (13, 17, "fasta", True, "in.fa"), (14, 31, "fasta", False, "in.fa.gz"),
                                                  (15, 5, "fastq", False, "in.fq.gz")])
def test_classify_seq_matches_oracle_and_reference(tmp_path, seed, k, fmt, crlf, name):
    pat, mat, text = make_case(seed, k=k, fmt=fmt, crlf=crlf, long_seq=3)
    r = run(BIN, tmp_path, pat, mat, text, fmt, name=name)
    assert r.returncode == 0, r.stderr[-500:].decode(errors="replace")
    want = s3.classify(pat, mat, text, fmt)
    assert r.stdout == want
    assert want.count(b"\n") > 60 and b"haplotype0" in want and b"haplotype1" in want
    if REF.exists():                                                   # the reference itself, live
        ref = run(REF, tmp_path, pat, mat, text, fmt, name=name)
        assert ref.returncode == 0 and ref.stdout == r.stdout


@pytest.mark.gpu
def test_classify_seq_rejects_lists_it_cannot_represent(tmp_path):
    pat, mat, text = make_case(16)
    r = run(BIN, tmp_path, pat + b"ACGTNACGTACGTACGTACGT\n", mat, text, "fasta")
    assert r.returncode != 0 and b"upper-case ACGT" in r.stderr and r.stdout == b""
