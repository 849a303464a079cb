#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by RUNNING THE UNTOUCHED REFERENCE.

    python tests/golden/make_golden.py

Needs oracle/_ref/{classify,mergeResult} (make -C oracle ref, which compiles the
reference sources where they lie under /root/reference).  Every case directory
holds the exact input files and `expected.tsv` = the reference binary's stdout;
`cmd.txt` records the arguments.  The as-shipped build (-g, no -O) is the
authority; the -O2 build is asserted to give identical bytes.
"""
import json
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))
import cases  # noqa: E402

REF = ROOT / "oracle" / "_ref"


def run_ref(case_dir: Path, args: list[str], binary="classify") -> bytes:
    r = subprocess.run([str(REF / binary)] + args, cwd=case_dir, capture_output=True)
    assert r.returncode == 0, (binary, args, r.stderr[-500:])
    return r.stdout


def classify_case(name, k, n_reads, seed, *, weight0="1.04", weight1=None, crlf=False, gz=False,
                  trailing_newline=True, twice=False, kmer_tail=False, adaptors=None, truncate_last=False,
                  max_len=140):
    d = HERE / name
    if d.exists():
        shutil.rmtree(d)
    d.mkdir(parents=True)
    c = cases.adversarial_case(k, n_reads, seed, max_len=max_len)
    pat, mat = c["pat_text"], c["mat_text"]
    if kmer_tail:                        # last line without '\n' is dropped (classify.cpp:41)
        pat = pat + pat[: k]             # a full-length k-mer, unterminated
        mat = mat + b"ACG"               # a short fragment, unterminated
    (d / "pat.mer").write_bytes(pat)
    (d / "mat.mer").write_bytes(mat)
    fq = "reads.fq.gz" if gz else "reads.fq"
    cases.write_fastq(d / fq, c["heads"], c["reads"], crlf=crlf, trailing_newline=trailing_newline)
    if truncate_last:                    # final record keeps only header + sequence lines
        raw = (d / fq).read_bytes()
        cut = raw.rstrip(b"\n").rfind(b"\n+")
        (d / fq).write_bytes(raw[:cut + 1])
    args = ["--hap0", "pat.mer", "--hap1", "mat.mer", "--read", fq, "--thread", "3"]
    if twice:
        args += ["--read", fq]
    if weight0:
        args += ["--weight0", weight0]
    if weight1:
        args += ["--weight1", weight1]
    if adaptors:
        args += ["--adaptor_f", adaptors[0], "--adaptor_r", adaptors[1]]
    out = run_ref(d, args)
    assert out == run_ref(d, args, "classify_O2"), "O0 and O2 reference builds disagree"
    (d / "expected.tsv").write_bytes(out)
    (d / "cmd.txt").write_text(json.dumps(args))
    print(f"{name}: {len(out.splitlines())} barcodes, {sum(f.stat().st_size for f in d.iterdir()) >> 10} KiB")


def main():
    if not (REF / "classify").exists():
        sys.exit("oracle/_ref/classify missing: run `make -C oracle ref` in a container that has /root/reference")
    classify_case("adv_k21", 21, 160, 1)
    classify_case("adv_k5", 5, 120, 2, max_len=60)
    classify_case("adv_k16", 16, 120, 3)
    classify_case("adv_k17", 17, 120, 4, weight0=None)
    classify_case("adv_k25", 25, 120, 5, weight0="0.9", weight1="1.3")
    classify_case("adv_k31", 31, 120, 6)
    classify_case("adv_k32", 32, 120, 7)
    classify_case("crlf_k21", 21, 100, 8, crlf=True)           # '\r' counts as a base (T), SURVEY A.9
    classify_case("gz_members_k21", 21, 140, 9, gz=True, trailing_newline=False)
    classify_case("twice_k21", 21, 100, 10, twice=True)          # a file listed twice counts twice
    classify_case("kmer_tail_k21", 21, 100, 11, kmer_tail=True)
    classify_case("custom_adaptor_k21", 21, 100, 12,
                  adaptors=("CTGTCTCTTATACACATCTTAGGAAGACAA", "TCTGCTGAGTCGAGAACGTCTCTG"))
    classify_case("truncated_tail_k21", 21, 100, 13, truncate_last=True)

    # mergeResult: bug-compatible sum (mergeResult.cpp:28-29)
    d = HERE / "merge"
    if d.exists():
        shutil.rmtree(d)
    d.mkdir()
    # Lines whose barcode is the empty string are left out: mergeResult reads them shifted by one
    # column (operator>> skips the leading tab) and then adds an UNINITIALISED int (hap1 is never
    # written when the stream hits end-of-line first, mergeResult.cpp:24-29), so the reference's
    # own output for them is stack garbage and cannot be a golden value.
    for src, dst in (("adv_k21", "a.tsv"), ("twice_k21", "b.tsv")):
        lines = (HERE / src / "expected.tsv").read_bytes().splitlines(keepends=True)
        (d / dst).write_bytes(b"".join(l for l in lines if not l.startswith(b"\t")))
    (d / "c.tsv").write_text("1_2_3\t1\t1\t3\n0_0_0\t-1\t5\t0\nzz\t-1\t0\t0\n")
    args = ["--input", "a.tsv", "--input", "b.tsv", "--input", "c.tsv", "--input", "c.tsv", "--weight0", "2.5"]
    out = run_ref(d, args, "mergeResult")
    (d / "expected.tsv").write_bytes(out)
    (d / "cmd.txt").write_text(json.dumps(args))
    print(f"merge: {len(out.splitlines())} lines")


if __name__ == "__main__":
    main()
