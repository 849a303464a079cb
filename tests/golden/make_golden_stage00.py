#!/usr/bin/env python
"""Generate tests/golden/stage00_* by RUNNING THE REFERENCE'S OWN stage-00 script
(/root/reference/00.build_unshare_kmers_by_jellyfish/build_unshared_kmers.sh with the
jellyfish 2.3.0 binary it vendors) on small synthetic parental reads.

    python tests/golden/make_golden_stage00.py

Every case directory holds the exact input files, `cmd.txt` (the script arguments) and
`expected/`: the two `*.unique.filter.mer` lists with their LINES SORTED (jellyfish dumps in
hash-table order, which depends on its -s/-t arguments; the set is the contract) and, for
--auto_bounds cases, `*.histo` and `*.bounds.txt` byte for byte.
"""
import gzip
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from hast_b200 import synth  # noqa: E402

SCRIPT = Path("/root/reference/00.build_unshare_kmers_by_jellyfish/build_unshared_kmers.sh")


def write_fasta(path: Path, reads: np.ndarray, width: int, gz: bool):
    """Multi-line FASTA: reads glued into records of 7 reads each, cut into lines of `width`, with
    a blank line inside some records (jellyfish skips it and joins across it)."""
    out = []
    for i in range(0, reads.shape[0], 7):
        seq = b"NN".join(r.tobytes() for r in reads[i:i + 7])      # 'N' breaks windows between reads
        out.append(b">rec%d some description\n" % i)
        for j in range(0, len(seq), width):
            out.append(seq[j:j + width] + b"\n")
            if j == 0 and (i // 7) % 3 == 0:
                out.append(b"\n")
    data = b"".join(out)
    with (gzip.open(path, "wb") if gz else open(path, "wb")) as f:
        f.write(data)


def case(name: str, k: int, genome: int, cov: float, args: list[str], seed: int, fasta_maternal=False, gz=False,
         two_paternal_files=False):
    d = HERE / name
    if d.exists():
        shutil.rmtree(d)
    (d / "expected").mkdir(parents=True)
    spec = synth.TrioSpec(genome_len=genome, het=0.01, k=k, seed=seed)
    r = synth.parent_reads(spec, cov, read_len=100, err=0.004, n_frac=0.02, lowercase_frac=0.03)
    ext = ".fq.gz" if gz else ".fq"
    pats = []
    if two_paternal_files:
        h = r["paternal"].shape[0] // 2
        synth.write_reads_fastq(d / ("pat_a" + ext), r["paternal"][:h], gz=gz)
        synth.write_reads_fastq(d / ("pat_b" + ext), r["paternal"][h:], gz=gz)
        pats = ["pat_a" + ext, "pat_b" + ext]
    else:
        synth.write_reads_fastq(d / ("pat" + ext), r["paternal"], gz=gz)
        pats = ["pat" + ext]
    if fasta_maternal:
        mat = "mat.fa.gz" if gz else "mat.fa"
        write_fasta(d / mat, r["maternal"], 60, gz)
    else:
        mat = "mat" + ext
        synth.write_reads_fastq(d / mat, r["maternal"], gz=gz)
    cmd = []
    for p in pats:
        cmd += ["--paternal", p]
    cmd += ["--maternal", mat, "--mer", str(k)] + args
    (d / "cmd.txt").write_text(" ".join(cmd) + "\n")
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        full = []
        for a in cmd:
            full.append(str(d / a) if (d / a).exists() else a)
        res = subprocess.run(["bash", str(SCRIPT)] + full + ["--thread", "2", "--memory", "1"], cwd=td,
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        for f in ("paternal.unique.filter.mer", "maternal.unique.filter.mer"):
            lines = sorted((td / f).read_bytes().splitlines())
            (d / "expected" / f).write_bytes(b"".join(x + b"\n" for x in lines))
            print(name, f, len(lines))
        if "--auto_bounds" in args:
            for f in ("paternal.histo", "maternal.histo", "paternal.bounds.txt", "maternal.bounds.txt"):
                shutil.copy(td / f, d / "expected" / f)
            print(name, (td / "paternal.bounds.txt").read_text().replace("\n", " "))


if __name__ == "__main__":
    assert SCRIPT.exists(), "needs /root/reference"
    case("stage00_auto_k21", 21, 3000, 24.0, ["--auto_bounds"], seed=3)
    case("stage00_bounds_k17_fasta_gz", 17, 2500, 20.0, ["--p-lower", "5", "--p-upper", "40", "--m-lower", "4", "--m-upper", "30"],
         seed=4, fasta_maternal=True, gz=True, two_paternal_files=True)
    case("stage00_auto_k31", 31, 2500, 22.0, ["--auto_bounds"], seed=5)
