"""GPU: the real drop-in boundary -- the reference's own, unchanged stage script
(01.classify_stlfr_reads/classify_stlfr_reads.sh:130-185) run once around the reference's classify binary and once
around bin/classify, on the same inputs (one plain, one gzip FASTQ): every file the stage leaves behind must be
identical.  Then `bin/classify --split-barcodes --partition-reads`, which folds the script's awk passes into the same
process, against those same files.

The script and the awk program are staged, verbatim, by `make -C oracle ref` into oracle/_ref/stage01 (git-ignored
reference artefacts that travel to the GPU box like the reference binaries); nothing here reads /root/reference."""
import os
import re
import shutil
import subprocess
from pathlib import Path

import pytest

from hast_b200 import synth

ROOT = Path(__file__).resolve().parent.parent
STAGE = ROOT / "oracle" / "_ref" / "stage01"
OURS = ROOT / "bin" / "classify"
pytestmark = pytest.mark.gpu

STAGE_FILES = ["phased.barcodes", "paternal.unique.barcodes", "maternal.unique.barcodes", "homozygous.unique.barcodes",
               "filter_reads.log"]
PARTS = ["paternal", "maternal", "homozygous", "nobarcode"]


def stage_outputs(d: Path, read_names):
    out = {n: (d / n).read_bytes() for n in STAGE_FILES}
    for r in read_names:
        stem = r[:-3] if r.endswith(".gz") else r
        for p in PARTS:
            f = d / f"{stem}.{p}.fastq"
            out[f.name] = f.read_bytes() if f.exists() else None
    return out


def run_script(stage_dir: Path, cwd: Path, pat, mat, reads, threads=4):
    cmd = ["bash", str(stage_dir / "classify_stlfr_reads.sh"), "--paternal_mer", pat, "--maternal_mer", mat, "--thread", str(threads)]
    for r in reads:
        cmd += ["--filial", r]
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0 and "__END__" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:] + (cwd / "phased.log").read_text()[-1500:]
    return r.stdout


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    if not (STAGE / "classify_stlfr_reads.sh").exists() or not (STAGE / "classify").exists():
        pytest.skip("oracle/_ref/stage01 not staged (run `make -C oracle ref` where /root/reference exists)")
    d = tmp_path_factory.mktemp("stage_inputs")
    spec = synth.config("small")
    spec.lowercase_frac = 0.01
    t = synth.make_trio(spec)
    pat, mat = t.write_kmer_lists(d)
    r1, _ = t.write_fastq(d, gz=False)
    _, r2 = t.write_fastq(d / "gz", gz=True)
    r2 = shutil.copy(r2, d / "child.r2.fq.gz")
    return pat, mat, [str(r1), str(r2)]


def test_reference_stage_script_runs_unchanged_around_bin_classify(inputs, tmp_path):
    pat, mat, reads = inputs
    a, b = tmp_path / "reference", tmp_path / "ours"
    a.mkdir(); b.mkdir()
    # the reference's directory as `make` leaves it
    out_a = run_script(STAGE, a, pat, mat, reads)
    # the same directory with the binary replaced: the script finds "$SPATH/classify" (classify_stlfr_reads.sh:112)
    mine = tmp_path / "stage_with_bin_classify"
    mine.mkdir()
    for f in ("classify_stlfr_reads.sh", "quartering_fastq.awk"):
        shutil.copy(STAGE / f, mine / f)
    (mine / "classify").write_text(f'#!/bin/bash\nexec "{OURS}" "$@"\n')
    (mine / "classify").chmod(0o755)
    out_b = run_script(mine, b, pat, mat, reads)
    names = [Path(r).name for r in reads]
    fa, fb = stage_outputs(a, names), stage_outputs(b, names)
    assert set(fa) == set(fb)
    for n in fa:
        assert fa[n] == fb[n], f"{n} differs between the reference binary and bin/classify under the reference's script"
    assert fa["phased.barcodes"].count(b"\n") > 900 and fa["child.r1.fq.paternal.fastq"]
    for s in ("step_9_done", "step_10_done", "step_11_done"):
        assert (a / s).exists() and (b / s).exists()
    # what the script prints about its own progress (counts of the three lists) agrees line for line, dates aside
    is_date = re.compile(r"^\w{3} \w{3} +\d+ \d\d:\d\d:\d\d ")                 # `date` lines
    strip = lambda o: [ln for ln in o.splitlines() if not is_date.match(ln) and "CMD :" not in ln and "in dir" not in ln]
    assert [l.split("/")[-1] for l in strip(out_a)] == [l.split("/")[-1] for l in strip(out_b)]

    # ---- the same stage in ONE process: bin/classify --split-barcodes --partition-reads ----------------------
    c = tmp_path / "folded"
    c.mkdir()
    cmd = [str(OURS), "--hap0", pat, "--hap1", mat, "--thread", "4", "--weight0", "1.04", "--split-barcodes",
           "--partition-reads", "--outdir", str(c)]
    for r in reversed(reads):                      # the script prepends every --filial (:86): same file order
        cmd += ["--read", r]
    r = subprocess.run(cmd, capture_output=True, cwd=c)
    assert r.returncode == 0, r.stderr[-1500:]
    (c / "phased.barcodes").write_bytes(r.stdout)
    fc = stage_outputs(c, names)
    for n in fa:
        assert fa[n] == fc[n], f"{n} differs between the script and bin/classify --partition-reads"


def test_script_skips_finished_steps_with_bin_classify_too(inputs, tmp_path):
    """step_*_done sentinels (classify_stlfr_reads.sh:146-152,167-169,186-191): a second run does nothing."""
    pat, mat, reads = inputs
    mine = tmp_path / "stage"
    mine.mkdir()
    for f in ("classify_stlfr_reads.sh", "quartering_fastq.awk"):
        shutil.copy(STAGE / f, mine / f)
    (mine / "classify").write_text(f'#!/bin/bash\nexec "{OURS}" "$@"\n')
    (mine / "classify").chmod(0o755)
    w = tmp_path / "w"
    w.mkdir()
    run_script(mine, w, pat, mat, reads[:1])
    before = {p.name: p.stat().st_mtime_ns for p in w.iterdir()}
    out = run_script(mine, w, pat, mat, reads[:1])
    assert "skip classify because step_9_done" in out and "skip extract reads" in out
    assert before == {p.name: p.stat().st_mtime_ns for p in w.iterdir()}
