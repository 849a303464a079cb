"""GPU: the drop-in process bin/classify against (1) the golden fixtures produced by the
untouched reference binary and (2), when oracle/_ref travelled with the repo, live runs
of that binary on a synthetic trio -- byte-identical stdout is the bar."""
import json
import os
import subprocess
from pathlib import Path

import pytest

import cases
import oracle as orc
from hast_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
CLASSIFY = ROOT / "bin" / "classify"
GOLDEN = Path(__file__).resolve().parent / "golden"
CLASSIFY_CASES = sorted(p.name for p in GOLDEN.iterdir() if (p / "pat.mer").exists())


def run_cli(args, cwd=None, env=None):
    e = dict(os.environ, HAST_GPUS="1")
    e.update(env or {})
    return subprocess.run([str(CLASSIFY)] + [str(a) for a in args], cwd=cwd, capture_output=True, env=e)


@pytest.mark.parametrize("packed", ["1", "0"], ids=["packed_h2d", "ascii_h2d"])
@pytest.mark.parametrize("name", CLASSIFY_CASES)
def test_cli_matches_golden(name, packed):
    """packed_h2d: the parser packs to 2 bits and submits through hast_submit_batch_packed (default);
    ascii_h2d: the ASCII bases travel and the GPU packs them (HAST_PACKED=0)."""
    d = GOLDEN / name
    args = json.loads((d / "cmd.txt").read_text())
    r = run_cli(args, cwd=d, env={"HAST_PACKED": packed})
    assert r.returncode == 0, r.stderr[-600:].decode(errors="replace")
    assert r.stdout == (d / "expected.tsv").read_bytes()
    if name == "adv_k21":                      # log lines users look for (classify.cpp:45,321)
        assert b"Recorded" in r.stderr and b"haplotype 0 specific 21-mers" in r.stderr
        assert b"INFO : erase a adaptor kmer from hap" in r.stderr


def test_cli_small_blocks_and_many_threads_same_bytes():
    d = GOLDEN / "adv_k21"
    args = json.loads((d / "cmd.txt").read_text())
    r = run_cli(args + ["-t", "7"], cwd=d, env={"HAST_BLOCK_MB": "0"})
    assert r.returncode == 0 and r.stdout == (d / "expected.tsv").read_bytes()


def test_cli_errors_are_nonzero(tmp_path):
    c = cases.adversarial_case(21, 60, seed=3)
    (tmp_path / "p.mer").write_bytes(c["pat_text"])
    (tmp_path / "m.mer").write_bytes(c["mat_text"])
    cases.write_fastq(tmp_path / "ok.fq", c["heads"], c["reads"])
    base = ["--hap0", tmp_path / "p.mer", "--hap1", tmp_path / "m.mer"]
    # read shorter than k: the reference dies with SIGABRT (kmer.h:171); we fail with a message
    cases.write_fastq(tmp_path / "short.fq", c["heads"][:3], [c["reads"][0], b"ACGTACGT", c["reads"][2]])
    r = run_cli(base + ["--read", tmp_path / "short.fq"])
    assert r.returncode != 0 and b"shorter than k" in r.stderr and r.stdout == b""
    # k-mer line of the wrong length (kmer.h:154)
    (tmp_path / "bad.mer").write_bytes(c["mat_text"] + b"ACGT\n")
    r = run_cli(["--hap0", tmp_path / "p.mer", "--hap1", tmp_path / "bad.mer", "--read", tmp_path / "ok.fq"])
    assert r.returncode != 0 and r.stdout == b""
    # adaptor shorter than k (chopRead2Kmer assert)
    r = run_cli(base + ["--read", tmp_path / "ok.fq", "--adaptor_f", "ACGT"])
    assert r.returncode != 0
    # missing input
    r = run_cli(base + ["--read", tmp_path / "nope.fq"])
    assert r.returncode != 0 and b"cannot open" in r.stderr


@pytest.fixture(scope="module")
def trio_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("trio")
    t = synth.make_trio(synth.config("small"))
    pat, mat = t.write_kmer_lists(d)
    r1, r2 = t.write_fastq(d, gz=False)
    g1, g2 = t.write_fastq(d, gz=True, stem="childgz")
    return t, d, pat, mat, (r1, r2), (g1, g2)


def expected_table(trio_files, tmp_path, reads, w0="1.04"):
    t, d, pat, mat, _, _ = trio_files
    if orc.ref_binary("classify_O2") is not None:        # the untouched reference, when it travelled
        return orc.run_ref_classify(pat, mat, reads, extra=["--weight0", w0])
    o = orc.Oracle()
    o.load_kmers_file(pat, 0)
    o.load_kmers_file(mat, 1)
    o.init_adaptor()
    o.set_weights(float(w0), 1.0)
    for r in reads:
        assert o.process_fastq(r) == 0
    return o.table(tmp_path)


def test_cli_synthetic_trio_plain_and_gz(trio_files, tmp_path):
    t, d, pat, mat, plain, gz = trio_files
    want = expected_table(trio_files, tmp_path, plain)
    stats = tmp_path / "stats.json"
    r = run_cli(["--hap0", pat, "--hap1", mat, "--read", plain[0], "--read", plain[1], "--weight0", "1.04",
                 "--stats-json", stats])
    assert r.returncode == 0, r.stderr[-500:]
    assert r.stdout == want
    s = json.loads(stats.read_text())
    assert s["reads"] == 2 * t.spec.n_pairs and s["barcodes"] == len(want.splitlines()) and s["kernel_launches"] > 0
    r = run_cli(["--hap0", pat, "--hap1", mat, "--read", gz[0], "--read", gz[1], "--weight0", "1.04", "-t", "2"])
    assert r.returncode == 0 and r.stdout == want


def test_cli_then_stage_script_awk_split(trio_files, tmp_path):
    """classify_stlfr_reads.sh:156-162 consumes the table with three awk one-liners on column 2."""
    t, d, pat, mat, plain, _ = trio_files
    r = run_cli(["--hap0", pat, "--hap1", mat, "--read", plain[0], "--read", plain[1], "--weight0", "1.04"])
    (tmp_path / "phased.barcodes").write_bytes(r.stdout)
    n = {}
    for name, cond in (("paternal", "$2==0"), ("maternal", "$2==1"), ("homozygous", '$2=="-1"')):
        out = subprocess.run(["awk", f"{cond}{{print $1}}", str(tmp_path / "phased.barcodes")], capture_output=True)
        n[name] = len(out.stdout.splitlines())
    assert sum(n.values()) == len(r.stdout.splitlines()) and n["paternal"] > 0 and n["maternal"] > 0


def test_cli_split_and_partition_equal_the_script_flow(trio_files, tmp_path):
    """bin/classify --partition-reads does classify_stlfr_reads.sh:148-185 in one process: the table on
    stdout, the three barcode lists (:156-162) and quartering_fastq.awk's outputs (:176-185) for every
    input, identical to running the awk programs (restated in oracle/stage01_post.py) on the table."""
    import sys
    sys.path.insert(0, str(ROOT / "oracle"))
    import stage01_post as post
    t, d, pat, mat, plain, gz = trio_files
    out = tmp_path / "out"
    out.mkdir()
    reads = [plain[0], gz[1]]                          # one plain, one gzip input
    r = run_cli(["--hap0", pat, "--hap1", mat, "--read", reads[0], "--read", reads[1], "--weight0", "1.04",
                 "--partition-reads", "--outdir", out])
    assert r.returncode == 0, r.stderr[-600:].decode(errors="replace")
    table = r.stdout
    want_lists = post.split_barcodes(table)
    names = ["paternal.unique.barcodes", "maternal.unique.barcodes", "homozygous.unique.barcodes"]
    for nm, want in zip(names, want_lists):
        assert (out / nm).read_bytes() == want
    assert all(len(w) > 0 for w in want_lists)
    log = b""
    import gzip as gz_mod
    for path in reads:
        raw = Path(path).read_bytes()
        is_gz = str(path).endswith(".gz")
        text = gz_mod.decompress(raw) if is_gz else raw
        prefix = Path(path).name[:-3] if is_gz else Path(path).name
        files, lg, _ = post.quartering(*want_lists, text, b"-" if is_gz else str(path).encode())
        log += lg
        for suffix in ("nobarcode", "paternal", "maternal", "homozygous"):
            p = out / f"{prefix}.{suffix}.fastq"
            assert (p.read_bytes() if p.exists() else None) == files.get(suffix), (prefix, suffix)
        assert sum(len(v) for v in files.values()) == len(text)      # every record lands somewhere
    assert (out / "filter_reads.log").read_bytes() == log
    # --split-barcodes alone writes the lists and no FASTQ
    out2 = tmp_path / "out2"
    out2.mkdir()
    r = run_cli(["--hap0", pat, "--hap1", mat, "--read", reads[0], "--weight0", "1.04", "--split-barcodes", "--outdir", out2])
    assert r.returncode == 0 and sorted(p.name for p in out2.iterdir()) == sorted(names)
