"""CPU: host-side logic of bin/classify and bin/mergeResult (no GPU needed).

CLI contract (classify.cpp:375-428), FASTQ framing / gzip / parseName / barcode
interning / output order through the HAST_PARSE_ONLY diagnostic, and mergeResult
against the golden produced by the reference binary.
"""
import json
import os
import subprocess
from collections import Counter
from pathlib import Path

import numpy as np
import pytest

import cases
import oracle as orc

ROOT = Path(__file__).resolve().parent.parent
CLASSIFY = ROOT / "bin" / "classify"
MERGE = ROOT / "bin" / "mergeResult"
GOLDEN = Path(__file__).resolve().parent / "golden"


def run(cmd, **kw):
    return subprocess.run([str(c) for c in cmd], capture_output=True, **kw)


@pytest.mark.parametrize("args", [[], ["-h"], ["--help"], ["-l", "x"], ["--bogus"], ["--hap0", "a", "--hap1", "b"],
                                  ["--hap0", "a", "--read", "r"], ["--hap0", "a", "--hap1", "b", "--read", "r", "-t", "0"]])
def test_usage_and_exit_255(args):
    r = run([CLASSIFY] + args)
    assert r.returncode == 255                      # `return -1` from main, classify.cpp:421-427
    assert b"Uasge" in r.stderr and r.stdout == b""  # usage goes to stderr, stdout stays pure
    ref = orc.ref_binary("classify")
    if ref is not None:
        assert run([ref] + args).returncode == 255


def parse_only(files, threads=3, block_mb=None, slice_bytes=None, serial=False):
    env = dict(os.environ, HAST_PARSE_ONLY="1")
    if block_mb is not None:
        env["HAST_BLOCK_MB"] = str(block_mb)
    if slice_bytes is not None:
        env["HAST_SLICE_BYTES"] = str(slice_bytes)
    if serial:
        env["HAST_SERIAL_READER"] = "1"
    cmd = [CLASSIFY, "--hap0", "x", "--hap1", "y", "-t", str(threads)]
    for f in files:
        cmd += ["--read", f]
    r = run(cmd, env=env)
    assert r.returncode == 0, r.stderr[-400:]
    return r.stdout


def expected_parse(heads, reads):
    n, b = Counter(), Counter()
    for h, r in zip(heads, reads):
        bc = orc.parse_name(h)
        n[bc] += 1
        b[bc] += len(r)
    return b"".join(b"%s\t%d\t%d\n" % (k, n[k], b[k]) for k in sorted(n))


@pytest.mark.parametrize("crlf", [False, True])
def test_front_end_framing_barcodes_and_order(tmp_path, crlf):
    c = cases.adversarial_case(21, 3000, seed=77)
    heads = c["heads"] + [b"@weird header no separators", b"@x/1#tail_bc", b"@a#b#c/d/e"]
    reads = c["reads"] + [b"ACGT" * 10, b"TTTT" * 12, b"GG" * 30]
    cases.write_fastq(tmp_path / "a.fq", heads[:1000], reads[:1000], crlf=crlf)
    cases.write_fastq(tmp_path / "b.fq.gz", heads[1000:], reads[1000:], crlf=crlf, trailing_newline=False)
    got = parse_only([tmp_path / "a.fq", tmp_path / "b.fq.gz"])
    if crlf:                                         # '\r' stays part of header and sequence
        heads = [h + b"\r" for h in heads]
        reads = [r + b"\r" for r in reads]
    assert got == expected_parse(heads, reads)
    # tiny blocks: many block boundaries, same answer; file listed twice counts twice
    assert parse_only([tmp_path / "a.fq", tmp_path / "b.fq.gz"], threads=2, block_mb=0) == got
    twice = parse_only([tmp_path / "a.fq", tmp_path / "a.fq"])
    assert twice == expected_parse(heads[:1000] * 2, reads[:1000] * 2)


def test_front_end_eof_semantics(tmp_path):
    # classify.cpp:257-269: a terminated header is processed with whatever follows; an unterminated one is dropped
    rec = b"@r1#1_1_1/1\nACGTACGTACGTACGTACGTACGT\n+\nFFFFFFFFFFFFFFFFFFFFFFFF\n"
    (tmp_path / "a.fq").write_bytes(rec + b"@r2#2_2_2/1\nACGTACGTACGTACGTACGTAAAA")          # seq unterminated, no + / qual
    assert parse_only([tmp_path / "a.fq"]) == b"1_1_1\t1\t24\n2_2_2\t1\t24\n"
    (tmp_path / "b.fq").write_bytes(rec + b"@r2#2_2_2/1")                                       # unterminated header
    assert parse_only([tmp_path / "b.fq"]) == b"1_1_1\t1\t24\n"
    (tmp_path / "c.fq").write_bytes(rec + b"@r2#2_2_2/1\n")                                     # header, then nothing: empty read
    assert parse_only([tmp_path / "c.fq"]) == b"1_1_1\t1\t24\n2_2_2\t1\t0\n"
    (tmp_path / "d.fq").write_bytes(b"")
    assert parse_only([tmp_path / "d.fq"]) == b""
    r = run([CLASSIFY, "--hap0", "x", "--hap1", "y", "--read", tmp_path / "missing.fq"],
            env=dict(os.environ, HAST_PARSE_ONLY="1"))
    assert r.returncode == 1 and b"cannot open" in r.stderr


@pytest.mark.parametrize("slice_bytes", [1, 7, 64, 257, 4096, 100_000])
def test_plain_files_sliced_across_threads_frame_like_getline(tmp_path, slice_bytes):
    """plain_slicer.h: many threads per plain file, records framed by newline COUNT (classify.cpp:257-269 validates
    neither '@' nor '+').  Quality lines that start with '@' or '+', '@' lines in sequence position, empty lines,
    records longer than a slice, CRLF, an unterminated tail: every slice size gives the serial reader's answer."""
    rng = np.random.default_rng(slice_bytes)
    heads, reads, quals = [], [], []
    for i in range(1500):
        bc = b"%d_%d_%d" % tuple(rng.integers(1, 40, 3))
        heads.append(b"@r%d#%s/%d" % (i, bc, 1 + i % 2) if i % 11 else b"@no barcode here %d" % i)
        L = int(rng.integers(0, 300)) if i % 97 else 9000            # some records far longer than the small slices
        reads.append(bytes(rng.choice(list(b"ACGTN@+"), L).astype(np.uint8)))
        quals.append(bytes(rng.choice(list(b"@+F#/I\r"), L).astype(np.uint8)))
    body = b"".join(h + b"\n" + r + b"\n+" + (h[1:] if i % 5 == 0 else b"") + b"\n" + q + b"\n"
                    for i, (h, r, q) in enumerate(zip(heads, reads, quals)))
    (tmp_path / "a.fq").write_bytes(body)
    (tmp_path / "b.fq").write_bytes(body + b"@tail#9_9_9/1\nACGT")          # unterminated sequence line at EOF
    (tmp_path / "c.fq").write_bytes(body + b"@tail#9_9_9/1")                  # unterminated header: dropped
    want_a = expected_parse(heads, reads)
    for f, extra in (("a.fq", ([], [])), ("b.fq", ([b"@tail#9_9_9/1"], [b"ACGT"])), ("c.fq", ([], []))):
        want = expected_parse(heads + extra[0], reads + extra[1])
        assert parse_only([tmp_path / f], serial=True) == want
        for threads in (1, 5):
            assert parse_only([tmp_path / f], threads=threads, slice_bytes=slice_bytes) == want
    two = parse_only([tmp_path / "a.fq", tmp_path / "b.fq"], threads=4, slice_bytes=slice_bytes)
    assert two == expected_parse(heads * 2 + [b"@tail#9_9_9/1"], reads * 2 + [b"ACGT"])
    assert want_a == parse_only([tmp_path / "a.fq"], threads=3)             # default slice size


@pytest.mark.parametrize("seed", range(6))
def test_sliced_reader_equals_serial_reader_on_byte_soup(tmp_path, seed):
    """Framing is by newline count alone, so ANY byte string must frame the same way in the sliced reader and in the
    serial one: random printable soup with newlines at random places, blank lines, very long lines, no final newline."""
    rng = np.random.default_rng(1000 + seed)
    alphabet = np.frombuffer(b"ACGTN@+#/_0123456789acgt \t\r", np.uint8)
    n = int(rng.integers(20_000, 200_000))
    soup = alphabet[rng.integers(0, alphabet.size, n)].copy()
    soup[rng.random(n) < (0.002 if seed % 2 else 0.05)] = ord("\n")
    if seed % 3 == 0:
        soup[-1] = ord("\n")
    (tmp_path / "s.fq").write_bytes(soup.tobytes())
    want = parse_only([tmp_path / "s.fq"], serial=True)
    for slice_bytes in (1, 13, 4099, int(rng.integers(2, 70_000))):
        assert parse_only([tmp_path / "s.fq"], threads=4, slice_bytes=slice_bytes) == want, slice_bytes


def test_parser_packs_what_the_device_expects(tmp_path):
    """HAST_PARSE_ONLY=2: the parser runs in its default PACKED mode and the tally reads every read back out of the 2-bit
    stream.  Checksums of the base codes ((byte >> 1) & 3, kmer.h:11: lower case, IUPAC letters and '\\r' included) and the
    containN flags (upper-case 'N' only, classify.cpp:182-185) per barcode must equal a direct computation, for plain and
    gzip input, tiny slices and blocks (reads straddle words, batches, slices)."""
    rng = np.random.default_rng(5)
    letters = np.frombuffer(b"ACGTACGTACGTNnacgtRY", np.uint8)
    heads, reads = [], []
    for i in range(4000):
        heads.append(b"@r%d#%d_%d_%d/1" % (i, *rng.integers(1, 30, 3)))
        reads.append(letters[rng.integers(0, letters.size, int(rng.integers(0, 260)))].tobytes())
    cases.write_fastq(tmp_path / "p.fq", heads, reads)
    cases.write_fastq(tmp_path / "p.fq.gz", heads, reads)
    mask = (1 << 64) - 1
    n, bsum, hsum, nn = Counter(), Counter(), Counter(), Counter()
    for h, r in zip(heads, reads):
        bc = orc.parse_name(h)
        n[bc] += 1
        bsum[bc] += len(r)
        x = 0
        for c in r:
            x = (x * 5 + ((c >> 1) & 3) + 1) & mask
        hsum[bc] = (hsum[bc] + x) & mask
        nn[bc] += b"N" in r
    want = b"".join(b"%s\t%d\t%d\t%d\t%d\n" % (k, n[k], bsum[k], hsum[k], nn[k]) for k in sorted(n))

    def run_check(files, **env):
        r = run([CLASSIFY, "--hap0", "x", "--hap1", "y", "-t", "3"] + [a for f in files for a in ("--read", f)],
                env=dict(os.environ, HAST_PARSE_ONLY="2", **env))
        assert r.returncode == 0, r.stderr[-400:]
        return r.stdout

    assert run_check([tmp_path / "p.fq"]) == want
    assert run_check([tmp_path / "p.fq"], HAST_SLICE_BYTES="777") == want
    assert run_check([tmp_path / "p.fq.gz"], HAST_BLOCK_MB="0", HAST_INFLATE_THREADS="3") == want
    assert run_check([tmp_path / "p.fq"], HAST_SERIAL_READER="1", HAST_BLOCK_MB="0") == want


def test_tiny_records_fill_a_batch_before_the_block_ends(tmp_path):
    """Records far shorter than the batch buffers were sized for: the block is parsed into several batches."""
    n = 700_000
    body = b"".join(b"@#%d/1\nA\n+\nF\n" % (i % 7) for i in range(n))
    (tmp_path / "t.fq").write_bytes(body)
    want = b"".join(b"%d\t%d\t%d\n" % (j, len(range(j, n, 7)), len(range(j, n, 7))) for j in range(7))
    assert parse_only([tmp_path / "t.fq"], threads=2) == want
    assert parse_only([tmp_path / "t.fq"], threads=2, serial=True) == want


def test_over_long_read_stops_the_run_with_its_name(tmp_path):
    """a read longer than one pass of the fused kernel is reported at once, by name."""
    (tmp_path / "l.fq").write_bytes(b"@ok#1_1_1/1\nACGT\n+\nFFFF\n@giant#2_2_2/1\n" + b"ACGT" * 7000 + b"\n+\n" + b"F" * 28000 + b"\n")
    r = run([CLASSIFY, "--hap0", "x", "--hap1", "y", "--read", tmp_path / "l.fq"], env=dict(os.environ, HAST_PARSE_ONLY="1"))
    assert r.returncode == 1 and b"@giant#2_2_2/1" in r.stderr and b"28000 bases" in r.stderr


def test_output_order_is_bytewise(tmp_path):
    # std::map<std::string>: "0_0_0" < "10_1_1" < "1_2_3" because '0' (0x30) < '_' (0x5F)  
    names = [b"1_2_3", b"10_1_1", b"0_0_0", b"Z", b"a", b"1_10_1", b"1_1_10"]
    heads = [b"@r#%s/1" % n for n in names]
    cases.write_fastq(tmp_path / "a.fq", heads, [b"ACGT" * 8] * len(names))
    got = [ln.split(b"\t")[0] for ln in parse_only([tmp_path / "a.fq"]).splitlines()]
    assert got == sorted(names) and got[:3] == [b"0_0_0", b"10_1_1", b"1_10_1"]


def test_merge_result_matches_golden_and_reference():
    d = GOLDEN / "merge"
    args = json.loads((d / "cmd.txt").read_text())
    r = run([MERGE] + args, cwd=d)
    assert r.returncode == 0 and r.stdout == (d / "expected.tsv").read_bytes()
    assert run([MERGE]).returncode == 255 and run([MERGE, "-i", "a.tsv"], cwd=d).returncode == 255   # no short -i
    assert run([MERGE, "--input", "nope.tsv"], cwd=d).returncode == 255
    ref = orc.ref_binary("mergeResult")
    if ref is not None:
        assert run([ref] + args, cwd=d).stdout == r.stdout


def test_merge_result_intended_mode(tmp_path):
    (tmp_path / "a.tsv").write_text("1_2_3\t1\t1\t3\n9_9_9\t0\t4\t1\n0_0_0\t-1\t7\t7\n")
    (tmp_path / "b.tsv").write_text("1_2_3\t0\t4\t1\n")
    r = run([MERGE, "--input", tmp_path / "a.tsv", "--input", tmp_path / "b.tsv", "--intended"])
    assert r.stdout == b"0_0_0\t-1\t7\t7\n1_2_3\t0\t5\t4\n9_9_9\t0\t4\t1\n"
    r = run([MERGE, "--input", tmp_path / "a.tsv", "--input", tmp_path / "b.tsv", "--intended", "--weight1", "2"])
    assert b"1_2_3\t1\t5\t4\n" in r.stdout
    out = tmp_path / "o.tsv"
    assert orc.merge_result([tmp_path / "a.tsv", tmp_path / "b.tsv"], out, 1.0, 2.0, intended=True) == 0
    assert out.read_bytes() == r.stdout


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the reference's own CPU path, oracle/_ref when the reference compiled here, else the
    oracle port): one JSON line with the keys the driver reads, no GPU needed."""
    import sys
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "1",
                        "--warmup", "0"], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr[-600:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "impl", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "pairs/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
