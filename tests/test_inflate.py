"""hast_b200/host/inflate.cpp (GzipInflater, the gzip decoder of the FASTQ readers; replaces the reference's
gzstream + zlib gzread, gzstream/gzstream.C:78-101) against zlib itself through bin/hast_gunzip: byte-identical
output on every kind of DEFLATE stream zlib can write, concatenated members, optional header fields, and a
clean error -- never a crash, never silent garbage -- on truncated or corrupted input."""
import gzip
import os
import struct
import subprocess
import zlib
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
TOOL = ROOT / "bin" / "hast_gunzip"


MODES = {"serial": [], "parallel": ["--threads", "3", "--chunk", "40000"], "parallel_big_chunks": ["--threads", "2", "--chunk", "700000"]}


def gunzip(path, mode="serial"):
    """serial = GzipInflater; parallel = ParallelGzip (inflate_par.cpp) with chunks small enough that even the small
    test streams are cut dozens of times."""
    return subprocess.run([str(TOOL)] + MODES[mode] + [str(path)], capture_output=True)


def gz_member(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=15, memlevel=8, fname=None, comment=None,
              extra=None, hcrc=False) -> bytes:
    """A gzip member written by hand around a raw zlib DEFLATE stream, so that every header field can be set."""
    c = zlib.compressobj(level, zlib.DEFLATED, -wbits, memlevel, strategy)
    raw = c.compress(data) + c.flush()
    flg = (4 if extra is not None else 0) | (8 if fname is not None else 0) | (16 if comment is not None else 0) | (2 if hcrc else 0)
    head = bytes([0x1F, 0x8B, 8, flg]) + struct.pack("<IBB", 0, 0, 255)
    if extra is not None:
        head += struct.pack("<H", len(extra)) + extra
    if fname is not None:
        head += fname + b"\0"
    if comment is not None:
        head += comment + b"\0"
    if hcrc:
        head += struct.pack("<H", zlib.crc32(head) & 0xFFFF)
    return head + raw + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data) & 0xFFFFFFFF)


def fastq_like(n_records: int, seed: int) -> bytes:
    rng = np.random.Generator(np.random.PCG64(seed))
    genome = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 50_000)]
    qual = np.frombuffer(b"FFFFFFFFFF:F,F;FFFEFGFFFF", np.uint8)
    out = []
    for i in range(n_records):
        s = int(rng.integers(0, genome.size - 100))
        out.append(b"@V300000001L1C001R%010d#%d_%d_%d/1\n" % (i, rng.integers(1, 1536), rng.integers(1, 1536), rng.integers(1, 1536)))
        out.append(genome[s:s + 100].tobytes() + b"\n+\n" + qual[rng.integers(0, qual.size, 100)].tobytes() + b"\n")
    return b"".join(out)


@pytest.fixture(scope="module")
def payloads():
    rng = np.random.Generator(np.random.PCG64(2))
    return {
        "fastq": fastq_like(40_000, 1),                                  # ~10 MB: several output chunks
        "random": rng.integers(0, 256, 3_000_000, dtype=np.uint8).tobytes(),   # incompressible: stored blocks
        "zeros": bytes(5_000_000),                                       # distance-1 runs, length-258 matches
        "period3": (b"ACG" * 400_000),                                   # distances 2..7
        "tiny": b"A",
        "empty": b"",
        "text": (b"the quick brown fox jumps over the lazy dog\n" * 50_000),
        "skewed": rng.choice(np.frombuffer(b"AAAAAAAAAAAAAAAAACGTN\n", np.uint8), 4_000_000).tobytes(),   # 1-2 bit codes
        "wide": rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes() + bytes(range(256)) * 2000,   # 15-bit codes
    }


def test_tool_built():
    assert TOOL.exists(), "bin/hast_gunzip missing: make host"


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("level", [0, 1, 4, 6, 9])
@pytest.mark.parametrize("name", ["fastq", "random", "zeros", "period3", "tiny", "empty", "text", "skewed", "wide"])
def test_matches_zlib(tmp_path, payloads, name, level, mode):
    data = payloads[name]
    p = tmp_path / "x.gz"
    p.write_bytes(gzip.compress(data, compresslevel=level))
    r = gunzip(p, mode)
    assert r.returncode == 0, r.stderr
    assert r.stdout == data
    if mode == "parallel" and name in ("fastq", "skewed") and level > 0:       # these compress to many 40 kB chunks
        assert b"block starts found" in r.stderr and int(r.stderr.split(b" block starts found")[0].split()[-1]) > 3, r.stderr


@pytest.mark.parametrize("strategy", [zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED])
@pytest.mark.parametrize("wbits,memlevel", [(15, 8), (9, 1), (12, 9)])
def test_strategies_and_windows(tmp_path, payloads, strategy, wbits, memlevel):
    """Fixed-Huffman blocks, literal-only streams, run-length matches, small windows, tiny blocks (memLevel 1)."""
    for name in ("fastq", "skewed", "zeros", "wide"):
        data = payloads[name][:1_500_000]
        p = tmp_path / "x.gz"
        p.write_bytes(gz_member(data, 6, strategy, wbits, memlevel))
        for mode in ("serial", "parallel"):
            r = gunzip(p, mode)
            assert r.returncode == 0, r.stderr
            assert r.stdout == data, (name, strategy, mode)


def test_members_headers_and_trailing_garbage(tmp_path, payloads):
    a, b, c = payloads["fastq"][:700_000], payloads["text"][:300_000], payloads["zeros"][:100_000]
    blob = (gz_member(a, fname=b"reads.fq", comment=b"made by hand", extra=b"BC\x02\x00\x10\x00", hcrc=True) +
            gz_member(b"") + gz_member(b, level=1) + gz_member(c, level=9, fname=b"x"))
    p = tmp_path / "multi.gz"
    for mode in MODES:
        p.write_bytes(blob)
        r = gunzip(p, mode)
        assert r.returncode == 0 and r.stdout == a + b + c, mode
        # zlib's gzread ignores whatever follows the last member; so do we
        p.write_bytes(blob + b"\0" * 700 + b"trailing junk")
        r = gunzip(p, mode)
        assert r.returncode == 0 and r.stdout == a + b + c, mode
    assert gzip.decompress(blob) == a + b + c


def test_many_small_members_like_bgzip(tmp_path, payloads):
    data = payloads["fastq"]
    blob = b"".join(gz_member(data[i:i + 60_000], extra=b"BC\x02\x00\xff\xff") for i in range(0, len(data), 60_000))
    p = tmp_path / "blocks.gz"
    p.write_bytes(blob)
    for mode in MODES:
        r = gunzip(p, mode)
        assert r.returncode == 0 and r.stdout == data, mode


@pytest.mark.parametrize("mode", ["serial", "parallel"])
def test_truncated_and_corrupted_input_is_an_error_not_a_crash(tmp_path, payloads, mode):
    data = payloads["fastq"][:400_000]
    blob = gzip.compress(data, compresslevel=6)
    p = tmp_path / "bad.gz"
    rng = np.random.Generator(np.random.PCG64(11))
    cuts = sorted(set([1, 2, 3, 9, 10, 11, 18, 40, len(blob) - 9, len(blob) - 8, len(blob) - 4, len(blob) - 1] +
                      rng.integers(12, len(blob) - 9, 40).tolist()))
    for cut in cuts:
        p.write_bytes(blob[:cut])
        r = gunzip(p, mode)
        assert r.returncode == 1 and b"error" in r.stderr, (cut, r.returncode, r.stderr[-200:])
        assert data.startswith(r.stdout)                           # whatever came out before the error is right
    # flipped bits: either the stream breaks or a check (CRC-32 / length) catches it; wrong data never passes
    for pos in rng.integers(10, len(blob) - 8, 60).tolist():
        bad = bytearray(blob)
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        p.write_bytes(bytes(bad))
        r = gunzip(p, mode)
        if r.returncode == 0:
            assert r.stdout == data, pos                           # a flip in dead bits (e.g. header MTIME) changes nothing
        else:
            assert r.returncode == 1 and b"error" in r.stderr
    # wrong CRC-32, wrong ISIZE
    for off in (8, 4):
        bad = bytearray(blob)
        bad[len(blob) - off] ^= 0x55
        p.write_bytes(bytes(bad))
        r = gunzip(p, mode)
        assert r.returncode == 1 and b"incorrect" in r.stderr
    p.write_bytes(b"plain text, not gzip\n")
    assert gunzip(p, mode).returncode == 1


@pytest.mark.parametrize("mode", ["serial", "parallel"])
def test_hand_made_invalid_streams(tmp_path, mode):
    head = bytes([0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, 255])
    p = tmp_path / "bad.gz"
    cases = {
        "reserved block type": bytes([0b111]),
        "stored length check": bytes([0b001, 5, 0, 0, 0]),
        "distance before start": bytes([0b011, 0b00000000 | 0x02, 0x20, 0x00]),   # fixed block, first symbol a match
    }
    for name, body in cases.items():
        p.write_bytes(head + body + bytes(16))
        r = gunzip(p, mode)
        assert r.returncode == 1, name
    p.write_bytes(bytes([0x1F, 0x8B, 7, 0, 0, 0, 0, 0, 0, 255]) + bytes(20))
    assert gunzip(p, mode).returncode == 1                               # unknown compression method
    p.write_bytes(bytes([0x1F, 0x8B, 8, 0x80, 0, 0, 0, 0, 0, 255]) + bytes(20))
    assert gunzip(p, mode).returncode == 1                               # reserved flag bits


def test_readers_use_it_and_agree_with_zlib(tmp_path, payloads):
    """bin/classify's front end (HAST_PARSE_ONLY) gives the same tally through GzipInflater and through zlib."""
    exe = ROOT / "bin" / "classify"
    if not exe.exists():
        pytest.skip("bin/classify not built")
    fq = tmp_path / "r.fq.gz"
    fq.write_bytes(gzip.compress(payloads["fastq"], compresslevel=6))
    (tmp_path / "k.mer").write_bytes(b"ACGTACGTACGTACGTACGTA\n")
    outs = []
    for force, threads in (("0", "1"), ("1", "1"), ("0", "3")):
        r = subprocess.run([str(exe), "-p", str(tmp_path / "k.mer"), "-m", str(tmp_path / "k.mer"), "-r", str(fq), "-t", "3"],
                           capture_output=True, env=dict(os.environ, HAST_PARSE_ONLY="1", HAST_ZLIB=force, HAST_INFLATE_THREADS=threads))
        assert r.returncode == 0, r.stderr[-300:]
        outs.append(r.stdout)
    assert outs[0] == outs[1] == outs[2] and len(outs[0]) > 100_000


def test_false_block_starts_are_dropped(tmp_path, payloads):
    """A gzip file stored inside stored blocks: the block finder sees perfectly valid dynamic-Huffman headers followed by
    text, but they are payload bytes of the outer stream.  The chunk before each of them runs past it on no block
    boundary, so the candidate is dropped and the output is still exact."""
    inner = gzip.compress(payloads["fastq"][:6_000_000], compresslevel=6)
    p = tmp_path / "nested.gz"
    p.write_bytes(gzip.compress(inner, compresslevel=0))
    r = gunzip(p, "parallel")
    assert r.returncode == 0, r.stderr
    assert r.stdout == inner
    dropped = int(r.stderr.split(b" dropped")[0].split()[-1])
    found = int(r.stderr.split(b" block starts found")[0].split()[-1])
    assert found > 3 and dropped == found, r.stderr
