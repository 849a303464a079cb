"""Seeded adversarial inputs shared by the CPU (oracle/golden) and GPU parity tests.

The cases follow the list SURVEY.md section 8(c) says the reference's behaviour
was verified on: overlapping parent sets, duplicate and reverse-complement
duplicate k-mers, adaptor erasure, N reads, lowercase reads, odd headers,
ragged read lengths, a last k-mer line without newline.
"""
from __future__ import annotations

import numpy as np

LET = np.frombuffer(b"ACTG", np.uint8)
COMP = {65: 84, 67: 71, 84: 65, 71: 67}

ADAPTOR_F = b"CTGTCTCTTATACACATCTTAGGAAGACAAGCACTGACGACATGA"
ADAPTOR_R = b"TCTGCTGAGTCGAGAACGTCTCTGTGAGCCAAGGAGTTGCTCTGG"


def revcomp_ascii(s: bytes) -> bytes:
    return bytes(COMP[c] for c in reversed(s))


def low_complexity_genome(rng, n: int = 4000) -> bytes:
    """Homopolymer runs and short tandem repeats with a few random stretches in between: every k-mer
    shares its few distinct m-mers (hence its minimizer) with many others, and most windows are their
    neighbours' reverse complement or a rotation of them."""
    out = bytearray()
    while len(out) < n:
        u = rng.random()
        if u < 0.25:
            out += bytes([int(LET[rng.integers(0, 4)])]) * int(rng.integers(20, 120))
        elif u < 0.8:
            unit = LET[rng.integers(0, 4, int(rng.integers(2, 7)))].tobytes()
            out += unit * int(rng.integers(8, 40))
        else:
            out += LET[rng.integers(0, 4, int(rng.integers(5, 60)))].tobytes()
    return bytes(out[:n])


def adversarial_case(k: int, n_reads: int, seed: int, min_len: int | None = None, max_len: int = 160,
                     n_barcodes: int = 37, with_adaptor: bool = True, low_complexity: bool = False):
    """Returns dict(pat_text, mat_text, reads=[bytes], heads=[bytes], bc_ids, bc_names).

    The k-mer lists are drawn from the reads themselves (so hits are common),
    partly in reverse-complement orientation, partly duplicated, partly shared
    by both parents; adaptor k-mers are planted in both lists and in the reads.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    min_len = k if min_len is None else min_len
    genome = low_complexity_genome(rng) if low_complexity else LET[rng.integers(0, 4, 4000)].tobytes()
    reads = []
    for i in range(n_reads):
        L = int(rng.integers(min_len, max_len + 1))
        s = int(rng.integers(0, len(genome) - L))
        r = bytearray(genome[s:s + L])
        if rng.random() < 0.3:
            r = bytearray(revcomp_ascii(bytes(r)))
        for _ in range(int(rng.integers(0, 3))):                      # substitutions
            r[int(rng.integers(0, L))] = int(LET[rng.integers(0, 4)])
        u = rng.random()
        if u < 0.06:
            r[int(rng.integers(0, L))] = ord("N")                     # whole read skipped
        elif u < 0.10:
            r = bytearray(bytes(r).lower())                           # lowercase works, incl. 'n' -> G
            if rng.random() < 0.5:
                r[int(rng.integers(0, L))] = ord("n")
        elif u < 0.13:
            r[int(rng.integers(0, L))] = ord("R")                     # IUPAC byte: packed via (c&6)>>1
        elif u < 0.16 and with_adaptor and L >= 45:
            a = ADAPTOR_F if rng.random() < 0.5 else ADAPTOR_R
            p = int(rng.integers(0, L - 45 + 1))
            r[p:p + 45] = a
        reads.append(bytes(r))

    def sample_kmers(n):
        out = []
        for _ in range(n):
            r = reads[int(rng.integers(0, n_reads))]
            if b"N" in r or len(r) < k:
                continue
            p = int(rng.integers(0, len(r) - k + 1))
            km = r[p:p + k].upper().replace(b"R", b"C").replace(b"N", b"G")   # as (c&6)>>1 packs them
            if rng.random() < 0.5:
                km = revcomp_ascii(km)
            out.append(km)
        return out

    shared = sample_kmers(max(4, n_reads // 4))
    pat = sample_kmers(n_reads) + shared
    mat = sample_kmers(n_reads) + shared
    pat += pat[:5] + [revcomp_ascii(x) for x in pat[5:10]]                # duplicates / RC duplicates
    decoys = [LET[rng.integers(0, 4, k)].tobytes() for _ in range(n_reads)]
    pat += decoys[: n_reads // 2]
    mat += decoys[n_reads // 2:]
    if with_adaptor and k <= 45:
        pat += [ADAPTOR_F[i:i + k] for i in range(0, 45 - k + 1, 3)]
        mat += [revcomp_ascii(ADAPTOR_R[i:i + k]) for i in range(1, 45 - k + 1, 4)]
        mat += [ADAPTOR_F[0:k]]
    rng.shuffle(pat)
    rng.shuffle(mat)
    pat_text = b"".join(x + b"\n" for x in pat)
    mat_text = b"".join(x + b"\n" for x in mat)

    names = [b"0_0_0", b"0", b"lib2_7_8_9", b""] + [b"%d_%d_%d" % tuple(rng.integers(1, 1537, 3)) for _ in range(n_barcodes)]
    names = list(dict.fromkeys(names))
    bc_ids = rng.integers(0, len(names), n_reads).astype(np.uint32)
    heads = []
    for i in range(n_reads):
        nm = names[bc_ids[i]]
        if nm == b"":
            heads.append(b"@V300R%07d#/1" % i)                        # empty barcode
        else:
            heads.append(b"@V300#junk/x_R%07d#%s/%d" % (i, nm, 1 + (i & 1)))   # LAST '#' and LAST '/' win
    return dict(k=k, pat_text=pat_text, mat_text=mat_text, reads=reads, heads=heads, bc_ids=bc_ids,
                bc_names=names)


def flatten(reads):
    lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
    off = np.zeros(len(reads) + 1, np.uint64)
    np.cumsum(lens, out=off[1:])
    bases = np.frombuffer(b"".join(reads), np.uint8)
    return bases, off


def write_fastq(path, heads, reads, crlf=False, trailing_newline=True):
    nl = b"\r\n" if crlf else b"\n"
    recs = [h + nl + r + nl + b"+" + nl + b"F" * len(r) for h, r in zip(heads, reads)]
    data = nl.join(recs) + (nl if trailing_newline else b"")
    if str(path).endswith(".gz"):
        import gzip
        half = len(recs) // 2                                         # two concatenated gzip members
        a = nl.join(recs[:half]) + nl
        b = nl.join(recs[half:]) + (nl if trailing_newline else b"")
        with open(path, "wb") as f:
            f.write(gzip.compress(a) + gzip.compress(b))
    else:
        with open(path, "wb") as f:
            f.write(data)
