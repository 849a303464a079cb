"""TEST INFRASTRUCTURE ONLY -- CPU restatement of HAST stage 00
(00.build_unshare_kmers_by_jellyfish/build_unshared_kmers.sh), the oracle for
bin/build_unshared_kmers and the hast_kc_* entry points.  Nothing in the product path
imports this.

The arithmetic of that stage lives in a third-party program the reference vendors as a
prebuilt binary, `jellyfish-linux` (jellyfish 2.3.0, `--version`), not in HAST source.  What
is restated here is jellyfish's published counting rule as the script uses it
(`count -m K -C`, `dump`, `dump -L/-U`, `histo`), and the script's own set algebra:

  sequences(path)        FASTA (multi-line records, blank lines skipped) or 4-line FASTQ, `.gz`
                         by suffix (the script pipes `zcat` for .gz input, :183-190,212-218)
  count_canonical(seqs)  `jellyfish count -m K -C`: every window of K bases all in ACGTacgt counts
                         once for its canonical form = the lexicographically smaller (A<C<G<T) of the
                         window and its reverse complement; any other byte breaks the window
  histo(counts)          `jellyfish histo` (low 1, high 10000, only non-empty bins, counts above
                         `high` collected in bin high+1)                 analysis_kmercount.sh:7-9
  find_bounds(histo)     find_bounds.awk:1-33 (first local minimum, then the global maximum after it)
  unshared(...)          build_unshared_kmers.sh:262-291: the "mix 2 copies of maternal + 1 of paternal"
                         trick is exact set algebra:
                             paternal.unique.filter = { x : cntP(x) in [PL,PU] and cntM(x) == 0 }
                             maternal.unique.filter = { x : cntM(x) in [ML,MU] and cntP(x) == 0 }

Pinned against the real thing: tests/test_stage00.py runs the reference's own script with its
vendored jellyfish binary on the same inputs whenever /root/reference is present (this
container) and compares histograms and bounds byte for byte and the two k-mer lists as sorted
line sets (jellyfish dumps in hash-table order, which depends on its -s/-t arguments); the
committed fixtures under tests/golden/stage00_* were produced by that script
(tests/golden/make_golden_stage00.py).
"""
from __future__ import annotations

import gzip
from collections import Counter

_COMP = bytes.maketrans(b"ACGT", b"TGCA")
_VALID = frozenset(b"ACGTacgt")


def sequences(path: str) -> list[bytes]:
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rb") as f:
        data = f.read()
    lines = data.split(b"\n")
    if data.endswith(b"\n"):
        lines.pop()
    out: list[bytes] = []
    first = next((ln for ln in lines if ln), b"")
    if first.startswith(b"@"):                       # FASTQ, strict 4-line records
        for i in range(0, len(lines) - 1, 4):
            out.append(lines[i + 1])
        return out
    cur: list[bytes] | None = None
    for ln in lines:                                 # FASTA
        if ln.startswith(b">"):
            if cur is not None:
                out.append(b"".join(cur))
            cur = []
        elif cur is not None and ln:
            cur.append(ln)
    if cur is not None:
        out.append(b"".join(cur))
    return out


def canonical(kmer: bytes) -> bytes:
    rc = kmer.translate(_COMP)[::-1]
    return kmer if kmer <= rc else rc


def count_canonical(seqs, k: int, counts: Counter | None = None) -> Counter:
    counts = Counter() if counts is None else counts
    for s in seqs:
        up = s.upper()
        run = 0
        for i, c in enumerate(s):
            run = run + 1 if c in _VALID else 0
            if run >= k:
                counts[canonical(up[i - k + 1:i + 1])] += 1
    return counts


def histo(counts: Counter, high: int = 10000) -> list[tuple[int, int]]:
    h: Counter = Counter()
    for c in counts.values():
        h[min(c, high + 1)] += 1
    return sorted(h.items())


def histo_text(h) -> bytes:
    return b"".join(b"%d %d\n" % (c, n) for c, n in h)


def find_bounds(h) -> dict:
    """find_bounds.awk: state 0 walks down to the first count bin that is not smaller than its
    predecessor (that line itself only flips the state), state 1 keeps the largest bin after it."""
    mn = mn_i = mx = mx_i = 0
    state = 0
    for i, c in h:
        if state == 0:
            if mn == 0 or c < mn:
                mn, mn_i = c, i
            else:
                state = 1
        else:
            if mx == 0 or c > mx:
                mx, mx_i = c, i
    up = 3 * mx_i - 2 * mn_i
    return {"MIN_INDEX": mn_i, "MAX_INDEX": mx_i, "LOWER_INDEX": mn_i + 1, "UPPER_INDEX": up - 1}


def bounds_text(b: dict) -> bytes:
    return b"".join(b"%s=%d\n" % (k.encode(), b[k]) for k in ("MIN_INDEX", "MAX_INDEX", "LOWER_INDEX", "UPPER_INDEX"))


def unshared(cp: Counter, cm: Counter, pl: int, pu: int, ml: int, mu: int):
    """-> (paternal.unique.filter.mer lines, maternal.unique.filter.mer lines), each sorted."""
    pat = sorted(x for x, c in cp.items() if pl <= c <= pu and x not in cm)
    mat = sorted(x for x, c in cm.items() if ml <= c <= mu and x not in cp)
    return pat, mat


def run(paternal_files, maternal_files, k: int = 21, pl: int = 9, pu: int = 33, ml: int = 9, mu: int = 33,
        auto_bounds: bool = False) -> dict:
    """The whole stage; returns the files the script leaves behind that later stages or people read."""
    cp, cm = Counter(), Counter()
    for p in paternal_files:
        count_canonical(sequences(p), k, cp)
    for p in maternal_files:
        count_canonical(sequences(p), k, cm)
    out = {}
    if auto_bounds:
        hp, hm = histo(cp), histo(cm)
        bp, bm = find_bounds(hp), find_bounds(hm)
        out.update({"paternal.histo": histo_text(hp), "maternal.histo": histo_text(hm),
                    "paternal.bounds.txt": bounds_text(bp), "maternal.bounds.txt": bounds_text(bm)})
        pl, pu, ml, mu = bp["LOWER_INDEX"], bp["UPPER_INDEX"], bm["LOWER_INDEX"], bm["UPPER_INDEX"]
    pat, mat = unshared(cp, cm, pl, pu, ml, mu)
    out["paternal.unique.filter.mer"] = b"".join(x + b"\n" for x in pat)
    out["maternal.unique.filter.mer"] = b"".join(x + b"\n" for x in mat)
    out["bounds"] = (pl, pu, ml, mu)
    out["distinct"] = (len(cp), len(cm))
    return out
