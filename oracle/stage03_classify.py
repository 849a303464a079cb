"""TEST INFRASTRUCTURE ONLY -- CPU restatement of HAST stage 03's per-sequence classifier
(03.mkoutput_by_fabulous2.0/src_main/classify.cpp), the oracle for bin/classify_seq.
Nothing in the product path imports this.

Pinned against the real thing: oracle/Makefile compiles the untouched reference source into
oracle/_ref/classify03, and tests/test_stage03.py compares this restatement, the reference binary
(when present) and bin/classify_seq on the same inputs, byte for byte.
"""
from __future__ import annotations

_COMP = {ord("A"): "T", ord("a"): "T", ord("G"): "C", ord("g"): "C", ord("C"): "G", ord("c"): "G",
         ord("T"): "A", ord("t"): "A", ord("N"): "N", ord("n"): "N"}          # classify.cpp:25-36


def _revcomp(s: bytes) -> bytes:                                              # :37-42, unknown -> '\0'
    return "".join(_COMP.get(c, "\0") for c in reversed(s)).encode("latin-1")


def _getlines(data: bytes):
    """Lines as `while(!getline(...).eof())` sees them: the piece after the last '\\n' is dropped."""
    parts = data.split(b"\n")
    return parts[:-1], parts[-1]


def load_kmers(data: bytes, index: int, k: int | None):
    """-> (set of byte strings, line count, k)   classify.cpp:52-72"""
    lines, tail = _getlines(data)
    if index == 0 and not lines:                     # the first getline of list 0 is not eof-checked
        lines = [tail]
    if index == 0:
        k = len(lines[0])
    s = set()
    for ln in lines:
        s.add(ln)
        s.add(_revcomp(ln))
    return s, len(lines), k


def _fmt(name: bytes, hc) -> bytes:                                           # PrintOutput :104-135
    best = second = 0.0
    hap = b""
    for i in range(2):
        if 0 < hc[i] < best and hc[i] > second:
            second = hc[i]
        if hc[i] > 0 and hc[i] > best:
            hap = b"haplotype%d" % i
            second = best
            best = hc[i]
    if second == 0 and best != 0:
        return b"%s\t%s\t%s\n" % (name, hap, b"%0.6f" % best)
    if best == 0 and second == 0:
        return name + b"\tambiguous\t0.0\n"
    if best / second > 1:
        return b"%s\t%s\t%s\n" % (name, hap, b"%0.6f" % best)
    return b"%s\tambiguous\t%s\n" % (name, b"%0.6f" % best)


def classify(hap0: bytes, hap1: bytes, reads: bytes, fmt: str = "fasta") -> bytes:
    s0, n0, k = load_kmers(hap0, 0, None)
    s1, n1, _ = load_kmers(hap1, 1, k)
    recs = []
    lines, _tail = _getlines(reads)
    if fmt == "fasta":                                                        # processFasta :272-300
        head, seq, n = b"", b"", 0
        for ln in lines:
            if not ln:
                continue
            assert ln[:1] not in (b"@", b"+")
            if ln[:1] == b">":
                if n > 0:
                    recs.append((head, seq))
                head, seq = ln, b""
                n += 1
            else:
                seq += ln
        recs.append((head, seq))
    else:                                                                     # processFastq :250-268
        i = 0
        while i < len(lines):
            head = lines[i]
            seq = lines[i + 1] if i + 1 < len(lines) else (_tail if i + 1 == len(lines) else b"")
            recs.append((head, seq))
            i += 4
    out = []
    for head, seq in recs:
        c = [0, 0]
        for i in range(len(seq) - k + 1):                                     # :210-214
            w = seq[i:i + k]
            if w in s0:
                c[0] += 1
            if w in s1:
                c[1] += 1
        out.append(_fmt(head[1:], (c[0] / n0, c[1] / n1)))
    return b"".join(out)
