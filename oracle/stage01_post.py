"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the post-processing half of
01.classify_stlfr_reads/classify_stlfr_reads.sh, used as the oracle for
hast_b200/host/partition.cpp (bin/quartering_fastq, bin/classify --split-barcodes /
--partition-reads).  Nothing in the product path imports this.

  split_barcodes(table)      classify_stlfr_reads.sh:157,160,163 -- the three awk one-liners
                             on phased.barcodes (default FS, `$2 == 0`, `$2 == 1`, `$2 == "-1"`)
  quartering(...)            quartering_fastq.awk:1-61 with `-F '#|/'` (script :181,183)

Pinned against the real thing: tests/test_partition.py runs the reference's own
quartering_fastq.awk and the script's awk one-liners (mawk 1.3.4 here) on the same inputs
whenever /root/reference is present, and compares every output file byte for byte.
"""
from __future__ import annotations

import re

_SEP = re.compile(rb"[#/]")


def _awk_strnum(tok: bytes):
    """awk 'looks like a number' for a field: returns float or None."""
    try:
        t = tok.decode("latin-1").strip(" \t")
        if not t or not re.fullmatch(r"[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", t):
            return None
        return float(t)
    except ValueError:
        return None


def split_barcodes(table: bytes):
    """-> (paternal, maternal, homozygous) file contents."""
    out = ([], [], [])
    for line in table.split(b"\n")[:-1] if table.endswith(b"\n") else table.split(b"\n"):
        f = line.split()                      # default FS: runs of blanks, leading blanks ignored
        f1 = f[0] if len(f) > 0 else b""
        f2 = f[1] if len(f) > 1 else b""
        num = _awk_strnum(f2) if len(f) > 1 else None
        # an unset $2 is the empty string AND zero: `$2 == 0` holds for a line with < 2 fields
        eq0 = (num == 0.0) if num is not None else (len(f) < 2 or f2 == b"0")
        eq1 = (num == 1.0) if num is not None else (f2 == b"1")
        if eq0:
            out[0].append(f1)
        if eq1:
            out[1].append(f1)
        if f2 == b"-1":
            out[2].append(f1)
    return tuple(b"".join(x + b"\n" for x in o) for o in out)


def _lines(data: bytes):
    if not data:
        return []
    ls = data.split(b"\n")
    if data.endswith(b"\n"):
        ls.pop()
    return ls


def quartering(paternal: bytes, maternal: bytes, homozygous: bytes, fastq: bytes, filename: bytes):
    """-> ({'nobarcode'|'paternal'|'maternal'|'homozygous': bytes} only for files that get created,
           filter_reads.log text appended by this run, stderr text)."""
    sets = []
    for lst in (paternal, maternal, homozygous):
        sets.append({_SEP.split(l)[0] for l in _lines(lst)})
    names = ["nobarcode", "paternal", "maternal", "homozygous"]
    outs = {}
    n = dict(total=0, no=0, pa=0, ma=0, ho=0)
    err = []
    log = b""
    rt = 0
    for fnr, line in enumerate(_lines(fastq), 1):
        if fnr == 1:
            log += filename + b"\n"
        if fnr % 4 == 1:
            n["total"] += 1
            f = _SEP.split(line) if line else []
            if len(f) > 1 and f[1] != b"0_0_0":
                if f[1] in sets[0]:
                    n["pa"] += 1; rt = 1
                elif f[1] in sets[1]:
                    n["ma"] += 1; rt = 2
                elif f[1] in sets[2]:
                    n["ho"] += 1; rt = 3
                else:
                    err.append(b"ERROR : unclassify barcode : " + f[1] + b"\n")
                    rt = -1
            else:
                n["no"] += 1; rt = 0
        if rt >= 0:
            outs.setdefault(names[rt], []).append(line + b"\n")
    log += (b"#Total reads                : %d \n#Reads without barcode      : %d \n"
            b"#Paternal reads             : %d \n#Maternal reads             : %d \n"
            b"#Homozygous reads           : %d \n" % (n["total"], n["no"], n["pa"], n["ma"], n["ho"]))
    return {k: b"".join(v) for k, v in outs.items()}, log, b"".join(err)
