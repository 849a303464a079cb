"""ctypes face of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's classify / mergeResult
(oracle/hast_oracle.c).  It is imported by the tests as the checker and never by
the product package ``hast_b200``.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB = ORACLE_DIR / "liboracle.so"
REF_DIR = ORACLE_DIR / "_ref"

ADAPTOR_F = b"CTGTCTCTTATACACATCTTAGGAAGACAAGCACTGACGACATGA"   # classify.cpp:312
ADAPTOR_R = b"TCTGCTGAGTCGAGAACGTCTCTGTGAGCCAAGGAGTTGCTCTGG"   # classify.cpp:313

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            subprocess.run(["make", "-C", str(ORACLE_DIR), "port"], check=True, capture_output=True)
        l = C.CDLL(str(LIB))
        vp, u64, i64 = C.c_void_p, C.c_uint64, C.c_long
        l.ho_base2int.argtypes = [C.c_ubyte]
        l.ho_int2base.restype = C.c_char
        l.ho_revcomp.restype = u64
        l.ho_revcomp.argtypes = [u64, C.c_int]
        l.ho_str2kmer.restype = u64
        l.ho_str2kmer.argtypes = [C.c_char_p, C.c_int]
        l.ho_chop.restype = i64
        l.ho_chop.argtypes = [C.c_char_p, i64, C.c_int, vp]
        l.ho_kmer2str.argtypes = [u64, C.c_int, C.c_char_p]
        l.ho_parse_name.argtypes = [C.c_char_p, i64, C.POINTER(i64), C.POINTER(i64)]
        l.ho_contain_n.argtypes = [C.c_char_p, i64]
        l.ho_get_hap.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                 C.c_double, C.c_double]
        l.ho_create.restype = vp
        l.ho_destroy.argtypes = [vp]
        l.ho_error.restype = C.c_char_p
        l.ho_error.argtypes = [vp]
        l.ho_load_kmers_mem.restype = i64
        l.ho_load_kmers_mem.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_int]
        l.ho_load_kmers_file.restype = i64
        l.ho_load_kmers_file.argtypes = [vp, C.c_char_p, C.c_int]
        l.ho_load_kmers_packed.restype = i64
        l.ho_load_kmers_packed.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int]
        l.ho_init_adaptor.restype = i64
        l.ho_init_adaptor.argtypes = [vp, C.c_char_p, C.c_char_p]
        l.ho_k.argtypes = [vp]
        l.ho_set_size.restype = C.c_size_t
        l.ho_set_size.argtypes = [vp, C.c_int]
        l.ho_lookup.argtypes = [vp, u64]
        l.ho_set_weights.argtypes = [vp, C.c_double, C.c_double]
        l.ho_process_read.argtypes = [vp, C.c_char_p, i64, C.c_char_p, i64]
        l.ho_process_fastq.argtypes = [vp, C.c_char_p]
        l.ho_print_file.argtypes = [vp, C.c_char_p]
        l.ho_n_barcodes.restype = C.c_size_t
        l.ho_n_barcodes.argtypes = [vp]
        l.ho_classify_batch.restype = C.c_longlong
        l.ho_classify_batch.argtypes = [vp, vp, vp, vp, C.c_size_t, vp, C.c_size_t, C.c_int]
        l.ho_merge_result.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_float, C.c_float, C.c_int, C.c_char_p]
        _lib = l
    return _lib


class Oracle:
    """g_kmers[2] + BarcodeCache of the reference, restated (classify.cpp:27,50-64)."""

    def __init__(self):
        self.l = lib()
        self.c = self.l.ho_create()

    def close(self):
        if self.c:
            self.l.ho_destroy(self.c)
            self.c = None

    def __del__(self):
        self.close()

    def err(self) -> str:
        return self.l.ho_error(self.c).decode()

    def load_kmers(self, text: bytes, index: int) -> int:
        return self.l.ho_load_kmers_mem(self.c, text, len(text), index)

    def load_kmers_file(self, path, index: int) -> int:
        return self.l.ho_load_kmers_file(self.c, str(path).encode(), index)

    def load_kmers_packed(self, kmers: np.ndarray, k: int, index: int) -> int:
        km = np.ascontiguousarray(kmers, np.uint64)
        return self.l.ho_load_kmers_packed(self.c, km.ctypes.data, km.size, k, index)

    def init_adaptor(self, fwd: bytes = ADAPTOR_F, rev: bytes = ADAPTOR_R) -> int:
        return self.l.ho_init_adaptor(self.c, fwd, rev)

    @property
    def k(self) -> int:
        return self.l.ho_k(self.c)

    def set_size(self, i: int) -> int:
        return self.l.ho_set_size(self.c, i)

    def lookup(self, canon: int) -> int:
        return self.l.ho_lookup(self.c, int(canon))

    def lookup_many(self, canon: np.ndarray) -> np.ndarray:
        return np.fromiter((self.l.ho_lookup(self.c, int(x)) for x in canon), dtype=np.uint8, count=len(canon))

    def set_weights(self, w0: float, w1: float):
        self.l.ho_set_weights(self.c, w0, w1)

    def process_read(self, head: bytes, seq: bytes) -> int:
        return self.l.ho_process_read(self.c, head, len(head), seq, len(seq))

    def process_fastq(self, path) -> int:
        return self.l.ho_process_fastq(self.c, str(path).encode())

    def print_file(self, path):
        assert self.l.ho_print_file(self.c, str(path).encode()) == 0

    def table(self, tmp_path) -> bytes:
        p = Path(tmp_path) / "oracle.out"
        self.print_file(p)
        return p.read_bytes()

    def classify_batch(self, bases: np.ndarray, read_off: np.ndarray, barcode_id: np.ndarray,
                       n_barcodes: int, nthreads: int = 4):
        bases = np.ascontiguousarray(bases, np.uint8).reshape(-1)
        off = np.ascontiguousarray(read_off, np.uint64)
        bc = np.ascontiguousarray(barcode_id, np.uint32)
        counts = np.zeros((n_barcodes, 2), np.int32)
        lookups = self.l.ho_classify_batch(self.c, bases.ctypes.data, off.ctypes.data, bc.ctypes.data, bc.size,
                                           counts.ctypes.data, n_barcodes, nthreads)
        return counts, lookups


def chop(read: bytes, k: int) -> np.ndarray:
    n = len(read) - k + 1
    out = np.zeros(max(n, 1), np.uint64)
    r = lib().ho_chop(read, len(read), k, out.ctypes.data)
    return out[:max(r, 0)]


def parse_name(head: bytes) -> bytes:
    s, n = C.c_long(), C.c_long()
    lib().ho_parse_name(head, len(head), C.byref(s), C.byref(n))
    return head[s.value:s.value + n.value]


def kmer2str(w: int, k: int) -> bytes:
    buf = C.create_string_buffer(k + 1)
    lib().ho_kmer2str(int(w), k, buf)
    return buf.value


def merge_result(inputs, out_path, w0=1.0, w1=1.0, intended=False) -> int:
    arr = (C.c_char_p * len(inputs))(*[str(p).encode() for p in inputs])
    return lib().ho_merge_result(arr, len(inputs), w0, w1, int(intended), str(out_path).encode())


def ref_binary(name: str = "classify") -> Path | None:
    p = REF_DIR / name
    return p if p.exists() else None


def run_ref_classify(pat, mat, reads, extra=(), binary="classify_O2", threads=4) -> bytes:
    """stdout of the UNTOUCHED reference binary built into oracle/_ref."""
    exe = ref_binary(binary)
    assert exe is not None
    cmd = [str(exe), "--hap0", str(pat), "--hap1", str(mat), "--thread", str(threads)]
    for r in reads:
        cmd += ["--read", str(r)]
    cmd += list(extra)
    return subprocess.run(cmd, check=True, capture_output=True).stdout
