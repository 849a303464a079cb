"""ctypes binding of the C ABI in ``include/hast_b200.h`` (libhast_b200.so).

This is the thin Python face of the engine used by the parity tests and by
``bench.py``; the production host program is the C++ ``bin/classify``.  There is
no fallback: if the shared library is missing or no CUDA device is present the
calls raise.  Reference seams are cited in the header next to each entry point
(classify.cpp / kmer.h of 01.classify_stlfr_reads).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = ROOT / "hast_b200" / "lib" / "libhast_b200.so"
HEADER_PATH = ROOT / "include" / "hast_b200.h"

HAST_OK = 0
E_ARG, E_CUDA, E_STATE, E_KMER_LINE, E_SHORT_READ, E_TABLE_FULL, E_NCCL, E_K = range(-1, -9, -1)


class HastError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"hast_b200 error {code}: {msg}")
        self.code = code


class TableInfo(C.Structure):
    _fields_ = [("k", C.c_int32), ("log2_buckets", C.c_int32), ("n_buckets", C.c_uint64),
                ("bytes", C.c_uint64), ("n_entries", C.c_uint64), ("n_displaced", C.c_uint64),
                ("n_overflow_buckets", C.c_uint64), ("size", C.c_uint64 * 2),
                ("filter_bytes", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("batches", "reads", "bases", "lookups", "reads_with_n", "reads_short",
                 "extra_probes", "kernel_launches", "h2d_bytes", "d2h_bytes", "filter_pass", "filter_loads",
                 "finish_reduce_us", "finish_d2h_us")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class KcInfo(C.Structure):
    _fields_ = [("k", C.c_int32), ("part", C.c_uint32), ("n_parts", C.c_uint32), ("n_slots", C.c_uint64),
                ("bytes", C.c_uint64), ("occupied", C.c_uint64), ("distinct", C.c_uint64 * 2), ("both", C.c_uint64),
                ("occurrences", C.c_uint64 * 2), ("windows", C.c_uint64), ("table_full", C.c_uint64)]


_vp, _u64, _u32, _i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int

# name -> (restype, argtypes): every symbol include/hast_b200.h declares
SIGNATURES = {
    "hast_abi_version": (_i32, []),
    "hast_device_count": (_i32, []),
    "hast_create": (_i32, [_i32, C.POINTER(_vp)]),
    "hast_destroy": (None, [_vp]),
    "hast_last_error": (C.c_char_p, [_vp]),
    "hast_device": (_i32, [_vp]),
    "hast_set_option": (_i32, [_vp, C.c_char_p, C.c_int64]),
    "hast_host_alloc": (_i32, [C.POINTER(_vp), C.c_size_t]),
    "hast_host_free": (_i32, [_vp]),
    "hast_table_begin": (_i32, [_vp, _i32, _u64]),
    "hast_table_add_text": (_i32, [_vp, _vp, _u64, _i32]),
    "hast_table_add_packed": (_i32, [_vp, _vp, _u64, _i32]),
    "hast_table_erase_seq": (_i32, [_vp, C.c_char_p, _u32, _vp, _vp, _u32, C.POINTER(_u32)]),
    "hast_table_info_get": (_i32, [_vp, C.POINTER(TableInfo)]),
    "hast_table_clone": (_i32, [_vp, _vp]),
    "hast_reserve_barcodes": (_i32, [_vp, _u64]),
    "hast_reset_counts": (_i32, [_vp]),
    "hast_submit_batch": (_i32, [_vp, _vp, _u64, _vp, _vp, _u32, C.POINTER(_u64)]),
    "hast_wait_copied": (_i32, [_vp, _u64]),
    "hast_submit_batch_device": (_i32, [_vp, _vp, _u64, _vp, _vp, _u32]),
    "hast_submit_batch_packed": (_i32, [_vp, _vp, _u64, _vp, _vp, _vp, _u32, C.POINTER(_u64)]),
    "hast_submit_batch_packed_device": (_i32, [_vp, _vp, _u64, _vp, _vp, _vp, _u32]),
    "hast_pack_bases": (_i32, [_vp, _u64, _vp, _u32, _vp, _vp, _i32]),
    "hast_sync": (_i32, [_vp]),
    "hast_finish": (_i32, [_vp, _vp, _u64]),
    "hast_stats_get": (_i32, [_vp, C.POINTER(Stats)]),
    "hast_counts_device_ptr": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_u64)]),
    "hast_timer_start": (_i32, [_vp]),
    "hast_timer_stop": (_i32, [_vp, C.POINTER(C.c_float)]),
    "hast_comm_init_all": (_i32, [C.POINTER(_vp), _i32]),
    "hast_comm_unique_id": (_i32, [_vp]),
    "hast_comm_init_rank": (_i32, [_vp, _i32, _i32, _vp]),
    "hast_extract_kmers": (_i32, [_vp, _vp, _u64, _vp, _u32, _vp, _u64, _vp]),
    "hast_lookup": (_i32, [_vp, _vp, _u64, _vp]),
    "hast_extract_kmers_device": (_i32, [_vp, _vp, _u64, _vp, _u32, _vp, _vp]),
    "hast_lookup_device": (_i32, [_vp, _vp, _u64, _vp]),
    "hast_gather_roofline": (_i32, [_vp, _u64, _u64, C.POINTER(C.c_float)]),
    "hast_kc_begin": (_i32, [_vp, _i32, _u64, _u32, _u32]),
    "hast_kc_add": (_i32, [_vp, _vp, _u64, _vp, _u32, _i32, C.POINTER(_u64)]),
    "hast_kc_add_device": (_i32, [_vp, _vp, _u64, _vp, _u32, _i32]),
    "hast_kc_info_get": (_i32, [_vp, C.POINTER(KcInfo)]),
    "hast_kc_histo": (_i32, [_vp, _i32, _u32, _vp]),
    "hast_kc_select": (_i32, [_vp, _i32, _u32, _u32, _i32, _vp, _u64, C.POINTER(_u64)]),
    "hast_kc_to_table": (_i32, [_vp, _vp, _u32, _u32, _u32, _u32]),
    "hast_kc_end": (_i32, [_vp]),
}

_lib = None


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """Load libhast_b200.so and attach the prototypes.  Raises if it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else Path(os.environ.get("HAST_B200_LIB", LIB_PATH))   # env: A/B builds of the library
    if not p.exists():
        raise FileNotFoundError(
            f"{p} not found: build it with `make lib` (or __graft_entry__.build()); "
            "there is no Python/CPU fallback for the classification path")
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def pack_bases(bases: np.ndarray, read_off: np.ndarray):
    """Host-side form of hast_submit_batch_packed: (uint32 words, 16 bases each, first base in the top
    two bits; has_n bit mask, one bit per read).  numpy restatement of what the C++ parser emits."""
    bases = np.ascontiguousarray(bases, np.uint8).reshape(-1)
    n = bases.size
    codes = ((bases >> 1) & 3).astype(np.uint32)
    pad = (-n) % 16
    if pad:
        codes = np.concatenate([codes, np.zeros(pad, np.uint32)])
    sh = (30 - 2 * np.arange(16, dtype=np.uint32)).astype(np.uint32)
    words = (codes.reshape(-1, 16) << sh).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    n_reads = read_off.size - 1
    is_n = np.concatenate([[0], np.cumsum(bases == ord("N"), dtype=np.int64)])
    off = read_off.astype(np.int64)
    flag = (is_n[off[1:]] - is_n[off[:-1]]) > 0
    bits = np.zeros(((n_reads + 31) // 32) * 32, np.uint8)
    bits[:n_reads] = flag
    has_n = np.packbits(bits.reshape(-1, 32)[:, ::-1], axis=1).view(">u4").astype(np.uint32).reshape(-1)
    return np.ascontiguousarray(words), np.ascontiguousarray(has_n if has_n.size else np.zeros(1, np.uint32))


def pack_bases_native(bases: np.ndarray, read_off: np.ndarray, threads: int = 1):
    """hast_pack_bases: the library's own host packer (no device needed) -> (words, has_n)"""
    lib = load_library()
    bases = np.ascontiguousarray(bases, np.uint8).reshape(-1)
    read_off = np.ascontiguousarray(read_off, np.uint32)
    n_reads = read_off.size - 1
    words = np.empty((bases.size + 15) // 16, np.uint32)
    has_n = np.zeros(max(1, (n_reads + 31) // 32), np.uint32)
    rc = lib.hast_pack_bases(_ptr(bases), bases.size, _ptr(read_off), n_reads, _ptr(words), _ptr(has_n), threads)
    if rc:
        raise HastError(rc, (lib.hast_last_error(None) or b"").decode())
    return words, has_n


class Engine:
    """One context = one GPU (the reference's MultiThread worker, classify.cpp:129-236)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self._ctx = C.c_void_p()
        self._pinned = []
        self._counts_pinned = None
        rc = self.lib.hast_create(device, C.byref(self._ctx))
        if rc:
            raise HastError(rc, (self.lib.hast_last_error(None) or b"").decode())

    # -- plumbing ---------------------------------------------------------
    def _ck(self, rc: int):
        if rc:
            raise HastError(rc, (self.lib.hast_last_error(self._ctx) or b"").decode())

    def close(self):
        if self._ctx:
            self.lib.hast_destroy(self._ctx)
            self._ctx = C.c_void_p()
            self._counts_pinned = None
            for p in self._pinned:
                self.lib.hast_host_free(p)
            self._pinned = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._ctx

    def set_option(self, name: str, value: int):
        self._ck(self.lib.hast_set_option(self._ctx, name.encode(), int(value)))

    # -- K1 ---------------------------------------------------------------
    def table_begin(self, k: int, expected_keys: int):
        self._ck(self.lib.hast_table_begin(self._ctx, k, int(expected_keys)))

    def table_add_text(self, text: bytes | np.ndarray, k: int, parent: int) -> int:
        """text = jellyfish-dump style list.  Only whole '\\n'-terminated lines are used
        (classify.cpp:41 drops an unterminated last line)."""
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else _arr(text, np.uint8)
        n_lines = buf.size // (k + 1)
        self._ck(self.lib.hast_table_add_text(self._ctx, _ptr(buf), n_lines, parent))
        return n_lines

    def table_add_packed(self, kmers: np.ndarray, parent: int):
        kmers = _arr(kmers, np.uint64)
        self._ck(self.lib.hast_table_add_packed(self._ctx, _ptr(kmers), kmers.size, parent))

    def table_erase_seq(self, seq: bytes):
        """Returns [(packed canonical k-mer, former tag)] in adaptor order."""
        cap = max(1, len(seq))
        er = np.zeros(cap, np.uint64)
        tg = np.zeros(cap, np.uint8)
        n = C.c_uint32()
        self._ck(self.lib.hast_table_erase_seq(self._ctx, seq, len(seq), _ptr(er), _ptr(tg), cap, C.byref(n)))
        return [(int(er[i]), int(tg[i])) for i in range(n.value)]

    def table_info(self) -> TableInfo:
        ti = TableInfo()
        self._ck(self.lib.hast_table_info_get(self._ctx, C.byref(ti)))
        return ti

    def table_clone_from(self, other: "Engine"):
        self._ck(self.lib.hast_table_clone(self._ctx, other._ctx))

    # -- K4 state ---------------------------------------------------------
    def reserve_barcodes(self, n: int):
        self._ck(self.lib.hast_reserve_barcodes(self._ctx, int(n)))

    def reset_counts(self):
        self._ck(self.lib.hast_reset_counts(self._ctx))

    # -- batches ----------------------------------------------------------
    def submit_batch(self, bases: np.ndarray, read_off: np.ndarray, barcode_id: np.ndarray) -> int:
        bases = _arr(bases, np.uint8).reshape(-1)
        read_off = _arr(read_off, np.uint32)
        barcode_id = _arr(barcode_id, np.uint32)
        n_reads = barcode_id.size
        assert read_off.size == n_reads + 1
        t = C.c_uint64()
        self._ck(self.lib.hast_submit_batch(self._ctx, _ptr(bases), bases.size, _ptr(read_off),
                                            _ptr(barcode_id), n_reads, C.byref(t)))
        # numpy buffers are pageable: keep them alive until the copy is done
        self._ck(self.lib.hast_wait_copied(self._ctx, t.value))
        return t.value

    def submit_batch_ptr(self, bases_ptr: int, n_bases: int, off_ptr: int, bc_ptr: int, n_reads: int) -> int:
        """Host pointers (e.g. pinned torch tensors); asynchronous, returns the ticket."""
        t = C.c_uint64()
        self._ck(self.lib.hast_submit_batch(self._ctx, bases_ptr, n_bases, off_ptr, bc_ptr, n_reads, C.byref(t)))
        return t.value

    def submit_batch_packed(self, bases: np.ndarray, read_off: np.ndarray, barcode_id: np.ndarray) -> int:
        """ASCII in, packed on the host here (pack_bases) and submitted through hast_submit_batch_packed."""
        bases = _arr(bases, np.uint8).reshape(-1)
        read_off = _arr(read_off, np.uint32)
        barcode_id = _arr(barcode_id, np.uint32)
        packed, has_n = pack_bases(bases, read_off)
        t = C.c_uint64()
        self._ck(self.lib.hast_submit_batch_packed(self._ctx, _ptr(packed), bases.size, _ptr(read_off),
                                                   _ptr(barcode_id), _ptr(has_n), barcode_id.size, C.byref(t)))
        self._ck(self.lib.hast_wait_copied(self._ctx, t.value))
        return t.value

    def submit_batch_packed_ptr(self, packed_ptr: int, n_bases: int, off_ptr: int, bc_ptr: int, hasn_ptr: int,
                                n_reads: int) -> int:
        t = C.c_uint64()
        self._ck(self.lib.hast_submit_batch_packed(self._ctx, packed_ptr, n_bases, off_ptr, bc_ptr, hasn_ptr,
                                                   n_reads, C.byref(t)))
        return t.value

    def submit_batch_packed_device(self, packed_ptr, n_bases, off_ptr, bc_ptr, hasn_ptr, n_reads):
        self._ck(self.lib.hast_submit_batch_packed_device(self._ctx, packed_ptr, n_bases, off_ptr, bc_ptr,
                                                          hasn_ptr, n_reads))

    def wait_copied(self, ticket: int):
        self._ck(self.lib.hast_wait_copied(self._ctx, ticket))

    def submit_batch_device(self, bases_ptr: int, n_bases: int, off_ptr: int, bc_ptr: int, n_reads: int):
        self._ck(self.lib.hast_submit_batch_device(self._ctx, bases_ptr, n_bases, off_ptr, bc_ptr, n_reads))

    def sync(self):
        self._ck(self.lib.hast_sync(self._ctx))

    def host_array(self, shape, dtype) -> np.ndarray:
        """numpy array over pinned host memory (hast_host_alloc); freed with the engine"""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        rc = self.lib.hast_host_alloc(C.byref(p), max(n, 1))
        if rc:
            raise HastError(rc, (self.lib.hast_last_error(None) or b"").decode())
        self._pinned.append(p)
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def finish(self, n_barcodes: int, want_counts: bool = True, pinned: bool = False) -> np.ndarray | None:
        """counts[n_barcodes][2].  pinned=True: the result lives in a pinned buffer owned by the engine that the NEXT
        pinned finish() overwrites (the read-back of 160 MB of counters then runs at PCIe speed instead of through
        the driver's pageable staging)."""
        out = None
        if want_counts and pinned:
            if self._counts_pinned is None or self._counts_pinned.shape[0] < n_barcodes:
                self._counts_pinned = self.host_array((max(n_barcodes, 1), 2), np.int32)
            out = self._counts_pinned[:n_barcodes]
        elif want_counts:
            out = np.zeros((n_barcodes, 2), np.int32)
        self._ck(self.lib.hast_finish(self._ctx, _ptr(out), n_barcodes))
        return out

    def stats(self) -> dict:
        s = Stats()
        self._ck(self.lib.hast_stats_get(self._ctx, C.byref(s)))
        return s.as_dict()

    def counts_device_ptr(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.hast_counts_device_ptr(self._ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def timer_start(self):
        self._ck(self.lib.hast_timer_start(self._ctx))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.hast_timer_stop(self._ctx, C.byref(ms)))
        return ms.value

    # -- multi-GPU ----------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.hast_comm_unique_id(buf)
        if rc:
            raise HastError(rc, (self.lib.hast_last_error(None) or b"").decode())
        return buf.raw

    def comm_init_rank(self, nranks: int, rank: int, uid: bytes):
        assert len(uid) == 128
        self._ck(self.lib.hast_comm_init_rank(self._ctx, nranks, rank, C.create_string_buffer(uid, 128)))

    # -- standalone K2 / K3 ---------------------------------------------------
    def extract_kmers(self, bases: np.ndarray, read_off: np.ndarray):
        """Returns (kmers[n_bases] with ~0 where no k-mer starts, has_n[n_reads])."""
        bases = _arr(bases, np.uint8).reshape(-1)
        read_off = _arr(read_off, np.uint32)
        n_reads = read_off.size - 1
        out = np.empty(bases.size, np.uint64)
        has_n = np.zeros(n_reads, np.uint8)
        self._ck(self.lib.hast_extract_kmers(self._ctx, _ptr(bases), bases.size, _ptr(read_off), n_reads,
                                             _ptr(out), out.size, _ptr(has_n)))
        return out, has_n

    def lookup(self, canonical: np.ndarray) -> np.ndarray:
        canonical = _arr(canonical, np.uint64)
        out = np.zeros(canonical.size, np.uint8)
        self._ck(self.lib.hast_lookup(self._ctx, _ptr(canonical), canonical.size, _ptr(out)))
        return out

    def extract_kmers_device(self, bases_ptr, n_bases, off_ptr, n_reads, out_ptr, has_n_ptr=None):
        self._ck(self.lib.hast_extract_kmers_device(self._ctx, bases_ptr, n_bases, off_ptr, n_reads, out_ptr, has_n_ptr))

    def lookup_device(self, canon_ptr, n, tags_ptr):
        self._ck(self.lib.hast_lookup_device(self._ctx, canon_ptr, n, tags_ptr))

    def gather_roofline(self, n_probes: int, span_bytes: int) -> float:
        g = C.c_float()
        self._ck(self.lib.hast_gather_roofline(self._ctx, int(n_probes), int(span_bytes), C.byref(g)))
        return g.value

    # -- stage 00: k-mer counting -----------------------------------------------
    def kc_begin(self, k: int, expected_distinct: int, part: int = 0, n_parts: int = 1):
        self._ck(self.lib.hast_kc_begin(self._ctx, k, int(expected_distinct), part, n_parts))

    def kc_add(self, bases: np.ndarray, seq_off: np.ndarray, parent: int):
        bases = _arr(bases, np.uint8).reshape(-1)
        seq_off = _arr(seq_off, np.uint32)
        t = C.c_uint64()
        self._ck(self.lib.hast_kc_add(self._ctx, _ptr(bases), bases.size, _ptr(seq_off), seq_off.size - 1, parent,
                                      C.byref(t)))
        self._ck(self.lib.hast_wait_copied(self._ctx, t.value))

    def kc_add_device(self, bases_ptr: int, n_bases: int, off_ptr: int, n_seqs: int, parent: int):
        self._ck(self.lib.hast_kc_add_device(self._ctx, bases_ptr, n_bases, off_ptr, n_seqs, parent))

    def kc_info(self) -> KcInfo:
        ki = KcInfo()
        self._ck(self.lib.hast_kc_info_get(self._ctx, C.byref(ki)))
        return ki

    def kc_histo(self, parent: int, high: int = 10000) -> np.ndarray:
        h = np.zeros(high + 2, np.uint64)
        self._ck(self.lib.hast_kc_histo(self._ctx, parent, high, _ptr(h)))
        return h

    def kc_select(self, parent: int, lower: int, upper: int, require_unique: bool = True) -> np.ndarray:
        """Sorted canonical k-mers in jellyfish's code (A0 C1 G2 T3)."""
        n = C.c_uint64()
        self._ck(self.lib.hast_kc_select(self._ctx, parent, lower, upper, int(require_unique), None, 0, C.byref(n)))
        out = np.zeros(max(1, n.value), np.uint64)
        self._ck(self.lib.hast_kc_select(self._ctx, parent, lower, upper, int(require_unique), _ptr(out), n.value,
                                         C.byref(n)))
        return out[:n.value]

    def kc_to_table(self, pl: int, pu: int, ml: int, mu: int, dst: "Engine | None" = None):
        self._ck(self.lib.hast_kc_to_table((dst or self)._ctx, self._ctx, pl, pu, ml, mu))

    def kc_end(self):
        self._ck(self.lib.hast_kc_end(self._ctx))


def kmers_to_text(kmers: np.ndarray, k: int, letters: bytes = b"ACGT") -> bytes:
    """Packed k-mers (first base in the highest used bits) -> one k-mer per line.  letters = b"ACGT" for
    jellyfish's code (hast_kc_select), b"ACTG" for kmer.h's."""
    lut = np.frombuffer(letters, np.uint8)
    km = np.ascontiguousarray(kmers, np.uint64)
    out = np.empty((km.size, k + 1), np.uint8)
    for j in range(k):
        out[:, j] = lut[((km >> np.uint64(2 * (k - 1 - j))) & np.uint64(3)).astype(np.intp)]
    out[:, k] = ord("\n")
    return out.tobytes()
