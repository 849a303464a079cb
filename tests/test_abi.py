"""CPU: the C-ABI library loads, exports every symbol include/hast_b200.h declares,
has no CPU fallback, and does not depend on anything under oracle/."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from hast_b200 import capi

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "hast_b200.h").read_text()


def declared_symbols():
    # prototypes look like:  <type> hast_name(args);
    return sorted(set(re.findall(r"\b(hast_[a-z0-9_]+)\s*\(", HEADER)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for need in ("hast_create", "hast_table_begin", "hast_table_add_text", "hast_table_erase_seq",
                 "hast_reserve_barcodes", "hast_submit_batch", "hast_finish", "hast_extract_kmers",
                 "hast_lookup", "hast_comm_init_all", "hast_comm_init_rank"):
        assert need in syms


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/hast_b200.h but not exported"
    assert lib.hast_abi_version() == int(re.search(r"#define HAST_ABI_VERSION (\d+)", HEADER).group(1))


def test_python_binding_covers_the_header():
    assert sorted(capi.SIGNATURES) == declared_symbols()


def test_every_entry_point_cites_the_reference():
    # the seams are the reference's own function boundaries: each block of the header names them
    for token in ("classify.cpp:30-46", "classify.cpp:314-339", "classify.cpp:186-219", "classify.cpp:220-229",
                  "kmer.h:169-194", "classify.cpp:195-202", "classify.cpp:164-181"):
        assert token in HEADER


def test_error_codes_match_binding():
    for name, val in re.findall(r"#define (HAST_E_[A-Z_]+)\s+(-\d+)", HEADER):
        assert getattr(capi, name.replace("HAST_", "")) == int(val)


def test_no_cpu_fallback_without_device():
    lib = capi.load_library()
    if lib.hast_device_count() > 0:
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    assert lib.hast_create(0, C.byref(ctx)) == capi.E_CUDA
    assert b"no CPU path" in lib.hast_last_error(None)
    with pytest.raises(capi.HastError):
        capi.Engine(0)


def test_product_does_not_link_or_mention_the_oracle():
    for f in ("hast_b200/lib/libhast_b200.so", "bin/classify", "bin/mergeResult"):
        p = ROOT / f
        assert p.exists(), f"{f} not built"
        out = subprocess.run(["ldd", str(p)], capture_output=True, text=True).stdout
        assert "oracle" not in out
        assert b"liboracle" not in p.read_bytes()
    for src in list((ROOT / "hast_b200").rglob("*.py")) + list((ROOT / "hast_b200").rglob("*.c*")) + \
            list((ROOT / "hast_b200").rglob("*.h")):
        text = src.read_text()
        assert "liboracle" not in text and "hast_oracle" not in text, src


def test_sass_uses_256bit_sector_loads():
    """The probe is ONE LDG.E.256 per bucket (sm_100a); guards against a silent split into 2x128."""
    out = subprocess.run(["cuobjdump", "-sass", str(ROOT / "hast_b200/lib/libhast_b200.so")],
                         capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout or "SM100a" in out.stdout or "sm_100" in out.stdout
    assert re.search(r"LDG\.E\.[A-Z0-9.]*256", out.stdout)


def test_bulk_packer_equals_reference_packing():
    """hast_pack_bases (csrc/host_pack.cpp; what hast_submit_batch runs under the option host_pack_threads) needs no
    device: it must give the words and containN bits of capi.pack_bases -- kmer.h:11 code, MSB-first, 'N' only upper
    case (classify.cpp:182-185) -- for ragged, empty and tiny reads, on any number of threads."""
    import numpy as np
    from hast_b200 import capi
    rng = np.random.default_rng(3)
    letters = np.frombuffer(b"ACGTNacgtnRY\r", np.uint8)
    for n_reads, L in ((1000, 100), (7, 3), (50_000, 151), (1, 1), (3, 16), (40, 0)):
        lens = rng.integers(0, 2 * L + 1, n_reads) if n_reads > 1 else np.array([L])
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
        bases = letters[rng.integers(0, letters.size, int(off[-1]))].copy()
        w0, h0 = capi.pack_bases(bases, off)
        for threads in (1, 3, 8):
            w1, h1 = capi.pack_bases_native(bases, off, threads)
            assert np.array_equal(w0, w1), (n_reads, L, threads)
            m = min(h0.size, h1.size)
            assert np.array_equal(h0[:m], h1[:m]) and not h0[m:].any() and not h1[m:].any(), (n_reads, L, threads)
