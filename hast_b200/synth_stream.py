"""Streamed synthetic trio for the large configurations (BASELINE.json configs[2] and configs[4]).

``synth.make_trio`` draws every read from host-side numpy generators and computes the parent-unique k-mer
sets by sorting ALL k-mers of the four haplotypes: fine up to the 100 Mbp / 20 M-pair configs[1], out of reach
for a 3.1 Gbp genome and 600 M pairs.  This module builds the same kind of data so that it scales:

* genome, SNPs, barcodes and read pairs are counter-based -- pure functions of (seed, index) -- so any rank
  generates any slice of the workload directly into device memory ("generate batches directly
  ... from (seed, pair index)"), and the CPU reproduces single pairs for the oracle side of the parity checks
  (``tools/synth_gen.h``, one header compiled for both);
* the parent-unique sets  pat = K(P1 u P2) \\ K(M1 u M2)  (and the mirror image) are still EXACT, but computed
  from the neighbourhoods of the SNPs: a window that no SNP of any haplotype touches is the ancestor's window in
  all four haplotypes and cannot be parent-unique; the windows that are touched ("candidates", ~4 G het k of
  them) go through the same sort/merge as before, and one streamed pass over the ancestor removes candidates
  that also occur, by coincidence, at an untouched position (those are in every haplotype).
  ``tests/test_synth_stream.py`` checks the result against the brute-force set difference.

Synthetic-data tooling, not part of the classification path.  Base codes: A0 C1 T2 G3 (kmer.h:11-12).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import torch

from .synth import LETTERS, ROOT, TOOLS_PATH, _canonical_kmers_torch

SYNTH_CUDA_PATH = ROOT / "hast_b200" / "lib" / "libhast_synth_cuda.so"


class SgParams(C.Structure):
    _fields_ = [("genome_len", C.c_uint64), ("n_barcodes", C.c_uint64), ("seed", C.c_uint64),
                ("read_len", C.c_uint32), ("thr_special", C.c_uint32), ("thr_n", C.c_uint32),
                ("thr_err", C.c_uint32 * 3), ("has_cdf", C.c_uint32), ("pad_", C.c_uint32)]


@dataclass
class StreamSpec:
    genome_len: int = 200_000
    het: float = 0.001            # SNP rate of each parental haplotype against the ancestor
    k: int = 21
    n_pairs: int = 20_000
    n_barcodes: int = 1_000
    read_len: int = 100
    err: float = 0.002
    n_frac: float = 0.005
    nobarcode_frac: float = 0.03
    zipf_alpha: float | None = None
    seed: int = 1
    read_seed: int = 51


def stream_config(name: str) -> StreamSpec:
    if name == "stream_tiny":
        return StreamSpec(genome_len=60_000, het=0.002, n_pairs=2_000, n_barcodes=100)
    if name == "stream_small":
        return StreamSpec(genome_len=2_000_000, het=0.0005, n_pairs=100_000, n_barcodes=5_000)
    if name == "cfg3":        # configs[2]: 3.1 Gbp trio, ~1 % parent-unique 21-mers (~62 M keys), 600 M pairs, 20 M barcodes
        return StreamSpec(genome_len=3_100_000_000, het=0.00024, n_pairs=600_000_000, n_barcodes=20_000_000)
    if name == "cfg3_100m":   # the same shape on the 100 Mbp genome (dev / small boxes)
        return StreamSpec(genome_len=100_000_000, het=0.001, n_pairs=20_000_000, n_barcodes=2_000_000)
    if name == "cfg5":        # configs[4]: human-scale table, 50 M heavy-tailed barcodes
        return StreamSpec(genome_len=3_100_000_000, het=0.00024, n_pairs=200_000_000, n_barcodes=50_000_000,
                          zipf_alpha=1.2)
    raise KeyError(name)


# ---- 64-bit mixing in torch (int64 arithmetic wraps; logical shifts spelled out) -------------------------
def _s64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


_M1, _M2, _GOLD = _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB), _s64(0x9E3779B97F4A7C15)


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    return (x >> s) & ((1 << (64 - s)) - 1)


def tmix(x: torch.Tensor) -> torch.Tensor:
    x = (x ^ _lsr(x, 30)) * _M1
    x = (x ^ _lsr(x, 27)) * _M2
    return x ^ _lsr(x, 31)


def _tag(seed: int, tag: int) -> int:
    x = (seed * 0x9E3779B97F4A7C15 + tag) & ((1 << 64) - 1)
    x ^= x >> 30; x = (x * 0xBF58476D1CE4E5B9) & ((1 << 64) - 1)
    x ^= x >> 27; x = (x * 0x94D049BB133111EB) & ((1 << 64) - 1)
    x ^= x >> 31
    return _s64(x)


def ancestor(G: int, seed: int, device) -> torch.Tensor:
    """i.i.d. uniform codes: base i = 2 bits of mix(tag ^ (i >> 5))."""
    out = torch.empty(G, dtype=torch.uint8, device=device)
    t = _tag(seed, 101)
    CH = 1 << 26
    sh = (2 * torch.arange(32, device=device, dtype=torch.int64))[None, :]
    for lo in range(0, G, CH):
        hi = min(G, lo + CH)
        w0, w1 = lo >> 5, (hi + 31) >> 5
        words = tmix(torch.arange(w0, w1, device=device, dtype=torch.int64) * _GOLD ^ t)
        codes = ((words[:, None] >> sh) & 3).to(torch.uint8).reshape(-1)
        out[lo:hi] = codes[lo - (w0 << 5): hi - (w0 << 5)]
    return out


def snp_sites(G: int, het: float, seed: int, which: int, device):
    """(sorted distinct positions, alt-code increments 1..3) of haplotype `which`."""
    n = int(round(G * het))
    j = torch.arange(n, device=device, dtype=torch.int64)
    pos = _lsr(tmix(j * _GOLD ^ _tag(seed, 110 + which)), 1) % G
    pos = torch.unique(pos)
    add = 1 + _lsr(tmix(pos * _GOLD ^ _tag(seed, 120 + which)), 1) % 3
    return pos, add.to(torch.uint8)


def _canon_at(hap: torch.Tensor, pos: torch.Tensor, k: int) -> torch.Tensor:
    """canonical k-mer (kmer.h:169-194 arithmetic) of the windows starting at `pos`."""
    fwd = torch.zeros_like(pos)
    rc = torch.zeros_like(pos)
    for j in range(k):
        c = hap[pos + j].to(torch.int64)
        fwd = (fwd << 2) | c
        rc |= (c ^ 2) << (2 * j)
    return torch.minimum(fwd, rc)


def _only_in_first(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """sorted unique a minus sorted unique b"""
    if a.numel() == 0 or b.numel() == 0:
        return a
    idx = torch.searchsorted(b, a).clamp_(max=b.numel() - 1)
    return a[b[idx] != a]


def parent_unique_sets(anc, haps: dict, sites: dict, k: int, chunk: int = 1 << 26):
    """Exact pat = K(P1 u P2) \\ K(M1 u M2) and mat = the mirror image (sorted int64 tensors), k <= 31."""
    G = anc.numel()
    dev = anc.device
    allpos = torch.unique(torch.cat([sites[n][0] for n in ("P1", "P2", "M1", "M2")]))
    # candidate window starts: every window that contains a SNP site of any haplotype
    cand = (allpos[:, None] - torch.arange(k, device=dev, dtype=torch.int64)[None, :]).reshape(-1)
    cand = torch.unique(cand[(cand >= 0) & (cand <= G - k)])
    kp = torch.unique(torch.cat([_canon_at(haps["P1"], cand, k), _canon_at(haps["P2"], cand, k)]))
    km = torch.unique(torch.cat([_canon_at(haps["M1"], cand, k), _canon_at(haps["M2"], cand, k)]))
    pat = _only_in_first(kp, km)
    mat = _only_in_first(km, kp)
    del kp, km
    # windows no SNP touches are the ancestor's in all four haplotypes: a candidate that also occurs there is shared
    surv = torch.cat([pat, mat]).sort().values
    shared = []
    for lo in range(0, G - k + 1, chunk):
        hi = min(G - k + 1, lo + chunk)
        x = _canonical_kmers_torch(anc[lo:hi + k - 1], k)
        a, b = torch.searchsorted(cand, torch.tensor([lo, hi], device=dev, dtype=torch.int64)).tolist()
        if b > a:
            x[cand[a:b] - lo] = -1
        idx = torch.searchsorted(surv, x).clamp_(max=max(surv.numel() - 1, 0))
        hit = surv[idx] == x if surv.numel() else torch.zeros_like(x, dtype=torch.bool)
        shared.append(x[hit])
        del x, idx, hit
    shared = torch.unique(torch.cat(shared)) if shared else surv[:0]
    return _only_in_first(pat, shared), _only_in_first(mat, shared)


_libs = {}


def _lib(cuda: bool):
    key = "cuda" if cuda else "cpu"
    if key not in _libs:
        path = SYNTH_CUDA_PATH if cuda else TOOLS_PATH
        if not path.exists():
            raise FileNotFoundError(f"{path} missing: run `make tools`")
        lib = C.CDLL(str(path))
        sfx = "device" if cuda else "host"
        vp, u64 = C.c_void_p, C.c_uint64
        getattr(lib, f"sg_gen_pairs_{sfx}").argtypes = [C.POINTER(SgParams), vp, vp, vp, vp, u64, u64, vp, vp]
        getattr(lib, f"sg_barcodes_{sfx}").argtypes = [C.POINTER(SgParams), vp, u64, u64, vp]
        if not cuda:
            lib.sg_barcode_names.argtypes = [u64, vp, vp, u64, vp, vp]
            lib.sg_barcode_names.restype = u64
        _libs[key] = lib
    return _libs[key]


class StreamTrio:
    """Haplotypes + parent-unique k-mer lists + a counter-based read-pair source."""

    def __init__(self, spec: StreamSpec, device: str = "cpu", sets: bool = True):
        self.spec, self.device = spec, torch.device(device)
        G, k = spec.genome_len, spec.k
        if k > 31:
            raise ValueError("StreamTrio supports k <= 31")
        dev = self.device
        anc = ancestor(G, spec.seed, dev)
        names = ("P1", "P2", "M1", "M2")
        sites = {n: snp_sites(G, spec.het, spec.seed, i, dev) for i, n in enumerate(names)}
        haps = {}
        for n in names if sets else ("P1", "M1"):
            h = anc.clone()
            pos, add = sites[n]
            h[pos] = (h[pos] + add) & 3
            haps[n] = h
        if sets:
            pat, mat = parent_unique_sets(anc, haps, sites, k)
            # file order of a k-mer list is arbitrary (jellyfish dumps in hash order): order by a hash of the key
            pat = pat[torch.argsort(tmix(pat ^ _tag(spec.seed, 131)))]
            mat = mat[torch.argsort(tmix(mat ^ _tag(spec.seed, 132)))]
            self.pat = pat.cpu().numpy().astype(np.uint64)
            self.mat = mat.cpu().numpy().astype(np.uint64)
        else:
            self.pat = self.mat = np.zeros(0, np.uint64)
        self.hap0, self.hap1 = haps["P1"], haps["M1"]        # the child = (P1, M1)
        del anc, haps, sites
        L, lam = spec.read_len, spec.read_len * spec.err
        tail = np.cumsum([np.exp(-lam) * lam ** j / np.prod(np.arange(1, j + 1)) for j in range(3)])   # P(<= j errors)
        two32 = float(1 << 32)
        self.params = SgParams(G, spec.n_barcodes, spec.read_seed, L,
                               min(int(spec.nobarcode_frac * two32), 0xFFFFFFFF), min(int(spec.n_frac * two32), 0xFFFFFFFF),
                               (C.c_uint32 * 3)(*[min(int((1.0 - t) * two32), 0xFFFFFFFF) for t in tail]),
                               1 if spec.zipf_alpha else 0, 0)
        self.cdf = None
        if spec.zipf_alpha:
            w = 1.0 / np.power(np.arange(1, spec.n_barcodes + 1, dtype=np.float64), spec.zipf_alpha)
            c = np.cumsum(w)
            c /= c[-1]
            thr = np.minimum(c * 18446744073709551616.0, 18446744073709549568.0).astype(np.uint64)
            thr[-1] = np.uint64(0xFFFFFFFFFFFFFFFF)
            self.cdf = torch.from_numpy(thr.view(np.int64)).to(dev)

    # ---- sizes -----------------------------------------------------------------------------------
    @property
    def n_barcodes(self) -> int:          # dense ids: B real barcodes + "0_0_0"
        return self.spec.n_barcodes + 1

    @property
    def is_cuda(self) -> bool:
        return self.device.type == "cuda"

    def _cdf_ptr(self):
        return self.cdf.data_ptr() if self.cdf is not None else None

    # ---- reads -------------------------------------------------------------------------------------
    def gen_pairs_into(self, lo: int, n: int, bases_ptr: int, bc_ptr: int):
        """pairs lo..lo+n-1 -> bases[2n, L] (rows 0..n-1 = r1, n..2n-1 = r2), bc[2n]; raw pointers on self.device"""
        fn = _lib(self.is_cuda).sg_gen_pairs_device if self.is_cuda else _lib(False).sg_gen_pairs_host
        rc = fn(C.byref(self.params), self._cdf_ptr(), self.hap0.data_ptr(), self.hap1.data_ptr(), None, lo, n,
                bases_ptr, bc_ptr)
        if rc:
            raise RuntimeError(f"synth_gen failed ({rc})")

    def gen_pairs(self, lo: int, n: int):
        L = self.spec.read_len
        bases = torch.empty((2 * n, L), dtype=torch.uint8, device=self.device)
        bc = torch.empty(2 * n, dtype=torch.int32, device=self.device)
        self.gen_pairs_into(lo, n, bases.data_ptr(), bc.data_ptr())
        return bases, bc

    def gen_pairs_idx(self, idx):
        """the pairs listed in idx (any order) -> (bases [2m, L] uint8, bc [2m]) as numpy arrays"""
        idx_t = torch.as_tensor(np.ascontiguousarray(idx, dtype=np.int64)).to(self.device)
        m, L = idx_t.numel(), self.spec.read_len
        bases = torch.empty((2 * m, L), dtype=torch.uint8, device=self.device)
        bc = torch.empty(2 * m, dtype=torch.int32, device=self.device)
        fn = _lib(self.is_cuda).sg_gen_pairs_device if self.is_cuda else _lib(False).sg_gen_pairs_host
        rc = fn(C.byref(self.params), self._cdf_ptr(), self.hap0.data_ptr(), self.hap1.data_ptr(), idx_t.data_ptr(),
                0, m, bases.data_ptr(), bc.data_ptr())
        if rc:
            raise RuntimeError(f"synth_gen failed ({rc})")
        if self.is_cuda:
            torch.cuda.synchronize(self.device)
        return bases.cpu().numpy(), bc.cpu().numpy().view(np.uint32)

    def barcodes_of(self, lo: int, n: int) -> torch.Tensor:
        """barcode id of pairs lo..lo+n-1 (int32 tensor on self.device)"""
        bc = torch.empty(n, dtype=torch.int32, device=self.device)
        fn = _lib(self.is_cuda).sg_barcodes_device if self.is_cuda else _lib(False).sg_barcodes_host
        rc = fn(C.byref(self.params), self._cdf_ptr(), lo, n, bc.data_ptr())
        if rc:
            raise RuntimeError(f"synth_gen failed ({rc})")
        return bc

    def pairs_of_barcodes(self, ids, lo: int = 0, hi: int | None = None, chunk: int = 1 << 26) -> np.ndarray:
        """indices of ALL pairs in [lo, hi) whose barcode id is in `ids` (a barcode-complete subsample)"""
        hi = self.spec.n_pairs if hi is None else hi
        want = torch.as_tensor(np.sort(np.asarray(ids, dtype=np.int64))).to(self.device)
        out = []
        for a in range(lo, hi, chunk):
            n = min(chunk, hi - a)
            bc = self.barcodes_of(a, n).to(torch.int64)
            pos = torch.searchsorted(want, bc).clamp_(max=want.numel() - 1)
            sel = (want[pos] == bc).nonzero().squeeze(1) + a
            out.append(sel.cpu().numpy())
        return np.concatenate(out) if out else np.zeros(0, np.int64)

    # ---- names / files ---------------------------------------------------------------------------------
    def barcode_name_blob(self):
        """(blob of NUL-terminated names, uint64 offsets) for ids 0..B; id B = "0_0_0".  Names are distinct
        a_b_c triples with a, b, c in 1..1536 (the stLFR barcode space)."""
        B = self.spec.n_barcodes
        off = np.zeros(B + 2, np.uint64)
        blob = np.zeros((B + 1) * 16, np.uint8)
        used = _lib(False).sg_barcode_names(B, blob.ctypes.data, off.ctypes.data, 0, None, None)
        return blob[:used].tobytes(), off[:B + 1]

    def kmer_text(self, which: int) -> bytes:
        """one k-mer per line, orientation chosen by a hash of the k-mer (the reference canonicalises on load)"""
        from .synth import revcomp_packed
        km = self.pat if which == 0 else self.mat
        k = self.spec.k
        flip = (tmix(torch.from_numpy(km.view(np.int64)) ^ 77).numpy() & 1).astype(bool)
        km = np.where(flip, revcomp_packed(km, k), km)
        t = torch.from_numpy(km.view(np.int64)).to(self.device)
        lut = torch.from_numpy(LETTERS.copy()).to(self.device)
        out = torch.empty((t.numel(), k + 1), dtype=torch.uint8, device=self.device)
        for j in range(k):
            out[:, j] = lut[(t >> (2 * (k - 1 - j))) & 3]
        out[:, k] = ord("\n")
        return out.cpu().numpy().tobytes()

    def write_kmer_lists(self, outdir):
        outdir = Path(outdir)
        outdir.mkdir(parents=True, exist_ok=True)
        paths = (str(outdir / "paternal.unique.filter.mer"), str(outdir / "maternal.unique.filter.mer"))
        for i, p in enumerate(paths):
            with open(p, "wb") as f:
                f.write(self.kmer_text(i))
        return paths

    def write_fastq(self, outdir, pair_idx=None, lo: int = 0, hi: int | None = None, gz: bool = False,
                    stem: str = "child", chunk: int = 1 << 21, names=None):
        """child.r1.fq[.gz] / child.r2.fq[.gz] of pairs [lo, hi) or of the listed pair indices."""
        from .synth import FastqWriter
        outdir = Path(outdir)
        outdir.mkdir(parents=True, exist_ok=True)
        blob, name_off = names if names is not None else self.barcode_name_blob()
        L = self.spec.read_len
        paths = [str(outdir / f"{stem}.r{m}.fq{'.gz' if gz else ''}") for m in (1, 2)]
        level = (6 if gz is True else int(gz)) if gz else 0
        if pair_idx is None:
            hi = self.spec.n_pairs if hi is None else hi
            pieces = ((np.arange(a, min(hi, a + chunk), dtype=np.uint64)) for a in range(lo, hi, chunk))
        else:
            pair_idx = np.asarray(pair_idx, dtype=np.uint64)
            pieces = (pair_idx[a:a + chunk] for a in range(0, pair_idx.size, chunk))
        with FastqWriter(paths[0], level) as w1, FastqWriter(paths[1], level) as w2:
            for idx in pieces:
                m = idx.size
                if not m:
                    continue
                bases, bc = self.gen_pairs_idx(idx.astype(np.int64))
                off = np.arange(m + 1, dtype=np.uint64) * np.uint64(L)
                w1.add(bases[:m], off, blob, name_off, bc[:m], idx, 1)
                w2.add(bases[m:], off, blob, name_off, bc[:m], idx, 2)
        return tuple(paths)
