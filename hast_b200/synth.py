"""Deterministic synthetic trio generator.

Produces exactly what stage 01 of HAST consumes: two parent-unique k-mer lists
(format of 00.build_unshare_kmers_by_jellyfish/build_unshared_kmers.sh:290-291,
one k-mer per line) and stLFR read pairs whose barcode sits in the read name
(``@...#a_b_c/1``, classify.cpp:109-111).  It is a data tool, not part of the
classification path; the same arrays feed the device interface, the FASTQ files
fed to ``bin/classify`` and to the reference binary, and the CPU oracle.

All random draws come from numpy ``Generator(PCG64(seed))`` on the host, so the
output is identical whether the heavy array work (k-mer set algebra, read
gathers) runs in numpy or, with ``device="cuda"``, in torch on the GPU.

Base codes follow the reference everywhere: A0 C1 T2 G3 (kmer.h:11-12).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

LETTERS = np.frombuffer(b"ACTG", dtype=np.uint8)      # kmer.h:12 int2base
ROOT = Path(__file__).resolve().parent.parent
TOOLS_PATH = ROOT / "hast_b200" / "lib" / "libhast_tools.so"


@dataclass
class TrioSpec:
    genome_len: int = 200_000
    het: float = 0.001            # SNP rate of each parental haplotype against the ancestor
    k: int = 21
    n_pairs: int = 20_000
    n_barcodes: int = 1_000
    read_len: int = 100
    err: float = 0.002            # substitution errors per base
    n_frac: float = 0.005         # reads that get one 'N'
    nobarcode_frac: float = 0.03  # pairs labelled 0_0_0
    zipf_alpha: float | None = None   # heavy-tailed reads-per-barcode (config 5)
    decoy_kmers: int = 0          # extra random k-mers per parent list (inflates the table)
    lowercase_frac: float = 0.0   # reads written in lowercase (kmer.h:11 is case-insensitive)
    seed: int = 1
    read_seed: int = 51           # reads only: ranks of a multi-GPU run draw different reads of one trio


@dataclass
class Trio:
    spec: TrioSpec
    pat: np.ndarray               # uint64 canonical packed parent-unique k-mers (hap0)
    mat: np.ndarray               # (hap1)
    r1: object                    # uint8 [n_pairs, L] ASCII, numpy or torch(cuda)
    r2: object
    pair_bc: np.ndarray           # uint32 [n_pairs] dense barcode id
    bc_triples: np.ndarray        # int64 [n_barcodes(+1), 3]; a row of zeros = "0_0_0"
    bc_hap: np.ndarray            # int8 [n_barcodes] true haplotype of each barcode (0 pat / 1 mat)
    _names: list | None = field(default=None, repr=False)

    @property
    def n_barcodes(self) -> int:
        return int(self.bc_triples.shape[0])

    def barcode_names(self) -> list[bytes]:
        if self._names is None:
            t = self.bc_triples
            self._names = [b"%d_%d_%d" % (a, b, c) for a, b, c in t.tolist()]
        return self._names

    # ---- k-mer lists ----------------------------------------------------
    def kmer_text(self, which: int, orient_seed: int = 77) -> bytes:
        """One k-mer per line.  Orientation is randomised (the reference canonicalises
        on load, classify.cpp:38,42), which exercises the device-side canonical step."""
        km = self.pat if which == 0 else self.mat
        k = self.spec.k
        rng = np.random.Generator(np.random.PCG64(orient_seed + which))
        flip = rng.random(km.size) < 0.5
        km = np.where(flip, revcomp_packed(km, k), km)
        out = np.empty((km.size, k + 1), np.uint8)
        for j in range(k):
            out[:, j] = LETTERS[((km >> np.uint64(2 * (k - 1 - j))) & np.uint64(3)).astype(np.intp)]
        out[:, k] = ord("\n")
        return out.tobytes()

    def write_kmer_lists(self, outdir) -> tuple[str, str]:
        outdir = Path(outdir)
        outdir.mkdir(parents=True, exist_ok=True)
        paths = (str(outdir / "paternal.unique.filter.mer"), str(outdir / "maternal.unique.filter.mer"))
        for i, p in enumerate(paths):
            with open(p, "wb") as f:
                f.write(self.kmer_text(i))
        return paths

    # ---- reads ------------------------------------------------------------
    def _np(self, x) -> np.ndarray:
        return x if isinstance(x, np.ndarray) else x.cpu().numpy()

    def batch(self, lo: int = 0, hi: int | None = None):
        """(bases, read_off, barcode_id) of pairs [lo, hi): r1 reads then r2 reads."""
        hi = self.spec.n_pairs if hi is None else hi
        L = self.spec.read_len
        a, b = self._np(self.r1[lo:hi]), self._np(self.r2[lo:hi])
        bases = np.concatenate([a.reshape(-1), b.reshape(-1)])
        n = 2 * (hi - lo)
        off = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
        bc = np.concatenate([self.pair_bc[lo:hi], self.pair_bc[lo:hi]]).astype(np.uint32)
        return bases, off, bc

    def write_fastq(self, outdir, gz: bool = False, lo: int = 0, hi: int | None = None,
                    stem: str = "child") -> tuple[str, str]:
        """child.r1.fq[.gz] / child.r2.fq[.gz] (names must contain r1/r2, HAST.sh:31,36)."""
        hi = self.spec.n_pairs if hi is None else hi
        outdir = Path(outdir)
        outdir.mkdir(parents=True, exist_ok=True)
        names = self.barcode_names()
        blob = b"\0".join(names) + b"\0"
        name_off = np.zeros(len(names), np.uint64)
        np.cumsum([len(x) + 1 for x in names[:-1]], out=name_off[1:])
        L = self.spec.read_len
        n = hi - lo
        off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
        read_no = np.arange(lo, hi, dtype=np.uint64)
        bc = np.ascontiguousarray(self.pair_bc[lo:hi], dtype=np.uint32)
        paths = []
        CH = 1 << 21
        for mate, arr in ((1, self.r1), (2, self.r2)):
            p = str(outdir / f"{stem}.r{mate}.fq{'.gz' if gz else ''}")
            with FastqWriter(p, (6 if gz is True else int(gz)) if gz else 0) as w:
                for a in range(0, n, CH):
                    b = min(n, a + CH)
                    seqs = np.ascontiguousarray(self._np(arr[lo + a:lo + b])).reshape(-1)
                    w.add(seqs, off[:b - a + 1], blob, name_off, bc[a:b], read_no[a:b], mate)
            paths.append(p)
        return tuple(paths)


def _tools():
    if not TOOLS_PATH.exists():
        raise FileNotFoundError(f"{TOOLS_PATH} missing: run `make tools`")
    lib = C.CDLL(str(TOOLS_PATH))
    lib.ff_write_fastq.restype = C.c_int
    lib.ff_write_fastq.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_char_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.ff_open.restype = C.c_void_p
    lib.ff_open.argtypes = [C.c_char_p, C.c_int]
    lib.ff_add.restype = C.c_int
    lib.ff_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_char_p, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_int]
    lib.ff_close.restype = C.c_int
    lib.ff_close.argtypes = [C.c_void_p]
    return lib


class FastqWriter:
    """Multi-threaded FASTQ writer of the synthetic generators (tools/fastq_fmt.c ff_open/ff_add/ff_close): plain text,
    or -- gz_level 1..9 -- ONE gzip member deflated on all cores."""

    def __init__(self, path, gz_level: int = 0):
        self.lib = _tools()
        self.path = str(path)
        self.h = self.lib.ff_open(self.path.encode(), int(gz_level))
        if not self.h:
            raise OSError(f"cannot write {path}")

    def add(self, seqs: np.ndarray, off: np.ndarray, blob: bytes, name_off: np.ndarray, bc: np.ndarray,
            read_no: np.ndarray, mate: int):
        seqs = np.ascontiguousarray(seqs, np.uint8).reshape(-1)
        off = np.ascontiguousarray(off, np.uint64)
        bc = np.ascontiguousarray(bc, np.uint32)
        read_no = np.ascontiguousarray(read_no, np.uint64)
        name_off = np.ascontiguousarray(name_off, np.uint64)
        if self.lib.ff_add(self.h, seqs.ctypes.data, off.ctypes.data, bc.size, blob, name_off.ctypes.data,
                           bc.ctypes.data, read_no.ctypes.data, mate):
            raise OSError(f"cannot write {self.path}")

    def close(self):
        if self.h:
            h, self.h = self.h, None
            if self.lib.ff_close(h):
                raise OSError(f"cannot write {self.path}")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# --------------------------------------------------------------------------
# packed k-mer helpers (numpy)
# --------------------------------------------------------------------------
def revcomp_packed(km: np.ndarray, k: int) -> np.ndarray:
    """kmer.h:196-223 on uint64 arrays."""
    w = km.astype(np.uint64) ^ np.uint64(0xAAAAAAAAAAAAAAAA)
    for sh, m in ((2, 0x3333333333333333), (4, 0x0F0F0F0F0F0F0F0F), (8, 0x00FF00FF00FF00FF),
                  (16, 0x0000FFFF0000FFFF), (32, 0x00000000FFFFFFFF)):
        m = np.uint64(m)
        w = ((w & m) << np.uint64(sh)) | ((w >> np.uint64(sh)) & m)
    if k < 32:
        w = w >> np.uint64(64 - 2 * k)
    return w


def canonical_kmers_np(codes: np.ndarray, k: int) -> np.ndarray:
    """All canonical k-mers of a code sequence (kmer.h:169-194 restated as array ops)."""
    n = codes.size - k + 1
    fwd = np.zeros(n, np.uint64)
    rc = np.zeros(n, np.uint64)
    for j in range(k):
        c = codes[j:j + n].astype(np.uint64)
        fwd = (fwd << np.uint64(2)) | c
        rc |= (c ^ np.uint64(2)) << np.uint64(2 * j)
    return np.minimum(fwd, rc)


def _canonical_kmers_torch(codes, k: int):
    import torch
    n = codes.numel() - k + 1
    fwd = torch.zeros(n, dtype=torch.int64, device=codes.device)
    rc = torch.zeros(n, dtype=torch.int64, device=codes.device)
    for j in range(k):
        c = codes[j:j + n].to(torch.int64)
        fwd = (fwd << 2) | c
        rc |= (c ^ 2) << (2 * j)
    return torch.minimum(fwd, rc)          # k <= 31: values < 2^62, signed order == unsigned order


def _unique_sets(p_list, m_list, device):
    """pat = K(P1 u P2) \\ K(M1 u M2) and the symmetric set, exact."""
    if device == "cpu":
        p = np.unique(np.concatenate(p_list))
        m = np.unique(np.concatenate(m_list))
        return np.setdiff1d(p, m, assume_unique=True), np.setdiff1d(m, p, assume_unique=True)
    import torch
    p = torch.unique(torch.cat(p_list))
    m = torch.unique(torch.cat(m_list))
    both = torch.cat([p << 1, (m << 1) | 1])        # k <= 31 keeps this inside int64
    both, _ = torch.sort(both)
    key = both >> 1
    same_next = torch.zeros_like(key, dtype=torch.bool)
    same_next[:-1] = key[:-1] == key[1:]
    same_prev = torch.zeros_like(key, dtype=torch.bool)
    same_prev[1:] = same_next[:-1]
    alone = ~(same_next | same_prev)
    is_m = (both & 1).bool()
    pat = key[alone & ~is_m].cpu().numpy().astype(np.uint64)
    mat = key[alone & is_m].cpu().numpy().astype(np.uint64)
    return pat, mat


def _haplotypes(spec: TrioSpec) -> dict:
    """Ancestor genome (i.i.d. uniform) and the four parental haplotypes P1 P2 M1 M2, each the ancestor
    with independent SNPs at rate spec.het.  Codes A0 C1 T2 G3."""
    G = spec.genome_len
    rng = lambda s: np.random.Generator(np.random.PCG64(spec.seed * 1000 + s))
    anc = rng(1).integers(0, 4, G, dtype=np.uint8)
    haps = {}
    for name, s in (("P1", 11), ("P2", 12), ("M1", 21), ("M2", 22)):
        g = rng(s)
        n_snp = g.binomial(G, spec.het)
        pos = g.integers(0, G, n_snp)
        h = anc.copy()
        h[pos] = (h[pos] + g.integers(1, 4, n_snp, dtype=np.uint8)) & 3
        haps[name] = h
    return haps


def parent_reads(spec: TrioSpec, coverage: float, read_len: int = 100, err: float = 0.003, n_frac: float = 0.01,
                 lowercase_frac: float = 0.01, seed: int = 61) -> dict:
    """Whole-genome shotgun reads of the two parents, the input of HAST stage 00
    (00.build_unshare_kmers_by_jellyfish/build_unshared_kmers.sh --paternal / --maternal): each parent is
    sequenced to `coverage` x in total, half from each of its haplotypes, either strand, substitution errors at
    rate `err`, a fraction of reads with one 'N', a fraction in lower case.
    -> {"paternal": uint8 [n, L] ASCII, "maternal": ...}"""
    haps = _haplotypes(spec)
    G, L = spec.genome_len, read_len
    out = {}
    for pi, (parent, names) in enumerate((("paternal", ("P1", "P2")), ("maternal", ("M1", "M2")))):
        g = np.random.Generator(np.random.PCG64(spec.seed * 1000 + seed + pi))
        n = int(coverage * G / L)
        which = g.integers(0, 2, n)
        start = g.integers(0, G - L + 1, n)
        codes = np.empty((n, L), np.uint8)
        for h in (0, 1):
            sel = np.nonzero(which == h)[0]
            codes[sel] = np.lib.stride_tricks.sliding_window_view(haps[names[h]], L)[start[sel]]
        rev = g.random(n) < 0.5
        codes[rev] = (codes[rev] ^ 2)[:, ::-1]
        n_err = int(g.binomial(n * L, err))
        flat = codes.reshape(-1)
        ep = g.integers(0, n * L, n_err)
        flat[ep] = (flat[ep] + g.integers(1, 4, n_err, dtype=np.uint8)) & 3
        reads = LETTERS[codes]
        with_n = np.nonzero(g.random(n) < n_frac)[0]
        reads[with_n, g.integers(0, L, with_n.size)] = ord("N")
        low = g.random(n) < lowercase_frac
        reads[low] |= 0x20
        out[parent] = reads
    return out


def write_reads_fastq(path, reads: np.ndarray, gz: bool = False, prefix: str = "r") -> str:
    """Plain 4-line FASTQ of an [n, L] ASCII array (parental WGS reads carry no barcode)."""
    import gzip
    n, L = reads.shape
    qual = b"F" * L
    chunks = []
    for i in range(n):
        chunks.append(b"@%s%d\n%s\n+\n%s\n" % (prefix.encode(), i, reads[i].tobytes(), qual))
    data = b"".join(chunks)
    with (gzip.open(path, "wb", compresslevel=1) if gz else open(path, "wb")) as f:
        f.write(data)
    return str(path)


def make_trio(spec: TrioSpec, device: str = "cpu", keep_reads_on_device: bool = False) -> Trio:
    if spec.k > 31 and device != "cpu":
        raise ValueError("torch path supports k <= 31")
    G, k, L = spec.genome_len, spec.k, spec.read_len
    base_seed = spec.seed
    rng = lambda s: np.random.Generator(np.random.PCG64(base_seed * 1000 + s))

    # 1. ancestor and the four parental haplotypes
    haps = _haplotypes(spec)

    # 2. parent-unique canonical k-mers (exact set difference)
    if device == "cpu":
        ks = {n: canonical_kmers_np(h, k) for n, h in haps.items()}
        pat, mat = _unique_sets([ks["P1"], ks["P2"]], [ks["M1"], ks["M2"]], "cpu")
        del ks
    else:
        import torch
        ks = {n: _canonical_kmers_torch(torch.from_numpy(h).to(device), k) for n, h in haps.items()}
        pat, mat = _unique_sets([ks["P1"], ks["P2"]], [ks["M1"], ks["M2"]], device)
        del ks
    g = rng(31)
    if spec.decoy_kmers:
        kmask = np.uint64((1 << (2 * k)) - 1) if k < 32 else np.uint64(0xFFFFFFFFFFFFFFFF)
        for which in (0, 1):
            d = g.integers(0, 1 << 62, spec.decoy_kmers, dtype=np.uint64) & kmask
            d = np.minimum(d, revcomp_packed(d, k))
            if which == 0:
                pat = np.unique(np.concatenate([pat, d]))
            else:
                mat = np.unique(np.concatenate([mat, d]))
    pat = pat[g.permutation(pat.size)]
    mat = mat[g.permutation(mat.size)]

    # 3. barcodes: distinct a_b_c triples, each tied to one child haplotype and 1-3 fragments
    g = rng(41)
    B = spec.n_barcodes
    codes = np.unique(g.integers(0, 1536 ** 3, int(B * 1.2) + 16))
    while codes.size < B:
        codes = np.unique(np.concatenate([codes, g.integers(0, 1536 ** 3, B)]))
    codes = codes[g.permutation(codes.size)[:B]]
    triples = np.stack([codes // (1536 * 1536) + 1, (codes // 1536) % 1536 + 1, codes % 1536 + 1], axis=1)
    bc_hap = g.integers(0, 2, B).astype(np.int8)
    nfrag = g.integers(1, 4, B)
    flen = np.minimum(g.integers(20_000, 60_001, (B, 3)), max(G // 2, 2 * L + 600))
    fstart = (g.random((B, 3)) * (G - flen)).astype(np.int64)

    # 4. read pairs
    g = rng(spec.read_seed)
    P = spec.n_pairs
    if spec.zipf_alpha:
        w = 1.0 / np.power(np.arange(1, B + 1, dtype=np.float64), spec.zipf_alpha)
        cdf = np.cumsum(w / w.sum())
        src = np.minimum(np.searchsorted(cdf, g.random(P)), B - 1)
        src = g.permutation(B)[src]
    else:
        src = g.integers(0, B, P)
    fi = (g.random(P) * nfrag[src]).astype(np.int64)
    fl = flen[src, fi]
    fs = fstart[src, fi]
    ins = g.integers(300, 501, P)
    ins = np.minimum(ins, fl)
    ins = np.maximum(ins, L)
    s1 = fs + (g.random(P) * (fl - ins + 1)).astype(np.int64)
    s1 = np.minimum(s1, G - ins)
    s2 = s1 + ins - L
    hap_of_pair = bc_hap[src]
    label = src.astype(np.uint32)
    special = g.random(P) < spec.nobarcode_frac
    has_special = bool(special.any())
    if has_special:
        label[special] = B
        triples = np.concatenate([triples, np.zeros((1, 3), np.int64)])

    n_err = int(g.binomial(2 * P * L, spec.err))
    err_pos = g.integers(0, 2 * P * L, n_err)
    err_add = g.integers(1, 4, n_err, dtype=np.uint8)
    n_reads_N = g.random(2 * P) < spec.n_frac
    n_pos = g.integers(0, L, 2 * P)
    lower = g.random(2 * P) < spec.lowercase_frac if spec.lowercase_frac else None

    use_torch = device != "cpu"
    if use_torch:
        import torch
        dev_h = [torch.from_numpy(haps["P1"]).to(device), torch.from_numpy(haps["M1"]).to(device)]
        lut = torch.from_numpy(LETTERS.copy()).to(device)
        codes_all = torch.empty((2 * P, L), dtype=torch.uint8, device=device)
        CH = 1 << 20
        hp = torch.from_numpy(hap_of_pair.astype(np.int64)).to(device)
        t1 = torch.from_numpy(s1).to(device)
        t2 = torch.from_numpy(s2).to(device)
        ar = torch.arange(L, device=device)
        for lo in range(0, P, CH):
            hi = min(P, lo + CH)
            for h in (0, 1):
                sel = (hp[lo:hi] == h).nonzero().squeeze(1)
                if sel.numel() == 0:
                    continue
                a = dev_h[h][(t1[lo:hi][sel][:, None] + ar)]
                b = dev_h[h][(t2[lo:hi][sel][:, None] + ar)]
                codes_all[lo + sel] = a
                codes_all[P + lo + sel] = (b ^ 2).flip(1)
        flat = codes_all.view(-1)
        ep = torch.from_numpy(err_pos).to(device)
        flat[ep] = (flat[ep] + torch.from_numpy(err_add).to(device)) & 3
        reads = codes_all                      # code -> letter in place, chunked (A65 C67 T84 G71)
        for lo in range(0, 2 * P, CH):
            c = reads[lo:lo + CH]
            reads[lo:lo + CH] = lut[c.to(torch.int64)]
        del codes_all
        nr = torch.from_numpy(np.nonzero(n_reads_N)[0]).to(device)
        reads[nr, torch.from_numpy(n_pos[n_reads_N]).to(device)] = ord("N")
        if lower is not None:
            lr = torch.from_numpy(np.nonzero(lower)[0]).to(device)
            reads[lr] = reads[lr] | 0x20
        r1, r2 = reads[:P], reads[P:]
        if not keep_reads_on_device:
            r1, r2 = r1.cpu().numpy(), r2.cpu().numpy()
    else:
        codes_all = np.empty((2 * P, L), np.uint8)
        win = {0: np.lib.stride_tricks.sliding_window_view(haps["P1"], L),
               1: np.lib.stride_tricks.sliding_window_view(haps["M1"], L)}
        CH = 1 << 18
        for lo in range(0, P, CH):
            hi = min(P, lo + CH)
            for h in (0, 1):
                sel = np.nonzero(hap_of_pair[lo:hi] == h)[0]
                if sel.size == 0:
                    continue
                codes_all[lo + sel] = win[h][s1[lo:hi][sel]]
                codes_all[P + lo + sel] = (win[h][s2[lo:hi][sel]] ^ 2)[:, ::-1]
        flat = codes_all.reshape(-1)
        flat[err_pos] = (flat[err_pos] + err_add) & 3
        reads = LETTERS[codes_all]
        del codes_all
        reads[np.nonzero(n_reads_N)[0], n_pos[n_reads_N]] = ord("N")
        if lower is not None:
            reads[lower] |= 0x20
        r1, r2 = reads[:P], reads[P:]

    return Trio(spec=spec, pat=pat.astype(np.uint64), mat=mat.astype(np.uint64), r1=r1, r2=r2,
                pair_bc=label, bc_triples=triples, bc_hap=bc_hap)


# The named configurations of BASELINE.json
def config(name: str) -> TrioSpec:
    if name == "tiny":
        return TrioSpec(genome_len=60_000, het=0.002, n_pairs=2_000, n_barcodes=100)
    if name == "small":
        return TrioSpec(genome_len=500_000, het=0.001, n_pairs=20_000, n_barcodes=1_000)
    if name == "cfg1":       # configs[0]
        return TrioSpec(genome_len=5_000_000, het=0.001, n_pairs=200_000, n_barcodes=10_000)
    if name == "cfg2":       # configs[1]: the single-GPU bench workload
        return TrioSpec(genome_len=100_000_000, het=0.001, n_pairs=20_000_000, n_barcodes=500_000)
    if name == "cfg3t":      # configs[2] TABLE scale on one GPU: the cfg2 trio and reads, k-mer lists padded with
        # random decoy k-mers to the ~62 M keys of a human trio, so that the exact table
        # (1 GiB) is HBM-resident and the pre-filter runs at its size cap
        return TrioSpec(genome_len=100_000_000, het=0.001, n_pairs=20_000_000, n_barcodes=500_000,
                        decoy_kmers=26_800_000)
    if name == "cfg3":       # configs[2] as ONE GPU of eight sees it: human-size k-mer lists (as cfg3t), all 20 M barcodes,
        # 600 M / 8 = 75 M read pairs (15 GB of bases, resident in HBM)
        return TrioSpec(genome_len=100_000_000, het=0.001, n_pairs=75_000_000, n_barcodes=20_000_000,
                        decoy_kmers=26_800_000)
    if name == "cfg3b":      # configs[2] BARCODE scale for the test suite: 20 M barcodes (160 MB of counters, larger
        # than L2), the cfg2 trio's own k-mer lists, 10 M pairs
        return TrioSpec(genome_len=100_000_000, het=0.001, n_pairs=10_000_000, n_barcodes=20_000_000)
    raise KeyError(name)
