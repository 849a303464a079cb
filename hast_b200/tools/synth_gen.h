/*
 * synth_gen.h -- counter-based stLFR read-pair generator ("generate batches directly
 * ... from (seed, pair index)").
 *
 * Synthetic-data tooling, NOT on the classification path.  Every read pair is a pure function of
 * (parameters, pair index), so any rank can produce any slice of a 600 M-pair workload without host
 * arrays, and the CPU (oracle side of the parity checks) reproduces single pairs bit for bit: this one
 * header is compiled by nvcc into the device generator (synth_gen.cu) and by g++ into the host one
 * (synth_gen_cpu.cpp).
 *
 * Model (the shape of hast_b200/synth.py): a barcode owns 1-3 fragments of
 * 20-60 kb of ONE child haplotype; a pair picks a barcode (uniform, or through an integer CDF for the
 * heavy-tailed configs[4]), a fragment, an insert of 300-500 bp; r1 = forward L bases, r2 = reverse
 * complement of the last L bases of the insert; substitution errors (Poisson, mean L * 0.2 %), a
 * fraction of reads with one 'N', a fraction of pairs labelled 0_0_0 (barcode id = n_barcodes).
 * Base codes are the reference's: A0 C1 T2 G3 (kmer.h:11-12).
 */
#ifndef HAST_SYNTH_GEN_H
#define HAST_SYNTH_GEN_H
#include <stdint.h>

#ifdef __CUDACC__
#define SG_HD __host__ __device__ __forceinline__
#else
#define SG_HD static inline
#endif

typedef struct sg_params {
    uint64_t genome_len;      /* G: length of each child haplotype (codes, one byte per base)        */
    uint64_t n_barcodes;      /* B: real barcodes 0..B-1; id B is "0_0_0"                            */
    uint64_t seed;
    uint32_t read_len;        /* L                                                                    */
    uint32_t thr_special;     /* pair is 0_0_0 iff (h & 0xffffffff) < thr_special                     */
    uint32_t thr_n;           /* read gets one 'N' iff (h & 0xffffffff) < thr_n                       */
    uint32_t thr_err[3];      /* read has > j errors iff (h & 0xffffffff) < thr_err[j]                */
    uint32_t has_cdf;         /* barcode drawn through cdf[] (heavy tail) instead of uniformly        */
    uint32_t pad_;
} sg_params;

SG_HD uint64_t sg_mix(uint64_t x) {             /* splitmix64 finaliser */
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
SG_HD uint64_t sg_mulhi(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
SG_HD uint64_t sg_stream(uint64_t seed, uint64_t tag, uint64_t i) {
    return sg_mix(sg_mix(seed * 0x9E3779B97F4A7C15ull + tag) ^ (i * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull));
}

typedef struct sg_pair {
    uint32_t barcode;         /* dense id, n_barcodes = special                                        */
    uint32_t hap;             /* child haplotype the pair is read from: 0 = paternal, 1 = maternal     */
    uint64_t s1, s2;          /* start of r1's window; start of the window r2 is the reverse complement of */
} sg_pair;

/* cdf: B ascending 64-bit thresholds (only when p->has_cdf): barcode = first b with h <= cdf[b] */
SG_HD sg_pair sg_pair_of(const sg_params* p, const uint64_t* cdf, uint64_t i) {
    sg_pair r;
    const uint64_t G = p->genome_len, L = p->read_len, B = p->n_barcodes;
    const uint64_t h0 = sg_stream(p->seed, 1, i);
    const int special = (uint32_t)h0 < p->thr_special;
    const uint64_t h1 = sg_stream(p->seed, 2, i);
    uint64_t src;
    if (p->has_cdf) {
        uint64_t lo = 0, hi = B - 1;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (cdf[mid] < h1) lo = mid + 1; else hi = mid; }
        src = lo;
    } else {
        src = sg_mulhi(h1, B);
    }
    const uint64_t hb = sg_stream(p->seed, 3, src);
    r.hap = (uint32_t)(hb & 1);
    const uint64_t nfrag = 1 + ((hb >> 1) % 3);
    const uint64_t f = sg_mulhi(sg_stream(p->seed, 4, i), nfrag);
    const uint64_t hf = sg_stream(p->seed, 5, src * 4 + f);
    uint64_t flen = 20000 + (hf % 40001);
    uint64_t fmax = G / 2; if (fmax < 2 * L + 600) fmax = 2 * L + 600; if (fmax > G) fmax = G;
    if (flen > fmax) flen = fmax;
    const uint64_t fstart = sg_mulhi(sg_mix(hf), G - flen + 1);
    const uint64_t h3 = sg_stream(p->seed, 6, i);
    uint64_t ins = 300 + (h3 % 201);
    if (ins > flen) ins = flen;
    if (ins < L) ins = L;
    uint64_t s1 = fstart + sg_mulhi(sg_mix(h3), flen - ins + 1);
    if (s1 > G - ins) s1 = G - ins;
    r.s1 = s1;
    r.s2 = s1 + ins - L;
    r.barcode = (uint32_t)(special ? B : src);
    return r;
}

/* Edits of read (pair i, mate m in {0,1}): up to 3 substitutions then at most one 'N'.
 * pos[j] < L, add[j] in 1..3; n_pos = L when the read gets no N. */
typedef struct sg_edits { uint32_t n_err; uint32_t pos[3]; uint32_t add[3]; uint32_t n_pos; } sg_edits;

SG_HD sg_edits sg_edits_of(const sg_params* p, uint64_t i, uint32_t mate) {
    sg_edits e;
    const uint32_t L = p->read_len;
    const uint64_t he = sg_stream(p->seed, 7 + mate, i);
    const uint32_t u = (uint32_t)he;
    e.n_err = (u < p->thr_err[0]) + (u < p->thr_err[1]) + (u < p->thr_err[2]);
    uint64_t h = he;
    for (int j = 0; j < 3; ++j) {
        h = sg_mix(h + 0x9E3779B97F4A7C15ull);
        e.pos[j] = (uint32_t)((h >> 8) % L);
        e.add[j] = 1 + (uint32_t)((h >> 40) % 3);
    }
    const uint64_t hn = sg_stream(p->seed, 9 + mate, i);
    e.n_pos = ((uint32_t)hn < p->thr_n) ? (uint32_t)((hn >> 32) % L) : L;
    return e;
}

/* base j (0..L-1) of the read as an ASCII letter; hap0/hap1 = child haplotype codes */
SG_HD uint8_t sg_base(const sg_params* p, const uint8_t* hap0, const uint8_t* hap1, const sg_pair* pr,
                      const sg_edits* e, uint32_t mate, uint32_t j) {
    const uint8_t* h = pr->hap ? hap1 : hap0;
    const uint32_t L = p->read_len;
    uint32_t c = mate ? (uint32_t)(h[pr->s2 + (L - 1 - j)] ^ 2u) : (uint32_t)h[pr->s1 + j];
    for (uint32_t t = 0; t < e->n_err; ++t)
        if (e->pos[t] == j) c = (c + e->add[t]) & 3u;
    if (e->n_pos == j) return (uint8_t)'N';
    return (uint8_t)(0x47544341u >> (8u * c));      /* "ACTG"[c], kmer.h:12 */
}

#endif
