/*
 * hast_oracle.c -- plain-C restatement of HAST stage 01, see hast_oracle.h.
 *
 * TEST INFRASTRUCTURE ONLY (checker for tests/, smoke() and bench.py's CPU
 * baseline legs).  Never linked into or called from the product path.
 *
 * Every function names the reference lines it follows; paths are relative to
 * /root/reference/01.classify_stlfr_reads/ .  The reference keeps k-mers in a
 * 128-bit {high,low} pair (kmer.h:63) but is only correct for k <= 32
 * (kmer.h:225-238 falls through for wider words), where high == 0; this
 * restatement therefore carries `low` alone in a uint64_t.
 */
#define _GNU_SOURCE
#include "hast_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* ------------------------------------------------------------------------ */
/* k-mer arithmetic                                                          */
/* ------------------------------------------------------------------------ */

/* kmer.h:11 */
int ho_base2int(unsigned char c) { return (c & 0x06) >> 1; }
/* kmer.h:12 */
char ho_int2base(int code) { return "ACTG"[code & 3]; }

/* kmer.h:129-148 createFilter: the 2k low bits */
static uint64_t ho_filter(int k) {
    return (2 * k < 64) ? ((((uint64_t)1) << (2 * k)) - 1) : ~(uint64_t)0;
}

/* kmer.h:196-223 fastReverseComp: complement every 2-bit group (^10b),
 * reverse the order of the groups, right-align the k used groups. */
uint64_t ho_revcomp(uint64_t w, int k) {
    w ^= 0xAAAAAAAAAAAAAAAAULL;
    w = ((w & 0x3333333333333333ULL) << 2) | ((w & 0xCCCCCCCCCCCCCCCCULL) >> 2);
    w = ((w & 0x0F0F0F0F0F0F0F0FULL) << 4) | ((w & 0xF0F0F0F0F0F0F0F0ULL) >> 4);
    w = ((w & 0x00FF00FF00FF00FFULL) << 8) | ((w & 0xFF00FF00FF00FF00ULL) >> 8);
    w = ((w & 0x0000FFFF0000FFFFULL) << 16) | ((w & 0xFFFF0000FFFF0000ULL) >> 16);
    w = ((w & 0x00000000FFFFFFFFULL) << 32) | ((w & 0xFFFFFFFF00000000ULL) >> 32);
    if (k < 32) w >>= (64 - 2 * k);
    return w;
}

/* kmer.h:153-166 */
uint64_t ho_str2kmer(const char *s, int k) {
    uint64_t word = 0;
    for (int i = 0; i < k; i++) word = (word << 2) | (uint64_t)ho_base2int((unsigned char)s[i]);
    uint64_t bal = ho_revcomp(word, k);
    return word < bal ? word : bal;
}

/* kmer.h:169-194 with nextKmer (109-114) and prevKmer (116-127) */
long ho_chop(const char *read, long len, int k, uint64_t *out) {
    if (len < k) return -1;                       /* kmer.h:171 assert */
    const uint64_t filter = ho_filter(k);
    uint64_t word = 0;
    for (int i = 0; i < k; i++) word = (word << 2) | (uint64_t)ho_base2int((unsigned char)read[i]);
    uint64_t bal = ho_revcomp(word, k);
    long n = 0;
    out[n++] = word < bal ? word : bal;
    for (long index = 1; index <= len - k; index++) {
        /* nextKmer(read[index-1+k]) */
        word = ((word << 2) & filter) | (uint64_t)ho_base2int((unsigned char)read[index - 1 + k]);
        /* prevKmer(bal_read[len-index-k]); bal_read[j] = comp(read[len-1-j])
         * (kmer.h:38-52), i.e. the complement of the base that just entered */
        uint64_t ch = (uint64_t)(ho_base2int((unsigned char)read[index - 1 + k]) ^ 0x02);
        bal = (bal >> 2) | (ch << (2 * (k - 1)));
        out[n++] = word < bal ? word : bal;
    }
    return n;
}

/* kmer.h:244-254, 14-25 */
void ho_kmer2str(uint64_t w, int k, char *out) {
    for (int i = 0; i < k; i++) {
        out[k - 1 - i] = ho_int2base((int)(w & 3));
        w >>= 2;
    }
    out[k] = 0;
}

/* ------------------------------------------------------------------------ */
/* header parsing, N scan, haplotype call                                    */
/* ------------------------------------------------------------------------ */

/* classify.cpp:112-119.  substr(s+1, e-s-1): a negative count converts to a
 * huge size_t, which std::string clamps to "through the end of the string". */
void ho_parse_name(const char *head, long len, long *start, long *blen) {
    long s = -1, e = -1;
    for (long i = 0; i < len; i++) {
        if (head[i] == '#') s = i;
        if (head[i] == '/') e = i;
    }
    long cnt = e - s - 1;
    *start = s + 1;
    *blen = (cnt < 0 || s + 1 + cnt > len) ? len - (s + 1) : cnt;
}

/* classify.cpp:182-185 */
int ho_contain_n(const char *seq, long len) {
    for (long i = 0; i < len; i++) if (seq[i] == 'N') return 1;
    return 0;
}

/* classify.cpp:66-86 */
int ho_get_hap(const char *barcode, int has0, int c0, int has1, int c1,
               size_t n0, size_t n1, double w0, double w1) {
    if (!strcmp(barcode, "0_0_0") || !strcmp(barcode, "0_0") || !strcmp(barcode, "0")) return -1;
    if (has0 && has1) {
        double df0 = (double)c0 / (double)n0;
        double df1 = (double)c1 / (double)n1;
        df0 *= w0;
        df1 *= w1;
        if (df0 > df1) return 0;
        if (df1 > df0) return 1;
        return -1;
    } else if (has0) {
        return c0 > 0 ? 0 : -1;
    } else if (has1) {
        return c1 > 0 ? 1 : -1;
    }
    return -1;
}

/* ------------------------------------------------------------------------ */
/* the two k-mer sets (classify.cpp:27) -- open addressing, linear probing   */
/* ------------------------------------------------------------------------ */

#define HO_EMPTY 0xFFFFFFFFFFFFFFFFULL   /* never a canonical k-mer: rc(GG..G)=CC..C is smaller */
#define HO_TOMB  0xFFFFFFFFFFFFFFFEULL   /* likewise: GG..GT -> rc AC..C is smaller            */

typedef struct { uint64_t *slot; size_t cap, size, used; } ho_set;

static uint64_t ho_mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static void ho_set_init(ho_set *s, size_t cap) {
    s->cap = cap; s->size = 0; s->used = 0;
    s->slot = (uint64_t *)malloc(cap * sizeof(uint64_t));
    for (size_t i = 0; i < cap; i++) s->slot[i] = HO_EMPTY;
}
static int ho_set_has(const ho_set *s, uint64_t key) {
    if (!s->cap) return 0;
    size_t m = s->cap - 1, i = ho_mix(key) & m;
    for (;;) {
        uint64_t v = s->slot[i];
        if (v == key) return 1;
        if (v == HO_EMPTY) return 0;
        i = (i + 1) & m;
    }
}
static void ho_set_insert(ho_set *s, uint64_t key);
static void ho_set_grow(ho_set *s) {
    ho_set old = *s;
    ho_set_init(s, old.cap ? old.cap * 2 : 1024);
    for (size_t i = 0; i < old.cap; i++)
        if (old.slot[i] != HO_EMPTY && old.slot[i] != HO_TOMB) ho_set_insert(s, old.slot[i]);
    free(old.slot);
}
static void ho_set_insert(ho_set *s, uint64_t key) {
    if ((s->used + 1) * 2 > s->cap) ho_set_grow(s);
    size_t m = s->cap - 1, i = ho_mix(key) & m;
    for (;;) {
        uint64_t v = s->slot[i];
        if (v == key) return;                       /* unordered_set::insert: no duplicates */
        if (v == HO_EMPTY) { s->slot[i] = key; s->size++; s->used++; return; }
        i = (i + 1) & m;
    }
}
static int ho_set_erase(ho_set *s, uint64_t key) {
    if (!s->cap) return 0;
    size_t m = s->cap - 1, i = ho_mix(key) & m;
    for (;;) {
        uint64_t v = s->slot[i];
        if (v == key) { s->slot[i] = HO_TOMB; s->size--; return 1; }
        if (v == HO_EMPTY) return 0;
        i = (i + 1) & m;
    }
}

/* ------------------------------------------------------------------------ */
/* BarcodeCache (classify.cpp:50-64): map<string, map<int,int>>              */
/* ------------------------------------------------------------------------ */

typedef struct {
    char *name;
    int has[3];    /* key present for hap -1, 0, 1  (index hap+1) */
    int cnt[3];
} ho_bc;

typedef struct {
    ho_bc *e; size_t n, cap;
    uint32_t *ht; size_t hcap;       /* open-addressing index: entry+1, 0 = empty */
} ho_bcmap;

static uint64_t ho_strhash(const char *s, size_t n) {
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 1099511628211ULL; }
    return ho_mix(h);
}
static void ho_bcmap_rehash(ho_bcmap *m, size_t hcap) {
    free(m->ht);
    m->hcap = hcap;
    m->ht = (uint32_t *)calloc(hcap, sizeof(uint32_t));
    for (size_t i = 0; i < m->n; i++) {
        size_t j = ho_strhash(m->e[i].name, strlen(m->e[i].name)) & (hcap - 1);
        while (m->ht[j]) j = (j + 1) & (hcap - 1);
        m->ht[j] = (uint32_t)(i + 1);
    }
}
static ho_bc *ho_bcmap_get(ho_bcmap *m, const char *s, size_t n) {
    if (!m->hcap) ho_bcmap_rehash(m, 1024);
    size_t j = ho_strhash(s, n) & (m->hcap - 1);
    while (m->ht[j]) {
        ho_bc *b = &m->e[m->ht[j] - 1];
        if (strlen(b->name) == n && !memcmp(b->name, s, n)) return b;
        j = (j + 1) & (m->hcap - 1);
    }
    if (m->n == m->cap) {
        m->cap = m->cap ? m->cap * 2 : 1024;
        m->e = (ho_bc *)realloc(m->e, m->cap * sizeof(ho_bc));
    }
    ho_bc *b = &m->e[m->n++];
    memset(b, 0, sizeof(*b));
    b->name = (char *)malloc(n + 1);
    memcpy(b->name, s, n);
    b->name[n] = 0;
    m->ht[j] = (uint32_t)m->n;
    if (m->n * 2 > m->hcap) { ho_bcmap_rehash(m, m->hcap * 2); b = &m->e[m->n - 1]; }
    return b;
}
/* classify.cpp:52-56 IncrBarcodeHaps */
static void ho_incr(ho_bcmap *m, const char *bc, size_t n, int hap, int incr) {
    ho_bc *b = ho_bcmap_get(m, bc, n);
    if (!b->has[hap + 1]) { b->has[hap + 1] = 1; b->cnt[hap + 1] = 0; }
    b->cnt[hap + 1] += incr;
}

/* ------------------------------------------------------------------------ */
/* classifier                                                                */
/* ------------------------------------------------------------------------ */

struct ho_classifier {
    int k;                 /* g_K, classify.cpp:29 */
    ho_set set[2];         /* g_kmers, classify.cpp:27 */
    double w0, w1;         /* g_hap0_fac / g_hap1_fac, classify.cpp:22-23 */
    ho_bcmap data;         /* BarcodeCache data, classify.cpp:439 */
    char err[256];
};

ho_classifier *ho_create(void) {
    ho_classifier *c = (ho_classifier *)calloc(1, sizeof(*c));
    c->w0 = c->w1 = 1.0;
    return c;
}
void ho_destroy(ho_classifier *c) {
    if (!c) return;
    free(c->set[0].slot); free(c->set[1].slot);
    for (size_t i = 0; i < c->data.n; i++) free(c->data.e[i].name);
    free(c->data.e); free(c->data.ht);
    free(c);
}
const char *ho_error(const ho_classifier *c) { return c->err; }
int ho_k(const ho_classifier *c) { return c->k; }
size_t ho_set_size(const ho_classifier *c, int index) { return c->set[index].size; }
void ho_set_weights(ho_classifier *c, double w0, double w1) { c->w0 = w0; c->w1 = w1; }
size_t ho_n_barcodes(const ho_classifier *c) { return c->data.n; }
int ho_lookup(const ho_classifier *c, uint64_t canonical) {
    return ho_set_has(&c->set[0], canonical) | (ho_set_has(&c->set[1], canonical) << 1);
}

/* classify.cpp:30-46.  std::getline semantics: the loop ends at the first
 * getline that reaches end-of-file, so a last line lacking '\n' is dropped
 * (classify.cpp:41); for index 0 the first line is taken unconditionally
 * (classify.cpp:35-39), even when it is that unterminated last line. */
long ho_load_kmers_mem(ho_classifier *c, const char *text, size_t n, int index) {
    long total = 0;
    size_t pos = 0;
    if (index == 0) {
        const char *nl = (const char *)memchr(text, '\n', n);
        size_t len = nl ? (size_t)(nl - text) : n;
        c->k = (int)len;
        if (len < 1 || len > 32) {
            snprintf(c->err, sizeof c->err, "k=%zu outside 1..32", len);
            return -1;
        }
        ho_set_insert(&c->set[0], ho_str2kmer(text, c->k));
        total++;
        pos = nl ? len + 1 : n;
    }
    while (pos < n) {
        const char *nl = (const char *)memchr(text + pos, '\n', n - pos);
        if (!nl) break;                                   /* unterminated tail: dropped */
        size_t len = (size_t)(nl - (text + pos));
        if ((int)len != c->k) {                           /* kmer.h:154 assert */
            snprintf(c->err, sizeof c->err, "k-mer line of length %zu, expected %d", len, c->k);
            return -1;
        }
        ho_set_insert(&c->set[index], ho_str2kmer(text + pos, c->k));
        total++;
        pos += len + 1;
    }
    return total;
}

/* Harness convenience for the large configurations (not a reference function): the k-mers of a list that the
 * generator already holds as packed words of the reference's code (kmer.h:11-12, first base in the highest used
 * bits), inserted exactly as load_kmers does after str2Kmer -- canonical form (kmer.h:153-166), then the set. */
long ho_load_kmers_packed(ho_classifier *c, const uint64_t *kmers, size_t n, int k, int index) {
    if (k < 1 || k > 32) { snprintf(c->err, sizeof c->err, "k=%d outside 1..32", k); return -1; }
    if (index == 0) c->k = k;
    else if (c->k != k) { snprintf(c->err, sizeof c->err, "k=%d, expected %d", k, c->k); return -1; }
    for (size_t i = 0; i < n; i++) {
        const uint64_t rc = ho_revcomp(kmers[i], k);
        ho_set_insert(&c->set[index], kmers[i] < rc ? kmers[i] : rc);
    }
    return (long)n;
}

static char *ho_slurp(const char *path, size_t *n) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)sz + 1);
    *n = fread(buf, 1, (size_t)sz, f);
    fclose(f);
    return buf;
}

long ho_load_kmers_file(ho_classifier *c, const char *path, int index) {
    size_t n;
    char *buf = ho_slurp(path, &n);
    if (!buf) { snprintf(c->err, sizeof c->err, "cannot open %s", path); return -1; }
    long r = ho_load_kmers_mem(c, buf, n, index);
    free(buf);
    return r;
}

/* classify.cpp:314-339 */
long ho_init_adaptor(ho_classifier *c, const char *fwd, const char *rev) {
    const char *ad[2] = { fwd, rev };
    long erased = 0;
    for (int a = 0; a < 2; a++) {
        long len = (long)strlen(ad[a]);
        if (len < c->k) { snprintf(c->err, sizeof c->err, "adaptor shorter than k"); return -1; }
        uint64_t *km = (uint64_t *)malloc((size_t)(len - c->k + 1) * sizeof(uint64_t));
        long n = ho_chop(ad[a], len, c->k, km);
        for (long i = 0; i < n; i++) {
            erased += ho_set_erase(&c->set[0], km[i]);
            erased += ho_set_erase(&c->set[1], km[i]);
        }
        free(km);
    }
    return erased;
}

/* classify.cpp:186-209 */
int ho_process_read(ho_classifier *c, const char *head, long hlen, const char *seq, long slen) {
    int vote[2] = { 0, 0 };
    long bs, bl;
    ho_parse_name(head, hlen, &bs, &bl);
    const char *bc = head + bs;
    if (ho_contain_n(seq, slen)) {
        ho_incr(&c->data, bc, (size_t)bl, -1, 1);
        return 0;
    }
    if (slen < c->k) {                                    /* kmer.h:171 assert -> SIGABRT */
        snprintf(c->err, sizeof c->err, "read of length %ld shorter than k=%d", slen, c->k);
        return -1;
    }
    uint64_t *km = (uint64_t *)malloc((size_t)(slen - c->k + 1) * sizeof(uint64_t));
    long n = ho_chop(seq, slen, c->k, km);
    for (long i = 0; i < n; i++) {
        if (ho_set_has(&c->set[0], km[i])) vote[0]++;
        if (ho_set_has(&c->set[1], km[i])) vote[1]++;
    }
    free(km);
    if (vote[0] > 0) ho_incr(&c->data, bc, (size_t)bl, 0, vote[0]);
    if (vote[1] > 0) ho_incr(&c->data, bc, (size_t)bl, 1, vote[1]);
    if (vote[0] == 0 && vote[1] == 0) ho_incr(&c->data, bc, (size_t)bl, -1, 1);
    return 0;
}

/* A std::getline look-alike over zlib (gz) or stdio (plain), reporting the
 * two stream bits the reference's loops test. */
typedef struct {
    gzFile gz; FILE *fp;
    char *buf; size_t pos, len, cap;
    int eofbit, failbit;
} ho_stream;

static int ho_fill(ho_stream *s) {
    s->pos = 0;
    if (s->gz) { int r = gzread(s->gz, s->buf, (unsigned)s->cap); s->len = r > 0 ? (size_t)r : 0; }
    else s->len = fread(s->buf, 1, s->cap, s->fp);
    return s->len > 0;
}
/* returns the line in *line (malloc-grown); sets eofbit/failbit like getline */
static void ho_getline(ho_stream *s, char **line, size_t *n, size_t *cap) {
    if (s->eofbit || s->failbit) { s->failbit = 1; return; }   /* sentry fails: string untouched */
    *n = 0;
    size_t got = 0;
    for (;;) {
        if (s->pos == s->len && !ho_fill(s)) {
            s->eofbit = 1;
            if (!got) s->failbit = 1;
            return;
        }
        char *p = s->buf + s->pos;
        size_t avail = s->len - s->pos;
        char *nl = (char *)memchr(p, '\n', avail);
        size_t take = nl ? (size_t)(nl - p) : avail;
        if (*n + take + 1 > *cap) { *cap = (*n + take + 1) * 2; *line = (char *)realloc(*line, *cap); }
        memcpy(*line + *n, p, take);
        *n += take;
        got += take + (nl ? 1 : 0);
        s->pos += take + (nl ? 1 : 0);
        if (nl) return;
    }
}

/* classify.cpp:238-278 */
int ho_process_fastq(ho_classifier *c, const char *path) {
    ho_stream s;
    memset(&s, 0, sizeof s);
    size_t pl = strlen(path);
    int gz = pl > 3 && !strcmp(path + pl - 3, ".gz");         /* classify.cpp:245-250 */
    if (gz) s.gz = gzopen(path, "rb"); else s.fp = fopen(path, "rb");
    if (!s.gz && !s.fp) { snprintf(c->err, sizeof c->err, "cannot open %s", path); return -1; }
    s.cap = 1 << 20;
    s.buf = (char *)malloc(s.cap);
    char *head = NULL, *seq = NULL, *tmp = NULL;
    size_t hn = 0, hc = 0, sn = 0, sc = 0, tn = 0, tc = 0;
    int rc = 0;
    for (;;) {
        ho_getline(&s, &head, &hn, &hc);
        if (s.eofbit) break;                                   /* classify.cpp:257 */
        sn = 0;
        ho_getline(&s, &seq, &sn, &sc);                        /* :258 */
        if (s.failbit) sn = 0;
        if (ho_process_read(c, head ? head : "", (long)hn, seq ? seq : "", (long)sn)) { rc = -1; break; }
        ho_getline(&s, &tmp, &tn, &tc);                        /* :267-268 */
        ho_getline(&s, &tmp, &tn, &tc);
    }
    free(head); free(seq); free(tmp); free(s.buf);
    if (s.gz) gzclose(s.gz); else fclose(s.fp);
    return rc;
}

static int ho_bc_cmp(const void *a, const void *b) {
    /* std::map<std::string>: char_traits<char>::compare == unsigned bytewise */
    return strcmp(((const ho_bc *)a)->name, ((const ho_bc *)b)->name);
}

/* classify.cpp:88-102 */
int ho_print(ho_classifier *c, FILE *out) {
    ho_bc *v = (ho_bc *)malloc((c->data.n ? c->data.n : 1) * sizeof(ho_bc));
    memcpy(v, c->data.e, c->data.n * sizeof(ho_bc));
    qsort(v, c->data.n, sizeof(ho_bc), ho_bc_cmp);
    for (size_t i = 0; i < c->data.n; i++) {
        int hap = ho_get_hap(v[i].name, v[i].has[1], v[i].cnt[1], v[i].has[2], v[i].cnt[2],
                             c->set[0].size, c->set[1].size, c->w0, c->w1);
        fprintf(out, "%s\t%d\t%d\t%d\n", v[i].name, hap,
                v[i].has[1] ? v[i].cnt[1] : 0, v[i].has[2] ? v[i].cnt[2] : 0);
    }
    free(v);
    return 0;
}
int ho_print_file(ho_classifier *c, const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    ho_print(c, f);
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* dense batch form of process_reads (classify.cpp:186-209)                  */
/* ------------------------------------------------------------------------ */

typedef struct {
    const ho_classifier *c; const uint8_t *bases; const uint64_t *off; const uint32_t *bc;
    size_t lo, hi, nbc; int32_t *counts; long long lookups; int bad;
} ho_job;

static void *ho_batch_worker(void *arg) {
    ho_job *j = (ho_job *)arg;
    const int k = j->c->k;
    uint64_t *km = NULL;
    size_t kcap = 0;
    for (size_t r = j->lo; r < j->hi; r++) {
        const char *seq = (const char *)j->bases + j->off[r];
        long len = (long)(j->off[r + 1] - j->off[r]);
        if (ho_contain_n(seq, len)) continue;
        if (len < k) { j->bad = 1; continue; }
        size_t n = (size_t)(len - k + 1);
        if (n > kcap) { kcap = n * 2; km = (uint64_t *)realloc(km, kcap * sizeof(uint64_t)); }
        ho_chop(seq, len, k, km);
        int v0 = 0, v1 = 0;
        for (size_t i = 0; i < n; i++) {
            v0 += ho_set_has(&j->c->set[0], km[i]);
            v1 += ho_set_has(&j->c->set[1], km[i]);
        }
        j->lookups += (long long)n;
        j->counts[2 * (size_t)j->bc[r] + 0] += v0;
        j->counts[2 * (size_t)j->bc[r] + 1] += v1;
    }
    free(km);
    return NULL;
}

long long ho_classify_batch(const ho_classifier *c, const uint8_t *bases, const uint64_t *read_off,
                            const uint32_t *barcode_id, size_t n_reads, int32_t *counts,
                            size_t n_barcodes, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    ho_job *jobs = (ho_job *)calloc((size_t)nthreads, sizeof(ho_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) {
        ho_job *j = &jobs[t];
        j->c = c; j->bases = bases; j->off = read_off; j->bc = barcode_id; j->nbc = n_barcodes;
        j->lo = n_reads * (size_t)t / (size_t)nthreads;
        j->hi = n_reads * (size_t)(t + 1) / (size_t)nthreads;
        /* private partial counts per worker, summed afterwards: BarcodeCache::Add,
         * classify.cpp:57-63,226-229 */
        j->counts = t == 0 ? counts : (int32_t *)calloc(n_barcodes * 2, sizeof(int32_t));
        if (t) pthread_create(&th[t], NULL, ho_batch_worker, j);
    }
    ho_batch_worker(&jobs[0]);
    long long lookups = jobs[0].lookups;
    int bad = jobs[0].bad;
    for (int t = 1; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        for (size_t i = 0; i < n_barcodes * 2; i++) counts[i] += jobs[t].counts[i];
        free(jobs[t].counts);
        lookups += jobs[t].lookups;
        bad |= jobs[t].bad;
    }
    free(jobs); free(th);
    return bad ? -1 : lookups;
}

/* ------------------------------------------------------------------------ */
/* mergeResult (mergeResult.cpp)                                             */
/* ------------------------------------------------------------------------ */

/* mergeResult.cpp:33-53: the same ladder as classify's getHap but without the
 * division by the set sizes, and with float weights promoted to double. */
static int ho_merge_get_hap(const char *barcode, int has0, int c0, int has1, int c1, float w0, float w1) {
    if (!strcmp(barcode, "0_0_0") || !strcmp(barcode, "0_0") || !strcmp(barcode, "0")) return -1;
    if (has0 && has1) {
        double df0 = (double)c0, df1 = (double)c1;
        df0 *= w0;
        df1 *= w1;
        if (df0 > df1) return 0;
        if (df1 > df0) return 1;
        return -1;
    } else if (has0) {
        return c0 > 0 ? 0 : -1;
    } else if (has1) {
        return c1 > 0 ? 1 : -1;
    }
    return -1;
}

int ho_merge_result(const char *const *inputs, int n_inputs, float w0, float w1, int intended,
                    const char *out_path) {
    ho_bcmap m;
    memset(&m, 0, sizeof m);
    for (int f = 0; f < n_inputs; f++) {
        size_t n;
        char *buf = ho_slurp(inputs[f], &n);
        if (!buf) return -1;                                   /* mergeResult.cpp:119-122 */
        size_t pos = 0;
        while (pos < n) {
            char *nl = (char *)memchr(buf + pos, '\n', n - pos);
            if (!nl) break;                                    /* :124 loop ends at eof */
            *nl = 0;
            /* mergeResult.cpp:21-30 AddLine: is>>barcode>>type>>hap0>>hap1.  operator>> skips
             * leading whitespace, so a line whose barcode is the empty string ("\t-1\t0\t0", which
             * classify prints for headers like "@r#/1") is read shifted by one column: barcode "-1",
             * type 0, hap0 0, and hap1 fails -> 0 (C++11 writes 0 on a failed extraction). */
            char bc[4096];
            int v[3] = { 0, 0, 0 };
            const char *q = buf + pos;
            size_t bl = 0;
            while (*q == ' ' || *q == '\t') q++;
            while (*q && *q != ' ' && *q != '\t' && bl < sizeof bc - 1) bc[bl++] = *q++;
            bc[bl] = 0;
            for (int i = 0; i < 3; i++) {
                char *e;
                while (*q == ' ' || *q == '\t') q++;
                long x = strtol(q, &e, 10);
                if (e == q) break;                              /* failed extraction: this and later stay 0 */
                v[i] = (int)x;
                q = e;
            }
            int h0 = v[1], h1 = v[2];
            if (intended) {
                ho_incr(&m, bc, strlen(bc), 0, h0);
                ho_incr(&m, bc, strlen(bc), 1, h1);
            } else {                                           /* :28-29, both into key 0 */
                ho_incr(&m, bc, strlen(bc), 0, h0);
                ho_incr(&m, bc, strlen(bc), 0, h1);
            }
            pos = (size_t)(nl - buf) + 1;
        }
        free(buf);
    }
    FILE *out = fopen(out_path, "wb");
    if (!out) return -1;
    qsort(m.e, m.n, sizeof(ho_bc), ho_bc_cmp);
    for (size_t i = 0; i < m.n; i++) {
        ho_bc *b = &m.e[i];
        int hap = ho_merge_get_hap(b->name, b->has[1], b->cnt[1], b->has[2], b->cnt[2], w0, w1);
        fprintf(out, "%s\t%d\t%d\t%d\n", b->name, hap, b->has[1] ? b->cnt[1] : 0, b->has[2] ? b->cnt[2] : 0);
    }
    fclose(out);
    for (size_t i = 0; i < m.n; i++) free(m.e[i].name);
    free(m.e); free(m.ht);
    return 0;
}
