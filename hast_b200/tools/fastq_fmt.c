/*
 * fastq_fmt.c -- fast FASTQ text emitter for the synthetic-trio generator.
 *
 * Synthetic-data tooling (not on the classification path): turns the generator's in-memory reads (the same
 * arrays the device interface takes) into stLFR-style FASTQ files so that the
 * reference binary in oracle/_ref and the product CLI can be fed identical
 * input.  Header layout follows the example at classify.cpp:109-111:
 *     @V300000001L1C001R0000000042#203_1533_1069/1
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* seqs[off[i]..off[i+1]) are the bases of read i; bc_names is a blob of
 * NUL-terminated barcode strings, bc_name_off[id] the start of barcode id. */
int ff_write_fastq(const char *path, int gz, const uint8_t *seqs, const uint64_t *off, size_t n,
                   const char *bc_names, const uint64_t *bc_name_off, const uint32_t *bc_id,
                   const uint64_t *read_no, int mate) {
    gzFile g = NULL;
    FILE *f = NULL;
    if (gz) g = gzopen(path, "wb1"); else f = fopen(path, "wb");
    if (!g && !f) return -1;
    size_t cap = 8u << 20, len = 0;
    char *buf = (char *)malloc(cap + 4096);
    for (size_t i = 0; i < n; i++) {
        size_t L = (size_t)(off[i + 1] - off[i]);
        if (len + 2 * L + 256 > cap) {
            if (g) gzwrite(g, buf, (unsigned)len); else fwrite(buf, 1, len, f);
            len = 0;
            if (2 * L + 256 > cap) { cap = 4 * L + 512; buf = (char *)realloc(buf, cap + 4096); }
        }
        len += (size_t)sprintf(buf + len, "@V300000001L1C001R%010llu#%s/%d\n",
                               (unsigned long long)read_no[i], bc_names + bc_name_off[bc_id[i]], mate);
        memcpy(buf + len, seqs + off[i], L); len += L;
        buf[len++] = '\n'; buf[len++] = '+'; buf[len++] = '\n';
        memset(buf + len, 'F', L); len += L;
        buf[len++] = '\n';
    }
    if (g) { gzwrite(g, buf, (unsigned)len); gzclose(g); } else { fwrite(buf, 1, len, f); fclose(f); }
    free(buf);
    return 0;
}

/* ---- multi-threaded writer (large synthetic files) -------------------------------------------------------
 * ff_open / ff_add / ff_close.  Each ff_add cuts its n reads into one range per thread; every range is
 * formatted and, when gz_level > 0, deflated on its own thread into raw DEFLATE blocks that end on a byte
 * boundary (Z_FULL_FLUSH), pigz-style; the pieces are written in order, so the file is ONE gzip member -- what
 * a sequencer ships and the hard case for a parallel decoder (no member or flush-point index to lean on:
 * block boundaries are not byte aligned in general, these happen to be every few tens of MB only).
 * CRC-32 of the member = crc32_combine over the pieces.                                                      */
#include <pthread.h>
#include <unistd.h>

typedef struct { FILE *f; int level; uLong crc; uint64_t isize; } ff_writer;

typedef struct {
    const uint8_t *seqs; const uint64_t *off; size_t lo, hi;
    const char *bc_names; const uint64_t *bc_name_off; const uint32_t *bc_id; const uint64_t *read_no;
    int mate, gz_level;
    char *out; size_t out_len; uLong crc; size_t text_len; int err;
} ff_job;

static void *ff_job_run(void *arg) {
    ff_job *j = (ff_job *)arg;
    size_t cap = 0;
    for (size_t i = j->lo; i < j->hi; i++) cap += 2 * (size_t)(j->off[i + 1] - j->off[i]) + 96;
    char *buf = (char *)malloc(cap + 64);
    if (!buf) { j->err = 1; return NULL; }
    size_t len = 0;
    for (size_t i = j->lo; i < j->hi; i++) {
        size_t L = (size_t)(j->off[i + 1] - j->off[i]);
        len += (size_t)sprintf(buf + len, "@V300000001L1C001R%010llu#%s/%d\n", (unsigned long long)j->read_no[i],
                               j->bc_names + j->bc_name_off[j->bc_id[i]], j->mate);
        memcpy(buf + len, j->seqs + j->off[i], L); len += L;
        buf[len++] = '\n'; buf[len++] = '+'; buf[len++] = '\n';
        memset(buf + len, 'F', L); len += L;
        buf[len++] = '\n';
    }
    j->text_len = len;
    if (!j->gz_level) { j->out = buf; j->out_len = len; return NULL; }
    uLong crc = crc32(0L, Z_NULL, 0);
    for (size_t p = 0; p < len; p += (1u << 30)) crc = crc32(crc, (const Bytef *)buf + p, (uInt)(len - p > (1u << 30) ? (1u << 30) : len - p));
    j->crc = crc;
    z_stream z;
    memset(&z, 0, sizeof z);
    if (deflateInit2(&z, j->gz_level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { free(buf); j->err = 1; return NULL; }
    size_t zcap = len + len / 8 + 4096;
    char *zb = (char *)malloc(zcap);
    if (!zb) { deflateEnd(&z); free(buf); j->err = 1; return NULL; }
    size_t in_done = 0, out_done = 0;
    for (;;) {                                                   /* zlib counts in 32 bits: feed in slices */
        size_t in_now = len - in_done; if (in_now > (1u << 30)) in_now = 1u << 30;
        size_t out_now = zcap - out_done; if (out_now > (1u << 30)) out_now = 1u << 30;
        const int last = in_done + in_now == len;
        z.next_in = (Bytef *)(buf + in_done); z.avail_in = (uInt)in_now;
        z.next_out = (Bytef *)(zb + out_done); z.avail_out = (uInt)out_now;
        int rc = deflate(&z, last ? Z_FULL_FLUSH : Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_BUF_ERROR) { j->err = 1; break; }
        in_done += in_now - z.avail_in;
        out_done += out_now - z.avail_out;
        if (last && z.avail_in == 0 && z.avail_out != 0) break;
        if (out_done == zcap) { j->err = 1; break; }
    }
    j->out_len = out_done;
    j->out = zb;
    deflateEnd(&z);
    free(buf);
    return NULL;
}

void *ff_open(const char *path, int gz_level) {
    ff_writer *w = (ff_writer *)calloc(1, sizeof *w);
    if (!w) return NULL;
    w->f = fopen(path, "wb");
    if (!w->f) { free(w); return NULL; }
    w->level = gz_level;
    w->crc = crc32(0L, Z_NULL, 0);
    if (gz_level) {
        static const unsigned char hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3};
        fwrite(hdr, 1, 10, w->f);
    }
    return w;
}

int ff_add(void *handle, const uint8_t *seqs, const uint64_t *off, size_t n, const char *bc_names,
           const uint64_t *bc_name_off, const uint32_t *bc_id, const uint64_t *read_no, int mate) {
    ff_writer *w = (ff_writer *)handle;
    if (!w || !w->f) return -1;
    long hw = sysconf(_SC_NPROCESSORS_ONLN);
    size_t t = (size_t)(hw > 0 ? hw : 1);
    if (t > 64) t = 64;
    if (t > n / 4096 + 1) t = n / 4096 + 1;
    ff_job jobs[64];
    pthread_t th[64];
    int started[64];
    for (size_t i = 0; i < t; i++) {
        ff_job j = {seqs, off, n * i / t, n * (i + 1) / t, bc_names, bc_name_off, bc_id, read_no, mate, w->level, NULL, 0, 0, 0, 0};
        jobs[i] = j;
        started[i] = pthread_create(&th[i], NULL, ff_job_run, &jobs[i]) == 0;
        if (!started[i]) ff_job_run(&jobs[i]);
    }
    int err = 0;
    for (size_t i = 0; i < t; i++) {
        if (started[i]) pthread_join(th[i], NULL);
        if (jobs[i].err) err = -1;
        if (!err && jobs[i].out_len && fwrite(jobs[i].out, 1, jobs[i].out_len, w->f) != jobs[i].out_len) err = -1;
        if (!err && w->level) {
            size_t left = jobs[i].text_len;                       /* crc32_combine takes a z_off_t length */
            uLong c = jobs[i].crc;
            (void)left;
            w->crc = crc32_combine(w->crc, c, (z_off_t)jobs[i].text_len);
            w->isize += jobs[i].text_len;
        }
        free(jobs[i].out);
    }
    return err;
}

int ff_close(void *handle) {
    ff_writer *w = (ff_writer *)handle;
    if (!w) return -1;
    int err = 0;
    if (w->f && w->level) {
        unsigned char tail[10] = {0x03, 0x00, 0, 0, 0, 0, 0, 0, 0, 0};   /* final empty fixed block, CRC-32, ISIZE */
        for (int i = 0; i < 4; i++) { tail[2 + i] = (unsigned char)(w->crc >> (8 * i)); tail[6 + i] = (unsigned char)(w->isize >> (8 * i)); }
        if (fwrite(tail, 1, 10, w->f) != 10) err = -1;
    }
    if (w->f && fclose(w->f)) err = -1;
    free(w);
    return err;
}
