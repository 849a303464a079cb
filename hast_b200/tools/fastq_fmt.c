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

/* ---- appending, multi-threaded variant (large synthetic files) ----------------------------------------
 * Appends reads to `path`.  The n reads are cut into one range per thread; each range is formatted and,
 * when gz_level > 0, deflated into its own gzip member in memory; the pieces are then appended in order.
 * A file of concatenated members is what `cat a.gz b.gz` or bgzip produce; zlib's gzread (the
 * reference's igzstream) walks them transparently.                                                      */
#include <pthread.h>
#include <unistd.h>

typedef struct {
    const uint8_t *seqs; const uint64_t *off; size_t lo, hi;
    const char *bc_names; const uint64_t *bc_name_off; const uint32_t *bc_id; const uint64_t *read_no;
    int mate, gz_level;
    char *out; size_t out_len; int err;
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
    if (!j->gz_level) { j->out = buf; j->out_len = len; return NULL; }
    z_stream z;
    memset(&z, 0, sizeof z);
    if (deflateInit2(&z, j->gz_level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) { free(buf); j->err = 1; return NULL; }
    size_t zcap = deflateBound(&z, (uLong)len) + 64;
    char *zb = (char *)malloc(zcap);
    if (!zb) { deflateEnd(&z); free(buf); j->err = 1; return NULL; }
    size_t in_done = 0;
    z.next_out = (Bytef *)zb;
    size_t out_left = zcap;
    int rc = Z_OK;
    while (rc != Z_STREAM_END) {                                /* zlib counts in 32 bits: feed in slices */
        size_t in_now = len - in_done; if (in_now > (1u << 30)) in_now = 1u << 30;
        size_t out_now = out_left > (1u << 30) ? (1u << 30) : out_left;
        z.next_in = (Bytef *)(buf + in_done); z.avail_in = (uInt)in_now; z.avail_out = (uInt)out_now;
        rc = deflate(&z, in_done + in_now == len ? Z_FINISH : Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR) { j->err = 1; break; }
        in_done += in_now - z.avail_in;
        out_left -= out_now - z.avail_out;
    }
    j->out_len = zcap - out_left;
    j->out = zb;
    deflateEnd(&z);
    free(buf);
    return NULL;
}

int ff_append_fastq(const char *path, int gz_level, const uint8_t *seqs, const uint64_t *off, size_t n,
                    const char *bc_names, const uint64_t *bc_name_off, const uint32_t *bc_id,
                    const uint64_t *read_no, int mate) {
    long hw = sysconf(_SC_NPROCESSORS_ONLN);
    size_t t = (size_t)(hw > 0 ? hw : 1);
    if (t > 64) t = 64;
    if (t > n / 4096 + 1) t = n / 4096 + 1;
    ff_job jobs[64];
    pthread_t th[64];
    for (size_t i = 0; i < t; i++) {
        ff_job j = {seqs, off, n * i / t, n * (i + 1) / t, bc_names, bc_name_off, bc_id, read_no, mate, gz_level, NULL, 0, 0};
        jobs[i] = j;
        if (pthread_create(&th[i], NULL, ff_job_run, &jobs[i])) { ff_job_run(&jobs[i]); th[i] = 0; }
    }
    FILE *f = fopen(path, "ab");
    int err = f ? 0 : -1;
    for (size_t i = 0; i < t; i++) {
        if (th[i]) pthread_join(th[i], NULL);
        if (jobs[i].err) err = -1;
        if (!err && jobs[i].out_len && fwrite(jobs[i].out, 1, jobs[i].out_len, f) != jobs[i].out_len) err = -1;
        free(jobs[i].out);
    }
    if (f) fclose(f);
    return err;
}
