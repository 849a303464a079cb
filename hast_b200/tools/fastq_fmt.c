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
