/*
 * hast_oracle.h -- CPU restatement of HAST stage 01 (classify + mergeResult).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library, and
 * only as the checker.  The product (libhast_b200.so, bin/classify) never
 * links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py)
 * against (1) every known answer in the reference's TestAll()
 * (classify.cpp:341-367) and (2) the byte-exact stdout of the untouched
 * reference binaries compiled into oracle/_ref/ (see oracle/Makefile), both
 * live when oracle/_ref exists and through committed fixtures in tests/golden/.
 *
 * All citations are relative to /root/reference/01.classify_stlfr_reads/ .
 */
#ifndef HAST_ORACLE_H
#define HAST_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- k-mer arithmetic (kmer/kmer.h) ------------------------------------ */
/* kmer.h:11   base -> 2-bit code, (c & 6) >> 1 : A0 C1 T2 G3, any byte maps */
int      ho_base2int(unsigned char c);
/* kmer.h:12   code -> letter, "ACTG"[code] */
char     ho_int2base(int code);
/* kmer.h:196-223 reverse complement of a packed k-mer, 1 <= k <= 32 */
uint64_t ho_revcomp(uint64_t w, int k);
/* kmer.h:153-166 canonical packed k-mer of exactly k letters */
uint64_t ho_str2kmer(const char *s, int k);
/* kmer.h:169-194 all len-k+1 canonical k-mers of a read; returns the count,
 * or -1 when len < k (the reference asserts, kmer.h:171) */
long     ho_chop(const char *read, long len, int k, uint64_t *out);
/* kmer.h:244-254 + kmer.h:14-25 packed k-mer -> letters (out has k+1 bytes) */
void     ho_kmer2str(uint64_t w, int k, char *out);

/* ---- header parsing / N scan (classify.cpp) ---------------------------- */
/* classify.cpp:112-119 barcode = head.substr(s+1, e-s-1), s/e = LAST '#'/'/' */
void     ho_parse_name(const char *head, long len, long *start, long *blen);
/* classify.cpp:182-185 true iff some byte == 'N' */
int      ho_contain_n(const char *seq, long len);
/* classify.cpp:66-86 haplotype call; has0/has1 = "key present in the map" */
int      ho_get_hap(const char *barcode, int has0, int c0, int has1, int c1,
                    size_t n0, size_t n1, double w0, double w1);

/* ---- classifier state (g_kmers[2], g_K, weights) ----------------------- */
typedef struct ho_classifier ho_classifier;

ho_classifier *ho_create(void);
void           ho_destroy(ho_classifier *c);
const char    *ho_error(const ho_classifier *c);

/* classify.cpp:30-46 load a k-mer list held in memory (one k-mer per line).
 * index 0 fixes k from its first line.  Returns the number of lines recorded
 * ("total_kmer"), or -1 on a line whose length != k (reference: assert). */
long   ho_load_kmers_mem(ho_classifier *c, const char *text, size_t n, int index);
long   ho_load_kmers_file(ho_classifier *c, const char *path, int index);
/* harness convenience: packed words (kmer.h code) instead of text; canonicalised and inserted like load_kmers */
long   ho_load_kmers_packed(ho_classifier *c, const uint64_t *kmers, size_t n, int k, int index);
/* classify.cpp:314-339 erase the adaptors' canonical k-mers from both sets;
 * returns the number of erased set members, -1 if an adaptor is shorter than k */
long   ho_init_adaptor(ho_classifier *c, const char *fwd, const char *rev);
int    ho_k(const ho_classifier *c);
size_t ho_set_size(const ho_classifier *c, int index);
/* membership bits of a canonical k-mer: bit0 = in hap0 set, bit1 = in hap1 */
int    ho_lookup(const ho_classifier *c, uint64_t canonical);
void   ho_set_weights(ho_classifier *c, double w0, double w1);

/* classify.cpp:186-209 one read given header and sequence (string level) */
int    ho_process_read(ho_classifier *c, const char *head, long hlen,
                       const char *seq, long slen);
/* classify.cpp:238-278 one FASTQ file (.gz by suffix) */
int    ho_process_fastq(ho_classifier *c, const char *path);
/* classify.cpp:93-102 table in bytewise barcode order */
int    ho_print(ho_classifier *c, FILE *out);
int    ho_print_file(ho_classifier *c, const char *path);
size_t ho_n_barcodes(const ho_classifier *c);

/* Dense form of process_reads used to check the device batch interface:
 * reads are bases[read_off[i] .. read_off[i+1]), barcode ids are dense.
 * counts is int32[n_barcodes][2], accumulated (not cleared).  Returns the
 * number of k-mer lookups done, or -1 if some read is shorter than k.
 * nthreads > 1 splits the reads over pthreads with private partial counts. */
long long ho_classify_batch(const ho_classifier *c, const uint8_t *bases,
                            const uint64_t *read_off, const uint32_t *barcode_id,
                            size_t n_reads, int32_t *counts, size_t n_barcodes,
                            int nthreads);

/* ---- mergeResult (mergeResult.cpp) ------------------------------------- */
/* mergeResult.cpp:21-30,33-69,83-131.  intended = 0 reproduces the shipped
 * behaviour (both count columns are summed into key 0);  intended = 1 sums
 * the columns separately and applies the ratio rule with float weights. */
int    ho_merge_result(const char *const *inputs, int n_inputs, float w0,
                       float w1, int intended, const char *out_path);

#ifdef __cplusplus
}
#endif
#endif
