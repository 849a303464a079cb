/*
 * hast_b200.h -- C ABI of the B200 read-classification engine (libhast_b200.so)
 *
 * The reference (BGI-Qingdao/HAST, 01.classify_stlfr_reads) has no plugin or
 * FFI interface: `classify` is one C++ translation unit.  This ABI cuts that
 * program along its own function boundaries; every entry point names the
 * reference code it replaces (paths relative to 01.classify_stlfr_reads/).
 * bin/classify (hast_b200/host/) is the drop-in process built on top of it.
 *
 * Conventions
 *   - plain C: opaque context, plain pointers and sizes, no C++/torch types.
 *   - every call returns 0 on success or a negative HAST_E_* code; the text of
 *     the last error is available from hast_last_error().
 *   - one context per GPU; a context is driven by one host thread at a time.
 *   - "host" pointers may be pageable or pinned (hast_host_alloc); copies from
 *     pinned memory are asynchronous and double-buffered on the device side.
 *   - there is NO CPU implementation behind this interface: without a CUDA
 *     device every compute entry point fails with HAST_E_CUDA.
 */
#ifndef HAST_B200_H
#define HAST_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HAST_ABI_VERSION 6

#define HAST_OK            0
#define HAST_E_ARG        -1   /* bad argument                                            */
#define HAST_E_CUDA       -2   /* CUDA runtime / driver failure, or no device             */
#define HAST_E_STATE      -3   /* call out of order (e.g. lookup before table build)      */
#define HAST_E_KMER_LINE  -4   /* k-mer list line whose length != k  (kmer.h:154 assert)  */
#define HAST_E_SHORT_READ -5   /* read shorter than k                (kmer.h:171 assert)  */
#define HAST_E_TABLE_FULL -6   /* table capacity exceeded: rebuild with a larger one      */
#define HAST_E_NCCL       -7   /* NCCL unavailable or failed                              */
#define HAST_E_K          -8   /* k outside 1..32                                         */

typedef struct hast_ctx hast_ctx;

typedef struct hast_table_info {
    int32_t  k;
    int32_t  log2_buckets;
    uint64_t n_buckets;        /* 32-byte buckets of four 8-byte slots                    */
    uint64_t bytes;
    uint64_t n_entries;        /* distinct canonical k-mers stored (either parent)        */
    uint64_t n_displaced;      /* entries that live outside their home bucket             */
    uint64_t n_overflow_buckets;
    uint64_t size[2];          /* g_kmers[i].size() after erases                          */
    uint64_t filter_bytes;     /* L2-resident Bloom pre-filter in front of the table      */
} hast_table_info;

typedef struct hast_stats {
    uint64_t batches;          /* submit calls                                            */
    uint64_t reads;            /* reads submitted                                         */
    uint64_t bases;            /* bytes of read sequence submitted                        */
    uint64_t lookups;          /* k-mer positions of non-N reads (two set probes = one)   */
    uint64_t reads_with_n;     /* reads skipped by containN                               */
    uint64_t reads_short;      /* reads shorter than k (an error in the reference)        */
    uint64_t extra_probes;     /* bucket reads beyond the first one of a lookup           */
    uint64_t kernel_launches;  /* launches of this library's own kernels                  */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t filter_pass;      /* lookups the pre-filter sent on to the exact table       */
    uint64_t filter_loads;     /* 8-byte pre-filter words fetched (minimizer sweep: < lookups) */
    uint64_t finish_reduce_us; /* hast_finish: the ncclReduce of the counts, CUDA-event time (cumulative) */
    uint64_t finish_d2h_us;    /* hast_finish: the read-back of the counts to the host (cumulative)        */
} hast_stats;

/* ---- library / context ------------------------------------------------- */
int          hast_abi_version(void);
int          hast_device_count(void);
/* MultiThread ctor/dtor, classify.cpp:164-181: one worker per GPU. */
int          hast_create(int device, hast_ctx **out);
void         hast_destroy(hast_ctx *ctx);
const char  *hast_last_error(const hast_ctx *ctx);      /* ctx may be NULL */
int          hast_device(const hast_ctx *ctx);
/* Tuning knobs, set before hast_table_begin: "kernel" (3 = pre-filtered fused
 * kernel whose filter word is chosen by the k-mer's minimizer, default for
 * k = 21/25/31; 1 = filter word chosen by a hash of the k-mer, what every
 * other k runs; 2 / 4 = as 1 / 3 with TMA-staged reads; 0 = direct table probe per
 * position; 1/2 <-> 3/4 takes effect at the next hast_table_begin), "filter_bits_per_key"
 * (default 16), "filter_max_bytes" (default 64 MiB: the filter is meant to stay
 * L2-resident).  None of them changes any result.  "seq_mode" = 1 switches the
 * fused kernel to the window rule of HAST stage 03 (03.mkoutput_by_fabulous2.0/
 * src_main/classify.cpp:203-218, string k-mers): a k-mer position votes iff all
 * its k bytes are upper-case A/C/G/T, other bytes do not silence the rest of the
 * sequence, and sequences shorter than k are not an error.
 * Launch-side knobs, effective at once: "reads_per_tile" (0, the default = as many
 * reads as fill one pass of the fused kernel; 1..416 fixes it), "l2_persist_bytes"
 * (L2 set aside for the pre-filter's evict_last loads, device-wide; default 64 MiB,
 * clamped to the device maximum), "l2_fetch_granularity" (32 / 64 / 128),
 * "host_pack_threads" (0 = hast_submit_batch copies the ASCII bytes; N > 0 = it
 * packs them to 2 bits on N host threads first and copies a quarter).           */
int          hast_set_option(hast_ctx *ctx, const char *name, int64_t value);
/* pinned host memory for batch buffers */
int          hast_host_alloc(void **ptr, size_t bytes);
int          hast_host_free(void *ptr);

/* ---- K1: parent-unique k-mer table ------------------------------------- */
/* load_kmers, classify.cpp:30-46 + Kmer::str2Kmer, kmer.h:153-166.
 * begin: fix k (1..32) and the expected number of distinct keys (sizes the
 * table at load <= 0.5); any previous table and all counts are dropped.      */
int hast_table_begin(hast_ctx *ctx, int k, uint64_t expected_keys);
/* add_text: `text` holds n_lines k-mers, each exactly k letters followed by
 * '\n' (the jellyfish-dump format of build_unshared_kmers.sh:290-291, already
 * cut at the last '\n' by the caller per classify.cpp:41).  Letters are packed
 * with (c & 6) >> 1 (kmer.h:11), canonicalised on the device and inserted with
 * parent tag `parent` (0 = hap0/paternal, 1 = hap1/maternal).                 */
int hast_table_add_text(hast_ctx *ctx, const char *text, uint64_t n_lines, int parent);
/* add_packed: the same for already packed k-mers (MSB-first, kmer.h:156-160);
 * they need not be canonical.                                                 */
int hast_table_add_packed(hast_ctx *ctx, const uint64_t *kmers, uint64_t n, int parent);
/* InitAdaptor, classify.cpp:314-339: erase every canonical k-mer of `seq`
 * from both parents' sets.  erased_out (may be NULL) receives up to cap packed
 * canonical k-mers that were members, with their former tag in tags_out, for
 * the "INFO : erase a adaptor kmer" log lines (classify.cpp:321-336).          */
int hast_table_erase_seq(hast_ctx *ctx, const char *seq, uint32_t len,
                         uint64_t *erased_out, uint8_t *tags_out, uint32_t cap,
                         uint32_t *n_erased);
/* g_kmers[i].size(), classify.cpp:70-71 (distinct, after erases) + layout.    */
int hast_table_info_get(hast_ctx *ctx, hast_table_info *out);
/* Replicate a built table on another GPU of the same process (peer copy over
 * NVLink) instead of rebuilding it there.                                     */
int hast_table_clone(hast_ctx *dst, hast_ctx *src);

/* ---- K4 state: dense per-barcode counters ------------------------------ */
/* BarcodeCache, classify.cpp:50-64, as int32 counts[n][2] on the device;
 * grows preserving contents.  Barcode strings stay with the caller.          */
int hast_reserve_barcodes(hast_ctx *ctx, uint64_t n_barcodes);
int hast_reset_counts(hast_ctx *ctx);

/* ---- K2+K3+K4 fused: one batch of reads -------------------------------- */
/* MultiThread::submit + process_reads, classify.cpp:186-219.
 * Read i is bases[read_off[i] .. read_off[i+1]) (ASCII, exactly the FASTQ
 * sequence line), its barcode is the dense id barcode_id[i] < n_barcodes.
 * Per read: any 'N' => no votes (classify.cpp:182-185,190-193); otherwise every
 * k-mer position probes the table once and adds the tag bits to the read's
 * two votes, which are then added to counts[barcode][0/1].
 * Host form: asynchronous; *ticket (may be NULL) identifies the batch.        */
int hast_submit_batch(hast_ctx *ctx, const uint8_t *bases, uint64_t n_bases,
                      const uint32_t *read_off, const uint32_t *barcode_id,
                      uint32_t n_reads, uint64_t *ticket);
/* Block until the host buffers of batch `ticket` may be reused.               */
int hast_wait_copied(hast_ctx *ctx, uint64_t ticket);
/* Device form: the same batch with all three arrays already resident in this
 * context's GPU memory (no copy); runs on the context's compute stream.       */
int hast_submit_batch_device(hast_ctx *ctx, const uint8_t *d_bases, uint64_t n_bases,
                             const uint32_t *d_read_off, const uint32_t *d_barcode_id,
                             uint32_t n_reads);
/* The same batch as the host parser can hand it over when it packs while it
 * parses: `packed` holds the concatenated reads at 2 bits per base with the
 * reference's code (c & 6) >> 1 (kmer.h:11), 16 bases per 32-bit word, first
 * base in the two most significant bits (kmer.h:156-160), reads back to back
 * (no per-read alignment), the last word zero-padded; read_off[] counts BASES;
 * bit (i & 31) of has_n[i >> 5] is set iff read i contains the byte 'N'
 * (containN, classify.cpp:182-185) -- such a read casts no votes.  A quarter of
 * the host-to-device bytes of hast_submit_batch; results are identical.       */
int hast_submit_batch_packed(hast_ctx *ctx, const uint32_t *packed, uint64_t n_bases,
                             const uint32_t *read_off, const uint32_t *barcode_id,
                             const uint32_t *has_n, uint32_t n_reads, uint64_t *ticket);
int hast_submit_batch_packed_device(hast_ctx *ctx, const uint32_t *d_packed, uint64_t n_bases,
                                    const uint32_t *d_read_off, const uint32_t *d_barcode_id,
                                    const uint32_t *d_has_n, uint32_t n_reads);
/* The packing step on its own, on the host (no device, no context): ASCII bases ->
 * the words and has_n bits hast_submit_batch_packed takes, on `threads` host
 * threads.  Same code as kmer.h:11 (base2int) / :156-160 (MSB-first order) and
 * containN, classify.cpp:182-185.  words_out: (n_bases + 15) / 16 words,
 * has_n_out: (n_reads + 31) / 32 words.  With the option "host_pack_threads" > 0
 * hast_submit_batch does this itself and copies a quarter of the bytes.          */
int hast_pack_bases(const uint8_t *bases, uint64_t n_bases, const uint32_t *read_off,
                    uint32_t n_reads, uint32_t *words_out, uint32_t *has_n_out, int threads);
int hast_sync(hast_ctx *ctx);

/* ---- finish: collect the per-barcode counts ---------------------------- */
/* MultiThread::wait + collectBarcodes + BarcodeCache::Add,
 * classify.cpp:220-229,57-63,276-277.  Synchronises; when a communicator is
 * attached, sums the partial counts of all ranks with ONE ncclReduce(int32,sum)
 * to rank 0 over NVLink; copies counts[n_barcodes][2] to counts_out on rank 0
 * (other ranks may pass NULL).  Fails with HAST_E_SHORT_READ if any read was
 * shorter than k.                                                             */
int hast_finish(hast_ctx *ctx, int32_t *counts_out, uint64_t n_barcodes);
int hast_stats_get(hast_ctx *ctx, hast_stats *out);
/* raw device pointer of the int32 counts[n][2] array (for zero-copy views)    */
int hast_counts_device_ptr(hast_ctx *ctx, void **ptr, uint64_t *n_barcodes);
/* CUDA event timing of the work submitted between the two marks (ms), on the
 * context's compute stream.                                                   */
int hast_timer_start(hast_ctx *ctx);
int hast_timer_stop(hast_ctx *ctx, float *ms);

/* ---- multi-GPU ---------------------------------------------------------- */
/* single process, one context per GPU (bin/classify): ncclCommInitAll         */
int hast_comm_init_all(hast_ctx **ctxs, int n);
/* one process per GPU (torchrun): 128-byte ncclUniqueId made on rank 0 and
 * distributed by the caller                                                   */
int hast_comm_unique_id(void *id128);
int hast_comm_init_rank(hast_ctx *ctx, int nranks, int rank, const void *id128);

/* ---- standalone K2 / K3 (parity tests, ncu evidence) ------------------- */
/* Kmer::chopRead2Kmer, kmer.h:169-194: canonical k-mers of every read, written
 * to kmers_out[kmer_off[i] + j] with kmer_off[i] = read_off[i] - i*(k-1)
 * (reads shorter than k or containing 'N' are still extracted when long
 * enough; has_n_out[i] = containN).  Host buffers.                            */
int hast_extract_kmers(hast_ctx *ctx, const uint8_t *bases, uint64_t n_bases,
                       const uint32_t *read_off, uint32_t n_reads,
                       uint64_t *kmers_out, uint64_t n_kmers_out, uint8_t *has_n_out);
/* g_kmers[0/1].find, classify.cpp:195-202: tag bits (bit0 hap0, bit1 hap1) of
 * n canonical packed k-mers.  Host buffers.                                   */
int hast_lookup(hast_ctx *ctx, const uint64_t *canonical, uint64_t n, uint8_t *tags_out);
/* Device-resident forms used by bench.py / ncu: time only the kernel.         */
int hast_extract_kmers_device(hast_ctx *ctx, const uint8_t *d_bases, uint64_t n_bases,
                              const uint32_t *d_read_off, uint32_t n_reads,
                              uint64_t *d_kmers_out, uint8_t *d_has_n_out);
int hast_lookup_device(hast_ctx *ctx, const uint64_t *d_canonical, uint64_t n, uint8_t *d_tags_out);
/* Random 32-byte-sector gather over the table's own memory: the measured
 * random-access roofline the lookup kernel is compared with.  Returns the
 * achieved GB/s of n_probes independent sector reads.                         */
int hast_gather_roofline(hast_ctx *ctx, uint64_t n_probes, uint64_t span_bytes, float *gbps);

/* ---- stage 00: parent-unique k-mer lists from parental reads ------------ */
/* Replaces 00.build_unshare_kmers_by_jellyfish/build_unshared_kmers.sh:163-291 (five
 * `jellyfish count -m K -C` runs, the dumps between them and the "2 copies of maternal +
 * 1 copy of paternal" mix, :262-291) by one device table of per-parent counts.
 * K-mers cross this part of the ABI in jellyfish's code (A0 C1 G2 T3, first base in the
 * highest used bits), canonical = the numerically smaller of a k-mer and its reverse
 * complement = the string `jellyfish dump` prints.                                    */
typedef struct hast_kc_info {
    int32_t  k;
    uint32_t part, n_parts;   /* this table keeps the keys of partition `part` of `n_parts` */
    uint64_t n_slots;         /* 16-byte slots                                          */
    uint64_t bytes;
    uint64_t occupied;        /* distinct canonical k-mers stored (either parent)       */
    uint64_t distinct[2];     /* ... seen in the paternal / maternal reads              */
    uint64_t both;            /* ... seen in both                                       */
    uint64_t occurrences[2];  /* sum of counts per parent                               */
    uint64_t windows;         /* valid k-mer windows streamed past (all partitions)     */
    uint64_t table_full;      /* insertions that found no slot: rebuild larger / more partitions */
} hast_kc_info;
/* begin: k in 1..32; the table gets the smallest power-of-two number of slots >= 2 *
 * expected_distinct (load <= 0.5).  n_parts > 1: only canonical k-mers whose partition
 * (top bits of a 64-bit mix) equals `part` are counted, the others are skipped -- stream
 * the same reads once per partition, or to several GPUs that each own one.            */
int hast_kc_begin(hast_ctx *ctx, int k, uint64_t expected_distinct, uint32_t part, uint32_t n_parts);
/* `jellyfish count -C` over one batch of sequences of parent `parent` (0 paternal,
 * 1 maternal): sequence i is bases[seq_off[i] .. seq_off[i+1]) (ASCII; a sequence line of a
 * FASTQ record or the joined lines of a FASTA record).  A sequence longer than 32768 bytes
 * must be cut by the caller into chunks that overlap by k-1 bytes.  Every window of k
 * bytes all in ACGTacgt counts once.  Host form: asynchronous, see hast_wait_copied.   */
int hast_kc_add(hast_ctx *ctx, const uint8_t *bases, uint64_t n_bases, const uint32_t *seq_off,
                uint32_t n_seqs, int parent, uint64_t *ticket);
int hast_kc_add_device(hast_ctx *ctx, const uint8_t *d_bases, uint64_t n_bases, const uint32_t *d_seq_off,
                       uint32_t n_seqs, int parent);
int hast_kc_info_get(hast_ctx *ctx, hast_kc_info *out);
/* `jellyfish histo -h high` of one parent (analysis_kmercount.sh:7-9): histo[c] = number
 * of distinct k-mers with count c for 1 <= c <= high, histo[high+1] = those above,
 * histo[0] = all distinct k-mers of the parent.  histo holds high + 2 entries.        */
int hast_kc_histo(hast_ctx *ctx, int parent, uint32_t high, uint64_t *histo);
/* The k-mers of `parent` with lower <= count <= upper (`jellyfish dump -L -U`), and, when
 * require_unique, absent from the other parent (build_unshared_kmers.sh:262-291), sorted
 * ascending, up to cap of them into out; *n receives how many there are.              */
int hast_kc_select(hast_ctx *ctx, int parent, uint32_t lower, uint32_t upper, int require_unique,
                   uint64_t *out, uint64_t cap, uint64_t *n);
/* Build the classification table of stage 01 (hast_table_begin + both parents' lists)
 * straight from the count table of `src`, on the device, without the text round trip;
 * bounds as in the script: paternal [pl,pu], maternal [ml,mu].  dst may equal src.     */
int hast_kc_to_table(hast_ctx *dst, hast_ctx *src, uint32_t pl, uint32_t pu, uint32_t ml, uint32_t mu);
int hast_kc_end(hast_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
