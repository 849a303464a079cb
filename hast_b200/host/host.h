// host.h -- host side of bin/classify: the reference's process interface
// (01.classify_stlfr_reads/classify.cpp) rebuilt as a streaming pipeline
//     reader (read/inflate) -> parser threads -> pinned batches -> GPU contexts
// on top of the C ABI in include/hast_b200.h.  Nothing here computes k-mers or
// looks anything up: that happens only on the device.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace hasthost {

// ---- command line (classify.cpp:375-428) -----------------------------------
struct Options {
    std::string hap0, hap1;
    std::vector<std::string> reads;
    int threads = 8;                       // -t: here the number of FASTQ parser threads
    double weight0 = 1.0, weight1 = 1.0;   // classify.cpp:22-23
    std::string adaptor_f = "CTGTCTCTTATACACATCTTAGGAAGACAAGCACTGACGACATGA";   // classify.cpp:312
    std::string adaptor_r = "TCTGCTGAGTCGAGAACGTCTCTGTGAGCCAAGGAGTTGCTCTGG";   // classify.cpp:313
    // extensions (long options only; the reference rejects them, so no clash)
    int gpus = 0;                          // 0 = default (one GPU)
    std::string stats_json;                // throughput side output
    bool split_barcodes = false;           // also write the three *.unique.barcodes lists (script :156-162)
    bool partition_reads = false;          // also partition every input FASTQ (script :176-185); implies split
    std::string outdir;                    // where those files go ("" = cwd, like the script)
    bool packed_h2d = true;                // parser packs to 2 bits before the copy (env HAST_PACKED=0: ASCII)
    size_t batch_bytes = 8u << 20;         // raw FASTQ text per parse block
};
// returns 0 to run, otherwise the process exit code (255 after printing usage)
int parse_options(int argc, char** argv, Options& opt);
void print_usage();

// ---- k-mer list files (classify.cpp:30-46) ----------------------------------
struct KmerList {
    std::string text;        // n_lines lines of k letters + '\n', ready for hast_table_add_text
    uint64_t n_lines = 0;    // "total_kmer" of classify.cpp:45
    int k = 0;
};
// index 0 derives k from the first line; index 1 takes k from the caller.
// Returns "" or an error message.
std::string load_kmer_list(const std::string& path, int index, int k_in, KmerList& out);

// ---- barcode interning ------------------------------------------------------
// Barcodes are arbitrary byte strings (tools/mark_library.sh:24 makes lib2_a_b_c);
// every distinct one seen gets a dense global id, whether or not it ever scores
// (classify.cpp:191,208 create the map entry through key -1).
class BarcodeIndex {
public:
    BarcodeIndex();
    ~BarcodeIndex();
    uint32_t intern(const char* s, size_t n);
    // the same in two steps, so that a parser can have the table line of record i+8 on its way while it
    // resolves record i (the table of a 20 M-barcode run is far larger than the caches)
    static uint64_t hash(const char* s, size_t n);
    void prefetch(uint64_t h) const;
    uint32_t intern_hashed(uint64_t h, const char* s, size_t n);
    uint32_t size() const { return next_id_.load(std::memory_order_acquire); }
    // names in id order (call after all parsing finished)
    void export_names(std::vector<std::string>& out) const;
private:
    struct Shard;
    struct Arena;
    static constexpr int kShards = 256;
    std::unique_ptr<Arena> arena_;                         // declared first: outlives the shards' tables
    std::unique_ptr<Shard[]> shards_;
    std::atomic<uint32_t> next_id_{0};
};

// ---- FASTQ blocks -------------------------------------------------------------
// A block holds whole records only (the reader cuts at a newline where the line
// count is a multiple of four, classify.cpp:257-269 framing).
struct TextBlock {
    std::vector<char> data;      // owned storage (stream readers); unused when view is set
    const char* view = nullptr;  // records live in a file mapping: view[0, len)
    size_t begin = 0;            // owned storage: the records are data[begin, begin + len)
    size_t len = 0;
    bool last_of_file = false;   // the final block may end in a partial record / unterminated line
    // optional: offsets of every '\n' of the block (relative to its start), made by the producer in the pass
    // that framed the block -- nl[nl_begin, nl_begin + nl_count)
    std::vector<uint32_t> nl;
    size_t nl_begin = 0, nl_count = 0;
    bool has_nl = false;
    std::shared_ptr<void> hold;  // keeps the memory behind `view` alive (a decoder's output buffer); reset() when parsed
    const char* text() const { return view ? view : data.data() + begin; }
};
// Offsets of every '\n' in p[0, n), appended to out from index `at` on (out is grown as needed); returns the
// number found.  One AVX2 pass.
size_t newline_index(const char* p, size_t n, std::vector<uint32_t>& out, size_t at);

// One parsed batch in pinned memory, laid out for hast_submit_batch.
struct Batch {
    uint8_t* bases = nullptr;   size_t cap_bases = 0;  uint64_t n_bases = 0;
    uint32_t* read_off = nullptr; uint32_t* barcode_id = nullptr; size_t cap_reads = 0;
    uint32_t n_reads = 0;
    uint32_t max_barcode = 0;
    std::string error;
    // packed mode (hast_submit_batch_packed): the parser emits 2-bit words + a containN bit per
    // read instead of the ASCII bases -- a quarter of the pinned-memory and PCIe traffic
    uint32_t* packed = nullptr;   size_t cap_words = 0;
    uint32_t* has_n = nullptr;    // (cap_reads + 31) / 32 words
};

// classify.cpp:112-119
inline void parse_name(const char* head, size_t len, size_t& start, size_t& blen) {
    // last '#' and last '/' of the header: the barcode sits at its end, so look from there
    long s = -1, e = -1;
    for (size_t i = len; i-- > 0;) {
        const char c = head[i];
        if (c == '/' && e < 0) e = (long)i;
        if (c == '#' && s < 0) s = (long)i;
        if (s >= 0 && e >= 0) break;
    }
    long cnt = e - s - 1;
    start = (size_t)(s + 1);
    blen = (cnt < 0 || (size_t)(s + 1 + cnt) > len) ? len - (size_t)(s + 1) : (size_t)cnt;
}

// Parse one block into a batch.  Returns false on a framing error that the
// reference would have died on (message in batch.error).
// resume (optional, in/out): offset inside the block where parsing starts; on return the offset where it
// stopped -- blk.len unless the batch filled up first (very short records), in which case the caller
// submits this batch and calls again with a fresh one.
bool parse_block(const TextBlock& blk, BarcodeIndex& index, Batch& out, size_t* resume = nullptr);
// Append `n` ASCII bases to a 2-bit MSB-first stream (kmer.h:11 code, kmer.h:156-160 order).
// `acc`/`nbits` carry the partially filled word between calls; returns true if an 'N' was seen.
bool pack_append(const char* seq, size_t n, uint32_t* words, size_t& n_words, uint64_t& acc, unsigned& nbits);

// ---- haplotype call + table output (classify.cpp:66-102) ----------------------
int get_hap(const std::string& barcode, int c0, int c1, uint64_t n0, uint64_t n1, double w0, double w1);
// order_out / hap_out (optional): the barcode ids in output order and their calls
void print_table(FILE* out, const std::vector<std::string>& names, const int32_t* counts, uint64_t n0,
                 uint64_t n1, double w0, double w1, std::vector<uint32_t>* order_out = nullptr,
                 std::vector<int8_t>* hap_out = nullptr);

// ---- stage-01 post-processing (classify_stlfr_reads.sh:156-185, quartering_fastq.awk) ----
// The three awk associative arrays of quartering_fastq.awk:12-18 as one table.
class BarcodeLists {
public:
    enum { kPaternal = 1, kMaternal = 2, kHomozygous = 3 };
    // one list line (no '\n'); the key is its $1 under FS '#|/'
    void add(const char* line, size_t n, int type);
    std::string load(const std::string& path, int type);      // "" or an error message
    int find(const char* key, size_t n) const;                // 0 = in no list
    size_t size() const { return count_; }
private:
    static constexpr uint32_t kEmpty = 0xFFFFFFFFu;
    struct Slot { uint64_t hash = 0; uint64_t off = 0; uint32_t len = kEmpty; uint8_t types = 0; };
    void grow();
    std::vector<Slot> slots_;
    std::vector<char> arena_;
    size_t count_ = 0;
};
struct PartitionStats {
    uint64_t total = 0, no_barcode = 0, paternal = 0, maternal = 0, homozygous = 0, unclassified = 0;
    uint64_t text_bytes = 0;
};
// {paternal,maternal,homozygous}.unique.barcodes from the table rows (barcode, call, counts) in
// output order, with the awk field rules of classify_stlfr_reads.sh:157,160,163; lists (optional)
// receives exactly the lines written.  counts_out = lines per file.
std::string write_barcode_lists(const std::string& dir, const std::vector<std::string>& names,
                                const std::vector<uint32_t>& order, const std::vector<int8_t>& haps,
                                const int32_t* counts, uint64_t counts_out[3], BarcodeLists* lists);
// quartering_fastq.awk over one FASTQ(.gz | "-"): PREFIX.{paternal,maternal,homozygous,nobarcode}.fastq
// and filter_reads.log (appended) in outdir ("" = cwd).  display_name is awk's FILENAME.
// threads <= 0: all cores (at most 16)
std::string partition_fastq(const std::string& input, const std::string& display_name, const std::string& prefix,
                            const std::string& outdir, const BarcodeLists& lists, PartitionStats& st, int threads = 0);
std::string partition_prefix(const std::string& path);       // basename, ".gz" stripped (:178-180)

// ---- the pipeline ---------------------------------------------------------------
struct RunStats {
    uint64_t reads = 0, bases = 0, barcodes = 0, lookups = 0, text_bytes = 0;
    double t_table = 0, t_reads = 0, t_finish = 0, t_print = 0, t_total = 0;
    int gpus = 0, parser_threads = 0;
    uint64_t table_bytes = 0, size0 = 0, size1 = 0, kernel_launches = 0;
    double t_split = 0, t_partition = 0;
    uint64_t partition_text_bytes = 0;
};
int run_classify(const Options& opt, RunStats& st);

}  // namespace hasthost
