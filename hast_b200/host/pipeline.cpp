// pipeline.cpp -- the streaming host pipeline of bin/classify.
//
//   plain files     : every parser thread claims slices of the file, pread()s and frames them itself
//                     (plain_slicer.h: exact four-line framing by newline count, any number of threads per file)
//   gzip / pipes    : reader threads inflate (inflate_par.h) into text blocks of whole records
//                                                         (processFastq, classify.cpp:238-269)
//   parser threads  : block -> pinned batch; barcodes interned to dense ids
//                     (parseName :112-119; MultiThread::submit's Buffer :121-127,211-219)
//   one thread/GPU  : hast_submit_batch (async H2D + fused kernel), batches taken from a
//                     shared queue, i.e. read batches are sharded over the GPUs
//                     (the reference round-robins buffers over worker threads, :214-218)
//   finish          : hast_finish on every GPU at once -> one ncclReduce -> counts on GPU 0
//                     (wait / collectBarcodes / Add, :220-229,57-63)
//
// The k-mer table is built once on GPU 0 and cloned to the others over NVLink.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <unistd.h>

#include "../../include/hast_b200.h"
#include "fastq_source.h"
#include "host.h"
#include "plain_slicer.h"

namespace hasthost {

namespace {

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class T>
class Queue {
public:
    explicit Queue(size_t cap) : cap_(cap) {}
    bool push(T v) {
        std::unique_lock<std::mutex> lk(mu_);
        not_full_.wait(lk, [&] { return q_.size() < cap_ || closed_; });
        if (closed_) return false;
        q_.push_back(std::move(v));
        not_empty_.notify_one();
        return true;
    }
    bool pop(T& v) {
        std::unique_lock<std::mutex> lk(mu_);
        not_empty_.wait(lk, [&] { return !q_.empty() || done_ || closed_; });
        if (closed_ || q_.empty()) return false;
        v = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    void finish() {                  // producers are done: consumers drain then stop
        std::lock_guard<std::mutex> lk(mu_);
        done_ = true;
        not_empty_.notify_all();
    }
    void abort() {                   // error: everybody stops now
        std::lock_guard<std::mutex> lk(mu_);
        closed_ = true;
        not_empty_.notify_all();
        not_full_.notify_all();
    }
private:
    std::mutex mu_;
    std::condition_variable not_full_, not_empty_;
    std::deque<T> q_;
    size_t cap_;
    bool done_ = false, closed_ = false;
};

struct Shared {
    std::mutex mu;
    std::string error;
    std::atomic<bool> failed{false};
    void fail(const std::string& e) {
        std::lock_guard<std::mutex> lk(mu);
        if (error.empty()) error = e;
        failed = true;
    }
};

void logtime() {                     // classify.cpp:17-21
    time_t t = time(nullptr);
    fprintf(stderr, "%s\n", ctime(&t));
}

std::string kmer_to_string(uint64_t w, int k) {     // Kmer::ToBaseStr + BaseStr2Str, kmer.h:244-254,14-25
    std::string s((size_t)k, 'A');
    for (int i = 0; i < k; ++i) { s[(size_t)(k - 1 - i)] = "ACTG"[w & 3]; w >>= 2; }
    return s;
}

}  // namespace

int run_classify(const Options& opt, RunStats& st) {
    const double t_start = now();
    // stdout carries the table and nothing else (classify_stlfr_reads.sh:148-149 redirects it into
    // phased.barcodes).  Libraries underneath (NCCL with NCCL_DEBUG set prints its version to
    // stdout) must not be able to write into it: keep a private handle for the table and point
    // file descriptor 1 at stderr for the rest of the process.
    fflush(stdout);
    FILE* table_out = nullptr;
    {
        const int fd = dup(STDOUT_FILENO);
        if (fd >= 0) table_out = fdopen(fd, "w");
        if (!table_out) table_out = stdout; else dup2(STDERR_FILENO, STDOUT_FILENO);
    }
    // HAST_PARSE_ONLY=1: host front end only (reader, framing, parseName, interning, ordering);
    // prints barcode \t reads \t bases.  A diagnostic for the host logic, it classifies nothing.
    const bool parse_only = getenv("HAST_PARSE_ONLY") != nullptr;
    // HAST_PARSE_ONLY=2: the parser also PACKS (the default mode of a real run) and the tally reads the reads back out of
    // the 2-bit stream: two more columns, a checksum of every read's base codes and the number of reads flagged containN.
    // Lets the CPU tests check the packed batches the GPU would receive.
    const bool parse_check = parse_only && atoi(getenv("HAST_PARSE_ONLY")) == 2;
    int n_dev = parse_only ? 1 : hast_device_count();
    if (n_dev <= 0) {
        fprintf(stderr, "ERROR : no CUDA device found; this build of classify has no CPU path\n");
        return 1;
    }
    // One GPU by default: measured on an 8 x B200 box (profiles/bench_r02_m_cfg5.json) the host streams gzip FASTQ at
    // 13-14 M pairs/s whether 1 or 8 GPUs classify it (one GPU's kernel takes 1500 M pairs/s), and 8 contexts cost
    // 4 s more to set up.  --gpus N shards the batches over N GPUs for callers whose input is faster than that.
    int n_gpu = parse_only ? 1 : std::min(opt.gpus > 0 ? opt.gpus : 1, n_dev);
    st.gpus = n_gpu;
    st.parser_threads = opt.threads;

    std::vector<hast_ctx*> ctx((size_t)n_gpu, nullptr);
    auto cleanup = [&] { for (hast_ctx* c : ctx) hast_destroy(c); };

    fprintf(stderr, "__START__\n use hap0 weight %g\n use hap1 weight %g\n", opt.weight0, opt.weight1);
    fprintf(stderr, " use %d GPU(s), %d parser thread(s)\n", n_gpu, opt.threads);
    logtime();

    // CUDA start-up stays on this thread, before any other thread exists: context creation maps and unmaps memory,
    // and with a dozen parser and decoder threads faulting pages at the same time it took 1.1-1.4 s instead of 0.3
    // (profiles/bench_r02_p_cli16m.json, the build that overlapped it).
    for (int g = 0; g < n_gpu && !parse_only; ++g)
        if (hast_create(g, &ctx[(size_t)g]) != HAST_OK) {
            fprintf(stderr, "ERROR : %s\n", hast_last_error(nullptr));
            cleanup();
            return 1;
        }

    // ---- k-mer table (load_kmers x2 + InitAdaptor, classify.cpp:433-437) -----------
    // (a lambda because it was also tried on a thread of its own, see below where it is called)
    auto build_table = [&]() -> std::string {
#define TCHECK(c, call)                                                     \
    do {                                                                    \
        if ((call) != HAST_OK) return std::string(hast_last_error(c));      \
    } while (0)
        KmerList l0, l1;
        fprintf(stderr, "__load hap0 kmers__\n");
        std::string e = load_kmer_list(opt.hap0, 0, 0, l0);
        if (!e.empty()) return e;
        fprintf(stderr, "Recorded %llu haplotype 0 specific %d-mers\n", (unsigned long long)l0.n_lines, l0.k);
        fprintf(stderr, "__load hap1 kmers__\n");
        e = load_kmer_list(opt.hap1, 1, l0.k, l1);
        if (!e.empty()) return e;
        fprintf(stderr, "Recorded %llu haplotype 1 specific %d-mers\n", (unsigned long long)l1.n_lines, l0.k);

        uint64_t expected = l0.n_lines + l1.n_lines;
        for (int attempt = 0;; ++attempt) {
            TCHECK(ctx[0], hast_table_begin(ctx[0], l0.k, expected));
            int rc = hast_table_add_text(ctx[0], l0.text.data(), l0.n_lines, 0);
            if (rc == HAST_OK) rc = hast_table_add_text(ctx[0], l1.text.data(), l1.n_lines, 1);
            if (rc == HAST_E_TABLE_FULL && attempt < 4) { expected = expected * 2 + 64; continue; }
            if (rc != HAST_OK) return std::string(hast_last_error(ctx[0]));
            break;
        }
        fprintf(stderr, "Adaptor forward :%s\nAdaptor reverse :%s\n", opt.adaptor_f.c_str(), opt.adaptor_r.c_str());
        for (const std::string* ad : {&opt.adaptor_f, &opt.adaptor_r}) {
            std::vector<uint64_t> er(ad->size() + 1);
            std::vector<uint8_t> tg(ad->size() + 1);
            uint32_t n = 0;
            TCHECK(ctx[0], hast_table_erase_seq(ctx[0], ad->data(), (uint32_t)ad->size(), er.data(), tg.data(),
                                                (uint32_t)er.size(), &n));
            for (uint32_t i = 0; i < n && i < er.size(); ++i)       // classify.cpp:319-337
                for (int h = 0; h < 2; ++h)
                    if (tg[i] & (1 << h))
                        fprintf(stderr, " INFO : erase a adaptor kmer from hap %d ; kmer= %s\n", h,
                                kmer_to_string(er[i], l0.k).c_str());
        }
        hast_table_info ti;
        TCHECK(ctx[0], hast_table_info_get(ctx[0], &ti));
        st.size0 = ti.size[0];
        st.size1 = ti.size[1];
        st.table_bytes = ti.bytes;
        fprintf(stderr, " table : %llu buckets (%.1f MiB), %llu distinct k-mers, %llu displaced, |S0|=%llu |S1|=%llu\n",
                (unsigned long long)ti.n_buckets, ti.bytes / 1048576.0, (unsigned long long)ti.n_entries,
                (unsigned long long)ti.n_displaced, (unsigned long long)ti.size[0], (unsigned long long)ti.size[1]);
        for (int g = 1; g < n_gpu; ++g) TCHECK(ctx[(size_t)g], hast_table_clone(ctx[(size_t)g], ctx[0]));
        if (n_gpu > 1) TCHECK(ctx[0], hast_comm_init_all(ctx.data(), n_gpu));
#undef TCHECK
        st.t_table = now() - t_start;
        logtime();
        return "";
    };

    // ---- streaming classification ---------------------------------------------------
    const double t_reads0 = now();
    const size_t block_bytes = std::max<size_t>(opt.batch_bytes, 1u << 16);
    const size_t n_batches = (size_t)opt.threads + 3 * (size_t)n_gpu + 1;
    Shared sh;
    BarcodeIndex index;
    Queue<TextBlock*> q_text((size_t)opt.threads + 8 + 2), q_text_free(1u << 20);
    Queue<Batch*> q_batch(n_batches), q_batch_free(1u << 20);
    std::vector<TextBlock> text_pool((size_t)opt.threads + 8 + 3);
    std::vector<Batch> batch_pool(n_batches);
    for (auto& t : text_pool) q_text_free.push(&t);
    bool alloc_ok = true;
    for (auto& b : batch_pool) {
        b.cap_bases = block_bytes + 4096;
        b.cap_reads = block_bytes / 24 + 16;
        void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr, *p3 = nullptr;
        const bool packed = (opt.packed_h2d && !parse_only) || parse_check;
        b.cap_words = b.cap_bases / 16 + 2;
        if (parse_only) {
            p0 = malloc(packed ? b.cap_words * 4 : b.cap_bases); p1 = malloc((b.cap_reads + 1) * 4); p2 = malloc(b.cap_reads * 4);
            if (packed) p3 = malloc((b.cap_reads / 32 + 2) * 4);
        } else if (hast_host_alloc(&p0, packed ? b.cap_words * 4 : b.cap_bases) ||
                   hast_host_alloc(&p1, (b.cap_reads + 1) * 4) || hast_host_alloc(&p2, b.cap_reads * 4) ||
                   (packed && hast_host_alloc(&p3, (b.cap_reads / 32 + 2) * 4))) {
            alloc_ok = false;
            break;
        }
        if (packed) { b.packed = (uint32_t*)p0; b.has_n = (uint32_t*)p3; }
        else b.bases = (uint8_t*)p0;
        b.read_off = (uint32_t*)p1; b.barcode_id = (uint32_t*)p2;
        q_batch_free.push(&b);
    }
    auto free_batches = [&] {
        for (auto& b : batch_pool) {
            if (parse_only) { free(b.bases); free(b.packed); free(b.has_n); free(b.read_off); free(b.barcode_id); }
            else {
                hast_host_free(b.bases); hast_host_free(b.packed); hast_host_free(b.has_n);
                hast_host_free(b.read_off); hast_host_free(b.barcode_id);
            }
        }
    };
    if (!alloc_ok) {
        fprintf(stderr, "ERROR : pinned host allocation failed: %s\n", hast_last_error(nullptr));
        free_batches(); cleanup();
        return 1;
    }
    auto abort_all = [&] { q_text.abort(); q_text_free.abort(); q_batch.abort(); q_batch_free.abort(); };

    // Plain regular files are read by the parser threads themselves, slice by slice (plain_slicer.h).  Everything
    // else -- gzip, pipes, standard input -- is a serial stream and gets a reader thread per file, several files
    // at once: inflating is the slowest stage of the whole program (~0.3-0.7 GB/s of text per core against
    // > 100 GB/s the GPUs classify), but the files of a run (r1/r2, lanes) are independent and the per-barcode sums
    // do not depend on the order in which reads arrive.
    std::atomic<uint64_t> text_bytes{0};
    std::vector<std::unique_ptr<PlainSlicer>> plain;
    std::vector<std::string> streams;
    for (const std::string& path : opt.reads) {
        const size_t n = path.size();
        const bool gz = n > 3 && path.compare(n - 3, 3, ".gz") == 0;      // classify.cpp:245-250
        if (!gz && path != "-" && !getenv("HAST_SERIAL_READER")) {
            std::unique_ptr<PlainSlicer> sl(new PlainSlicer());
            size_t slice_bytes = block_bytes - 8192;
            if (const char* sb = getenv("HAST_SLICE_BYTES")) slice_bytes = (size_t)std::max(1L, atol(sb));   // tests: slices of a few bytes
            const std::string e = sl->open(path, slice_bytes);
            if (!e.empty()) { fprintf(stderr, "ERROR : %s\n", e.c_str()); free_batches(); cleanup(); return 1; }
            if (sl->usable()) { plain.push_back(std::move(sl)); continue; }
        }
        streams.push_back(path);
    }
    // The table is built BEFORE the pipeline threads start.  Overlapping the two was measured (profiles/bench_r02_{f,p,q}_cli*.json):
    // it saves 0.3 s on a 4 M-pair input when it works, but with a dozen parser and sixteen decoder threads competing for the
    // cores and the memory system the table build itself took anything from 0.4 to 1.7 s instead of a steady 0.4-0.5 s, and
    // for inputs of useful size the overlap is worth 2 % at best.
    if (!parse_only) {
        const std::string e = build_table();
        if (!e.empty()) { fprintf(stderr, "ERROR : %s\n", e.c_str()); free_batches(); cleanup(); return 1; }
    }
    std::atomic<size_t> next_file{0};
    const int n_readers = (int)std::min<size_t>({streams.size(), (size_t)opt.threads, (size_t)8});
    // threads per gzip stream (inflate_par.h: one replays, the others entropy-decode).  Measured on a 16-core box with
    // two streams (profiles/r02_d_gz_threads.txt): 5 / 6 / 7 / 8 threads per stream -> 12.6 / 10.7 / 12.9 / 13.9 M pairs/s;
    // the decoders may outnumber the cores -- they and the parsers sleep when their queues are empty
    int inflate_threads = (int)std::max(2u, std::min(8u, std::thread::hardware_concurrency() / (unsigned)std::max(1, n_readers)));
    if (const char* e = getenv("HAST_INFLATE_THREADS")) inflate_threads = std::max(1, atoi(e));
    std::atomic<int> readers_left{n_readers};
    std::vector<std::thread> readers;
    if (n_readers == 0) q_text.finish();
    for (int r = 0; r < n_readers; ++r)
        readers.emplace_back([&] {
            for (;;) {
                const size_t fi = next_file.fetch_add(1);
                if (fi >= streams.size() || sh.failed) break;
                const std::string& path = streams[fi];
                fprintf(stderr, "__process read: %s\n", path.c_str());
                FastqSource src;
                std::string e = src.open(path, inflate_threads);
                if (!e.empty()) { sh.fail(e); abort_all(); break; }
                bool stop = false;
                for (;;) {
                    TextBlock* blk = nullptr;
                    if (!q_text_free.pop(blk)) { stop = true; break; }
                    std::string err;
                    const bool more = src.next(*blk, block_bytes - 8192, err);
                    if (!err.empty()) { sh.fail(err); abort_all(); stop = true; break; }
                    if (!more) { q_text_free.push(blk); break; }
                    if (!q_text.push(blk)) { stop = true; break; }
                }
                text_bytes += src.bytes_out();
                if (stop) break;
            }
            if (--readers_left == 0) q_text.finish();
        });

    std::atomic<int> parsers_left{opt.threads};
    std::atomic<size_t> plain_cursor{0};
    std::vector<std::thread> parsers;
    for (int t = 0; t < opt.threads; ++t)
        parsers.emplace_back([&] {
            // block -> one batch (several when the records are far shorter than the buffers were sized for)
            auto parse_and_push = [&](const TextBlock& blk) -> bool {
                size_t pos = 0;
                while (pos < blk.len) {
                    Batch* b = nullptr;
                    if (!q_batch_free.pop(b)) return false;
                    if (!parse_block(blk, index, *b, &pos)) { sh.fail(b->error); abort_all(); return false; }
                    if (!q_batch.push(b)) return false;
                }
                return true;
            };
            bool alive = true;
            TextBlock own;                                        // plain files: this thread's slice buffer
            while (alive) {
                const size_t fi = plain_cursor.load(std::memory_order_acquire);
                if (fi >= plain.size() || sh.failed) break;
                bool first = false;
                if (!plain[fi]->next(own, &first)) {          // file used up: move everybody on to the next one
                    if (first) fprintf(stderr, "__process read: %s\n", plain[fi]->path().c_str());   // an empty file
                    size_t expect = fi;
                    plain_cursor.compare_exchange_strong(expect, fi + 1);
                    continue;
                }
                if (first) {
                    fprintf(stderr, "__process read: %s\n", plain[fi]->path().c_str());
                    text_bytes += plain[fi]->size();
                }
                if (own.len) alive = parse_and_push(own);
            }
            TextBlock* blk = nullptr;
            while (alive && q_text.pop(blk)) {
                alive = parse_and_push(*blk);
                blk->hold.reset();                                // a decoder's buffer goes back to its pool
                q_text_free.push(blk);
            }
            if (--parsers_left == 0) q_batch.finish();
        });

    std::atomic<uint64_t> n_reads{0}, n_bases{0};
    std::vector<uint64_t> po_tally;                      // parse-only: reads, bases per barcode id
    std::vector<uint64_t> po_check;                      // HAST_PARSE_ONLY=2: checksum of the packed base codes, containN reads
    std::vector<std::thread> gpu_threads;
    for (int g = 0; g < n_gpu; ++g)
        gpu_threads.emplace_back([&, g] {
            uint64_t reserved = 0;
            std::deque<std::pair<uint64_t, Batch*>> inflight;
            Batch* b = nullptr;
            while (parse_only && q_batch.pop(b)) {
                for (uint32_t i = 0; i < b->n_reads; ++i) {
                    const uint32_t id = b->barcode_id[i];
                    if (po_tally.size() < 2 * ((size_t)id + 1)) po_tally.resize(2 * ((size_t)id + 1), 0);
                    po_tally[2 * (size_t)id] += 1;
                    po_tally[2 * (size_t)id + 1] += b->read_off[i + 1] - b->read_off[i];
                    if (parse_check) {
                        if (po_check.size() < 2 * ((size_t)id + 1)) po_check.resize(2 * ((size_t)id + 1), 0);
                        uint64_t h = 0;
                        for (uint32_t g = b->read_off[i]; g < b->read_off[i + 1]; ++g)
                            h = h * 5u + ((b->packed[g >> 4] >> (30u - 2u * (g & 15u))) & 3u) + 1u;
                        po_check[2 * (size_t)id] += h;
                        po_check[2 * (size_t)id + 1] += (b->has_n[i >> 5] >> (i & 31u)) & 1u;
                    }
                }
                n_reads += b->n_reads;
                n_bases += b->n_bases;
                q_batch_free.push(b);
            }
            if (parse_only) return;
            hast_ctx* c = ctx[(size_t)g];
            while (q_batch.pop(b)) {
                const uint64_t need = (uint64_t)b->max_barcode + 1;
                if (need > reserved) {
                    reserved = std::max<uint64_t>(need, index.size());
                    if (hast_reserve_barcodes(c, reserved) != HAST_OK) { sh.fail(hast_last_error(c)); abort_all(); break; }
                }
                uint64_t ticket = 0;
                const int src = b->packed
                    ? hast_submit_batch_packed(c, b->packed, b->n_bases, b->read_off, b->barcode_id, b->has_n, b->n_reads, &ticket)
                    : hast_submit_batch(c, b->bases, b->n_bases, b->read_off, b->barcode_id, b->n_reads, &ticket);
                if (src != HAST_OK) {
                    sh.fail(hast_last_error(c)); abort_all(); break;
                }
                n_reads += b->n_reads;
                n_bases += b->n_bases;
                inflight.emplace_back(ticket, b);
                while (inflight.size() > 2) {                  // double buffering: two batches in flight per GPU
                    hast_wait_copied(c, inflight.front().first);
                    q_batch_free.push(inflight.front().second);
                    inflight.pop_front();
                }
            }
            hast_sync(c);
            for (auto& f : inflight) q_batch_free.push(f.second);
        });

    for (auto& t : readers) t.join();
    for (auto& t : parsers) t.join();
    for (auto& t : gpu_threads) t.join();
    if (sh.failed) {
        fprintf(stderr, "ERROR : %s\n", sh.error.c_str());
        free_batches(); cleanup();
        return 1;
    }
    st.reads = n_reads;
    st.bases = n_bases;
    st.text_bytes = text_bytes;
    st.t_reads = now() - t_reads0;
    logtime();
    fprintf(stderr, "__process read done__\n");

    if (parse_only) {
        std::vector<std::string> names;
        index.export_names(names);
        po_tally.resize(2 * names.size(), 0);
        std::vector<uint32_t> order(names.size());
        for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return names[a] < names[b]; });
        po_check.resize(2 * names.size(), 0);
        for (uint32_t id : order) {
            fprintf(table_out, "%s\t%llu\t%llu", names[id].c_str(), (unsigned long long)po_tally[2 * (size_t)id],
                    (unsigned long long)po_tally[2 * (size_t)id + 1]);
            if (parse_check)
                fprintf(table_out, "\t%llu\t%llu", (unsigned long long)po_check[2 * (size_t)id],
                        (unsigned long long)po_check[2 * (size_t)id + 1]);
            fputc('\n', table_out);
        }
        fflush(table_out);
        free_batches();
        st.t_total = now() - t_start;
        return 0;
    }

    // ---- collect (collectBarcodes / Add, classify.cpp:226-229,57-63) -------------------
    const double t_fin0 = now();
    const uint64_t n_bc = index.size();
    st.barcodes = n_bc;
    // pinned: the read-back of the counters (160 MB at 20 M barcodes) then runs at PCIe speed instead of through
    // the driver's pageable staging buffers
    struct PinnedCounts {
        int32_t* p = nullptr;
        ~PinnedCounts() { hast_host_free(p); }
        int32_t* data() const { return p; }
    } counts;
    {
        void* p = nullptr;
        if (hast_host_alloc(&p, std::max<uint64_t>(n_bc, 1) * 2 * sizeof(int32_t)) != HAST_OK) {
            fprintf(stderr, "ERROR : pinned host allocation failed: %s\n", hast_last_error(nullptr));
            free_batches(); cleanup();
            return 1;
        }
        counts.p = static_cast<int32_t*>(p);
        memset(counts.p, 0, std::max<uint64_t>(n_bc, 1) * 2 * sizeof(int32_t));
    }
    {
        std::vector<int> rcs((size_t)n_gpu, 0);
        std::vector<std::thread> fin;
        for (int g = 0; g < n_gpu; ++g)
            fin.emplace_back([&, g] {
                hast_ctx* c = ctx[(size_t)g];
                int rc = hast_reserve_barcodes(c, n_bc);
                if (rc == HAST_OK) rc = hast_finish(c, g == 0 ? counts.data() : nullptr, n_bc);
                rcs[(size_t)g] = rc;
            });
        for (auto& t : fin) t.join();
        for (int g = 0; g < n_gpu; ++g)
            if (rcs[(size_t)g] != HAST_OK) {
                fprintf(stderr, "ERROR : %s\n", hast_last_error(ctx[(size_t)g]));
                free_batches(); cleanup();
                return 1;
            }
    }
    for (int g = 0; g < n_gpu; ++g) {
        hast_stats hs;
        if (hast_stats_get(ctx[(size_t)g], &hs) == HAST_OK) { st.lookups += hs.lookups; st.kernel_launches += hs.kernel_launches; }
    }
    st.t_finish = now() - t_fin0;

    // ---- output (printBarcodeInfos, classify.cpp:447) ---------------------------------
    fprintf(stderr, "__print result__\n");
    const double t_pr0 = now();
    std::vector<std::string> names;
    index.export_names(names);
    std::vector<uint32_t> order;
    std::vector<int8_t> haps;
    print_table(table_out, names, counts.data(), st.size0, st.size1, opt.weight0, opt.weight1,
                opt.split_barcodes ? &order : nullptr, opt.split_barcodes ? &haps : nullptr);
    fflush(table_out);
    st.t_print = now() - t_pr0;
    logtime();
    free_batches();
    cleanup();

    // ---- optional: the rest of classify_stlfr_reads.sh (:156-185) in the same process ----
    if (opt.split_barcodes) {
        const double t0 = now();
        BarcodeLists lists;
        uint64_t n3[3];
        const std::string e = write_barcode_lists(opt.outdir, names, order, haps, counts.data(), n3,
                                                  opt.partition_reads ? &lists : nullptr);
        if (!e.empty()) { fprintf(stderr, "ERROR : %s\n", e.c_str()); return 1; }
        fprintf(stderr, "final paternal barcode : %llu\nfinal maternal barcodes : %llu\nfinal homozygous barcodes : %llu\n",
                (unsigned long long)n3[0], (unsigned long long)n3[1], (unsigned long long)n3[2]);
        st.t_split = now() - t0;
        if (opt.partition_reads) {
            const double t1 = now();
            fprintf(stderr, "phase reads ...\n");
            for (const std::string& path : opt.reads) {
                const bool gz = path.size() > 3 && path.compare(path.size() - 3, 3, ".gz") == 0;
                PartitionStats ps;
                // the script pipes gzip input into `awk ... -`, whose FILENAME is then "-" (:181)
                const std::string e2 = partition_fastq(path, gz ? "-" : path, partition_prefix(path), opt.outdir,
                                                       lists, ps, opt.threads);
                if (!e2.empty()) { fprintf(stderr, "ERROR : %s\n", e2.c_str()); return 1; }
                st.partition_text_bytes += ps.text_bytes;
            }
            st.t_partition = now() - t1;
            fprintf(stderr, "phase reads done\n");
            logtime();
        }
    }
    fprintf(stderr, "__END__\n");
    st.t_total = now() - t_start;
    return 0;
}

}  // namespace hasthost
