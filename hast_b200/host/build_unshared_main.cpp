// build_unshared_main.cpp -- `build_unshared_kmers`, the drop-in for HAST stage 00
// (00.build_unshare_kmers_by_jellyfish/build_unshared_kmers.sh, called by HAST.sh:141-148).
//
//     build_unshared_kmers --paternal p.fq[.gz] [--paternal ...] --maternal m.fq[.gz] [--maternal ...]
//                          [--mer 21] [--thread 8] [--memory 10] [--auto_bounds]
//                          [--p-lower 9 --p-upper 33 --m-lower 9 --m-upper 33]
//
// Same arguments, same messages on stdout, same files left in the working directory that later
// stages (or people) read: paternal.unique.filter.mer, maternal.unique.filter.mer and, with
// --auto_bounds, {paternal,maternal}.histo and {paternal,maternal}.bounds.txt.  The script's five
// `jellyfish count` runs, its FASTA dumps and the "2 x maternal + 1 x paternal" mix (:163-291)
// become ONE pass of the parental reads through a count table on the GPU (hast_kc_*,
// csrc/kcount.cuh); the intermediate *.mer.fa / *.jf files are not produced.  The k-mer lists
// hold the same LINES as the script's; their order is ascending (jellyfish dumps in the order
// of its hash table, which depends on --memory and --thread).
//
// What stays on the host: reading / inflating the files, cutting FASTA / FASTQ records into
// sequence chunks (overlapping by k-1 so that every window is seen once), find_bounds.awk's
// arithmetic on the histogram, and writing the text files.
//
// A count table for both human parents does not fit one GPU: the key space is cut into
// partitions (--parts, chosen from the input size and the free device memory by default), each
// GPU owns one partition per pass and sees every read; the reads are streamed once per pass.
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hast_b200.h"
#include "inflate.h"
#include "inflate_par.h"

namespace {

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void usage() {                                             // build_unshared_kmers.sh:6-37
    puts("Usage    :");
    puts("    ./build_unshared_kmers.sh [OPTION]");
    puts("");
    puts("Build parental unshared-kmers based on paternal and maternal NGS reads by jellyfish.");
    puts("");
    puts("Options  :");
    puts("        --paternal    paternal NGS reads file in FASTA/FASTQ format.");
    puts("                      file in gzip format can be accepted, but filename must end by \".gz\".");
    puts("        --maternal    maternal NGS reads file in FASTA/FASTQ format.");
    puts("                      file in gzip format can be accepted, but filename must end by \".gz\".");
    puts("        --thread      thread number.");
    puts("                      [ optional, default 8 threads. ]");
    puts("        --memory      x (GB) of memory to initial hash table by jellyfish.");
    puts("                      ( note: real memory used may be greater than this. )");
    puts("                      [ optional, default 20GB. ]");
    puts("        --mer         mer-size");
    puts("                      [ optional, default 21. ]");
    puts("        --m-lower     maternal kmer frequency table will ignore kmers with count < m-lower.");
    puts("                      [ optional, default 9. ]");
    puts("        --m-upper     maternal kmer frequency table will ignore kmers with count > m-upper.");
    puts("                      [ optional, default 33. ]");
    puts("        --p-lower     paternal kmer frequency table will ignore kmers with count < p-lower.");
    puts("                      [ optional, default 9. ]");
    puts("        --p-upper     paternal kmer frequency table will ignore kmers with count > p-upper.");
    puts("                      [ optional, default 33. ]");
    puts("        --auto_bounds automatically calcuate lower and upper bounds based on kmer analysis.");
    puts("                      [ optional, default not trigger; no parameter. ]");
    puts("                      ( note : if auto_bounds is on, it will overwrite --*-lower and --*-upper  ]");
    puts("                      ( !!! WARN : default bounds is seted for 30X WGS reads , if your data is not close to 30X, please use your own bounds or simply open auto_bounds !!! ) ");
    puts("        --help        print this usage message.");
    puts("        (this build, long options only: --gpus N, --parts N, --expected-distinct N, --stats-json FILE)");
}

void print_date() {
    time_t t = time(nullptr);
    char buf[128];
    strftime(buf, sizeof buf, "%a %b %e %H:%M:%S %Z %Y", localtime(&t));
    puts(buf);
}

bool ends_with_gz(const std::string& s) { return s.size() >= 3 && s.compare(s.size() - 3, 3, ".gz") == 0; }

// ---- sequence batches -----------------------------------------------------------------------
struct SeqBatch {
    uint8_t* bases = nullptr;   size_t cap_bases = 0;  size_t n_bases = 0;
    uint32_t* off = nullptr;    size_t cap_seqs = 0;   uint32_t n_seqs = 0;
    int parent = 0;
};

template <class T>
class Queue {
public:
    void push(T v) { { std::lock_guard<std::mutex> lk(mu_); q_.push_back(v); } cv_.notify_one(); }
    bool pop(T& v) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !q_.empty() || done_; });
        if (q_.empty()) return false;
        v = q_.front();
        q_.pop_front();
        return true;
    }
    void finish() { { std::lock_guard<std::mutex> lk(mu_); done_ = true; } cv_.notify_all(); }
    void reopen() { std::lock_guard<std::mutex> lk(mu_); done_ = false; }
private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<T> q_;
    bool done_ = false;
};

// Cuts the records of one FASTA / FASTQ file into sequence chunks.  jellyfish's reader: the format is
// decided by the first byte of the file ('>' or '@'); a FASTA record's lines are joined (blank lines
// dropped); a FASTQ record's sequence runs until the '+' line and its quality until it is as long.
class SeqReader {
public:
    SeqReader(int k, size_t chunk, Queue<SeqBatch*>& free_q, Queue<SeqBatch*>& full_q, int parent, int inflate_threads)
        : k_(k), chunk_(chunk), free_(free_q), full_(full_q), parent_(parent), inflate_threads_(inflate_threads) {}

    std::string run(const std::string& path) {
        std::unique_ptr<hasthost::GzipInflater> inf(new hasthost::GzipInflater());
        if (!inf->open(path).empty() || !inf->is_gzip() || getenv("HAST_ZLIB")) inf.reset();
        std::unique_ptr<hasthost::ParallelGzip> pinf;
        if (inf && inflate_threads_ > 1) {                 // one gzip stream on several threads (inflate_par.h)
            pinf.reset(new hasthost::ParallelGzip(inflate_threads_));
            if (!pinf->open(path).empty()) pinf.reset(); else inf.reset();
        }
        gzFile gz = nullptr;
        if (!inf && !pinf) {
            gz = gzopen(path.c_str(), "rb");               // passes plain text through
            if (!gz) return "cannot open " + path;
            gzbuffer(gz, 1u << 20);
        }
        std::vector<char> buf(4u << 20);
        std::string line;
        enum { kStart, kFastaSeq, kFqSeq, kFqQual } st = kStart;
        bool fastq = false;
        size_t seq_len = 0, qual_len = 0;
        auto handle = [&](const char* p, size_t n) {       // one line, without its '\n'
            switch (st) {
                case kStart:
                    if (!n) return;
                    if (p[0] == '@') { fastq = true; st = kFqSeq; seq_len = 0; begin_seq(); }
                    else if (p[0] == '>') { fastq = false; st = kFastaSeq; begin_seq(); }
                    return;
                case kFastaSeq:
                    if (n && p[0] == '>') { end_seq(); begin_seq(); return; }
                    append(p, n);
                    return;
                case kFqSeq:
                    if (n && p[0] == '+') { end_seq(); st = kFqQual; qual_len = 0; if (seq_len == 0) st = kStart; return; }
                    append(p, n);
                    seq_len += n;
                    return;
                case kFqQual:
                    qual_len += n;
                    if (qual_len >= seq_len) st = kStart;
                    return;
            }
        };
        auto feed = [&](const char* p, size_t n) {
            text_bytes_ += (uint64_t)n;
            const char* const end = p + n;
            while (p < end) {
                const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
                if (!nl) { line.append(p, (size_t)(end - p)); break; }
                if (line.empty()) handle(p, (size_t)(nl - p));
                else { line.append(p, (size_t)(nl - p)); handle(line.data(), line.size()); line.clear(); }
                p = nl + 1;
            }
        };
        if (pinf) {
            const uint8_t* p;
            size_t n;
            while (pinf->next(&p, &n)) feed((const char*)p, n);
            if (!pinf->error().empty()) return "inflate failed on " + path + ": " + pinf->error();
        } else if (inf) {                                  // gzip members: the readers' own decoder (inflate.h)
            const uint8_t* p;
            size_t n;
            while (inf->next(&p, &n)) feed((const char*)p, n);
            if (!inf->error().empty()) return "inflate failed on " + path + ": " + inf->error();
        } else
        for (;;) {
            const int r = gzread(gz, buf.data(), (unsigned)buf.size());
            if (r < 0) { int e = 0; std::string m = gzerror(gz, &e); gzclose(gz); return "gzread failed on " + path + ": " + m; }
            if (r == 0) break;
            feed(buf.data(), (size_t)r);
        }
        if (!line.empty()) handle(line.data(), line.size());
        if (st == kFastaSeq || st == kFqSeq) end_seq();
        if (gz) gzclose(gz);
        (void)fastq;
        flush();
        return "";
    }
    uint64_t text_bytes() const { return text_bytes_; }
    uint64_t sequences() const { return n_records_; }

private:
    void need_batch() {
        if (cur_) return;
        free_.pop(cur_);
        cur_->n_bases = 0;
        cur_->n_seqs = 0;
        cur_->off[0] = 0;
        cur_->parent = parent_;
    }
    void begin_seq() { in_seq_ = true; piece_ = 0; ++n_records_; need_batch(); }
    void close_piece() {                                   // the bytes since the last offset become one chunk
        if (!cur_) return;
        if (cur_->n_bases > cur_->off[cur_->n_seqs]) cur_->off[++cur_->n_seqs] = (uint32_t)cur_->n_bases;
        piece_ = 0;
    }
    void end_seq() {
        close_piece();
        in_seq_ = false;
        if (cur_ && (cur_->n_bases + chunk_ + 64 > cur_->cap_bases || cur_->n_seqs + 2 >= cur_->cap_seqs)) flush();
    }
    void append(const char* p, size_t n) {
        while (n) {
            need_batch();
            size_t room = chunk_ - piece_;
            const size_t m = std::min(n, room);
            memcpy(cur_->bases + cur_->n_bases, p, m);
            cur_->n_bases += m;
            piece_ += m;
            p += m;
            n -= m;
            if (piece_ == chunk_) {                        // chunk full: the next one restarts k-1 bytes back
                const size_t ov = (size_t)(k_ - 1);
                char tail[64];
                memcpy(tail, cur_->bases + cur_->n_bases - ov, ov);
                close_piece();
                if (cur_->n_bases + chunk_ + 64 > cur_->cap_bases || cur_->n_seqs + 2 >= cur_->cap_seqs) flush();
                need_batch();
                memcpy(cur_->bases + cur_->n_bases, tail, ov);
                cur_->n_bases += ov;
                piece_ = ov;
            }
        }
    }
    void flush() {
        if (!cur_) return;
        close_piece();
        if (cur_->n_seqs) full_.push(cur_); else free_.push(cur_);
        cur_ = nullptr;
    }

    int k_;
    size_t chunk_;
    Queue<SeqBatch*>& free_;
    Queue<SeqBatch*>& full_;
    int parent_;
    int inflate_threads_;
    SeqBatch* cur_ = nullptr;
    size_t piece_ = 0;
    bool in_seq_ = false;
    uint64_t text_bytes_ = 0, n_records_ = 0;
};

// find_bounds.awk:1-33 on the non-empty bins of `jellyfish histo`
struct Bounds { long min_index = 0, max_index = 0, lower = 0, upper = 0; };
Bounds find_bounds(const std::vector<uint64_t>& h) {
    uint64_t mn = 0, mx = 0;
    Bounds b;
    int state = 0;
    for (size_t i = 1; i < h.size(); ++i) {
        const uint64_t c = h[i];
        if (!c) continue;                                  // histo prints non-empty bins only
        if (state == 0) {
            if (mn == 0 || c < mn) { mn = c; b.min_index = (long)i; }
            else state = 1;
        } else if (mx == 0 || c > mx) { mx = c; b.max_index = (long)i; }
    }
    b.lower = b.min_index + 1;
    b.upper = 3 * b.max_index - 2 * b.min_index - 1;
    return b;
}

bool write_file(const std::string& path, const std::string& data) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
    return fclose(f) == 0 && ok;
}

struct Opt {
    long mer = 21, cpu = 8, memory = 10, plower = 9, pupper = 33, mlower = 9, mupper = 33;
    bool auto_bounds = false;
    std::vector<std::string> paternal, maternal;
    int gpus = 0;
    long parts = 0;
    uint64_t expected_distinct = 0;
    std::string stats_json;
};

}  // namespace

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    if (argc == 1) { usage(); return 0; }                 // :56-59
    Opt o;
    {
        std::string cmd = "CMD :";
        cmd += argv[0];
        for (int i = 1; i < argc; ++i) { cmd += ' '; cmd += argv[i]; }
        puts(cmd.c_str());
    }
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](long& dst) { if (i + 1 < argc) dst = atol(argv[++i]); };
        if (a == "-h" || a == "--help") { usage(); return 0; }
        else if (a == "--memory") val(o.memory);
        else if (a == "--thread") val(o.cpu);
        else if (a == "--m-lower") val(o.mlower);
        else if (a == "--m-upper") val(o.mupper);
        else if (a == "--p-lower") val(o.plower);
        else if (a == "--p-upper") val(o.pupper);
        else if (a == "--mer") val(o.mer);
        else if (a == "--auto_bounds") o.auto_bounds = true;
        else if (a == "--paternal") { if (i + 1 < argc) o.paternal.insert(o.paternal.begin(), argv[++i]); }   // :104-107 prepends
        else if (a == "--maternal") { if (i + 1 < argc) o.maternal.insert(o.maternal.begin(), argv[++i]); }
        else if (a == "--gpus") { long v = 0; val(v); o.gpus = (int)v; }
        else if (a == "--parts") val(o.parts);
        else if (a == "--expected-distinct") { long v = 0; val(v); o.expected_distinct = (uint64_t)v; }
        else if (a == "--stats-json") { if (i + 1 < argc) o.stats_json = argv[++i]; }
        else { printf("invalid params : \"%s\" . exit ... \n", a.c_str()); return 0; }                        // :117-120: bare `exit`
    }
    auto join = [](const std::vector<std::string>& v) { std::string s; for (auto& x : v) { s += x; s += ' '; } return s; };
    puts("HAST starting with : ");
    printf("    paternal input : %s\n", join(o.paternal).c_str());
    printf("    maternal input : %s\n", join(o.maternal).c_str());
    printf("    memory         : %ld GB\n", o.memory);
    printf("    thread         : %ld \n", o.cpu);
    printf("    mer            : %ld \n", o.mer);
    printf("    lower(maternal): %ld\n", o.mlower);
    printf("    upper(maternal): %ld\n", o.mupper);
    printf("    lower(paternal): %ld\n", o.plower);
    printf("    upper(paternal): %ld\n", o.pupper);
    printf("    auto_bounds    : %d\n", o.auto_bounds ? 1 : 0);
    if (o.memory < 1 || o.cpu < 1 || o.paternal.empty() || o.maternal.empty() || o.mer < 11 || o.mlower < 1 ||
        o.mupper > 100000000 || o.plower < 1 || o.pupper > 100000000) {                                       // :141-152
        puts("ERROR : arguments invalid ... exit!!! ");
        return 1;
    }
    if (o.mer > 32) { puts("ERROR : this build packs k-mers into 64 bits: --mer must be <= 32 ... exit!!! "); return 1; }
    for (const auto* v : {&o.maternal, &o.paternal})
        for (const auto& x : *v) {
            struct stat sb;
            if (stat(x.c_str(), &sb) != 0) { printf("ERROR : input file \"%s\" is not exist ! exit ...\n", x.c_str()); return 1; }
        }
    for (const auto* v : {&o.maternal, &o.paternal}) {     // :166-186
        int gz = 0;
        for (const auto& x : *v) {
            const int g = ends_with_gz(x) ? 2 : 1;
            if (gz && gz != g) { puts("ERROR : please don't mixed gz input with non-gz input."); return 1; }
            gz = g;
        }
    }
    print_date();
    puts("extract unique mers by jellyfish ...");
    const char* kPat = "paternal.unique.filter.mer";
    const char* kMat = "maternal.unique.filter.mer";
    auto count_lines = [](const char* path) -> long long {
        FILE* f = fopen(path, "rb");
        if (!f) return -1;
        std::vector<char> b(1 << 20);
        long long n = 0;
        size_t r;
        while ((r = fread(b.data(), 1, b.size(), f)) > 0)
            for (size_t i = 0; i < r; ++i) n += b[i] == '\n';
        fclose(f);
        return n;
    };
    auto report = [&] {                                    // :296-305
        puts("final paternal unique kmer is : ");
        printf("%lld %s\n", count_lines(kPat), kPat);
        puts("final maternal unique kmer is : ");
        printf("%lld %s\n", count_lines(kMat), kMat);
        puts("extract unique mers done...");
        print_date();
    };
    if (access("step_08_done", F_OK) == 0 && access(kPat, F_OK) == 0 && access(kMat, F_OK) == 0) {   // resume marker, :281-294
        puts("skip extract *aternal.unique.filter.mer  because step_08_done file already exist ...");
        report();
        return 0;
    }

    const double t_start = now();
    int n_dev = hast_device_count();
    if (n_dev <= 0) { puts("ERROR : no CUDA device found; this build has no CPU path"); return 1; }
    const int n_gpu = o.gpus > 0 ? std::min(o.gpus, n_dev) : n_dev;
    std::vector<hast_ctx*> ctx((size_t)n_gpu, nullptr);
    auto cleanup = [&] { for (auto* c : ctx) hast_destroy(c); };
    for (int g = 0; g < n_gpu; ++g)
        if (hast_create(g, &ctx[(size_t)g]) != HAST_OK) { printf("ERROR : %s\n", hast_last_error(nullptr)); cleanup(); return 1; }

    // ---- sizing: worst case every window is a new k-mer; text is about half sequence ----------
    const int k = (int)o.mer;
    uint64_t est_windows = 0;
    for (const auto* v : {&o.maternal, &o.paternal})
        for (const auto& x : *v) {
            struct stat sb;
            stat(x.c_str(), &sb);
            est_windows += (uint64_t)sb.st_size * (ends_with_gz(x) ? 4 : 1) / 2;
        }
    // a quarter of the windows as distinct k-mers (30x reads: ~1/5); the run restarts larger when the table fills up
    uint64_t expected = o.expected_distinct ? o.expected_distinct : std::max<uint64_t>(est_windows / 4, 1u << 16);
    const char* free_env = getenv("HAST_KC_TABLE_MB");     // cap of one GPU's table (tests / small devices)
    uint64_t cap_slots = (uint64_t)1 << 32;                // 64 GiB of 16-byte slots
    if (free_env) { cap_slots = 1024; while (cap_slots * 2 * 16 <= (uint64_t)atol(free_env) << 20) cap_slots *= 2; }
    long parts = o.parts > 0 ? o.parts : 1;
    if (o.parts <= 0) while (2 * expected / (uint64_t)parts > cap_slots) parts *= 2;
    parts = (parts + n_gpu - 1) / n_gpu * n_gpu;

    const size_t kChunk = 16384;                           // bytes per sequence chunk (device limit 32768)
    const size_t kBatchBases = 32u << 20;
    const int n_batches = 2 * n_gpu + (int)std::min<long>(o.cpu, 16) + 2;
    std::vector<SeqBatch> pool((size_t)n_batches);
    Queue<SeqBatch*> q_free, q_full;
    for (auto& b : pool) {
        b.cap_bases = kBatchBases + kChunk + 4096;
        b.cap_seqs = kBatchBases / 32 + 16;
        void *p0 = nullptr, *p1 = nullptr;
        if (hast_host_alloc(&p0, b.cap_bases) || hast_host_alloc(&p1, (b.cap_seqs + 1) * 4)) {
            printf("ERROR : pinned host allocation failed: %s\n", hast_last_error(nullptr));
            cleanup();
            return 1;
        }
        b.bases = (uint8_t*)p0;
        b.off = (uint32_t*)p1;
        q_free.push(&b);
    }
    auto free_pool = [&] { for (auto& b : pool) { hast_host_free(b.bases); hast_host_free(b.off); } };

    std::vector<uint64_t> histo[2];
    std::vector<uint64_t> lists[2];
    uint64_t windows = 0, distinct[2] = {0, 0}, occupied = 0, text_bytes = 0, records = 0;
    double t_count = 0;
    const uint32_t kHigh = 10000;                          // jellyfish histo default --high
    Bounds bp, bm;
    bool bounds_known = !o.auto_bounds;

    // Two sweeps when the bounds come from the histogram AND the table is partitioned: the histogram
    // needs every partition before any list can be cut.  With one pass per sweep the table is kept.
    struct Files { const std::vector<std::string>* v; int parent; };
    const Files inputs[2] = {{&o.paternal, 0}, {&o.maternal, 1}};

    for (int attempt = 0;; ++attempt) {
        // every GPU owns one partition per pass: recomputed per attempt, a retry may have raised `parts`
        const long passes = parts / n_gpu;
        bool full = false;
        for (auto& h : histo) h.assign(kHigh + 2, 0);
        for (auto& l : lists) l.clear();
        windows = occupied = 0;
        distinct[0] = distinct[1] = 0;
        const int sweeps = (o.auto_bounds && passes > 1) ? 2 : 1;
        for (int sweep = 0; sweep < sweeps && !full; ++sweep) {
            for (long pass = 0; pass < passes && !full; ++pass) {
                for (int g = 0; g < n_gpu; ++g)
                    if (hast_kc_begin(ctx[(size_t)g], k, expected / (uint64_t)parts + 1024, (uint32_t)(pass * n_gpu + g),
                                      (uint32_t)parts) != HAST_OK) {
                        printf("ERROR : %s\n", hast_last_error(ctx[(size_t)g]));
                        free_pool(); cleanup();
                        return 1;
                    }
                // readers: one thread per input file, at most --thread at once
                const double t0 = now();
                std::vector<std::pair<std::string, int>> files;
                for (const auto& in : inputs) for (const auto& f : *in.v) files.emplace_back(f, in.parent);
                std::atomic<size_t> next{0};
                std::atomic<int> left{(int)std::min<size_t>(files.size(), (size_t)o.cpu)};
                std::mutex err_mu;
                std::string err;
                std::vector<std::thread> readers;
                q_full.reopen();
                const int n_readers = left;
                int inflate_threads = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency() * 3u / 4u / (unsigned)std::max(n_readers, 1)));
                if (const char* e = getenv("HAST_INFLATE_THREADS")) inflate_threads = std::max(1, atoi(e));
                for (int r = 0; r < n_readers; ++r)
                    readers.emplace_back([&] {
                        for (;;) {
                            const size_t i = next.fetch_add(1);
                            if (i >= files.size()) break;
                            SeqReader rd(k, kChunk, q_free, q_full, files[i].second, inflate_threads);
                            const std::string e = rd.run(files[i].first);
                            std::lock_guard<std::mutex> lk(err_mu);
                            if (!e.empty() && err.empty()) err = e;
                            if (sweep == 0 && pass == 0) { text_bytes += rd.text_bytes(); records += rd.sequences(); }
                        }
                        if (--left == 0) q_full.finish();
                    });
                SeqBatch* b = nullptr;
                std::string gpu_err;
                while (q_full.pop(b)) {
                    std::vector<uint64_t> tickets((size_t)n_gpu, 0);
                    for (int g = 0; g < n_gpu && gpu_err.empty(); ++g)
                        if (hast_kc_add(ctx[(size_t)g], b->bases, b->n_bases, b->off, b->n_seqs, b->parent, &tickets[(size_t)g]) != HAST_OK)
                            gpu_err = hast_last_error(ctx[(size_t)g]);
                    for (int g = 0; g < n_gpu && gpu_err.empty(); ++g) hast_wait_copied(ctx[(size_t)g], tickets[(size_t)g]);
                    q_free.push(b);
                }
                for (auto& t : readers) t.join();
                if (!err.empty() || !gpu_err.empty()) {
                    printf("ERROR : %s\n", (err.empty() ? gpu_err : err).c_str());
                    free_pool(); cleanup();
                    return 1;
                }
                for (int g = 0; g < n_gpu; ++g) hast_sync(ctx[(size_t)g]);
                t_count += now() - t0;
                // collect this pass
                for (int g = 0; g < n_gpu && !full; ++g) {
                    hast_ctx* c = ctx[(size_t)g];
                    hast_kc_info ki;
                    if (hast_kc_info_get(c, &ki) != HAST_OK) { printf("ERROR : %s\n", hast_last_error(c)); free_pool(); cleanup(); return 1; }
                    if (ki.table_full) { full = true; break; }
                    if (sweep == 0) {
                        windows = ki.windows;
                        occupied += ki.occupied;
                        distinct[0] += ki.distinct[0];
                        distinct[1] += ki.distinct[1];
                        if (o.auto_bounds) {
                            std::vector<uint64_t> h(kHigh + 2);
                            for (int p = 0; p < 2; ++p) {
                                if (hast_kc_histo(c, p, kHigh, h.data()) != HAST_OK) { printf("ERROR : %s\n", hast_last_error(c)); free_pool(); cleanup(); return 1; }
                                for (size_t i = 0; i < h.size(); ++i) histo[p][i] += h[i];
                            }
                        }
                    }
                    const bool last_of_sweep0 = sweep == 0 && pass == passes - 1 && g == n_gpu - 1;
                    if (o.auto_bounds && !bounds_known && last_of_sweep0) {
                        bp = find_bounds(histo[0]);
                        bm = find_bounds(histo[1]);
                        o.plower = bp.lower; o.pupper = bp.upper; o.mlower = bm.lower; o.mupper = bm.upper;
                        bounds_known = true;
                    }
                }
                if (full) break;
                // lists can be cut as soon as the bounds are known: always in the last sweep, and in a single
                // sweep whenever one pass holds everything (all GPUs still hold their tables here)
                if (bounds_known && (sweep == sweeps - 1)) {
                    for (int g = 0; g < n_gpu; ++g) {
                        hast_ctx* c = ctx[(size_t)g];
                        const long lo[2] = {o.plower, o.mlower}, hi[2] = {o.pupper, o.mupper};
                        for (int p = 0; p < 2; ++p) {
                            uint64_t n = 0;
                            const uint32_t l = (uint32_t)std::max<long>(lo[p], 0), u = (uint32_t)std::max<long>(std::min<long>(hi[p], 0xFFFFFFFFl), 0);
                            if (hi[p] < lo[p] || hi[p] < 1) continue;
                            if (hast_kc_select(c, p, l, u, 1, nullptr, 0, &n) != HAST_OK) { printf("ERROR : %s\n", hast_last_error(c)); free_pool(); cleanup(); return 1; }
                            const size_t at = lists[p].size();
                            lists[p].resize(at + n);
                            if (n && hast_kc_select(c, p, l, u, 1, lists[p].data() + at, n, &n) != HAST_OK) { printf("ERROR : %s\n", hast_last_error(c)); free_pool(); cleanup(); return 1; }
                        }
                    }
                }
            }
        }
        if (!full) break;
        if (attempt >= 6) { puts("ERROR : the k-mer count table keeps overflowing; pass --expected-distinct"); free_pool(); cleanup(); return 1; }
        expected *= 4;                                     // more distinct k-mers than estimated: larger tables / more passes
        if (o.parts <= 0) { parts = 1; while (2 * expected / (uint64_t)parts > cap_slots) parts *= 2; parts = (parts + n_gpu - 1) / n_gpu * n_gpu; }
        bounds_known = !o.auto_bounds;
        printf(" count table too small, retrying with room for %llu distinct k-mers in %ld partition(s)\n",
               (unsigned long long)expected, parts);
    }
    for (int g = 0; g < n_gpu; ++g) hast_kc_end(ctx[(size_t)g]);
    free_pool();

    // ---- files ---------------------------------------------------------------------------------
    if (o.auto_bounds) {                                   // analysis_kmercount.sh:7-13
        const char* names[2] = {"paternal", "maternal"};
        const Bounds* bs[2] = {&bp, &bm};
        for (int p = 0; p < 2; ++p) {
            std::string h;
            char line[64];
            for (size_t i = 1; i < histo[p].size(); ++i)
                if (histo[p][i]) { snprintf(line, sizeof line, "%zu %llu\n", i, (unsigned long long)histo[p][i]); h += line; }
            char b[160];
            snprintf(b, sizeof b, "MIN_INDEX=%ld\nMAX_INDEX=%ld\nLOWER_INDEX=%ld\nUPPER_INDEX=%ld\n", bs[p]->min_index,
                     bs[p]->max_index, bs[p]->lower, bs[p]->upper);
            if (!write_file(std::string(names[p]) + ".histo", h) || !write_file(std::string(names[p]) + ".bounds.txt", b)) {
                puts("ERROR : cannot write the histogram / bounds files");
                cleanup();
                return 1;
            }
        }
    }
    printf("  the real used kmer-count bounds of maternal is [ %ld , %ld ] \n", o.mlower, o.mupper);   // :258-259
    printf("  the real used kmer-count bounds of paternal is [ %ld , %ld ] \n", o.plower, o.pupper);
    const char* outs[2] = {kPat, kMat};
    for (int p = 0; p < 2; ++p) {
        if (parts > 1) std::sort(lists[p].begin(), lists[p].end());     // partitions are sorted one by one
        std::string text;
        text.resize(lists[p].size() * (size_t)(k + 1));
        char* w = &text[0];
        for (uint64_t x : lists[p]) {
            for (int j = k - 1; j >= 0; --j) { w[j] = "ACGT"[x & 3]; x >>= 2; }
            w[k] = '\n';
            w += k + 1;
        }
        if (!write_file(outs[p], text)) { printf("ERROR : cannot write %s\n", outs[p]); cleanup(); return 1; }
    }
    {
        time_t t = time(nullptr);
        FILE* f = fopen("step_08_done", "a");
        if (f) { fputs(ctime(&t), f); fclose(f); }
    }
    const double t_total = now() - t_start;
    if (!o.stats_json.empty()) {
        FILE* f = fopen(o.stats_json.c_str(), "w");
        if (f) {
            fprintf(f, "{\"k\": %d, \"gpus\": %d, \"partitions\": %ld, \"records\": %llu, \"text_bytes\": %llu, \"windows\": %llu, "
                       "\"distinct_paternal\": %llu, \"distinct_maternal\": %llu, \"distinct_union\": %llu, "
                       "\"paternal_unique_filter\": %zu, \"maternal_unique_filter\": %zu, "
                       "\"bounds\": [%ld, %ld, %ld, %ld], \"t_count_s\": %.4f, \"t_total_s\": %.4f}\n",
                    k, n_gpu, parts, (unsigned long long)records, (unsigned long long)text_bytes, (unsigned long long)windows,
                    (unsigned long long)distinct[0], (unsigned long long)distinct[1], (unsigned long long)occupied,
                    lists[0].size(), lists[1].size(), o.plower, o.pupper, o.mlower, o.mupper, t_count, t_total);
            fclose(f);
        }
    }
    cleanup();
    report();
    return 0;
}
