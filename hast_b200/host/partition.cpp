// partition.cpp -- the read partitioner and the barcode-list split of stage 01.
//
// Replaces, byte for byte, the second half of classify_stlfr_reads.sh:
//   :156-162  three awk passes over phased.barcodes -> {paternal,maternal,homozygous}.unique.barcodes
//   :176-185  `[gzip -dc |] awk -v prefix=NAME -F '#|/' -f quartering_fastq.awk P M H INPUT`
//             -> NAME.{paternal,maternal,homozygous,nobarcode}.fastq + filter_reads.log
// (quartering_fastq.awk:1-61).  The awk program is single-threaded and, for gzip
// input, sits behind a second full `gzip -dc` of every FASTQ; here the inflate runs
// in a reader thread (the FastqSource of the classify pass) and the routing is a
// memchr scan plus one hash probe per record.
//
// awk semantics that are reproduced on purpose:
//   * fields are cut at EVERY '#' or '/' of the header line (FS = '#|/'), so the barcode
//     is $2 = the text between the first and the second separator -- not parseName's
//     last-'#'/last-'/' rule (classify.cpp:112-119); the two agree on stLFR names
//   * a list line contributes its $1 under the same FS; lookups go paternal, maternal,
//     homozygous in that order (quartering_fastq.awk:25-34)
//   * `NF > 1 && $2 != "0_0_0"` else no-barcode (:24,41-44); a barcode in no list is
//     reported on stderr and the record is DROPPED (:35-39)
//   * the type of a record is decided on lines with FNR % 4 == 1 and sticks to the
//     following lines (:22,46-54); every printed line gets "\n" appended
//   * an output file exists only if at least one line was routed to it (awk opens on
//     first print, truncating); filter_reads.log is appended to (:20,57-61)
#include <fcntl.h>
#include <unistd.h>

#include <cerrno>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "fastq_source.h"
#include "host.h"
#include "plain_slicer.h"

namespace hasthost {

namespace {

inline uint64_t hash_bytes(const char* s, size_t n) {          // FNV-1a, 64 bit, finished with a mix
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 0x100000001b3ull; }
    h ^= h >> 32;
    return h * 0x9E3779B97F4A7C15ull;
}

inline size_t first_sep(const char* s, size_t n) {             // index of the first '#' or '/', n if none
    for (size_t i = 0; i < n; ++i)
        if (s[i] == '#' || s[i] == '/') return i;
    return n;
}

class OutFile {                                                 // awk's `print > file`: created on first use
public:
    OutFile(std::string path) : path_(std::move(path)) { buf_.reserve(kFlush + (1u << 16)); }
    ~OutFile() { close(); }
    bool put(const char* p, size_t n) {                         // one line, "\n" appended (ORS)
        buf_.insert(buf_.end(), p, p + n);
        buf_.push_back('\n');
        used_ = true;
        return buf_.size() < kFlush || flush();
    }
    bool flush() {
        if (!used_ || buf_.empty()) return true;
        if (fd_ < 0) {
            if (opened_) { err_ = path_ + " already closed"; return false; }
            opened_ = true;
            fd_ = ::open(path_.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
            if (fd_ < 0) { err_ = "cannot open " + path_ + " for writing: " + strerror(errno); return false; }
        }
        size_t off = 0;
        while (off < buf_.size()) {
            ssize_t w = ::write(fd_, buf_.data() + off, buf_.size() - off);
            if (w < 0) {
                if (errno == EINTR) continue;
                err_ = "write to " + path_ + " failed: " + strerror(errno);
                return false;
            }
            off += (size_t)w;
        }
        buf_.clear();
        return true;
    }
    bool close() {
        bool ok = flush();
        if (fd_ >= 0) { ::close(fd_); fd_ = -1; }
        return ok;
    }
    const std::string& error() const { return err_; }
private:
    static constexpr size_t kFlush = 4u << 20;
    std::string path_, err_;
    std::vector<char> buf_;
    int fd_ = -1;
    bool used_ = false, opened_ = false;
};

}  // namespace

// ---- BarcodeLists: the three awk associative arrays as one open-addressing table ----
void BarcodeLists::add(const char* line, size_t n, int type) {
    n = first_sep(line, n);                                     // $1 under FS '#|/'
    if (slots_.empty() || (count_ + 1) * 2 > slots_.size()) grow();
    const uint64_t h = hash_bytes(line, n);
    size_t i = h & (slots_.size() - 1);
    for (;; i = (i + 1) & (slots_.size() - 1)) {
        Slot& s = slots_[i];
        if (s.len == kEmpty) {
            s.off = arena_.size();
            s.len = (uint32_t)n;
            s.hash = h;
            s.types = (uint8_t)(1u << type);
            arena_.insert(arena_.end(), line, line + n);
            ++count_;
            return;
        }
        if (s.hash == h && s.len == n && memcmp(arena_.data() + s.off, line, n) == 0) {
            s.types |= (uint8_t)(1u << type);                   // the same key may sit in several lists
            return;
        }
    }
}

void BarcodeLists::grow() {
    std::vector<Slot> old;
    old.swap(slots_);
    slots_.assign(old.empty() ? 1024 : old.size() * 2, Slot{});
    for (const Slot& s : old) {
        if (s.len == kEmpty) continue;
        size_t i = s.hash & (slots_.size() - 1);
        while (slots_[i].len != kEmpty) i = (i + 1) & (slots_.size() - 1);
        slots_[i] = s;
    }
}

int BarcodeLists::find(const char* key, size_t n) const {
    if (slots_.empty()) return 0;
    const uint64_t h = hash_bytes(key, n);
    for (size_t i = h & (slots_.size() - 1);; i = (i + 1) & (slots_.size() - 1)) {
        const Slot& s = slots_[i];
        if (s.len == kEmpty) return 0;
        if (s.hash == h && s.len == n && memcmp(arena_.data() + s.off, key, n) == 0) {
            if (s.types & (1u << kPaternal)) return kPaternal;   // quartering_fastq.awk:25,28,31
            if (s.types & (1u << kMaternal)) return kMaternal;
            return kHomozygous;
        }
    }
}

std::string BarcodeLists::load(const std::string& path, int type) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return "cannot open barcode list " + path;
    std::vector<char> buf(1u << 20);
    std::string carry;
    size_t n;
    while ((n = fread(buf.data(), 1, buf.size(), f)) > 0) {
        size_t s = 0;
        for (;;) {
            const char* nl = (const char*)memchr(buf.data() + s, '\n', n - s);
            if (!nl) { carry.append(buf.data() + s, n - s); break; }
            const size_t e = (size_t)(nl - buf.data());
            if (!carry.empty()) {
                carry.append(buf.data() + s, e - s);
                add(carry.data(), carry.size(), type);
                carry.clear();
            } else {
                add(buf.data() + s, e - s, type);
            }
            s = e + 1;
        }
    }
    if (!carry.empty()) add(carry.data(), carry.size(), type);   // awk reads an unterminated last line too
    fclose(f);
    return "";
}

// classify_stlfr_reads.sh:156-162 -- the three lists straight from the calls.
//   awk '{if($2 == 0) print $1;}'  /  '{if($2 == 1) ...}'  /  '{if($2 == "-1") ...}'   phased.barcodes
// run with the default FS (runs of blanks), so $1 / $2 are the first two blank-separated tokens of the
// row `barcode \t call \t c0 \t c1`: the barcode and the call unless the barcode is empty or holds
// blanks.  `$2 == 0` is numeric when $2 looks like a number (strnum), a string comparison otherwise.
namespace {
bool awk_num_equals(const std::string& tok, double v) {
    if (tok.empty()) return false;
    char* end = nullptr;
    const double d = strtod(tok.c_str(), &end);
    if (end == tok.c_str() || *end != '\0') return tok == (v == 0 ? "0" : "1");   // not a number: string compare
    return d == v;
}
}  // namespace

std::string write_barcode_lists(const std::string& dir, const std::vector<std::string>& names,
                                const std::vector<uint32_t>& order, const std::vector<int8_t>& haps,
                                const int32_t* counts, uint64_t counts_out[3], BarcodeLists* lists) {
    static const char* kNames[3] = {"paternal.unique.barcodes", "maternal.unique.barcodes",
                                    "homozygous.unique.barcodes"};
    static const int kType[3] = {BarcodeLists::kPaternal, BarcodeLists::kMaternal, BarcodeLists::kHomozygous};
    counts_out[0] = counts_out[1] = counts_out[2] = 0;
    FILE* f[3];
    for (int i = 0; i < 3; ++i) {
        const std::string p = dir.empty() ? kNames[i] : dir + "/" + kNames[i];
        f[i] = fopen(p.c_str(), "wb");
        if (!f[i]) {
            for (int j = 0; j < i; ++j) fclose(f[j]);
            return "cannot open " + p + " for writing";
        }
    }
    std::vector<std::string> tok;
    for (size_t i = 0; i < order.size(); ++i) {
        const uint32_t id = order[i];
        const std::string& nm = names[id];
        const std::string* t1;
        std::string t2;
        bool simple = !nm.empty();
        for (char c : nm) if (c == ' ' || c == '\t') { simple = false; break; }
        if (simple) {
            t1 = &nm;
            t2 = std::to_string((int)haps[i]);
        } else {                                               // the general awk tokenisation of the row
            tok.clear();
            size_t s = 0;
            while (s < nm.size()) {
                while (s < nm.size() && (nm[s] == ' ' || nm[s] == '\t')) ++s;
                size_t e = s;
                while (e < nm.size() && nm[e] != ' ' && nm[e] != '\t') ++e;
                if (e > s) tok.emplace_back(nm, s, e - s);
                s = e;
            }
            tok.push_back(std::to_string((int)haps[i]));
            tok.push_back(std::to_string(counts[2 * (size_t)id]));
            tok.push_back(std::to_string(counts[2 * (size_t)id + 1]));
            t1 = &tok[0];
            t2 = tok[1];
        }
        const bool sel[3] = {awk_num_equals(t2, 0.0), awk_num_equals(t2, 1.0), t2 == "-1"};
        for (int w = 0; w < 3; ++w) {
            if (!sel[w]) continue;
            fwrite(t1->data(), 1, t1->size(), f[w]);
            fputc('\n', f[w]);
            ++counts_out[w];
            if (lists) lists->add(t1->data(), t1->size(), kType[w]);
        }
    }
    for (int i = 0; i < 3; ++i) fclose(f[i]);
    return "";
}

// quartering_fastq.awk on one input file, on several threads.
//
// The awk program is a serial pass; here blocks of whole records (plain files: slices of the mapping,
// plain_slicer.h; gzip / pipes: what the reader thread inflates) are routed by `threads` workers into four private
// buffers each, and committed in input order: a short critical section hands every block its byte offsets in the
// four output files (and adds up the statistics and the "unclassify barcode" messages, in order), the bytes
// themselves are then written with pwrite() by all workers at once.  The files come out byte-identical to awk's.
namespace {

struct Routed {
    std::vector<char> buf[4];
    PartitionStats st;
    std::string msgs;
    bool any_line = false;
};

// one block of whole records -> r (quartering_fastq.awk:22-55)
void route_block(const char* p, size_t len, const BarcodeLists& lists, Routed& r) {
    for (auto& b : r.buf) b.clear();
    r.st = PartitionStats{};
    r.msgs.clear();
    r.any_line = len > 0;
    uint64_t fnr = 0;
    int type = 0;          // awk: uninitialised read_type compares equal to 0; set before use anyway
    size_t s = 0;
    while (s < len) {
        const char* nl = (const char*)memchr(p + s, '\n', len - s);
        const size_t e = nl ? (size_t)(nl - p) : len;          // an unterminated last line is a record too
        ++fnr;
        if (fnr % 4 == 1) {                                    // quartering_fastq.awk:22
            ++r.st.total;
            const char* h = p + s;
            const size_t hl = e - s;
            const size_t a = first_sep(h, hl);
            if (a < hl) {                                      // NF > 1
                const size_t b = a + 1 + first_sep(h + a + 1, hl - a - 1);
                const char* key = h + a + 1;
                const size_t kl = b - a - 1;
                if (kl == 5 && memcmp(key, "0_0_0", 5) == 0) {
                    ++r.st.no_barcode; type = 0;
                } else {
                    const int t = lists.find(key, kl);
                    if (t == BarcodeLists::kPaternal) { ++r.st.paternal; type = 1; }
                    else if (t == BarcodeLists::kMaternal) { ++r.st.maternal; type = 2; }
                    else if (t == BarcodeLists::kHomozygous) { ++r.st.homozygous; type = 3; }
                    else {
                        r.msgs += "ERROR : unclassify barcode : ";
                        r.msgs.append(key, kl);
                        r.msgs += '\n';
                        ++r.st.unclassified; type = -1;
                    }
                }
            } else {
                ++r.st.no_barcode; type = 0;
            }
        }
        if (type >= 0) {
            std::vector<char>& o = r.buf[type];
            o.insert(o.end(), p + s, p + e);
            o.push_back('\n');                                 // ORS
        }
        s = e + 1;
    }
}

bool pwrite_all(int fd, const char* p, size_t n, uint64_t off, std::string& err, const std::string& path) {
    while (n) {
        const ssize_t w = ::pwrite(fd, p, n, (off_t)off);
        if (w < 0) {
            if (errno == EINTR) continue;
            err = "write to " + path + " failed: " + strerror(errno);
            return false;
        }
        p += w; n -= (size_t)w; off += (uint64_t)w;
    }
    return true;
}

}  // namespace

std::string partition_fastq(const std::string& input, const std::string& display_name, const std::string& prefix,
                            const std::string& outdir, const BarcodeLists& lists, PartitionStats& st, int threads) {
    st = PartitionStats{};
    if (threads <= 0) threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const std::string base = outdir.empty() ? prefix : outdir + "/" + prefix;
    const std::string paths[4] = {base + ".nobarcode.fastq", base + ".paternal.fastq", base + ".maternal.fastq",
                                  base + ".homozygous.fastq"};
    // input: slices of a mapped plain file, or blocks from a reader thread
    const size_t n_in = input.size();
    const bool gz = n_in > 3 && input.compare(n_in - 3, 3, ".gz") == 0;
    PlainSlicer slicer;
    bool sliced = false;
    if (!gz && input != "-" && !getenv("HAST_SERIAL_READER")) {
        size_t slice_bytes = (size_t)4 << 20;
        if (const char* sb = getenv("HAST_SLICE_BYTES")) slice_bytes = (size_t)std::max(1L, atol(sb));
        const std::string e = slicer.open(input, slice_bytes);
        if (!e.empty()) return e;
        sliced = slicer.usable();
    }
    FastqSource src;
    if (!sliced) {
        const std::string e = src.open(input, std::max(1, std::min(8, threads / 2)));
        if (!e.empty()) return e;
    }

    std::mutex mu;
    std::condition_variable cv;
    // reader-thread mode
    const size_t n_blocks = (size_t)threads + 2;
    std::vector<TextBlock> blocks(sliced ? 0 : n_blocks);
    std::deque<std::pair<size_t, uint64_t>> ready;             // (block id, sequence number)
    std::deque<size_t> free_ids;
    for (size_t i = 0; i < blocks.size(); ++i) free_ids.push_back(i);
    bool read_done = sliced, failed = false;
    std::string rerr, werr;
    // ordered commit
    uint64_t next_seq = 0, off[4] = {0, 0, 0, 0}, text_bytes = 0;
    int fds[4] = {-1, -1, -1, -1};
    bool any_line = false;

    std::thread reader;
    if (!sliced)
        reader = std::thread([&] {
            uint64_t seq = 0;
            for (;;) {
                size_t id;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return !free_ids.empty() || failed; });
                    if (failed) break;
                    id = free_ids.front();
                    free_ids.pop_front();
                }
                std::string e;
                const bool more = src.next(blocks[id], (size_t)8 << 20, e);
                std::lock_guard<std::mutex> lk(mu);
                if (!e.empty()) { rerr = e; failed = true; }
                if (!more || !e.empty()) break;
                ready.emplace_back(id, seq++);
                cv.notify_all();
            }
            std::lock_guard<std::mutex> lk(mu);
            read_done = true;
            cv.notify_all();
        });

    auto worker = [&] {
        Routed r;
        TextBlock own;
        for (;;) {
            const char* p = nullptr;
            size_t len = 0, id = 0;
            uint64_t seq = 0;
            if (sliced) {
                if (!slicer.next(own, nullptr, &seq)) break;
                p = own.text();
                len = own.len;
            } else {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !ready.empty() || read_done || failed; });
                if (failed || ready.empty()) break;
                id = ready.front().first;
                seq = ready.front().second;
                ready.pop_front();
                p = blocks[id].text();
                len = blocks[id].len;
            }
            route_block(p, len, lists, r);
            uint64_t my_off[4];
            bool stop = false;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return next_seq == seq || failed; });
                if (failed) stop = true;
                else {
                    for (int i = 0; i < 4 && !stop; ++i) {
                        my_off[i] = off[i];
                        if (r.buf[i].empty()) continue;
                        if (fds[i] < 0) {                          // awk's `print > file`: created on first use
                            fds[i] = ::open(paths[i].c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
                            if (fds[i] < 0) { werr = "cannot open " + paths[i] + " for writing: " + strerror(errno); failed = stop = true; }
                        }
                        off[i] += r.buf[i].size();
                    }
                    if (!r.msgs.empty()) fwrite(r.msgs.data(), 1, r.msgs.size(), stderr);
                    st.total += r.st.total; st.no_barcode += r.st.no_barcode; st.paternal += r.st.paternal;
                    st.maternal += r.st.maternal; st.homozygous += r.st.homozygous; st.unclassified += r.st.unclassified;
                    any_line |= r.any_line;
                    text_bytes += len;
                    ++next_seq;
                }
                if (!sliced) { blocks[id].hold.reset(); free_ids.push_back(id); }
                cv.notify_all();
            }
            if (stop) break;
            for (int i = 0; i < 4; ++i) {
                std::string e;
                if (!r.buf[i].empty() && !pwrite_all(fds[i], r.buf[i].data(), r.buf[i].size(), my_off[i], e, paths[i])) {
                    std::lock_guard<std::mutex> lk(mu);
                    if (werr.empty()) werr = e;
                    failed = true;
                    cv.notify_all();
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!read_done) failed = failed || !werr.empty();
        cv.notify_all();
    }
    if (reader.joinable()) {
        { std::lock_guard<std::mutex> lk(mu); if (!werr.empty()) failed = true; cv.notify_all(); }
        reader.join();
    }
    for (int i = 0; i < 4; ++i)
        if (fds[i] >= 0 && ::close(fds[i]) != 0 && werr.empty()) werr = "close of " + paths[i] + " failed: " + strerror(errno);
    if (!rerr.empty()) return rerr;
    if (!werr.empty()) return werr;

    // filter_reads.log (quartering_fastq.awk:19-21,56-61), appended
    const std::string logp = outdir.empty() ? "filter_reads.log" : outdir + "/filter_reads.log";
    FILE* lg = fopen(logp.c_str(), "ab");
    if (!lg) return "cannot open " + logp;
    if (any_line) fprintf(lg, "%s\n", display_name.c_str());
    fprintf(lg, "#Total reads                : %llu \n", (unsigned long long)st.total);
    fprintf(lg, "#Reads without barcode      : %llu \n", (unsigned long long)st.no_barcode);
    fprintf(lg, "#Paternal reads             : %llu \n", (unsigned long long)st.paternal);
    fprintf(lg, "#Maternal reads             : %llu \n", (unsigned long long)st.maternal);
    fprintf(lg, "#Homozygous reads           : %llu \n", (unsigned long long)st.homozygous);
    fclose(lg);
    st.text_bytes = text_bytes;
    return "";
}

// `name=`basename $x`; name=${name%%.gz}`, classify_stlfr_reads.sh:178-180
std::string partition_prefix(const std::string& path) {
    size_t s = path.find_last_of('/');
    std::string name = s == std::string::npos ? path : path.substr(s + 1);
    if (name.size() >= 3 && name.compare(name.size() - 3, 3, ".gz") == 0) name.resize(name.size() - 3);
    return name;
}

}  // namespace hasthost
