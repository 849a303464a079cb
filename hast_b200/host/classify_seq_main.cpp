// classify_seq_main.cpp -- `classify_seq`, the drop-in for HAST stage 03's per-sequence
// classifier (03.mkoutput_by_fabulous2.0/src_main/classify.cpp, installed there as bin/classify
// and called by mkoutput_by_fabulous2.0.sh:121-123 on the phased-bubble FASTA).
//
//     classify_seq --hap hap0.kmer --hap hap1.kmer --read seq.fa [--read more.fa]
//                  [--thread N] [--format fasta|fastq]  > phasing.out
//
// Same argv, same stdout (`name \t haplotype0|haplotype1|ambiguous \t score`, one line per
// sequence in input order, a block per --read file), usage + exit 255 on bad arguments.
// The k-mer work runs on the GPU through the C ABI: the two lists go into the same table as
// in stage 01, every sequence is cut into chunks that overlap by k-1 bases, each chunk is a
// "read" whose "barcode" is the sequence id, and the fused kernel runs with the stage-03
// window rule (hast_set_option "seq_mode").  What stays on the host is the reference's
// arithmetic on the two counts: the division by the number of list LINES (classify.cpp:68,
// 215-216) and the call (:104-135).
//
// Reference behaviour restated here, with its line numbers:
//   :52-72    k = length of the first line of the first list; a list contributes its line count
//             (duplicates included, an unterminated last line dropped) as the divisor;
//             lines of another length can never match
//   :203-218  every window of k bytes is looked up as a STRING in both sets (which hold each
//             k-mer and its reverse complement), so only upper-case ACGT windows can match
//   :104-135  best / second-best rule; equal non-zero scores give haplotype0 (second stays 0)
//   :272-300  FASTA: blank lines skipped, sequence lines concatenated, name = header minus '>';
//             a line starting with '@' or '+' is a format error (exit 1)
//   :250-268  FASTQ: strict 4-line records, name = header minus '@'; '>' is a format error
// Not reproduced: k-mer lists with letters other than upper-case ACGT (string matching would
// honour them; the 2-bit table cannot) -- such a list is rejected with an error.
#include <getopt.h>
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hast_b200.h"

namespace {

void usage() {
    fputs("Uasge :\n\tclassify_read --hap hap0.kmer --hap hap1.kmer --read read.fa [--read read_2.fa] "
          "[--thread t_num (8 default) ] [--format fasta/fastq (default fasta)] \n"
          "notice : --read accept file in gzip format , but file must end by \".gz\"\n"
          "warn   : --read default only accept fasta read.\n"
          "         add --format fastq if --read refer to fastq file.\n", stderr);
}


// getline over a plain or gzip file (gzread passes plain text through)
class Lines {
public:
    bool open(const std::string& path) {
        gz_ = gzopen(path.c_str(), "rb");
        if (gz_) gzbuffer(gz_, 1u << 20);
        return gz_ != nullptr;
    }
    ~Lines() { if (gz_) gzclose(gz_); }
    // returns false when the read hit end of file (the reference's `getline(...).eof()`): a final
    // line without '\n' is therefore reported as false with its text in `out`
    bool next(std::string& out) {
        out.clear();
        for (;;) {
            if (pos_ == len_) {
                const int r = gzread(gz_, buf_, sizeof buf_);
                if (r <= 0) return false;
                pos_ = 0;
                len_ = (size_t)r;
            }
            const char* nl = (const char*)memchr(buf_ + pos_, '\n', len_ - pos_);
            if (nl) {
                out.append(buf_ + pos_, (size_t)(nl - (buf_ + pos_)));
                pos_ = (size_t)(nl - buf_) + 1;
                return true;
            }
            out.append(buf_ + pos_, len_ - pos_);
            pos_ = len_;
        }
    }
private:
    gzFile gz_ = nullptr;
    char buf_[1 << 16];
    size_t pos_ = 0, len_ = 0;
};

struct List { std::string text; long long lines = 0; };

// load_kmers, classify.cpp:52-72
int load_list(const std::string& path, int index, int& k, List& out) {
    Lines in;
    if (!in.open(path)) { fprintf(stderr, "ERROR : cannot open %s\n", path.c_str()); return 1; }
    std::string line;
    bool first = index == 0;
    for (;;) {
        const bool terminated = in.next(line);
        if (!terminated && !first) break;                  // `while(!getline().eof())` drops the tail
        if (first) { k = (int)line.size(); first = false; }
        ++out.lines;
        if ((int)line.size() == k) {
            for (char c : line)
                if (c != 'A' && c != 'C' && c != 'G' && c != 'T') {
                    fprintf(stderr, "ERROR : %s holds a k-mer with letters other than upper-case ACGT (%s); "
                                    "the GPU table cannot represent it\n", path.c_str(), line.c_str());
                    return 1;
                }
            out.text += line;
            out.text += '\n';
        }
        if (!terminated) break;
    }
    fprintf(stderr, "Recorded %lld haplotype %d specific %d-mers\n", out.lines, index, k);
    return 0;
}

// PrintOutput, classify.cpp:104-135
void print_call(const std::string& name, const double hc[2]) {
    double best = 0, second = 0;
    int hap = -1;
    for (int i = 0; i < 2; ++i) {
        if (hc[i] > 0 && hc[i] < best && hc[i] > second) second = hc[i];
        if (hc[i] > 0 && hc[i] > best) { hap = i; second = best; best = hc[i]; }
    }
    if (second == 0 && best != 0) printf("%s\thaplotype%d\t%0.6f\n", name.c_str(), hap, best);
    else if (best == 0 && second == 0) printf("%s\t%s\t0.0\n", name.c_str(), "ambiguous");
    else if (best / second > 1) printf("%s\thaplotype%d\t%0.6f\n", name.c_str(), hap, best);
    else printf("%s\t%s\t%0.6f\n", name.c_str(), "ambiguous", best);
}

struct Batcher {
    hast_ctx* ctx;
    int k;
    size_t chunk;                       // bases per chunk; consecutive chunks overlap by k-1
    std::vector<uint8_t> bases;
    std::vector<uint32_t> off{0}, id;
    static constexpr size_t kFlushBases = 64u << 20;
    int add_sequence(const std::string& seq, uint32_t seq_id) {
        const size_t step = chunk - (size_t)(k - 1);
        size_t s = 0;
        do {                                               // at least one (possibly empty) chunk per sequence
            const size_t n = std::min(chunk, seq.size() - s);
            bases.insert(bases.end(), seq.begin() + (long)s, seq.begin() + (long)(s + n));
            off.push_back((uint32_t)bases.size());
            id.push_back(seq_id);
            if (s + n >= seq.size()) break;
            s += step;
        } while (true);
        return bases.size() >= kFlushBases ? flush() : 0;
    }
    int flush() {
        if (id.empty()) return 0;
        uint64_t ticket = 0;
        int rc = hast_submit_batch(ctx, bases.data(), bases.size(), off.data(), id.data(), (uint32_t)id.size(), &ticket);
        if (rc == HAST_OK) rc = hast_wait_copied(ctx, ticket);      // pageable vectors are reused below
        bases.clear(); off.assign(1, 0); id.clear();
        return rc;
    }
};

int fail_ctx(hast_ctx* c) { fprintf(stderr, "ERROR : %s\n", hast_last_error(c)); return 1; }

int process_file(hast_ctx* ctx, const std::string& path, bool fasta, int k, const long long total[2]) {
    Lines in;
    if (!in.open(path)) { fprintf(stderr, "ERROR : cannot open %s\n", path.c_str()); return 1; }
    if (hast_reset_counts(ctx) != HAST_OK) return fail_ctx(ctx);
    std::vector<std::string> names;
    Batcher b{ctx, k, (size_t)8192};
    uint64_t reserved = 0;
    auto submit = [&](const std::string& head, const std::string& seq) -> int {
        const uint32_t sid = (uint32_t)names.size();
        names.push_back(head.empty() ? std::string() : head.substr(1));   // classify.cpp:206
        if (names.size() > reserved) {
            reserved = std::max<uint64_t>(names.size() * 2, 1u << 16);
            if (hast_reserve_barcodes(ctx, reserved) != HAST_OK) return fail_ctx(ctx);
        }
        if (b.add_sequence(seq, sid) != HAST_OK) return fail_ctx(ctx);
        return 0;
    };
    std::string head, seq, tmp;
    if (fasta) {                                            // processFasta, classify.cpp:272-300
        long long id = 0;
        for (;;) {
            const bool ok = in.next(tmp);
            if (!ok) break;                                 // the line that hits EOF is dropped, like the reference
            if (tmp.empty()) continue;
            if (tmp[0] == '@' || tmp[0] == '+') {
                fputs("fasta detected . ERROR . please use \"--format fastq\". exit ... \n", stderr);
                return 1;
            }
            if (tmp[0] == '>') {
                if (id > 0 && submit(head, seq)) return 1;
                std::swap(head, tmp);
                seq.clear();
                ++id;
            } else {
                seq += tmp;
            }
        }
        if (id == 0) { fprintf(stderr, "ERROR : no FASTA record in %s\n", path.c_str()); return 1; }
        if (submit(head, seq)) return 1;
    } else {                                                // processFastq, classify.cpp:250-268
        for (;;) {
            if (!in.next(head)) break;
            if (!head.empty() && head[0] == '>') {
                fputs("fasta detected . ERROR . please use \"--format fasta\". exit ... \n", stderr);
                return 1;
            }
            in.next(seq);
            in.next(tmp);
            in.next(tmp);
            if (submit(head, seq)) return 1;
        }
    }
    if (b.flush() != HAST_OK) return fail_ctx(ctx);
    const uint64_t n = names.size();
    std::vector<int32_t> counts(std::max<uint64_t>(n, 1) * 2, 0);
    if (n) {
        if (hast_reserve_barcodes(ctx, n) != HAST_OK) return fail_ctx(ctx);
        if (hast_finish(ctx, counts.data(), n) != HAST_OK) return fail_ctx(ctx);
    }
    for (uint64_t i = 0; i < n; ++i) {
        double hc[2];
        for (int j = 0; j < 2; ++j) {
            hc[j] = (double)counts[2 * i + j];              // classify.cpp:213 `hapCounts[j] ++` on a double
            hc[j] /= total[j];                              // :215-216, int divisor
        }
        print_call(names[i], hc);
    }
    fflush(stdout);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    static struct option lo[] = {{"hap", required_argument, nullptr, 'p'},    {"read", required_argument, nullptr, 'r'},
                                 {"format", required_argument, nullptr, 'f'}, {"thread", required_argument, nullptr, 't'},
                                 {"help", no_argument, nullptr, 'h'},         {nullptr, 0, nullptr, 0}};
    std::vector<std::string> haps, reads;
    std::string format = "fasta";
    int threads = 8;
    for (;;) {
        const int c = getopt_long(argc, argv, "p:r:t:f:h", lo, nullptr);
        if (c < 0) break;
        switch (c) {
            case 'p': haps.emplace_back(optarg); break;
            case 'r': reads.emplace_back(optarg); break;
            case 'f': format = optarg; break;
            case 't': threads = atoi(optarg); break;
            default: usage(); return 255;
        }
    }
    if (haps.size() != 2 || reads.empty() || threads < 1) { usage(); return 255; }
    if (format != "fasta" && format != "fastq") {
        fprintf(stderr, " ERROR : invalid format : [%s] . exit ...\n", format.c_str());
        return 255;
    }
    // stdout carries the table only: keep libraries underneath from writing into it
    fflush(stdout);
    fprintf(stderr, "__START__\n");
    if (hast_device_count() <= 0) {
        fprintf(stderr, "ERROR : no CUDA device found; this build has no CPU path\n");
        return 1;
    }
    int k = 0;
    List lists[2];
    for (int i = 0; i < 2; ++i) {
        fprintf(stderr, "__load hap%d kmers__\n", i);
        if (load_list(haps[(size_t)i], i, k, lists[i])) return 1;
    }
    if (k < 1 || k > 32) { fprintf(stderr, "ERROR : k = %d is outside 1..32\n", k); return 1; }
    hast_ctx* ctx = nullptr;
    if (hast_create(0, &ctx) != HAST_OK) { fprintf(stderr, "ERROR : %s\n", hast_last_error(nullptr)); return 1; }
    int rc = 0;
    const uint64_t n_keys = (lists[0].text.size() + lists[1].text.size()) / (size_t)(k + 1);
    if (hast_set_option(ctx, "seq_mode", 1) != HAST_OK || hast_table_begin(ctx, k, std::max<uint64_t>(n_keys, 16)) != HAST_OK ||
        hast_table_add_text(ctx, lists[0].text.data(), lists[0].text.size() / (size_t)(k + 1), 0) != HAST_OK ||
        hast_table_add_text(ctx, lists[1].text.data(), lists[1].text.size() / (size_t)(k + 1), 1) != HAST_OK)
        rc = fail_ctx(ctx);
    const long long total[2] = {lists[0].lines, lists[1].lines};
    for (size_t i = 0; i < reads.size() && rc == 0; ++i) {
        fprintf(stderr, "__process read: %s\n", reads[i].c_str());
        rc = process_file(ctx, reads[i], format == "fasta", k, total);
        fprintf(stderr, "__process read done__\n");
    }
    hast_destroy(ctx);
    if (rc == 0) fprintf(stderr, "__END__\n");
    return rc;
}
