// fastq_source.cpp -- see fastq_source.h
#include "fastq_source.h"

#include <emmintrin.h>
#include <fcntl.h>
#include <unistd.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>

namespace hasthost {

size_t count_newlines(const char* p, size_t n) {
    size_t c = 0, i = 0;
    const __m128i nl = _mm_set1_epi8('\n');
    for (; i + 16 <= n; i += 16) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p + i));
        c += (size_t)__builtin_popcount((unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(v, nl)));
    }
    for (; i < n; ++i) c += p[i] == '\n';
    return c;
}

FastqSource::~FastqSource() {
    if (gz_) gzclose(gz_);
    if (fd_ >= 0) close(fd_);
}

std::string FastqSource::open(const std::string& path, int inflate_threads) {
    path_ = path;
    const size_t n = path.size();
    const bool gz = n > 3 && path[n - 3] == '.' && path[n - 2] == 'g' && path[n - 1] == 'z';   // classify.cpp:245-250
    if (path == "-") {                                 // standard input (plain text), as `awk ... -`
        fd_ = dup(STDIN_FILENO);
        if (fd_ < 0) return std::string("cannot read standard input: ") + strerror(errno);
    } else if (gz) {
        const char* force = getenv("HAST_ZLIB");
        if (!(force && force[0] == '1') && inflate_threads > 1) {
            std::unique_ptr<ParallelGzip> pinf(new ParallelGzip(inflate_threads));
            if (pinf->open(path).empty() && pinf->is_gzip()) {
                pinf_ = std::move(pinf);
                return "";
            }
        }
        if (!(force && force[0] == '1')) {
            std::unique_ptr<GzipInflater> inf(new GzipInflater());
            if (inf->open(path).empty() && inf->is_gzip()) {
                inf_ = std::move(inf);
                return "";
            }
        }
        gz_ = gzopen(path.c_str(), "rb");
        if (!gz_) return "cannot open " + path;
        gzbuffer(gz_, 1u << 20);
    } else {
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) return "cannot open " + path + ": " + strerror(errno);
#ifdef POSIX_FADV_SEQUENTIAL
        posix_fadvise(fd_, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
    }
    return "";
}

size_t FastqSource::raw_read(char* dst, size_t n, std::string& err) {
    if (inf_ || pinf_) {
        size_t got = 0;
        while (got < n) {
            if (!chunk_left_) {
                const bool more = inf_ ? inf_->next(&chunk_, &chunk_left_) : pinf_->next(&chunk_, &chunk_left_);
                if (!more) {
                    const std::string& e = inf_ ? inf_->error() : pinf_->error();
                    if (!e.empty()) { err = "inflate failed on " + path_ + ": " + e; return 0; }
                    break;
                }
            }
            const size_t m = std::min(n - got, chunk_left_);
            memcpy(dst + got, chunk_, m);
            got += m;
            chunk_ += m;
            chunk_left_ -= m;
            if (got >= (1u << 20)) break;              // hand back what one call of read() would: whole chunks, not the whole block
        }
        return got;
    }
    if (gz_) {
        const unsigned want = (unsigned)std::min<size_t>(n, 1u << 30);
        const int r = gzread(gz_, dst, want);
        if (r < 0) {
            int e = 0;
            err = std::string("gzread failed on ") + path_ + ": " + gzerror(gz_, &e);
            return 0;
        }
        return (size_t)r;
    }
    for (;;) {
        const ssize_t r = ::read(fd_, dst, n);
        if (r >= 0) return (size_t)r;
        if (errno == EINTR) continue;
        err = "read failed on " + path_ + ": " + strerror(errno);
        return 0;
    }
}

// Parallel decoder: its output pieces are handed on as they are (no copy).  A record that straddles two pieces is
// made contiguous by writing the tail of the previous piece into the free space in front of the next one; the
// pass that finds the record boundary also builds the newline index the parser needs.
bool FastqSource::next_from_pieces(TextBlock& blk, std::string& err) {
    blk.hold.reset();
    blk.view = nullptr;
    blk.begin = 0;
    blk.has_nl = false;
    blk.last_of_file = false;
    for (;;) {
        uint8_t* data = nullptr;
        size_t n = 0;
        std::shared_ptr<void> hold;
        if (!pinf_->next_owned(&data, &n, &hold)) {
            if (!pinf_->error().empty()) { err = "inflate failed on " + path_ + ": " + pinf_->error(); return false; }
            eof_ = true;
            if (carry_.empty()) return false;
            blk.data.assign(carry_.begin(), carry_.end());    // what is left: at most a partial record
            blk.len = carry_.size();
            blk.last_of_file = true;
            bytes_out_ += blk.len;
            carry_.clear();
            return true;
        }
        char* base;
        size_t len;
        if (carry_.size() <= ParallelGzip::kFront) {
            base = reinterpret_cast<char*>(data) - carry_.size();
            if (!carry_.empty()) memcpy(base, carry_.data(), carry_.size());
            len = carry_.size() + n;
        } else {                                              // records larger than the front space: the copying way
            blk.data.resize(carry_.size() + n);
            memcpy(blk.data.data(), carry_.data(), carry_.size());
            memcpy(blk.data.data() + carry_.size(), data, n);
            hold.reset();
            base = blk.data.data();
            len = carry_.size() + n;
        }
        if (len > 0xFFFFFFFFull) { err = "FASTQ record larger than 4 GiB in " + path_; return false; }
        const size_t c = newline_index(base, len, blk.nl, 0);
        const size_t whole = c & ~(size_t)3;                  // newlines that close whole records
        if (!whole) {                                         // not one whole record yet: keep everything, read on
            carry_.assign(base, base + len);
            continue;
        }
        const size_t cut = (size_t)blk.nl[whole - 1] + 1;
        carry_.assign(base + cut, base + len);
        if (hold) { blk.view = base; blk.hold = std::move(hold); }
        blk.len = cut;
        blk.nl_begin = 0;
        blk.nl_count = whole;
        blk.has_nl = true;
        bytes_out_ += cut;
        return true;
    }
}

bool FastqSource::next(TextBlock& blk, size_t target, std::string& err) {
    if (pinf_ && !getenv("HAST_GZ_COPY")) return next_from_pieces(blk, err);
    blk.hold.reset();
    blk.view = nullptr;
    blk.has_nl = false;
    if (eof_ && carry_.empty()) return false;
    size_t cap = std::max(blk.data.size(), target + carry_.size() + 4096);
    blk.data.resize(cap);
    size_t len = carry_.size();
    if (len) memcpy(blk.data.data(), carry_.data(), len);
    uint64_t lines = carry_lines_;
    carry_.clear();
    carry_lines_ = 0;
    blk.last_of_file = false;
    blk.begin = 0;
    for (;;) {
        while (!eof_ && len < target) {
            const size_t n = raw_read(blk.data.data() + len, cap - len, err);
            if (!err.empty()) return false;
            if (n == 0) { eof_ = true; break; }
            lines += count_newlines(blk.data.data() + len, n);
            len += n;
        }
        if (eof_) {                                   // whatever is left, partial record included
            blk.len = len;
            blk.last_of_file = true;
            bytes_out_ += len;
            return len > 0;
        }
        if (lines >= 4) break;
        target *= 2;                                  // a record longer than the block: keep reading
        cap = target + 4096;
        blk.data.resize(cap);
    }
    // cut after the last newline that closes a whole record
    const char* d = blk.data.data();
    const uint32_t leftover = (uint32_t)(lines & 3u);
    const char* q = (const char*)memrchr(d, '\n', len);
    for (uint32_t i = 0; i < leftover; ++i) q = (const char*)memrchr(d, '\n', (size_t)(q - d));
    const size_t cut = (size_t)(q - d) + 1;
    carry_.assign(d + cut, d + len);
    carry_lines_ = leftover;
    blk.len = cut;
    bytes_out_ += cut;
    return true;
}

}  // namespace hasthost
