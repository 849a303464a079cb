// inflate_par.cpp -- see inflate_par.h.
#include "inflate_par.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>          // crc32(), crc32_combine() only
#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "crc32_clmul.h"
#include "inflate.h"

namespace hasthost {

namespace {

// table entry layout of inflate.cpp (GzipInflater::build_table, pack_literals = false)
constexpr uint32_t kKindMask = 3u << 6;
constexpr uint32_t kLit = 0u << 6, kBase = 1u << 6, kEob = 2u << 6, kSub = 3u << 6;
constexpr int kLitlenRoot = GzipInflater::kLitlenRoot, kDistRoot = GzipInflater::kDistRoot;
constexpr size_t kWindow = GzipInflater::kWindow;
constexpr uint16_t kPlaceholder = 0x8000;
constexpr size_t kMaxChunkOut = (size_t)1 << 29;            // symbols; a runaway decoder stops here

inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

struct BitReader {
    const uint8_t* base;
    size_t n;
    size_t pos = 0;            // next byte to load
    uint64_t bb = 0;
    unsigned bc = 0;

    BitReader(const uint8_t* b, size_t len) : base(b), n(len) {}
    void seek(uint64_t bit) {
        pos = (size_t)(bit >> 3);
        bb = 0;
        bc = 0;
        refill();
        drop((unsigned)(bit & 7));
    }
    inline void refill() {
        if (pos + 8 <= n) {
            bb |= load64(base + pos) << bc;
            pos += (63 - bc) >> 3;
            bc |= 56;
        } else {
            while (bc <= 55) {                             // zero bits past the end; overrun() reports it
                const uint64_t byte = pos < n ? base[pos] : 0;
                bb |= byte << bc;
                bc += 8;
                ++pos;
            }
        }
    }
    inline void drop(unsigned k) { bb >>= k; bc -= k; }
    inline uint32_t take(unsigned k) {                     // k <= 32
        refill();
        const uint32_t v = (uint32_t)(bb & ((1ull << k) - 1));
        drop(k);
        return v;
    }
    inline void align() { drop(bc & 7u); }
    uint64_t bitpos() const { return (uint64_t)pos * 8 - bc; }
    bool overrun() const { return bitpos() > (uint64_t)n * 8; }
};

struct Tables {
    uint32_t litlen[GzipInflater::kLitlenCap];
    uint32_t dist[GzipInflater::kDistCap];
};

// Dynamic block header after BFINAL/BTYPE (RFC 1951 3.2.7), every rule enforced.
bool parse_dynamic_header(BitReader& r, Tables& t) {
    const int hlit = (int)r.take(5) + 257, hdist = (int)r.take(5) + 1, hclen = (int)r.take(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t cl[19] = {0};
    for (int i = 0; i < hclen; ++i) cl[order[i]] = (uint8_t)r.take(3);
    uint16_t pre[128];
    {
        uint16_t count[8] = {0};
        for (int i = 0; i < 19; ++i) count[cl[i]]++;
        count[0] = 0;
        int left = 1;
        for (int l = 1; l <= 7; ++l) { left <<= 1; left -= count[l]; if (left < 0) return false; }
        if (left > 0) return false;
        uint16_t next_code[8];
        uint32_t code = 0;
        for (int l = 1; l <= 7; ++l) { code = (code + count[l - 1]) << 1; next_code[l] = (uint16_t)code; }
        for (int i = 0; i < 19; ++i) {
            const int l = cl[i];
            if (!l) continue;
            uint32_t c = next_code[l]++, rev = 0;
            for (int b = 0; b < l; ++b) { rev = (rev << 1) | (c & 1u); c >>= 1; }
            for (uint32_t j = rev; j < 128; j += 1u << l) pre[j] = (uint16_t)(i | (l << 8));
        }
    }
    uint8_t lens[320];
    const int total = hlit + hdist;
    int i = 0;
    while (i < total) {
        r.refill();
        const uint16_t e = pre[r.bb & 127];
        const int sym = e & 0xFF;
        r.drop((unsigned)(e >> 8));
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        int rep, val = 0;
        if (sym == 16) {
            if (i == 0) return false;
            val = lens[i - 1];
            rep = 3 + (int)(r.bb & 3); r.drop(2);
        } else if (sym == 17) {
            rep = 3 + (int)(r.bb & 7); r.drop(3);
        } else {
            rep = 11 + (int)(r.bb & 127); r.drop(7);
        }
        if (i + rep > total) return false;
        while (rep--) lens[i++] = (uint8_t)val;
        if (r.overrun()) return false;
    }
    if (r.overrun() || lens[256] == 0) return false;
    return GzipInflater::build_table(lens, hlit, kLitlenRoot, t.litlen, GzipInflater::kLitlenCap, true, false) &&
           GzipInflater::build_table(lens + hlit, hdist, kDistRoot, t.dist, GzipInflater::kDistCap, false, false);
}

void build_fixed(Tables& t) {
    uint8_t lens[288];
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    GzipInflater::build_table(lens, 288, kLitlenRoot, t.litlen, GzipInflater::kLitlenCap, true, false);
    uint8_t d[32];
    for (int i = 0; i < 32; ++i) d[i] = 5;
    GzipInflater::build_table(d, 32, kDistRoot, t.dist, GzipInflater::kDistCap, false, false);
}

inline bool is_text(uint32_t c) { return (c >= 32 && c <= 126) || c == '\n' || c == '\r' || c == '\t'; }

#define HASTP_LOOKUP(e, table, root, r)                                                   \
    do {                                                                                  \
        e = table[(r).bb & ((1u << root) - 1u)];                                          \
        if ((e & kKindMask) == kSub) {                                                    \
            (r).drop(root);                                                               \
            e = table[(e >> 16) + ((r).bb & ((1u << ((e >> 8) & 31u)) - 1u))];            \
        }                                                                                 \
    } while (0)

// Does a block that starts at `bit` look real?  Full header validation, then up to max_syms symbols that must all
// be valid codes and printable text.
bool plausible_block(const uint8_t* in, size_t n, uint64_t bit, Tables& t, int max_syms) {
    BitReader r(in, n);
    r.seek(bit + 3);
    if (!parse_dynamic_header(r, t)) return false;
    for (int s = 0; s < max_syms; ++s) {
        r.refill();
        uint32_t e;
        HASTP_LOOKUP(e, t.litlen, kLitlenRoot, r);
        if ((e & kKindMask) == kLit) {
            if (!is_text((e >> 8) & 0xFFu)) return false;
            r.drop(e & 15u);
            continue;
        }
        if ((e & kKindMask) == kEob) {
            if ((e >> 16) != 0) return false;
            r.drop(e & 63u);
            if (r.overrun()) return false;
            const uint32_t h = r.take(3);                  // what follows must at least be a legal block type
            return (h >> 1) != 3;
        }
        r.drop(e & 63u);
        uint32_t d;
        HASTP_LOOKUP(d, t.dist, kDistRoot, r);
        if ((d & kKindMask) != kBase) return false;
        r.drop(d & 63u);
        if (r.overrun()) return false;
    }
    return true;
}

// First plausible dynamic-Huffman block header in [from_bit, to_bit), or -1.
int64_t find_block_start(const uint8_t* in, size_t n, uint64_t from_bit, uint64_t to_bit) {
    Tables t;
    const uint64_t last = n >= 16 ? (uint64_t)(n - 16) * 8 : 0;     // a block needs more than this anyway
    to_bit = std::min(to_bit, last);
    for (uint64_t bit = from_bit; bit < to_bit; ++bit) {
        const uint64_t v = load64(in + (bit >> 3)) >> (bit & 7);
        if ((v & 7u) != 4u) continue;                      // BFINAL = 0, BTYPE = 10
        if (((v >> 3) & 31u) > 29u || ((v >> 8) & 31u) > 29u) continue;
        if (plausible_block(in, n, bit, t, 4096)) return (int64_t)bit;
    }
    return -1;
}

struct Segment {
    size_t begin, end;         // output range inside the chunk
    bool member_end;           // ends with a gzip trailer
    uint32_t crc, isize;       // the trailer's values
    uint32_t crc_got = 0;      // CRC-32 of the resolved bytes (phase 3b)
};

}  // namespace

struct ParallelGzip::Chunk {
    uint64_t start_bit = 0;
    bool at_file_start = false;        // begins with the first gzip header of the file
    bool found = false;
    std::unique_ptr<uint16_t[]> sym;   // [kWindow placeholders][output symbols]; kept across batches
    size_t sym_cap = 0;
    size_t n_out = 0;
    uint64_t end_bit = 0;
    std::vector<Segment> segs;
    uint32_t max_reach = 0;            // deepest reference into the unknown window (first segment only)
    bool eof = false;
    int targets_passed = 0;
    std::string err;
    std::vector<uint8_t> window;       // the 32 KiB that precede this chunk (phase 3a)
    std::vector<uint8_t> bytes;        // resolved output (phase 3b)
};

namespace {

bool parse_gzip_header(BitReader& r, std::string& err) {
    const uint32_t id1 = r.take(8), id2 = r.take(8), cm = r.take(8), flg = r.take(8);
    if (id1 != 0x1F || id2 != 0x8B) { err = "not in gzip format"; return false; }
    if (cm != 8) { err = "unknown compression method"; return false; }
    if (flg & 0xE0) { err = "unknown header flags set"; return false; }
    r.take(32);
    r.take(16);
    if (flg & 4) {
        uint32_t xlen = r.take(16);
        while (xlen--) { r.take(8); if (r.overrun()) break; }
    }
    for (int bit : {8, 16})
        if (flg & bit)
            for (;;) {
                const uint32_t c = r.take(8);
                if (!c || r.overrun()) break;
            }
    if (flg & 2) r.take(16);
    if (r.overrun()) { err = "unexpected end of file"; return false; }
    return true;
}

// Decode from c.start_bit until (a) a block boundary that is exactly the next still-reachable target, (b) once all
// targets were run over, the first block boundary at or after soft_stop (0 = none), or (c) the end of the last member.
void decode_chunk(const uint8_t* in, size_t n, ParallelGzip::Chunk& c, const std::vector<uint64_t>& targets,
                  uint64_t soft_stop) {
    BitReader r(in, n);
    r.seek(c.start_bit);
    std::unique_ptr<Tables> dyn(new Tables), fixed;
    if (c.sym_cap < kWindow + ((size_t)4 << 20)) {
        c.sym_cap = kWindow + ((size_t)4 << 20);
        c.sym.reset(new uint16_t[c.sym_cap]);
    }
    size_t cap = c.sym_cap;
    for (size_t i = 0; i < kWindow; ++i) c.sym[i] = (uint16_t)(kPlaceholder | i);
    uint16_t* buf = c.sym.get();
    uint16_t* out = buf + kWindow;
    size_t ti = 0;
    bool in_member = !c.at_file_start;
    bool window_reachable = !c.at_file_start;      // false once a member began inside this chunk
    size_t member_base = 0;                         // output index where the current member began (if !window_reachable)
    size_t seg_begin = 0;
    bool first_header = c.at_file_start;
    auto fail = [&](const char* m) { c.err = m; };
    auto grow = [&](size_t need) {
        const size_t used = (size_t)(out - buf);
        if (used + need <= cap) return true;
        if (used - kWindow > kMaxChunkOut) { fail("chunk output too large for the parallel decoder (set HAST_INFLATE_THREADS=1)"); return false; }
        cap = std::max(cap * 2, used + need);
        std::unique_ptr<uint16_t[]> bigger(new uint16_t[cap]);
        memcpy(bigger.get(), buf, used * sizeof(uint16_t));
        c.sym = std::move(bigger);
        c.sym_cap = cap;
        buf = c.sym.get();
        out = buf + used;
        return true;
    };

    for (;;) {
        if (!in_member) {
            r.align();
            const uint64_t byte = r.bitpos() >> 3;
            if (byte >= n) { c.eof = true; break; }
            if (!first_header && (n - byte < 2 || in[byte] != 0x1F || in[byte + 1] != 0x8B)) { c.eof = true; break; }   // trailing garbage: ignored
            first_header = false;
            if (!parse_gzip_header(r, c.err)) break;
            in_member = true;
            window_reachable = false;
            member_base = (size_t)(out - buf) - kWindow;
        }
        const uint64_t bp = r.bitpos();
        while (ti < targets.size() && bp > targets[ti]) { ++ti; ++c.targets_passed; }
        if (ti < targets.size()) {
            if (bp == targets[ti]) break;
        } else if (soft_stop && bp >= soft_stop) {
            break;
        }
        const uint32_t h = r.take(3);
        if (r.overrun()) { fail("unexpected end of file"); break; }
        const bool last = (h & 1u) != 0;
        const uint32_t type = h >> 1;
        if (type == 3) { fail("invalid block type"); break; }
        if (type == 0) {
            r.align();
            const uint32_t len = r.take(16), nlen = r.take(16);
            if (r.overrun()) { fail("unexpected end of file"); break; }
            if ((len ^ 0xFFFFu) != nlen) { fail("invalid stored block lengths"); break; }
            size_t byte = (size_t)(r.bitpos() >> 3);
            if (byte + len > n) { fail("unexpected end of file"); break; }
            if (!grow(len + 16)) break;
            for (uint32_t i = 0; i < len; ++i) out[i] = in[byte + i];
            out += len;
            r.seek((uint64_t)(byte + len) * 8);
        } else {
            const Tables* t;
            if (type == 1) {
                if (!fixed) { fixed.reset(new Tables); build_fixed(*fixed); }
                t = fixed.get();
            } else {
                if (!parse_dynamic_header(r, *dyn)) { fail("invalid dynamic block header"); break; }
                t = dyn.get();
            }
            const uint32_t* const lt = t->litlen;
            const uint32_t* const dt = t->dist;
            uint16_t* cap_end = buf + cap - 272;
            bool ok = true;
            for (;;) {
                if (out > cap_end) {
                    if (!grow(1u << 20)) { ok = false; break; }
                    cap_end = buf + cap - 272;
                }
                if (r.pos >= n && r.overrun()) { fail("unexpected end of file"); ok = false; break; }
                r.refill();
                uint32_t e;
                HASTP_LOOKUP(e, lt, kLitlenRoot, r);
                if ((e & kKindMask) == kLit) {             // up to three literals out of one refill (56 bits)
                    *out++ = (uint16_t)((e >> 8) & 0xFFu);
                    r.drop(e & 15u);
                    HASTP_LOOKUP(e, lt, kLitlenRoot, r);
                    if ((e & kKindMask) == kLit) {
                        *out++ = (uint16_t)((e >> 8) & 0xFFu);
                        r.drop(e & 15u);
                        HASTP_LOOKUP(e, lt, kLitlenRoot, r);
                        if ((e & kKindMask) == kLit) {
                            *out++ = (uint16_t)((e >> 8) & 0xFFu);
                            r.drop(e & 15u);
                            continue;
                        }
                    }
                }
                if ((e & kKindMask) == kEob) {
                    if ((e >> 16) != 0) { fail("invalid literal/length code"); ok = false; }
                    r.drop(e & 63u);
                    break;
                }
                const uint64_t saved = r.bb;
                const uint32_t cl = (e >> 8) & 31u, drop = e & 63u;
                r.drop(drop);
                const uint32_t length = (e >> 16) + (uint32_t)((saved >> cl) & ((1u << (drop - cl)) - 1u));
                r.refill();                                // literals may have used the budget of the distance
                uint32_t d;
                HASTP_LOOKUP(d, dt, kDistRoot, r);
                if ((d & kKindMask) != kBase) { fail("invalid distance code"); ok = false; break; }
                const uint64_t saved2 = r.bb;
                const uint32_t cl2 = (d >> 8) & 31u, drop2 = d & 63u;
                r.drop(drop2);
                const uint32_t dist = (d >> 16) + (uint32_t)((saved2 >> cl2) & ((1u << (drop2 - cl2)) - 1u));
                const size_t idx = (size_t)(out - buf) - kWindow;
                if (window_reachable) {
                    if (dist > idx && dist - idx > c.max_reach) c.max_reach = (uint32_t)(dist - idx);
                } else if (dist > idx - member_base) {
                    fail("invalid distance too far back");
                    ok = false;
                    break;
                }
                const uint16_t* src = out - dist;
                uint16_t* const end = out + length;
                if (dist >= 8) {                           // most matches are short: 16 symbols unconditionally
                    memcpy(out, src, 16);
                    memcpy(out + 8, src + 8, 16);
                    if (length > 16) {
                        out += 16; src += 16;
                        do { memcpy(out, src, 16); out += 8; src += 8; } while (out < end);
                    }
                } else if (dist >= 4) {
                    do { memcpy(out, src, 8); out += 4; src += 4; } while (out < end);
                } else {
                    do { *out++ = *src++; } while (out < end);
                }
                out = end;
            }
            if (!ok) break;
        }
        if (last) {
            r.align();
            const uint32_t crc = r.take(32), isize = r.take(32);
            if (r.overrun()) { fail("unexpected end of file"); break; }
            const size_t idx = (size_t)(out - buf) - kWindow;
            c.segs.push_back(Segment{seg_begin, idx, true, crc, isize});
            seg_begin = idx;
            in_member = false;
        }
    }
    c.end_bit = r.bitpos();
    c.n_out = (size_t)(out - buf) - kWindow;
    if (c.n_out > seg_begin) c.segs.push_back(Segment{seg_begin, c.n_out, false, 0, 0});
}

// 16-bit symbols -> bytes: literals are kept, placeholders read the window.  Sixteen at a time when none of them is
// a placeholder; otherwise through a 64 Ki-entry table (identity below 0x8000, the window above), which has no branch
// to mispredict -- in FASTQ the constant part of every read name is a placeholder all the way through a chunk.
void resolve_symbols(const uint16_t* sym, size_t n, const uint8_t* lut, uint8_t* dst) {
    size_t k = 0;
    for (; k + 16 <= n; k += 16) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(sym + k));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(sym + k + 8));
        if (_mm_movemask_epi8(_mm_or_si128(a, b)) & 0xAAAA) {            // some value has bit 15 set
            for (size_t j = k; j < k + 16; ++j) dst[j] = lut[sym[j]];
        } else {
            _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + k), _mm_packus_epi16(a, b));
        }
    }
    for (; k < n; ++k) dst[k] = lut[sym[k]];
}

template <class F>
void parallel_for(int threads, size_t n, F&& f) {
    if (n == 0) return;
    std::atomic<size_t> next{0};
    auto body = [&] { for (size_t i; (i = next.fetch_add(1)) < n;) f(i); };
    const int t = (int)std::min<size_t>((size_t)std::max(threads, 1), n);
    std::vector<std::thread> pool;
    for (int k = 1; k < t; ++k) pool.emplace_back(body);
    body();
    for (auto& th : pool) th.join();
}

}  // namespace

ParallelGzip::ParallelGzip(int threads, size_t chunk_bytes)
    : threads_(std::max(threads, 1)), chunk_bytes_(std::max<size_t>(chunk_bytes, 4096)) {}

ParallelGzip::~ParallelGzip() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_room_.notify_all();
    if (coord_.joinable()) coord_.join();
    if (map_) munmap(const_cast<uint8_t*>(map_), map_len_);
    if (fd_ >= 0) close(fd_);
}

std::string ParallelGzip::open(const std::string& path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) return "cannot open " + path + ": " + strerror(errno);
    struct stat sb;
    if (fstat(fd_, &sb) != 0 || !S_ISREG(sb.st_mode)) return "not a regular file: " + path;
    if (sb.st_size > 0) {
        void* p = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (p == MAP_FAILED) return "cannot map " + path + ": " + strerror(errno);
        map_ = (const uint8_t*)p;
        map_len_ = (size_t)sb.st_size;
    }
    in_ = map_;
    n_in_ = map_len_;
    return "";
}

void ParallelGzip::open_memory(const uint8_t* data, size_t n) {
    in_ = data;
    n_in_ = n;
}

void ParallelGzip::start() { coord_ = std::thread([this] { run(); }); }

void ParallelGzip::push(std::vector<uint8_t>&& piece) {
    if (piece.empty()) return;
    std::unique_lock<std::mutex> lk(mu_);
    cv_room_.wait(lk, [&] { return stop_ || ready_bytes_ < ((size_t)256 << 20); });
    if (stop_) return;
    ready_bytes_ += piece.size();
    ready_.push_back(std::move(piece));
    cv_out_.notify_one();
}

void ParallelGzip::finish(const std::string& err) {
    std::lock_guard<std::mutex> lk(mu_);
    err_pending_ = err;
    done_ = true;
    cv_out_.notify_all();
}

bool ParallelGzip::next(const uint8_t** data, size_t* len) {
    *data = nullptr;
    *len = 0;
    if (!coord_.joinable() && !done_) {
        if (n_in_ == 0) return false;
        start();
    }
    std::unique_lock<std::mutex> lk(mu_);
    cv_out_.wait(lk, [&] { return !ready_.empty() || done_; });
    if (ready_.empty()) {
        err_ = err_pending_;
        return false;
    }
    if (current_.capacity() && spare_.size() < 64) spare_.push_back(std::move(current_));
    current_ = std::move(ready_.front());
    ready_.pop_front();
    ready_bytes_ -= current_.size();
    cv_room_.notify_one();
    *data = current_.data();
    *len = current_.size();
    return true;
}

void ParallelGzip::run() {
    const uint8_t* const in = in_;
    const size_t n = n_in_;
    const uint64_t n_bits = (uint64_t)n * 8;
    uint64_t next_start = 0;
    bool at_file_start = true;
    std::vector<uint8_t> window(kWindow, 0);
    uint32_t crc_run = (uint32_t)crc32(0L, Z_NULL, 0);
    uint64_t member_out = 0;
    const size_t W = (size_t)std::max(2 * threads_, 2);
    std::vector<Chunk> chunks(W);
    const bool prof = getenv("HAST_PAR_PROF") != nullptr;
    double t_ph[5] = {0, 0, 0, 0, 0};
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    struct ProfOut { bool on; double* t; ~ProfOut() { if (on) fprintf(stderr, "phases: find %.3f decode %.3f window %.3f resolve %.3f emit %.3f s\n", t[0], t[1], t[2], t[3], t[4]); } } prof_out{prof, t_ph};

    for (;;) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (stop_) return;
        }
        const size_t base = (size_t)(next_start >> 3);
        for (Chunk& c : chunks) {                           // buffers stay, results go
            c.found = false; c.at_file_start = false; c.n_out = 0; c.end_bit = 0; c.segs.clear(); c.max_reach = 0;
            c.eof = false; c.targets_passed = 0; c.err.clear();
        }
        chunks[0].start_bit = next_start;
        chunks[0].at_file_start = at_file_start;
        chunks[0].found = true;
        const uint64_t batch_end = std::min<uint64_t>((uint64_t)(base + W * chunk_bytes_) * 8, n_bits);
        const uint64_t soft_stop = batch_end < n_bits ? batch_end : 0;
        double tp = now();
        // phase 1: block starts
        parallel_for(threads_, W - 1, [&](size_t j) {
            const size_t k = j + 1;
            const uint64_t from = (uint64_t)(base + k * chunk_bytes_) * 8, to = std::min<uint64_t>(from + (uint64_t)chunk_bytes_ * 8, n_bits);
            if (from >= n_bits) return;
            const int64_t s = find_block_start(in, n, from, to);
            if (s >= 0) { chunks[k].start_bit = (uint64_t)s; chunks[k].found = true; }
        });
        t_ph[0] += now() - tp; tp = now();
        std::vector<size_t> order;                         // chunks with a start, in stream order
        for (size_t k = 0; k < W; ++k) if (chunks[k].found) order.push_back(k);
        stats_.chunks += W;
        stats_.starts_found += order.size() - 1;
        ++stats_.batches;
        // phase 2: decode every chunk to the next start
        parallel_for(threads_, order.size(), [&](size_t oi) {
            std::vector<uint64_t> targets;
            for (size_t o = oi + 1; o < order.size(); ++o) targets.push_back(chunks[order[o]].start_bit);
            decode_chunk(in, n, chunks[order[oi]], targets, soft_stop);
        });
        t_ph[1] += now() - tp; tp = now();
        // phase 3: the chain of chunks whose starts were confirmed by their predecessor
        std::vector<size_t> chain;
        bool eof = false;
        for (size_t oi = 0; oi < order.size();) {
            Chunk& c = chunks[order[oi]];
            chain.push_back(order[oi]);
            if (!c.err.empty()) { finish(c.err); return; }
            stats_.starts_dropped += (uint64_t)c.targets_passed;
            const size_t nxt = oi + 1 + (size_t)c.targets_passed;
            if (c.eof) { eof = true; break; }
            if (nxt < order.size() && c.end_bit == chunks[order[nxt]].start_bit) { oi = nxt; continue; }
            break;                                         // soft stop: the next batch starts exactly here
        }
        Chunk& last = chunks[chain.back()];
        // 3a: windows, in stream order
        for (size_t ci : chain) {
            Chunk& c = chunks[ci];
            c.window = window;
            const uint16_t* sym = c.sym.get() + kWindow;
            if (c.n_out >= kWindow) {
                for (size_t i = 0; i < kWindow; ++i) {
                    const uint16_t v = sym[c.n_out - kWindow + i];
                    window[i] = v < kPlaceholder ? (uint8_t)v : c.window[v - kPlaceholder];
                }
            } else {
                std::vector<uint8_t> w(kWindow);
                memcpy(w.data(), c.window.data() + c.n_out, kWindow - c.n_out);
                for (size_t i = 0; i < c.n_out; ++i) {
                    const uint16_t v = sym[i];
                    w[kWindow - c.n_out + i] = v < kPlaceholder ? (uint8_t)v : c.window[v - kPlaceholder];
                }
                window.swap(w);
            }
        }
        t_ph[2] += now() - tp; tp = now();
        // 3b: bytes and per-segment CRC-32, in parallel
        parallel_for(threads_, chain.size(), [&](size_t i) {
            Chunk& c = chunks[chain[i]];
            {
                std::lock_guard<std::mutex> lk(mu_);       // recycled output buffers keep their pages
                if (!spare_.empty()) { c.bytes = std::move(spare_.back()); spare_.pop_back(); }
            }
            if (c.bytes.capacity() < c.n_out) { std::vector<uint8_t>().swap(c.bytes); c.bytes.reserve(c.n_out + c.n_out / 8); }
            c.bytes.resize(c.n_out);
            const uint16_t* sym = c.sym.get() + kWindow;
            std::vector<uint8_t> lut(65536);
            for (size_t v = 0; v < 256; ++v) lut[v] = (uint8_t)v;
            memcpy(lut.data() + kPlaceholder, c.window.data(), kWindow);
            const uint8_t* w = lut.data();
            uint8_t* dst = c.bytes.data();
            // resolve and checksum block by block, while the bytes are still in cache
            for (Segment& sg : c.segs) {
                uint32_t crc = (uint32_t)crc32(0L, Z_NULL, 0);
                for (size_t p = sg.begin; p < sg.end;) {
                    const size_t m = std::min<size_t>(sg.end - p, 65536);
                    resolve_symbols(sym + p, m, w, dst + p);
                    crc = hast_crc32(crc, dst + p, m);
                    p += m;
                }
                sg.crc_got = crc;
            }
        });
        t_ph[3] += now() - tp; tp = now();
        // 3c: member checks and hand-over, in stream order
        for (size_t ci : chain) {
            Chunk& c = chunks[ci];
            if (!c.at_file_start && c.max_reach > std::min<uint64_t>(member_out, kWindow)) { finish("invalid distance too far back"); return; }
            for (const Segment& s : c.segs) {
                const uint64_t len = s.end - s.begin;
                crc_run = (uint32_t)crc32_combine(crc_run, s.crc_got, (z_off_t)len);
                member_out += len;
                if (s.member_end) {
                    if (crc_run != s.crc) { finish("incorrect data check"); return; }
                    if ((uint32_t)(member_out & 0xFFFFFFFFu) != s.isize) { finish("incorrect length check"); return; }
                    crc_run = (uint32_t)crc32(0L, Z_NULL, 0);
                    member_out = 0;
                }
            }
            push(std::move(c.bytes));
        }
        t_ph[4] += now() - tp;
        if (eof) { finish(""); return; }
        if (last.end_bit >= n_bits) { finish("unexpected end of file"); return; }   // ran out of input inside a member
        next_start = last.end_bit;
        at_file_start = false;
    }
}

}  // namespace hasthost
