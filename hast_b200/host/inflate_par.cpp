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
constexpr size_t kMaxChunkOut = (size_t)1 << 31;            // bytes; a runaway decoder stops here
constexpr size_t kSlack = 258 + 64;                         // wild copies may run past the end of a chunk's output
// token: literal run, then (unless kNoMatch) one match
constexpr uint32_t kNoMatch = 1u << 23;
inline uint32_t make_token(uint32_t lits, uint32_t length, uint32_t dist) { return (lits << 24) | ((length - 3u) << 15) | (dist - 1u); }

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
bool parse_dynamic_header(BitReader& r, Tables& t, bool pack_literals = false) {
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
    return GzipInflater::build_table(lens, hlit, kLitlenRoot, t.litlen, GzipInflater::kLitlenCap, true, pack_literals) &&
           GzipInflater::build_table(lens + hlit, hdist, kDistRoot, t.dist, GzipInflater::kDistCap, false, false);
}

void build_fixed(Tables& t) {
    uint8_t lens[288];
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    GzipInflater::build_table(lens, 288, kLitlenRoot, t.litlen, GzipInflater::kLitlenCap, true, true);
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
    size_t tok_begin, tok_end; // tokens of the segment
    size_t lit_begin;          // its first literal byte
    size_t out_len;            // bytes it decodes to
    bool starts_member;        // a gzip header precedes it: the window starts empty
    bool member_end;           // ends with a gzip trailer
    uint32_t crc, isize;       // the trailer's values
    uint32_t crc_got = 0;      // CRC-32 of the replayed bytes
    size_t out_begin = 0;      // offset of its bytes in the chunk's output (set by the replay)
};

}  // namespace

struct ParallelGzip::Chunk {
    uint64_t start_bit = 0;
    bool at_file_start = false;        // begins with the first gzip header of the file
    bool found = false;
    std::vector<uint32_t> tok;         // buffers are kept across batches
    std::vector<uint8_t> lit;
    size_t n_tok = 0, n_lit = 0, n_out = 0;
    uint64_t end_bit = 0;
    std::vector<Segment> segs;
    bool eof = false;
    int targets_passed = 0;
    std::string err;
    ParallelGzip::Buffer bytes;        // [kWindow of history][n_out][kSlack]  (replay)
};

// One batch of chunks in flight: block starts are searched, then every chunk is decoded to tokens; the worker
// that finishes the last search launches the decodes, the one that finishes the last decode wakes the coordinator.
struct ParallelGzip::Batch {
    std::vector<Chunk> chunks;
    std::vector<size_t> order;         // chunks with a start, in stream order
    uint64_t soft_stop = 0;
    std::atomic<int> finds_left{0}, decodes_left{0};
    bool ready = false;                // under pool mutex
};

// Worker pool of one stream: plain FIFO with a front door for the short CRC jobs.
struct ParallelGzip::Pool {
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::deque<std::function<void()>> q;
    std::vector<std::thread> th;
    bool stop = false;
    explicit Pool(int n) {
        for (int i = 0; i < n; ++i)
            th.emplace_back([this] {
                for (;;) {
                    std::function<void()> f;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return stop || !q.empty(); });
                        if (q.empty()) return;
                        f = std::move(q.front());
                        q.pop_front();
                    }
                    f();
                }
            });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
            q.clear();
        }
        cv.notify_all();
        for (auto& t : th) t.join();
    }
    void submit(std::function<void()> f, bool front = false) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (front) q.push_front(std::move(f)); else q.push_back(std::move(f));
        }
        cv.notify_one();
    }
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

// Entropy-decode from c.start_bit into tokens + literal bytes -- no output bytes are produced, so no window is
// needed -- until (a) a block boundary that is exactly the next still-reachable target, (b) once all targets were
// run over, the first block boundary at or after soft_stop (0 = none), or (c) the end of the last member.
void decode_chunk(const uint8_t* in, size_t n, ParallelGzip::Chunk& c, const std::vector<uint64_t>& targets,
                  uint64_t soft_stop, size_t chunk_bytes) {
    BitReader r(in, n);
    r.seek(c.start_bit);
    std::unique_ptr<Tables> dyn(new Tables), fixed;
    if (c.tok.size() < chunk_bytes * 2) c.tok.resize(chunk_bytes * 2);
    if (c.lit.size() < chunk_bytes * 6) c.lit.resize(chunk_bytes * 6);
    uint32_t* tok = c.tok.data();
    uint8_t* lit = c.lit.data();
    size_t tok_cap = c.tok.size(), lit_cap = c.lit.size();
    size_t nt = 0, nl = 0;
    uint64_t match_sum = 0;                         // bytes produced by matches so far
    uint32_t run = 0;                               // literals not yet attached to a token
    size_t ti = 0;
    bool in_member = !c.at_file_start;
    bool first_header = c.at_file_start;
    bool seg_starts_member = false;
    size_t seg_tok = 0, seg_lit = 0;
    uint64_t seg_match = 0;
    auto fail = [&](const char* m) { c.err = m; };
    auto grow = [&](size_t need_tok, size_t need_lit) {
        if (nt + need_tok > tok_cap) { c.tok.resize(std::max(tok_cap * 2, nt + need_tok)); tok = c.tok.data(); tok_cap = c.tok.size(); }
        if (nl + need_lit > lit_cap) { c.lit.resize(std::max(lit_cap * 2, nl + need_lit)); lit = c.lit.data(); lit_cap = c.lit.size(); }
    };
    auto flush_run = [&] {
        while (run) { const uint32_t k = std::min(run, 255u); tok[nt++] = (k << 24) | kNoMatch; run -= k; }
    };
    auto close_segment = [&](bool member_end, uint32_t crc, uint32_t isize) {
        grow(8 + run / 255, 0);
        flush_run();
        const size_t out_len = (nl - seg_lit) + (size_t)(match_sum - seg_match);
        if (nt > seg_tok || member_end || seg_starts_member)
            c.segs.push_back(Segment{seg_tok, nt, seg_lit, out_len, seg_starts_member, member_end, crc, isize});
        seg_tok = nt; seg_lit = nl; seg_match = match_sum;
        seg_starts_member = false;
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
            seg_starts_member = true;
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
            grow(len / 255 + 8 + run / 255, len + 16);
            memcpy(lit + nl, in + byte, len);
            nl += len;
            run += len;
            flush_run();
            r.seek((uint64_t)(byte + len) * 8);
        } else {
            const Tables* t;
            if (type == 1) {
                if (!fixed) { fixed.reset(new Tables); build_fixed(*fixed); }
                t = fixed.get();
            } else {
                if (!parse_dynamic_header(r, *dyn, true)) { fail("invalid dynamic block header"); break; }
                t = dyn.get();
            }
            const uint32_t* const lt = t->litlen;
            const uint32_t* const dt = t->dist;
            bool ok = true;
            // the hot loop keeps the bit reader and the output cursors in locals (registers); inputs shorter than
            // the 8-byte loads need are handled by the bounds-checked reader below
            uint64_t bb = r.bb;
            unsigned bc = r.bc;
            size_t pos = r.pos;
            const size_t fast_end = n >= 40 ? n - 40 : 0;      // an iteration refills at most three times (7 bytes each): every 8-byte load stays in bounds
            constexpr uint32_t kRootMask = (1u << kLitlenRoot) - 1u, kDistMask = (1u << kDistRoot) - 1u;
#define HT_REFILL()   do { bb |= load64(in + pos) << bc; pos += (63 - bc) >> 3; bc |= 56; } while (0)
#define HT_DROP(k)    do { bb >>= (k); bc -= (k); } while (0)
#define HT_SUB(e, table, root)                                                          \
    do {                                                                                \
        HT_DROP(root);                                                                  \
        e = table[(e >> 16) + (bb & ((1u << ((e >> 8) & 31u)) - 1u))];                  \
    } while (0)
#define HT_LITERALS(e)                                    \
    do {                                                  \
        const uint32_t v_ = e >> 8;                       \
        memcpy(lit + nl, &v_, 4);                         \
        const uint32_t k_ = 1u + ((e >> 4) & 3u);         \
        nl += k_;                                         \
        run += k_;                                        \
        HT_DROP(e & 15u);                                 \
    } while (0)
            uint32_t e;
            bool eob = false;
            if (pos < fast_end) {
                HT_REFILL();
                e = lt[bb & kRootMask];
                for (;;) {
                    if (nt + 8 > tok_cap || nl + 32 > lit_cap) grow(1u << 16, 1u << 18);
                    if (pos >= fast_end) break;                // the careful loop below finishes the block
                    if ((e & kKindMask) == kSub) HT_SUB(e, lt, kLitlenRoot);
                    if ((e & kKindMask) == kLit) {             // up to nine literals out of one refill (56 bits)
                        HT_LITERALS(e);
                        e = lt[bb & kRootMask];
                        if ((e & kKindMask) == kSub) HT_SUB(e, lt, kLitlenRoot);
                        if ((e & kKindMask) == kLit) {
                            HT_LITERALS(e);
                            e = lt[bb & kRootMask];
                            if ((e & kKindMask) == kSub) HT_SUB(e, lt, kLitlenRoot);
                            if ((e & kKindMask) == kLit) {
                                HT_LITERALS(e);
                                HT_REFILL();
                                e = lt[bb & kRootMask];
                                continue;
                            }
                        }
                        // e came out of the bits that are left: a length / distance pair needs up to 48 of them
                        if (bc < 48) {
                            // e's code bits are still at the bottom of bb, so a refill does not disturb it
                            HT_REFILL();
                        }
                    }
                    if ((e & kKindMask) == kEob) {
                        if ((e >> 16) != 0) { fail("invalid literal/length code"); ok = false; }
                        HT_DROP(e & 63u);
                        eob = true;
                        break;
                    }
                    const uint64_t saved = bb;
                    const uint32_t cl = (e >> 8) & 31u, drop = e & 63u;
                    HT_DROP(drop);
                    const uint32_t length = (e >> 16) + (uint32_t)((saved >> cl) & ((1u << (drop - cl)) - 1u));
                    uint32_t d = dt[bb & kDistMask];
                    if ((d & kKindMask) == kSub) HT_SUB(d, dt, kDistRoot);
                    if ((d & kKindMask) != kBase) { fail("invalid distance code"); ok = false; break; }
                    const uint64_t saved2 = bb;
                    const uint32_t cl2 = (d >> 8) & 31u, drop2 = d & 63u;
                    HT_DROP(drop2);
                    const uint32_t dist = (d >> 16) + (uint32_t)((saved2 >> cl2) & ((1u << (drop2 - cl2)) - 1u));
                    HT_REFILL();
                    e = lt[bb & kRootMask];                    // next symbol's entry: its load overlaps the bookkeeping
                    if (run > 255) {
                        if (nt + run / 255 + 8 > tok_cap) grow(run / 255 + (1u << 16), 0);
                        while (run > 255) { tok[nt++] = (255u << 24) | kNoMatch; run -= 255; }
                    }
                    tok[nt++] = make_token(run, length, dist);
                    run = 0;
                    match_sum += length;
                }
            }
            r.bb = bb; r.bc = bc; r.pos = pos;
#undef HT_REFILL
#undef HT_DROP
#undef HT_SUB
#undef HT_LITERALS
#define HASTP_LITERALS(e)                                 \
    do {                                                  \
        const uint32_t v_ = e >> 8;                       \
        memcpy(lit + nl, &v_, 4);                         \
        const uint32_t k_ = 1u + ((e >> 4) & 3u);         \
        nl += k_;                                         \
        run += k_;                                        \
        r.drop(e & 15u);                                  \
    } while (0)
            while (ok && !eob) {                               // last bytes of the input: bounds-checked reader
                if (nt + 8 > tok_cap || nl + 32 > lit_cap) grow(1u << 16, 1u << 18);
                if (r.pos >= n && r.overrun()) { fail("unexpected end of file"); ok = false; break; }
                r.refill();
                HASTP_LOOKUP(e, lt, kLitlenRoot, r);
                if ((e & kKindMask) == kLit) {
                    HASTP_LITERALS(e);
                    continue;
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
                r.refill();
                uint32_t d;
                HASTP_LOOKUP(d, dt, kDistRoot, r);
                if ((d & kKindMask) != kBase) { fail("invalid distance code"); ok = false; break; }
                const uint64_t saved2 = r.bb;
                const uint32_t cl2 = (d >> 8) & 31u, drop2 = d & 63u;
                r.drop(drop2);
                const uint32_t dist = (d >> 16) + (uint32_t)((saved2 >> cl2) & ((1u << (drop2 - cl2)) - 1u));
                if (run > 255) {
                    if (nt + run / 255 + 8 > tok_cap) grow(run / 255 + (1u << 16), 0);
                    while (run > 255) { tok[nt++] = (255u << 24) | kNoMatch; run -= 255; }
                }
                tok[nt++] = make_token(run, length, dist);
                run = 0;
                match_sum += length;
            }
#undef HASTP_LITERALS
            if (!ok) break;
            if (nl + match_sum > kMaxChunkOut) { fail("chunk output too large for the parallel decoder (set HAST_INFLATE_THREADS=1)"); break; }
        }
        if (last) {
            r.align();
            const uint32_t crc = r.take(32), isize = r.take(32);
            if (r.overrun()) { fail("unexpected end of file"); break; }
            close_segment(true, crc, isize);
            in_member = false;
        }
    }
    c.end_bit = r.bitpos();
    close_segment(false, 0, 0);
    c.n_tok = nt;
    c.n_lit = nl;
    c.n_out = nl + (size_t)match_sum;
}

// Tokens -> bytes at `out`; the `avail` bytes in front of `out` are the member's history (at most the window counts).
// Returns the end of the output or nullptr (a distance that reaches in front of the member).
uint8_t* replay_tokens(const uint32_t* tok, size_t n_tok, const uint8_t* lit, uint8_t* out, uint64_t avail) {
    const uint8_t* const begin = out;
    for (size_t i = 0; i < n_tok; ++i) {
        const uint32_t t = tok[i];
        const uint32_t ll = t >> 24;
        memcpy(out, lit, 16);                              // most literal runs are short
        if (ll > 16) {
            uint8_t* o = out + 16;
            const uint8_t* s = lit + 16;
            uint8_t* const e = out + ll;
            do { memcpy(o, s, 16); o += 16; s += 16; } while (o < e);
        }
        out += ll;
        lit += ll;
        if (t & kNoMatch) continue;
        const uint32_t length = ((t >> 15) & 255u) + 3u, dist = (t & 0x7FFFu) + 1u;
        if ((uint64_t)dist > avail + (uint64_t)(out - begin)) return nullptr;
        const uint8_t* src = out - dist;
        uint8_t* const end = out + length;
        if (dist >= 16) {
            memcpy(out, src, 16);
            if (length > 16) {
                out += 16; src += 16;
                do { memcpy(out, src, 16); out += 16; src += 16; } while (out < end);
            }
        } else if (dist >= 8) {
            do { memcpy(out, src, 8); out += 8; src += 8; } while (out < end);
        } else if (dist == 1) {
            memset(out, src[0], length);
        } else {
            do { *out++ = *src++; } while (out < end);
        }
        out = end;
    }
    return out;
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

// Output buffers are recycled: a fresh 5 MB buffer costs more in page faults and zeroing than replaying into it.
// Sizes are rounded up generously so that a recycled buffer nearly always fits the next chunk.
ParallelGzip::Buffer ParallelGzip::take_buffer(size_t need) {
    {
        std::lock_guard<std::mutex> lk(pool_->mu);
        std::vector<Buffer>& spare = pool_->spare;
        for (size_t i = spare.size(); i-- > 0;)
            if (spare[i].cap >= need) {
                Buffer b = std::move(spare[i]);
                spare.erase(spare.begin() + (long)i);
                return b;
            }
        if (spare.size() >= 8) spare.erase(spare.begin());     // none fits: let the oldest small one go
    }
    return Buffer(std::max<size_t>(need + need / 2, (size_t)8 << 20));
}

void ParallelGzip::recycle(Buffer&& b) {
    if (!b.cap) return;
    std::lock_guard<std::mutex> lk(pool_->mu);
    if (pool_->spare.size() < 64) pool_->spare.push_back(std::move(b));
}

ParallelGzip::Buffer::Buffer(size_t n) {
    const size_t huge = (size_t)2 << 20;
    cap_map = (n + huge - 1) / huge * huge + huge;
    void* m = mmap(nullptr, cap_map, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) { cap_map = 0; return; }
    map = m;
    p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(m) + huge - 1) & ~(uintptr_t)(huge - 1));
    cap = cap_map - (size_t)(p - static_cast<uint8_t*>(m));
#ifdef MADV_HUGEPAGE
    madvise(p, cap & ~(huge - 1), MADV_HUGEPAGE);
#endif
}
ParallelGzip::Buffer::~Buffer() { if (map) munmap(map, cap_map); }
ParallelGzip::Buffer::Buffer(Buffer&& o) noexcept : p(o.p), cap(o.cap), map(o.map), cap_map(o.cap_map) { o.p = nullptr; o.cap = 0; o.map = nullptr; o.cap_map = 0; }
ParallelGzip::Buffer& ParallelGzip::Buffer::operator=(Buffer&& o) noexcept {
    if (this != &o) {
        if (map) munmap(map, cap_map);
        p = o.p; cap = o.cap; map = o.map; cap_map = o.cap_map;
        o.p = nullptr; o.cap = 0; o.map = nullptr; o.cap_map = 0;
    }
    return *this;
}

void ParallelGzip::push(Piece&& piece) {
    if (!piece.len) { recycle(std::move(piece.buf)); return; }
    std::unique_lock<std::mutex> lk(mu_);
    cv_room_.wait(lk, [&] { return stop_ || ready_bytes_ < ((size_t)256 << 20); });
    if (stop_) return;
    ready_bytes_ += piece.len;
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
    Buffer old = std::move(current_.buf);
    current_ = std::move(ready_.front());
    ready_.pop_front();
    ready_bytes_ -= current_.len;
    cv_room_.notify_one();
    lk.unlock();
    recycle(std::move(old));
    *data = current_.buf.p + current_.off;
    *len = current_.len;
    return true;
}

bool ParallelGzip::next_owned(uint8_t** data, size_t* len, std::shared_ptr<void>* hold) {
    const uint8_t* p;
    if (!next(&p, len)) { *data = nullptr; return false; }
    static_assert(kFront <= kWindow, "the history area in front of a piece is what the caller may overwrite");
    *data = current_.buf.p + current_.off;
    std::shared_ptr<BufferPool> pool = pool_;
    *hold = std::shared_ptr<void>(new Buffer(std::move(current_.buf)), [pool](void* v) {
        Buffer* b = static_cast<Buffer*>(v);
        {
            std::lock_guard<std::mutex> lk(pool->mu);
            if (b->cap && pool->spare.size() < 64) pool->spare.push_back(std::move(*b));
        }
        delete b;
    });
    current_ = Piece{};
    return true;
}

// Launch one batch: search the block starts of chunks 1..W-1 (chunk 0 starts where the previous batch ended), then
// decode every chunk that has a start up to the next one's.
void ParallelGzip::launch(Batch& b, Pool& pool, uint64_t next_start, bool at_file_start) {
    const uint8_t* const in = in_;
    const size_t n = n_in_;
    const uint64_t n_bits = (uint64_t)n * 8;
    const size_t W = b.chunks.size();
    const size_t base = (size_t)(next_start >> 3);
    for (Chunk& c : b.chunks) {                                // buffers stay, results go
        c.found = false; c.at_file_start = false; c.n_out = c.n_tok = c.n_lit = 0; c.end_bit = 0; c.segs.clear();
        c.eof = false; c.targets_passed = 0; c.err.clear();
    }
    b.chunks[0].start_bit = next_start;
    b.chunks[0].at_file_start = at_file_start;
    b.chunks[0].found = true;
    const uint64_t batch_end = std::min<uint64_t>((uint64_t)(base + W * chunk_bytes_) * 8, n_bits);
    b.soft_stop = batch_end < n_bits ? batch_end : 0;
    b.order.clear();
    {
        std::lock_guard<std::mutex> lk(pool.mu);
        b.ready = false;
    }
    auto launch_decodes = [this, &b, &pool, in, n] {
        for (size_t k = 0; k < b.chunks.size(); ++k) if (b.chunks[k].found) b.order.push_back(k);
        b.decodes_left.store((int)b.order.size());
        for (size_t oi = 0; oi < b.order.size(); ++oi)
            pool.submit([this, &b, &pool, in, n, oi] {
                std::vector<uint64_t> targets;
                for (size_t o = oi + 1; o < b.order.size(); ++o) targets.push_back(b.chunks[b.order[o]].start_bit);
                const auto t0 = std::chrono::steady_clock::now();
                decode_chunk(in, n, b.chunks[b.order[oi]], targets, b.soft_stop, chunk_bytes_);
                decode_ns_.fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(),
                                     std::memory_order_relaxed);
                if (b.decodes_left.fetch_sub(1) == 1) {
                    { std::lock_guard<std::mutex> lk(pool.mu); b.ready = true; }
                    pool.cv_done.notify_all();
                }
            });
    };
    if (W == 1) { launch_decodes(); return; }
    b.finds_left.store((int)(W - 1));
    for (size_t k = 1; k < W; ++k)
        pool.submit([this, &b, in, n, n_bits, base, k, launch_decodes] {
            const uint64_t from = (uint64_t)(base + k * chunk_bytes_) * 8, to = std::min<uint64_t>(from + (uint64_t)chunk_bytes_ * 8, n_bits);
            if (from < n_bits) {
                const int64_t s = find_block_start(in, n, from, to);
                if (s >= 0) { b.chunks[k].start_bit = (uint64_t)s; b.chunks[k].found = true; }
            }
            if (b.finds_left.fetch_sub(1) == 1) launch_decodes();
        });
}

void ParallelGzip::run() {
    const size_t n = n_in_;
    const uint64_t n_bits = (uint64_t)n * 8;
    uint32_t crc_run = (uint32_t)crc32(0L, Z_NULL, 0);
    uint64_t member_out = 0;
    // this thread replays; the others search and decode.  Two batches: one being decoded while the other is replayed.
    const int n_workers = std::max(threads_ - 1, 1);
    const size_t W = (size_t)std::max(2 * n_workers, 2);
    Batch batches[2];
    for (Batch& b : batches) b.chunks = std::vector<Chunk>(W);
    std::vector<uint8_t> history(kWindow, 0);                  // last 32 KiB of output so far
    const bool prof = getenv("HAST_PAR_PROF") != nullptr;
    double t_ph[4] = {0, 0, 0, 0};
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    uint64_t n_tok_total = 0, n_lit_total = 0;
    struct ProfOut { bool on; double* t; uint64_t* a; uint64_t* b2; std::atomic<uint64_t>* dn; ~ProfOut() { if (on) fprintf(stderr, "coordinator: wait-for-decode %.3f replay %.3f crc-wait %.3f emit %.3f s; workers' entropy decode %.3f s; %llu tokens, %llu literal bytes\n", t[0], t[1], t[2], t[3], dn->load() * 1e-9, (unsigned long long)*a, (unsigned long long)*b2); } } prof_out{prof, t_ph, &n_tok_total, &n_lit_total, &decode_ns_};
    std::atomic<int> crc_left{0};
    Pool pool(n_workers);                                      // declared last: joins its threads before the batches go
    int cur = 0;
    launch(batches[cur], pool, 0, true);
    for (;;) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (stop_) return;
        }
        Batch& b = batches[cur];
        double tp = now();
        {
            std::unique_lock<std::mutex> lk(pool.mu);
            pool.cv_done.wait(lk, [&] { return b.ready; });
        }
        t_ph[0] += now() - tp; tp = now();
        stats_.chunks += W;
        stats_.starts_found += b.order.size() - 1;
        ++stats_.batches;
        // the chain of chunks whose starts were confirmed by their predecessor
        std::vector<size_t> chain;
        bool eof = false;
        for (size_t oi = 0; oi < b.order.size();) {
            Chunk& c = b.chunks[b.order[oi]];
            chain.push_back(b.order[oi]);
            if (!c.err.empty()) { finish(c.err); return; }
            stats_.starts_dropped += (uint64_t)c.targets_passed;
            const size_t nxt = oi + 1 + (size_t)c.targets_passed;
            if (c.eof) { eof = true; break; }
            if (nxt < b.order.size() && c.end_bit == b.chunks[b.order[nxt]].start_bit) { oi = nxt; continue; }
            break;                                             // soft stop: the next batch starts exactly here
        }
        const uint64_t next_start = b.chunks[chain.back()].end_bit;
        if (!eof && next_start >= n_bits) { finish("unexpected end of file"); return; }   // ran out of input inside a member
        if (!eof) launch(batches[cur ^ 1], pool, next_start, false);   // the workers move on while this thread replays
        // replay, in stream order; the CRC of every finished chunk is computed by the workers (front of their queue)
        for (size_t ci : chain) {
            Chunk& c = b.chunks[ci];
            const size_t need = kWindow + c.n_out + kSlack;
            c.bytes = take_buffer(need);                       // recycled output buffers keep their pages
            if (c.lit.size() < c.n_lit + 64) c.lit.resize(c.n_lit + 64);      // the literal copy reads 16 bytes at a time
            if (!c.bytes.p) { finish("out of memory"); return; }
            uint8_t* const base = c.bytes.p + kWindow;
            n_tok_total += c.n_tok; n_lit_total += c.n_lit;
            memcpy(c.bytes.p, history.data(), kWindow);
            uint8_t* out = base;
            for (Segment& sg : c.segs) {
                if (sg.starts_member) member_out = 0;
                sg.out_begin = (size_t)(out - base);
                uint8_t* const end = replay_tokens(c.tok.data() + sg.tok_begin, sg.tok_end - sg.tok_begin, c.lit.data() + sg.lit_begin,
                                                   out, std::min<uint64_t>(member_out, kWindow));
                if (!end || (size_t)(end - out) != sg.out_len) { finish(end ? "internal: replay length mismatch" : "invalid distance too far back"); return; }
                member_out += sg.out_len;
                out = end;
                if (sg.member_end) member_out = 0;
            }
            if (c.n_out >= kWindow) memcpy(history.data(), base + c.n_out - kWindow, kWindow);
            else if (c.n_out) {
                memmove(history.data(), history.data() + c.n_out, kWindow - c.n_out);
                memcpy(history.data() + kWindow - c.n_out, base, c.n_out);
            }
            crc_left.fetch_add(1);
            pool.submit([&c, &crc_left, &pool, base] {
                for (Segment& sg : c.segs) sg.crc_got = hast_crc32((uint32_t)crc32(0L, Z_NULL, 0), base + sg.out_begin, sg.out_len);
                if (crc_left.fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(pool.mu); pool.cv_done.notify_all(); }
            }, true);
        }
        t_ph[1] += now() - tp; tp = now();
        {
            std::unique_lock<std::mutex> lk(pool.mu);
            pool.cv_done.wait(lk, [&] { return crc_left.load() == 0; });
        }
        t_ph[2] += now() - tp; tp = now();
        // member checks and hand-over, in stream order
        uint64_t m_out = member_out_checked_;
        for (size_t ci : chain) {
            Chunk& c = b.chunks[ci];
            for (const Segment& s : c.segs) {
                if (s.starts_member) { crc_run = (uint32_t)crc32(0L, Z_NULL, 0); m_out = 0; }
                crc_run = (uint32_t)crc32_combine(crc_run, s.crc_got, (z_off_t)s.out_len);
                m_out += s.out_len;
                if (s.member_end) {
                    if (crc_run != s.crc) { finish("incorrect data check"); return; }
                    if ((uint32_t)(m_out & 0xFFFFFFFFu) != s.isize) { finish("incorrect length check"); return; }
                    crc_run = (uint32_t)crc32(0L, Z_NULL, 0);
                    m_out = 0;
                }
            }
            push(Piece{std::move(c.bytes), kWindow, c.n_out});
        }
        member_out_checked_ = m_out;
        t_ph[3] += now() - tp;
        if (eof) { finish(""); return; }
        cur ^= 1;
    }
}

}  // namespace hasthost
