// inflate.cpp -- see inflate.h.  Written from RFC 1951 (DEFLATE) and RFC 1952 (gzip).
#include "inflate.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>          // crc32() only (crc32_clmul.h falls back to it)

#include "crc32_clmul.h"

#include <cerrno>
#include <cstring>
#include <cstdlib>

namespace hasthost {

namespace {

// Table entry (32 bits), bits 6..7 = kind:
//   literal(s)    bits 0..3 bits to drop, bits 4..5 number of literals - 1, bits 8..31 up to three literal bytes
//                 (first in the low byte): a root entry whose remaining index bits already determine the next one or
//                 two literal codes carries them too, so FASTQ text (2-4 bit codes for ACGT and the common quality
//                 letters) decodes two or three bytes per table lookup
//   base          bits 0..5 bits to drop = code length + extra bits, bits 8..12 code length alone (the extra bits
//                 sit at saved >> codelen), bits 16..31 base length / distance
//   end of block  bits 0..5 bits to drop, bits 16..31 zero; 0xFFFF there marks an invalid code
//   subtable      bits 0..5 root bits, bits 8..12 subtable bits, bits 16..31 subtable start
constexpr uint32_t kKindMask = 3u << 6;
constexpr uint32_t kLit = 0u << 6, kBase = 1u << 6, kEob = 2u << 6, kSub = 3u << 6;
constexpr uint32_t kInvalid = kEob | (0xFFFFu << 16) | 1u;     // drops one bit, flagged invalid
constexpr int kLitlenRoot = GzipInflater::kLitlenRoot, kDistRoot = GzipInflater::kDistRoot;
constexpr size_t kLitlenCap = GzipInflater::kLitlenCap, kDistCap = GzipInflater::kDistCap;
constexpr size_t kTail = 64;                                   // private, zero-padded copy of the last input bytes
constexpr size_t kSlack = 258 + 64;                            // a match may run past the chunk limit

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115,
                               131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537,
                                2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }   // little-endian hosts
inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

}  // namespace

GzipInflater::GzipInflater()
    : out_(kWindow + kChunk + kSlack + 64), litlen_(kLitlenCap), dist_(kDistCap), fixed_litlen_(kLitlenCap), fixed_dist_(kDistCap) {}

GzipInflater::~GzipInflater() {
    if (map_) munmap(const_cast<uint8_t*>(map_), map_len_);
    if (fd_ >= 0) close(fd_);
}

std::string GzipInflater::open(const std::string& path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) return "cannot open " + path + ": " + strerror(errno);
    struct stat sb;
    if (fstat(fd_, &sb) != 0 || !S_ISREG(sb.st_mode)) return "not a regular file: " + path;
    if (sb.st_size > 0) {
        void* p = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (p == MAP_FAILED) return "cannot map " + path + ": " + strerror(errno);
        map_ = (const uint8_t*)p;
        map_len_ = (size_t)sb.st_size;
        madvise(p, map_len_, MADV_SEQUENTIAL);
    }
    open_memory(map_, map_len_);
    return "";
}

void GzipInflater::open_memory(const uint8_t* data, size_t n) {
    in_base_ = data;
    n_in_ = n;
    const size_t t = n < kTail ? n : kTail;
    tail_src_ = data + (n - t);
    tail_.assign(t + 64, 0);
    if (t) memcpy(tail_.data(), tail_src_, t);
    in_ = data;
    in_end_ = tail_src_;
    in_tail_ = false;
    if (in_ >= in_end_) remap_tail();
    bitbuf_ = 0;
    bitcnt_ = 0;
    phase_ = n ? Phase::kHeader : Phase::kDone;
    produced_ = 0;
    total_out_ = 0;
    members_ = 0;
    err_.clear();
}

void GzipInflater::remap_tail() {
    consumed_before_tail_ = (uint64_t)(tail_src_ - in_base_);
    in_ = tail_.data() + (in_ - tail_src_);
    in_end_ = tail_.data() + (tail_.size() - 64);
    in_tail_ = true;
}

bool GzipInflater::fail(const char* msg) {
    if (err_.empty()) err_ = msg;
    phase_ = Phase::kError;
    return false;
}

inline void GzipInflater::refill() {
    if (in_ >= in_end_ && !in_tail_) remap_tail();
    if (in_tail_ && in_ + 8 > tail_.data() + tail_.size()) return;        // far past the end: leave zeros, overrun is reported
    bitbuf_ |= load64(in_) << bitcnt_;
    in_ += (63 - bitcnt_) >> 3;
    bitcnt_ |= 56;
}

inline uint32_t GzipInflater::take(int n) {
    refill();
    const uint32_t v = (uint32_t)(bitbuf_ & ((1ull << n) - 1));
    bitbuf_ >>= n;
    bitcnt_ -= (unsigned)n;
    return v;
}

inline void GzipInflater::align_byte() {
    const unsigned d = bitcnt_ & 7u;
    bitbuf_ >>= d;
    bitcnt_ -= d;
}

bool GzipInflater::input_overrun() const { return in_tail_ && in_ - (bitcnt_ >> 3) > in_end_; }

uint64_t GzipInflater::bytes_in() const {
    const uint64_t pos = in_tail_ ? consumed_before_tail_ + (uint64_t)(in_ - tail_.data()) : (uint64_t)(in_ - in_base_);
    return pos - (bitcnt_ >> 3);
}

bool GzipInflater::parse_header() {
    align_byte();
    if (bytes_in() >= n_in_) { phase_ = Phase::kDone; return true; }
    if (members_ > 0) {
        // after the first member: anything that is not another gzip member is ignored, like zlib's gzread
        if (n_in_ - bytes_in() < 2) { phase_ = Phase::kDone; return true; }
        refill();
        if ((bitbuf_ & 0xFFFF) != 0x8B1F) { phase_ = Phase::kDone; return true; }
    }
    const uint32_t id1 = take(8), id2 = take(8), cm = take(8), flg = take(8);
    if (id1 != 0x1F || id2 != 0x8B) return fail("not in gzip format");
    if (cm != 8) return fail("unknown compression method");
    if (flg & 0xE0) return fail("unknown header flags set");
    take(32);                                            // MTIME
    take(16);                                            // XFL, OS
    if (flg & 4) {                                       // FEXTRA
        uint32_t xlen = take(16);
        while (xlen--) { take(8); if (input_overrun()) return fail("unexpected end of file"); }
    }
    for (int bit : {8, 16})                              // FNAME, FCOMMENT
        if (flg & bit)
            for (;;) {
                const uint32_t c = take(8);
                if (input_overrun()) return fail("unexpected end of file");
                if (!c) break;
            }
    if (flg & 2) take(16);                               // FHCRC
    if (input_overrun()) return fail("unexpected end of file");
    crc_ = (uint32_t)crc32(0L, Z_NULL, 0);
    member_out_ = 0;
    phase_ = Phase::kBlockHeader;
    return true;
}

bool GzipInflater::parse_trailer() {
    align_byte();
    const uint32_t crc = take(32);
    const uint32_t isize = take(32);
    if (input_overrun()) return fail("unexpected end of file");
    if (crc != crc_) return fail("incorrect data check");
    if (isize != (uint32_t)(member_out_ & 0xFFFFFFFFu)) return fail("incorrect length check");
    ++members_;
    phase_ = Phase::kHeader;
    return true;
}

bool GzipInflater::build_table(const uint8_t* lens, int n, int root, uint32_t* table, size_t cap, bool litlen,
                               bool pack_literals) {
    uint16_t count[16] = {0};
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    count[0] = 0;
    int max_len = 0;
    for (int l = 1; l <= 15; ++l) if (count[l]) max_len = l;
    const size_t root_size = (size_t)1 << root;
    for (size_t i = 0; i < root_size; ++i) table[i] = kInvalid;
    if (max_len == 0) return true;                       // no codes: any use of the table is an error
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;                      // over-subscribed
    }
    if (left > 0 && max_len != 1) return false;          // incomplete (a single one-bit code is tolerated, as in zlib)
    uint16_t next_code[16];
    {
        uint32_t code = 0;
        for (int l = 1; l <= 15; ++l) { code = (code + count[l - 1]) << 1; next_code[l] = (uint16_t)code; }
    }
    auto entry = [&](int sym, int len_bits) -> uint32_t {  // len_bits = code bits this lookup still has to drop
        if (litlen) {
            if (sym < 256) return kLit | (uint32_t)len_bits | ((uint32_t)sym << 8);
            if (sym == 256) return kEob | (uint32_t)len_bits;
            if (sym > 285) return kInvalid;
            const uint32_t ex = kLenExtra[sym - 257];
            return kBase | (uint32_t)(len_bits + ex) | ((uint32_t)len_bits << 8) | ((uint32_t)kLenBase[sym - 257] << 16);
        }
        if (sym > 29) return kInvalid;
        const uint32_t ex = kDistExtra[sym];
        return kBase | (uint32_t)(len_bits + ex) | ((uint32_t)len_bits << 8) | ((uint32_t)kDistBase[sym] << 16);
    };
    // pass A: longest code under every root prefix that needs a subtable
    uint8_t submax[1 << kLitlenRoot];
    if (max_len > root) {
        memset(submax, 0, root_size);
        uint16_t nc[16];
        memcpy(nc, next_code, sizeof nc);
        for (int i = 0; i < n; ++i) {
            const int l = lens[i];
            if (l <= root) { if (l) nc[l]++; continue; }
            const uint32_t rev = reverse_bits(nc[l]++, l);
            uint8_t& m = submax[rev & (root_size - 1)];
            if (l > m) m = (uint8_t)l;
        }
    }
    size_t next_free = root_size;
    for (int i = 0; i < n; ++i) {
        const int l = lens[i];
        if (!l) continue;
        const uint32_t rev = reverse_bits(next_code[l]++, l);
        if (l <= root) {
            const uint32_t e = entry(i, l);
            for (size_t j = rev; j < root_size; j += (size_t)1 << l) table[j] = e;
        } else {
            const size_t prefix = rev & (root_size - 1);
            uint32_t& ptr = table[prefix];
            const int sub_bits = submax[prefix] - root;
            if ((ptr & kKindMask) != kSub) {
                const size_t size = (size_t)1 << sub_bits;
                if (next_free + size > cap) return false;
                for (size_t j = 0; j < size; ++j) table[next_free + j] = kInvalid;
                ptr = kSub | (uint32_t)root | ((uint32_t)sub_bits << 8) | ((uint32_t)next_free << 16);
                next_free += size;
            }
            const size_t start = ptr >> 16;
            const uint32_t e = entry(i, l - root);
            for (size_t j = rev >> root; j < ((size_t)1 << sub_bits); j += (size_t)1 << (l - root)) table[start + j] = e;
        }
    }
    static const bool no_pack = getenv("HAST_NOPACK") != nullptr;
    if (litlen && pack_literals && !no_pack) {
        // pack the literals that follow a short literal code into its root entries
        uint32_t single[1 << kLitlenRoot];
        memcpy(single, table, root_size * sizeof(uint32_t));
        for (size_t i = 0; i < root_size; ++i) {
            uint32_t e = single[i];
            if ((e & kKindMask) != kLit) continue;
            int used = (int)(e & 15u), n_lit = 1;
            uint32_t bytes = e >> 8;
            while (n_lit < 3 && used < root) {
                const uint32_t e2 = single[i >> used];           // the upper index bits read as zero
                if ((e2 & kKindMask) != kLit || (int)(e2 & 15u) > root - used) break;
                bytes |= (e2 >> 8) << (8 * n_lit);
                used += (int)(e2 & 15u);
                ++n_lit;
            }
            table[i] = kLit | (uint32_t)used | ((uint32_t)(n_lit - 1) << 4) | (bytes << 8);
        }
    }
    return true;
}

void GzipInflater::build_fixed() {
    if (fixed_built_) return;
    uint8_t lens[288];
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    build_table(lens, 288, kLitlenRoot, fixed_litlen_.data(), kLitlenCap, true, true);
    uint8_t d[32];
    for (int i = 0; i < 32; ++i) d[i] = 5;
    build_table(d, 32, kDistRoot, fixed_dist_.data(), kDistCap, false, false);
    fixed_built_ = true;
}

bool GzipInflater::build_dynamic() {
    const int hlit = (int)take(5) + 257, hdist = (int)take(5) + 1, hclen = (int)take(4) + 4;
    if (hlit > 286 || hdist > 30) return fail("too many length or distance symbols");
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t cl[19] = {0};
    for (int i = 0; i < hclen; ++i) cl[order[i]] = (uint8_t)take(3);
    // code-length code: at most 7 bits, must be complete
    uint16_t pre[128];
    {
        uint16_t count[8] = {0};
        for (int i = 0; i < 19; ++i) count[cl[i]]++;
        count[0] = 0;
        int left = 1;
        for (int l = 1; l <= 7; ++l) { left <<= 1; left -= count[l]; if (left < 0) return fail("invalid code lengths set"); }
        if (left > 0) return fail("invalid code lengths set");
        uint16_t next_code[8];
        uint32_t code = 0;
        for (int l = 1; l <= 7; ++l) { code = (code + count[l - 1]) << 1; next_code[l] = (uint16_t)code; }
        for (int i = 0; i < 19; ++i) {
            const int l = cl[i];
            if (!l) continue;
            const uint32_t rev = reverse_bits(next_code[l]++, l);
            for (uint32_t j = rev; j < 128; j += 1u << l) pre[j] = (uint16_t)(i | (l << 8));
        }
    }
    uint8_t lens[320];
    const int total = hlit + hdist;
    int i = 0;
    while (i < total) {
        refill();
        const uint16_t e = pre[bitbuf_ & 127];
        const int sym = e & 0xFF, l = e >> 8;
        bitbuf_ >>= l;
        bitcnt_ -= (unsigned)l;
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        int rep, val = 0;
        if (sym == 16) {
            if (i == 0) return fail("invalid bit length repeat");
            val = lens[i - 1];
            rep = 3 + (int)(bitbuf_ & 3); bitbuf_ >>= 2; bitcnt_ -= 2;
        } else if (sym == 17) {
            rep = 3 + (int)(bitbuf_ & 7); bitbuf_ >>= 3; bitcnt_ -= 3;
        } else {
            rep = 11 + (int)(bitbuf_ & 127); bitbuf_ >>= 7; bitcnt_ -= 7;
        }
        if (i + rep > total) return fail("invalid bit length repeat");
        while (rep--) lens[i++] = (uint8_t)val;
        if (input_overrun()) return fail("unexpected end of file");
    }
    if (input_overrun()) return fail("unexpected end of file");
    if (lens[256] == 0) return fail("invalid code -- missing end-of-block");
    if (!build_table(lens, hlit, kLitlenRoot, litlen_.data(), kLitlenCap, true, true)) return fail("invalid literal/lengths set");
    if (!build_table(lens + hlit, hdist, kDistRoot, dist_.data(), kDistCap, false, false)) return fail("invalid distances set");
    cur_litlen_ = litlen_.data();
    cur_dist_ = dist_.data();
    return true;
}

bool GzipInflater::read_block_header() {
    const uint32_t h = take(3);
    if (input_overrun()) return fail("unexpected end of file");
    last_block_ = (h & 1u) != 0;
    switch (h >> 1) {
        case 0: {
            align_byte();
            const uint32_t len = take(16), nlen = take(16);
            if (input_overrun()) return fail("unexpected end of file");
            if ((len ^ 0xFFFFu) != nlen) return fail("invalid stored block lengths");
            stored_left_ = len;
            // hand the whole bytes still in the bit buffer back to the byte stream
            in_ -= bitcnt_ >> 3;
            bitbuf_ = 0;
            bitcnt_ = 0;
            phase_ = Phase::kStored;
            return true;
        }
        case 1:
            build_fixed();
            cur_litlen_ = fixed_litlen_.data();
            cur_dist_ = fixed_dist_.data();
            phase_ = Phase::kHuffman;
            return true;
        case 2:
            if (!build_dynamic()) return false;
            phase_ = Phase::kHuffman;
            return true;
        default:
            return fail("invalid block type");
    }
}

// One Huffman-coded block, until its end-of-block symbol or until `out` reaches `limit`
// (a match that starts before the limit completes into the slack behind it).
//
// Loop invariant at the top: the bit buffer was refilled after the last bits were dropped (>= 56 valid
// bits) and `e` is the literal/length root entry for its low bits -- a refill only ORs bits in above the
// valid ones, so an entry looked up before it stays right.  The entry of the NEXT symbol is loaded
// before a match is copied, so that the table load overlaps the copy.
bool GzipInflater::decode_huffman(uint8_t*& out_ref, uint8_t* limit) {
    const uint32_t* const lt = cur_litlen_;
    const uint32_t* const dt = cur_dist_;
    uint8_t* out = out_ref;
    uint8_t* const seg = out;                            // member_out_ counts the member's bytes before `seg`
    uint64_t bb = bitbuf_;
    unsigned bc = bitcnt_;
    const uint8_t* in = in_;
    const uint8_t* in_end = in_end_;
    const char* why = nullptr;
    constexpr uint32_t kRootMask = (1u << kLitlenRoot) - 1u, kDistMask = (1u << kDistRoot) - 1u;

#define HAST_REFILL()                      \
    do {                                   \
        bb |= load64(in) << bc;            \
        in += (63 - bc) >> 3;              \
        bc |= 56;                          \
    } while (0)
#define HAST_DROP(n)  do { bb >>= (n); bc -= (n); } while (0)
    // the input region `in` walks: switch to the padded private copy of the last bytes, or stop at its end
#define HAST_CHECK_INPUT()                                                                   \
    do {                                                                                     \
        if (in >= in_end) {                                                                  \
            if (!in_tail_) {                                                                 \
                in_ = in; remap_tail(); in = in_; in_end = in_end_;                          \
            } else if (in - (bc >> 3) > in_end || in + 24 > tail_.data() + tail_.size()) {   \
                why = "unexpected end of file";                                              \
                goto done;                                                                   \
            }                                                                                \
        }                                                                                    \
    } while (0)
#define HAST_LITERALS(e)                              \
    do {                                              \
        const uint32_t v_ = e >> 8;                   \
        memcpy(out, &v_, 4);                          \
        out += 1u + ((e >> 4) & 3u);                  \
        HAST_DROP(e & 15u);                           \
    } while (0)
    // a code longer than the root: drop the root bits, refill (keeps the 48-bit budget of a length +
    // distance pair), continue in the subtable
#define HAST_SUBTABLE(e, table, root)                                                   \
    do {                                                                                \
        HAST_DROP(root);                                                                \
        HAST_REFILL();                                                                  \
        e = table[(e >> 16) + (bb & ((1u << ((e >> 8) & 31u)) - 1u))];                  \
    } while (0)

    uint32_t e;
    HAST_CHECK_INPUT();
    HAST_REFILL();
    e = lt[bb & kRootMask];
    for (;;) {
        if (out >= limit) break;
        HAST_CHECK_INPUT();
        if ((e & kKindMask) == kSub) HAST_SUBTABLE(e, lt, kLitlenRoot);
        if ((e & kKindMask) == kLit) {
            HAST_LITERALS(e);                            // >= 41 bits left
            e = lt[bb & kRootMask];
            if ((e & kKindMask) == kSub) HAST_SUBTABLE(e, lt, kLitlenRoot);
            if ((e & kKindMask) == kLit) {
                HAST_LITERALS(e);                        // >= 26 bits left
                e = lt[bb & kRootMask];
                if ((e & kKindMask) == kSub) HAST_SUBTABLE(e, lt, kLitlenRoot);
                if ((e & kKindMask) == kLit) {
                    HAST_LITERALS(e);
                    HAST_REFILL();
                    e = lt[bb & kRootMask];
                    continue;
                }
            }
            HAST_REFILL();                               // e is a length or the end of the block: full budget again
        }
        if ((e & kKindMask) == kEob) {
            if ((e >> 16) != 0) { why = "invalid literal/length code"; goto done; }
            HAST_DROP(e & 63u);
            phase_ = last_block_ ? Phase::kTrailer : Phase::kBlockHeader;
            break;
        }
        {
            // length (<= 20 bits) and distance (<= 28 bits) out of the >= 48 bits at hand
            const uint64_t saved = bb;
            const uint32_t cl = (e >> 8) & 31u, drop = e & 63u;
            HAST_DROP(drop);
            const uint32_t length = (e >> 16) + (uint32_t)((saved >> cl) & ((1u << (drop - cl)) - 1u));
            uint32_t d = dt[bb & kDistMask];
            if ((d & kKindMask) == kSub) {
                HAST_DROP(kDistRoot);
                d = dt[(d >> 16) + (bb & ((1u << ((d >> 8) & 31u)) - 1u))];
            }
            if ((d & kKindMask) != kBase) { why = "invalid distance code"; goto done; }
            const uint64_t saved2 = bb;
            const uint32_t cl2 = (d >> 8) & 31u, drop2 = d & 63u;
            HAST_DROP(drop2);
            const uint32_t dist = (d >> 16) + (uint32_t)((saved2 >> cl2) & ((1u << (drop2 - cl2)) - 1u));
            HAST_REFILL();
            e = lt[bb & kRootMask];                      // next symbol's entry: its load overlaps the copy
            if ((uint64_t)dist > member_out_ + (uint64_t)(out - seg)) { why = "invalid distance too far back"; goto done; }
            const uint8_t* src = out - dist;
            uint8_t* const end = out + length;
            if (dist >= 8) {                             // most matches are short: two words unconditionally
                memcpy(out, src, 8);
                memcpy(out + 8, src + 8, 8);
                if (length > 16) {
                    out += 16; src += 16;
                    do { memcpy(out, src, 8); out += 8; src += 8; } while (out < end);
                }
            } else if (dist == 1) {
                memset(out, src[0], length);
            } else {
                do { *out++ = *src++; } while (out < end);
            }
            out = end;
        }
    }
done:
#undef HAST_REFILL
#undef HAST_DROP
#undef HAST_CHECK_INPUT
#undef HAST_LITERALS
#undef HAST_SUBTABLE
    bitbuf_ = bb;
    bitcnt_ = bc;
    in_ = in;
    out_ref = out;
    if (why) return fail(why);
    return true;
}

bool GzipInflater::next(const uint8_t** data, size_t* len) {
    *data = nullptr;
    *len = 0;
    if (phase_ == Phase::kDone || phase_ == Phase::kError) return false;
    uint8_t* const buf = out_.data();
    if (produced_) memmove(buf, buf + produced_, kWindow);       // the last 32 KiB stay in front of the new chunk
    uint8_t* const chunk = buf + kWindow;
    uint8_t* const limit = chunk + kChunk;
    uint8_t* out = chunk;
    uint8_t* seg = chunk;                                         // start of the bytes not yet added to crc_ / member_out_
    auto account = [&] {
        if (out > seg) {
            crc_ = hast_crc32(crc_, seg, (size_t)(out - seg));
            member_out_ += (uint64_t)(out - seg);
        }
        seg = out;
    };
    while (out < limit && phase_ != Phase::kDone && phase_ != Phase::kError) {
        switch (phase_) {
            case Phase::kHeader:
                account();
                parse_header();
                break;
            case Phase::kBlockHeader:
                read_block_header();
                break;
            case Phase::kStored: {
                while (stored_left_ && out < limit) {
                    if (in_ >= in_end_) {
                        if (!in_tail_) remap_tail();
                        if (in_ >= in_end_) { fail("unexpected end of file"); break; }
                    }
                    size_t n = stored_left_;
                    n = std::min<size_t>(n, (size_t)(in_end_ - in_));
                    n = std::min<size_t>(n, (size_t)(limit - out));
                    memcpy(out, in_, n);
                    out += n;
                    in_ += n;
                    stored_left_ -= (uint32_t)n;
                }
                if (!stored_left_ && phase_ == Phase::kStored) phase_ = last_block_ ? Phase::kTrailer : Phase::kBlockHeader;
                break;
            }
            case Phase::kHuffman:
                account();                                        // decode_huffman measures the member from `out`
                decode_huffman(out, limit);
                account();
                break;
            case Phase::kTrailer:
                account();
                parse_trailer();
                break;
            default:
                break;
        }
    }
    account();
    if (phase_ == Phase::kError) return false;
    produced_ = (size_t)(out - chunk);
    total_out_ += produced_;
    *data = chunk;
    *len = produced_;
    return produced_ > 0;
}

}  // namespace hasthost
