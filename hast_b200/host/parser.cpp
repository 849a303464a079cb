// parser.cpp -- FASTQ text block -> device batch (bases, read offsets, barcode ids).
//
// Framing follows processFastq, classify.cpp:257-269: strict four-line records,
// only lines 1 and 2 are used, no '@' / '+' validation.  At end of file a header
// without its terminating newline is dropped (the loop condition tests eof after
// getline, :257) while a terminated header is processed together with whatever
// follows it as the sequence line, even an empty string (which the reference
// then dies on in chopRead2Kmer, kmer.h:171; here the device flags it and
// hast_finish fails).
#include <algorithm>
#include <cstring>
#include <immintrin.h>
#include <string>
#include <vector>

#include "host.h"

namespace hasthost {

// 4 ASCII bases (little-endian load, first base in the low byte) -> 8 bits, first base on top:
// codes = (w >> 1) & 0x03030303 leaves base i in byte i, the multiply gathers the four 2-bit
// fields into the top byte without carries (same construction as pack4 in csrc/kmer.cuh).
static inline uint32_t pack4_host(uint32_t w) { return (((w >> 1) & 0x03030303u) * 0x40100401u) >> 24; }
static inline bool any_N4_host(uint32_t w) {
    const uint32_t x = w ^ 0x4E4E4E4Eu;
    return ((x - 0x01010101u) & ~x & 0x80808080u) != 0u;
}

// 16 bases per step with SSSE3: codes = (byte >> 1) & 3, pmaddubsw folds base pairs (x4 + y), pmaddwd folds
// the pairs of pairs (x16 + y), a byte shuffle lines the four 8-bit groups up first-base-on-top.
__attribute__((target("ssse3"))) static size_t pack_run_ssse3(const char* seq, size_t n, uint32_t* words, size_t& n_words,
                                                              uint64_t& acc, unsigned& nbits, bool& has_n) {
    const __m128i three = _mm_set1_epi8(3), w1 = _mm_set1_epi16(0x0104), w2 = _mm_set1_epi32(0x00010010);
    const __m128i gather = _mm_set_epi8(-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12);
    const __m128i big_n = _mm_set1_epi8('N');
    __m128i any_n = _mm_setzero_si128();
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(seq + i));
        any_n = _mm_or_si128(any_n, _mm_cmpeq_epi8(v, big_n));
        const __m128i codes = _mm_and_si128(_mm_srli_epi16(v, 1), three);
        const __m128i quads = _mm_madd_epi16(_mm_maddubs_epi16(codes, w1), w2);
        const uint32_t w = (uint32_t)_mm_cvtsi128_si32(_mm_shuffle_epi8(quads, gather));
        acc = (acc << 32) | w;
        words[n_words++] = (uint32_t)(acc >> nbits);     // nbits < 32 pending bits stay pending
    }
    if (_mm_movemask_epi8(any_n)) has_n = true;
    return i;
}

bool pack_append(const char* seq, size_t n, uint32_t* words, size_t& n_words, uint64_t& acc, unsigned& nbits) {
    bool has_n = false;
    size_t i = 0;
    static const bool have_ssse3 = __builtin_cpu_supports("ssse3");
    if (have_ssse3 && n >= 16) i = pack_run_ssse3(seq, n, words, n_words, acc, nbits, has_n);
    for (; i + 4 <= n; i += 4) {
        uint32_t w;
        memcpy(&w, seq + i, 4);
        has_n |= any_N4_host(w);
        acc = (acc << 8) | pack4_host(w);
        nbits += 8;
        if (nbits >= 32) { words[n_words++] = (uint32_t)(acc >> (nbits - 32)); nbits -= 32; }
    }
    for (; i < n; ++i) {
        const unsigned c = (unsigned char)seq[i];
        has_n |= c == 'N';
        acc = (acc << 2) | ((c >> 1) & 3u);
        nbits += 2;
        if (nbits >= 32) { words[n_words++] = (uint32_t)(acc >> (nbits - 32)); nbits -= 32; }
    }
    return has_n;
}

// Offsets of every '\n' of a block, found in one vector pass (the record loop below then walks this index
// instead of calling memchr four times per record).
__attribute__((target("avx2"))) static size_t newline_index_avx2(const char* p, size_t n, std::vector<uint32_t>& nl, size_t at) {
    const __m256i c = _mm256_set1_epi8('\n');
    size_t i = 0, k = at;
    uint32_t* out = nl.data();
    size_t cap = nl.size();
    for (; i + 32 <= n; i += 32) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + i));
        uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, c));
        if (!m) continue;
        if (k + 32 > cap) { nl.resize(std::max<size_t>(cap * 2, k + 4096)); out = nl.data(); cap = nl.size(); }
        do {
            out[k++] = (uint32_t)i + (uint32_t)__builtin_ctz(m);
            m &= m - 1;
        } while (m);
    }
    if (k + 32 > cap) { nl.resize(k + 64); out = nl.data(); }
    for (; i < n; ++i)
        if (p[i] == '\n') out[k++] = (uint32_t)i;
    return k - at;
}
size_t newline_index(const char* p, size_t n, std::vector<uint32_t>& nl, size_t at) {
    if (nl.size() < at + n / 48 + 64) nl.resize(at + n / 48 + 64);
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return newline_index_avx2(p, n, nl, at);
    size_t k = at;
    for (size_t i = 0; i < n; ++i)
        if (p[i] == '\n') {
            if (k >= nl.size()) nl.resize(nl.size() * 2);
            nl[k++] = (uint32_t)i;
        }
    return k - at;
}

// Longest read the device path takes: the smaller pass of the fused kernels (FusedSmem<true>::kCap = 24576 bases with
// TMA staging, 40960 without) minus its 16-byte alignment slack.  Checked here so that an over-long read stops the run
// at once, with its name, instead of after all input has been streamed (hast_finish would report it too).
static constexpr size_t kMaxReadBases = 24560;

bool parse_block(const TextBlock& blk, BarcodeIndex& index, Batch& out, size_t* resume) {
    const bool packed = out.packed != nullptr;
    size_t n_words = 0;
    uint64_t acc = 0;
    unsigned nbits = 0;
    const size_t skip = resume ? *resume : 0;             // records before this offset went into earlier batches
    const char* const base = blk.text() + skip;
    const char* p = base;
    const size_t len = blk.len - skip;
    const char* const end = p + len;
    if (resume) *resume = blk.len;
    if (len > 0xFFFFFFFFull) { out.error = "FASTQ block larger than 4 GiB"; return false; }
    static thread_local std::vector<uint32_t> nl_index;   // one per parser thread, reused from block to block
    const uint32_t* nls;
    size_t n_nl;
    if (blk.has_nl && skip == 0) {                        // the producer indexed the block while framing it
        nls = blk.nl.data() + blk.nl_begin;
        n_nl = blk.nl_count;
    } else {
        n_nl = newline_index(base, len, nl_index, 0);
        nls = nl_index.data();
    }
    size_t li = 0;                                        // next unused entry of nls: the newline that ends line li
    out.n_reads = 0;
    out.n_bases = 0;
    out.max_barcode = 0;
    out.error.clear();
    uint32_t n = 0;
    uint64_t nb = 0;
    // barcode interning runs kRing records behind the framing, with the table line prefetched in between
    constexpr uint32_t kRing = 8;
    struct Pending { const char* s; size_t len; uint64_t h; } ring[kRing];
    auto resolve = [&](uint32_t r) {
        const Pending& q = ring[r % kRing];
        const uint32_t id = index.intern_hashed(q.h, q.s, q.len);
        if (id > out.max_barcode) out.max_barcode = id;
        out.barcode_id[r] = id;
    };
    while (p < end) {
        if (li >= n_nl) break;                            // unterminated header at EOF: dropped
        const char* nl = base + nls[li];
        const char* head = p;
        const size_t hlen = (size_t)(nl - p);
        p = nl + 1;
        const char* seq = p;
        size_t slen;
        if (li + 1 < n_nl) { nl = base + nls[li + 1]; slen = (size_t)(nl - p); p = nl + 1; }
        else { slen = (size_t)(end - p); p = end; }
        if (n >= out.cap_reads || nb + slen > (packed ? out.cap_words * 16 - 16 : out.cap_bases)) {
            // the batch is full before the block is used up (records far shorter than the sizing assumes):
            // hand this batch over and let the caller come back for the rest of the block
            if (n == 0 || !resume) { out.error = "FASTQ record larger than the batch buffers (raise HAST_BLOCK_MB)"; return false; }
            *resume = skip + (size_t)(head - base);
            break;
        }
        if (slen > kMaxReadBases) {                       // one pass of the fused kernel holds a read whole (csrc/fused.cuh)
            out.error = "read '" + std::string(head, std::min<size_t>(hlen, 200)) + "' has " + std::to_string(slen) +
                        " bases; this build handles reads up to " + std::to_string(kMaxReadBases) + " bases";
            return false;
        }
        size_t bs, bl;
        parse_name(head, hlen, bs, bl);                   // classify.cpp:112-119
        if (n >= kRing) resolve(n - kRing);
        Pending& q = ring[n % kRing];
        q.s = head + bs;
        q.len = bl;
        q.h = BarcodeIndex::hash(q.s, q.len);
        index.prefetch(q.h);
        if (packed) {
            if ((n & 31u) == 0) out.has_n[n >> 5] = 0;
            if (pack_append(seq, slen, out.packed, n_words, acc, nbits)) out.has_n[n >> 5] |= 1u << (n & 31u);
        } else {
            memcpy(out.bases + nb, seq, slen);
        }
        out.read_off[n] = (uint32_t)nb;
        nb += slen;
        ++n;
        p = li + 3 < n_nl ? base + nls[li + 3] + 1 : end;    // past the '+' line and the quality line
        li += 4;
    }
    for (uint32_t r = n >= kRing ? n - kRing : 0; r < n; ++r) resolve(r);
    if (packed && nbits) out.packed[n_words++] = (uint32_t)(acc << (32 - nbits));   // zero-padded last word
    out.read_off[n] = (uint32_t)nb;
    out.n_reads = n;
    out.n_bases = nb;
    return true;
}

}  // namespace hasthost
