// parser.cpp -- FASTQ text block -> device batch (bases, read offsets, barcode ids).
//
// Framing follows processFastq, classify.cpp:257-269: strict four-line records,
// only lines 1 and 2 are used, no '@' / '+' validation.  At end of file a header
// without its terminating newline is dropped (the loop condition tests eof after
// getline, :257) while a terminated header is processed together with whatever
// follows it as the sequence line, even an empty string (which the reference
// then dies on in chopRead2Kmer, kmer.h:171; here the device flags it and
// hast_finish fails).
#include <cstring>

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

bool pack_append(const char* seq, size_t n, uint32_t* words, size_t& n_words, uint64_t& acc, unsigned& nbits) {
    bool has_n = false;
    size_t i = 0;
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

bool parse_block(const TextBlock& blk, BarcodeIndex& index, Batch& out) {
    const bool packed = out.packed != nullptr;
    size_t n_words = 0;
    uint64_t acc = 0;
    unsigned nbits = 0;
    const char* p = blk.data.data();
    const char* const end = p + blk.len;
    out.n_reads = 0;
    out.n_bases = 0;
    out.max_barcode = 0;
    out.error.clear();
    uint32_t n = 0;
    uint64_t nb = 0;
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!nl) break;                                   // unterminated header at EOF: dropped
        const char* head = p;
        const size_t hlen = (size_t)(nl - p);
        p = nl + 1;
        const char* seq = p;
        size_t slen;
        nl = p < end ? (const char*)memchr(p, '\n', (size_t)(end - p)) : nullptr;
        if (nl) { slen = (size_t)(nl - p); p = nl + 1; }
        else { slen = (size_t)(end - p); p = end; }
        if (n >= out.cap_reads || nb + slen > (packed ? out.cap_words * 16 - 16 : out.cap_bases)) {
            out.error = "FASTQ records too small for the batch buffers (raise HAST_BLOCK_MB?)";
            return false;
        }
        size_t bs, bl;
        parse_name(head, hlen, bs, bl);                   // classify.cpp:112-119
        const uint32_t id = index.intern(head + bs, bl);
        if (id > out.max_barcode) out.max_barcode = id;
        if (packed) {
            if ((n & 31u) == 0) out.has_n[n >> 5] = 0;
            if (pack_append(seq, slen, out.packed, n_words, acc, nbits)) out.has_n[n >> 5] |= 1u << (n & 31u);
        } else {
            memcpy(out.bases + nb, seq, slen);
        }
        out.read_off[n] = (uint32_t)nb;
        out.barcode_id[n] = id;
        nb += slen;
        ++n;
        for (int i = 0; i < 2 && p < end; ++i) {          // '+' line and quality line
            nl = (const char*)memchr(p, '\n', (size_t)(end - p));
            p = nl ? nl + 1 : end;
        }
    }
    if (packed && nbits) out.packed[n_words++] = (uint32_t)(acc << (32 - nbits));   // zero-padded last word
    out.read_off[n] = (uint32_t)nb;
    out.n_reads = n;
    out.n_bases = nb;
    return true;
}

}  // namespace hasthost
