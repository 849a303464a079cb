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

bool parse_block(const TextBlock& blk, BarcodeIndex& index, Batch& out) {
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
        if (n >= out.cap_reads || nb + slen > out.cap_bases) {
            out.error = "FASTQ records too small for the batch buffers (raise HAST_BLOCK_MB?)";
            return false;
        }
        size_t bs, bl;
        parse_name(head, hlen, bs, bl);                   // classify.cpp:112-119
        const uint32_t id = index.intern(head + bs, bl);
        if (id > out.max_barcode) out.max_barcode = id;
        memcpy(out.bases + nb, seq, slen);
        out.read_off[n] = (uint32_t)nb;
        out.barcode_id[n] = id;
        nb += slen;
        ++n;
        for (int i = 0; i < 2 && p < end; ++i) {          // '+' line and quality line
            nl = (const char*)memchr(p, '\n', (size_t)(end - p));
            p = nl ? nl + 1 : end;
        }
    }
    out.read_off[n] = (uint32_t)nb;
    out.n_reads = n;
    out.n_bases = nb;
    return true;
}

}  // namespace hasthost
