// inflate.h -- streaming gzip (RFC 1952 / DEFLATE RFC 1951) decoder for the FASTQ readers.
//
// The reference reads `.gz` input through gzstream (gzstream/gzstream.C:78-101: gzread of 299 bytes
// at a time on top of zlib).  Inflating is the slowest stage of the whole drop-in process by three
// orders of magnitude, so the readers use this decoder instead of zlib's:
// whole file mapped, 64-bit bit buffer refilled branch-free, two-level Huffman tables whose
// entries carry base value + extra-bit count, several literals per refill, word-wise match copy,
// 32 KiB of history kept in front of every output chunk.  Concatenated members (bgzip, `cat a.gz
// b.gz`) are decoded back to back like zlib's gzread does; CRC-32 and ISIZE of every member are
// verified (zlib's crc32).  Written from RFC 1951/1952; shares no code with zlib.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace hasthost {

class GzipInflater {
public:
    GzipInflater();
    ~GzipInflater();
    GzipInflater(const GzipInflater&) = delete;
    GzipInflater& operator=(const GzipInflater&) = delete;

    // Map `path`.  Returns "" or an error message.  is_gzip() tells whether it starts with the gzip magic.
    std::string open(const std::string& path);
    // Decode from memory instead (tests).  The buffer must stay alive.
    void open_memory(const uint8_t* data, size_t n);
    bool is_gzip() const { return n_in_ >= 2 && in_base_[0] == 0x1F && in_base_[1] == 0x8B; }

    // Next chunk of decompressed bytes: *data points into an internal buffer that stays valid until
    // the next call.  Returns true with *len > 0 while there is output; false at the end of the last
    // member or on error (error() non-empty).
    bool next(const uint8_t** data, size_t* len);
    const std::string& error() const { return err_; }
    uint64_t bytes_in() const;
    uint64_t bytes_out() const { return total_out_; }

    static constexpr size_t kWindow = 32768;
    // Two-level decode table of one canonical Huffman code (entry layout: inflate.cpp).  `pack_literals`: root
    // entries of short literal codes also carry the one or two literals that follow.  Shared with the
    // parallel decoder (inflate_par.cpp).  Returns false for an over-subscribed or incomplete code.
    static constexpr int kLitlenRoot = 11, kDistRoot = 8;
    static constexpr size_t kLitlenCap = 4096, kDistCap = 1024;
    static bool build_table(const uint8_t* lens, int n, int root_bits, uint32_t* table, size_t cap, bool litlen,
                            bool pack_literals);
    static constexpr size_t kChunk = 1u << 20;

private:
    enum class Phase { kHeader, kBlockHeader, kStored, kHuffman, kTrailer, kDone, kError };
    bool fail(const char* msg);
    bool parse_header();
    bool parse_trailer();
    bool read_block_header();
    bool build_dynamic();
    void build_fixed();
    bool decode_huffman(uint8_t*& out, uint8_t* out_limit);
    void remap_tail();
    void refill();
    uint32_t take(int n);           // n <= 32 bits, after refill
    void align_byte();
    bool input_overrun() const;

    // input
    int fd_ = -1;
    const uint8_t* map_ = nullptr;  size_t map_len_ = 0;
    const uint8_t* in_base_ = nullptr;  size_t n_in_ = 0;
    const uint8_t* in_ = nullptr;       // next byte to load into the bit buffer
    const uint8_t* in_end_ = nullptr;   // end of the valid bytes of the region `in_` walks
    const uint8_t* tail_src_ = nullptr; // where the padded private copy of the last bytes starts (in the source)
    std::vector<uint8_t> tail_;
    bool in_tail_ = false;
    uint64_t consumed_before_tail_ = 0;
    uint64_t bitbuf_ = 0;
    unsigned bitcnt_ = 0;

    // output: [history kWindow][chunk kChunk][slack]
    std::vector<uint8_t> out_;
    size_t produced_ = 0;           // bytes of the chunk produced by the previous call (history source)
    uint64_t total_out_ = 0;

    // member / block state
    Phase phase_ = Phase::kHeader;
    bool last_block_ = false;
    uint32_t stored_left_ = 0;
    uint32_t crc_ = 0;
    uint64_t member_out_ = 0;
    uint64_t members_ = 0;
    // a match interrupted by the end of the output chunk
    uint32_t pending_len_ = 0, pending_off_ = 0;

    std::vector<uint32_t> litlen_, dist_;
    bool fixed_built_ = false;
    std::vector<uint32_t> fixed_litlen_, fixed_dist_;
    const uint32_t* cur_litlen_ = nullptr;
    const uint32_t* cur_dist_ = nullptr;
    std::string err_;
};

}  // namespace hasthost
