// inflate_par.h -- gzip decoding of ONE stream on several threads.
//
// A DEFLATE stream is serial: a block can only be found by decoding everything before it, and its
// matches copy from the 32 KiB that precede it.  FASTQ text lets both be worked around (the idea of
// Kerbiriou & Chikhi's pugz, re-derived here):
//   1. the compressed file is cut into chunks at arbitrary byte offsets; inside each chunk the first
//      dynamic-Huffman block header is found by trying every bit position: a header must pass every
//      validity rule of RFC 1951 (complete pre-code, complete literal/length and distance codes, an
//      end-of-block code) and the symbols that follow must decode to printable text;
//   2. every chunk is decoded from its block to the next chunk's block without knowing the window
//      before it: output symbols are 16 bits wide, a value >= 0x8000 means "byte number v - 0x8000 of
//      the 32 KiB before this chunk", and the output buffer simply starts with those 32768
//      placeholders, so matches need no special case;
//   3. in stream order the last 32 KiB of every chunk are resolved against the previous chunk's, which
//      gives every chunk its window; then all chunks are resolved to bytes in parallel, CRC-32 is
//      computed per member segment and stitched with crc32_combine, and CRC / ISIZE of every gzip
//      member are checked exactly as in the serial decoder.
// A block start is only trusted if the previous chunk's decoder arrives at exactly that bit position on
// a block boundary; a candidate it runs past is dropped and its chunk is decoded by the predecessor.
// Output bytes are identical to GzipInflater's (tests/test_inflate.py runs both on every case).
#pragma once
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace hasthost {

class ParallelGzip {
public:
    explicit ParallelGzip(int threads, size_t chunk_bytes = 1u << 20);
    ~ParallelGzip();
    ParallelGzip(const ParallelGzip&) = delete;
    ParallelGzip& operator=(const ParallelGzip&) = delete;

    std::string open(const std::string& path);             // "" or an error message
    void open_memory(const uint8_t* data, size_t n);        // tests; the buffer must stay alive
    bool is_gzip() const { return n_in_ >= 2 && in_[0] == 0x1F && in_[1] == 0x8B; }
    // Next piece of decompressed bytes (valid until the next call).  false at the end or on error.
    bool next(const uint8_t** data, size_t* len);
    const std::string& error() const { return err_; }

    struct Stats { uint64_t chunks = 0, starts_found = 0, starts_dropped = 0, batches = 0; };
    Stats stats() const { return stats_; }
    struct Chunk;                                           // one piece of the stream (inflate_par.cpp)

private:
    void start();
    void run();                                             // coordinator thread
    void push(std::vector<uint8_t>&& piece);
    void finish(const std::string& err);

    int threads_;
    size_t chunk_bytes_;
    int fd_ = -1;
    const uint8_t* map_ = nullptr;  size_t map_len_ = 0;
    const uint8_t* in_ = nullptr;   size_t n_in_ = 0;

    std::thread coord_;
    std::mutex mu_;
    std::condition_variable cv_out_, cv_room_;
    std::deque<std::vector<uint8_t>> ready_;
    size_t ready_bytes_ = 0;
    bool done_ = false, stop_ = false;
    std::vector<uint8_t> current_;
    std::vector<std::vector<uint8_t>> spare_;   // output buffers handed back by next(), reused by the workers
    std::string err_, err_pending_;
    Stats stats_;
};

}  // namespace hasthost
