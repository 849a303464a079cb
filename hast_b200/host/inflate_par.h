// inflate_par.h -- gzip decoding of ONE stream on several threads.
//
// A DEFLATE stream is serial: a block can only be found by decoding everything before it, and its
// matches copy from the 32 KiB that precede it.  FASTQ text lets both be worked around (the idea of
// Kerbiriou & Chikhi's pugz, re-derived here):
//   1. the compressed file is cut into chunks at arbitrary byte offsets; inside each chunk the first
//      dynamic-Huffman block header is found by trying every bit position: a header must pass every
//      validity rule of RFC 1951 (complete pre-code, complete literal/length and distance codes, an
//      end-of-block code) and the symbols that follow must decode to printable text;
//   2. every chunk is ENTROPY-decoded from its block to the next chunk's block on a worker thread: the
//      Huffman codes are resolved into a stream of 32-bit tokens (literal-run length, match length,
//      match distance) plus the literal bytes.  No output byte is produced, so the 32 KiB window in front
//      of the chunk -- which nobody knows yet -- is not needed: this is the expensive part of inflating
//      (bit buffer, table lookups, ~1.5 bits per output byte) and it is what runs in parallel;
//   3. one thread per stream replays the tokens in stream order into bytes -- plain LZ77 copying, several
//      times faster than entropy decoding -- so every match reads real history; the workers meanwhile
//      decode the next batch of chunks.  CRC-32 is computed per member segment on the workers
//      (carry-less multiply) and stitched with crc32_combine; CRC / ISIZE of every gzip member are
//      checked exactly as in the serial decoder before the bytes are handed out.
// (The first version decoded to 16-bit symbols with window placeholders, pugz-style, and resolved them in a
// second parallel pass: 2.3-2.7x the CPU of the serial decoder per byte.  Tokens + serial replay cost ~1.2x.)
// A block start is only trusted if the previous chunk's decoder arrives at exactly that bit position on
// a block boundary; a candidate it runs past is dropped and its chunk is decoded by the predecessor.
// Output bytes are identical to GzipInflater's (tests/test_inflate.py runs both on every case).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <functional>
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
    // The same without the copy a caller would otherwise make: the piece's buffer changes hands.  *data points at
    // *len decoded bytes with at least kFront writable bytes in front of them (room for the tail of the previous
    // piece, so that a record straddling two pieces can be made contiguous in place); *hold keeps the buffer alive
    // and hands it back to this decoder's pool -- also after the decoder itself is gone -- when it is released.
    static constexpr size_t kFront = 32768;
    bool next_owned(uint8_t** data, size_t* len, std::shared_ptr<void>* hold);
    const std::string& error() const { return err_; }

    struct Stats { uint64_t chunks = 0, starts_found = 0, starts_dropped = 0, batches = 0; };
    Stats stats() const { return stats_; }
    // page-aligned, huge-page-advised output buffer; recycled between chunks
    struct Buffer {
        uint8_t* p = nullptr;
        size_t cap = 0;
        Buffer() = default;
        explicit Buffer(size_t n);
        ~Buffer();
        Buffer(Buffer&& o) noexcept;
        Buffer& operator=(Buffer&& o) noexcept;
        Buffer(const Buffer&) = delete;
        Buffer& operator=(const Buffer&) = delete;
    private:
        void* map = nullptr;
        size_t cap_map = 0;
    };
    struct Chunk;                                           // one piece of the stream (inflate_par.cpp)
    struct Batch;
    struct Pool;

private:
    struct Piece { Buffer buf; size_t off = 0, len = 0; };   // decoded bytes: buf.p[off, off + len)
    Buffer take_buffer(size_t need);
    void recycle(Buffer&& b);
    void start();
    void run();                                             // coordinator thread: sequencing, replay, checks
    void launch(Batch& b, Pool& pool, uint64_t next_start, bool at_file_start);
    void push(Piece&& piece);
    void finish(const std::string& err);

    int threads_;
    size_t chunk_bytes_;
    int fd_ = -1;
    const uint8_t* map_ = nullptr;  size_t map_len_ = 0;
    const uint8_t* in_ = nullptr;   size_t n_in_ = 0;

    std::thread coord_;
    std::mutex mu_;
    std::condition_variable cv_out_, cv_room_;
    std::deque<Piece> ready_;
    size_t ready_bytes_ = 0;
    bool done_ = false, stop_ = false;
    Piece current_;
    uint64_t member_out_checked_ = 0;
    std::atomic<uint64_t> decode_ns_{0};       // time the workers spent in entropy decoding (HAST_PAR_PROF)
    struct BufferPool { std::mutex mu; std::vector<Buffer> spare; };
    std::shared_ptr<BufferPool> pool_ = std::make_shared<BufferPool>();   // output buffers handed back, reused by the replay
    std::string err_, err_pending_;
    Stats stats_;
};

}  // namespace hasthost
