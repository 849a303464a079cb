// fastq_source.h -- sequential block reader over one FASTQ file (plain or gzip).
#pragma once
#include <zlib.h>

#include <string>
#include <vector>

#include "host.h"
#include "inflate.h"
#include "inflate_par.h"

namespace hasthost {

// Replaces processFastq's istream + getline loop (classify.cpp:238-269) and the
// vendored gzstream (gzstream.C:78-101, 299-byte gzread calls): large reads /
// inflates into recycled buffers, cut so that every block holds whole four-line
// records.  ".gz" is decided by the file-name suffix alone (classify.cpp:245-250);
// gzip members are decoded by GzipInflater (inflate.h; concatenated members walked like
// zlib's gzread does); a ".gz" file that holds plain text is passed through by zlib's
// gzread, exactly like the reference's igzstream.  HAST_ZLIB=1 forces zlib for gzip too.
class FastqSource {
public:
    FastqSource() = default;
    ~FastqSource();
    // returns "" or an error message
    // inflate_threads > 1: a gzip file is decoded by ParallelGzip (inflate_par.h) on that many threads
    std::string open(const std::string& path, int inflate_threads = 1);
    // Fills blk with the next run of whole records (blk.len > 0) and returns true;
    // returns false at end of file.  The block that ends the file has last_of_file set
    // and may end in a partial record / unterminated line.  err is set on I/O failure.
    bool next(TextBlock& blk, size_t target_bytes, std::string& err);
    uint64_t bytes_out() const { return bytes_out_; }
private:
    size_t raw_read(char* dst, size_t n, std::string& err);
    bool next_from_pieces(TextBlock& blk, std::string& err);
    std::string path_;
    gzFile gz_ = nullptr;
    std::unique_ptr<GzipInflater> inf_;
    std::unique_ptr<ParallelGzip> pinf_;
    const uint8_t* chunk_ = nullptr;   // unread part of the inflater's current chunk
    size_t chunk_left_ = 0;
    int fd_ = -1;
    bool eof_ = false;
    std::vector<char> carry_;      // bytes after the last record boundary handed out
    uint32_t carry_lines_ = 0;     // complete lines inside carry_ (0..3)
    uint64_t bytes_out_ = 0;
};

// number of '\n' in [p, p+n)
size_t count_newlines(const char* p, size_t n);

}  // namespace hasthost
