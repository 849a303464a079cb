// plain_slicer.h -- one plain FASTQ file read by MANY threads at once, framed exactly.
//
// The reference frames records by counting lines: line 4n is a header whatever it starts with
// (processFastq, classify.cpp:257-269; no '@' / '+' validation).  So a thread that lands in the middle of a
// file cannot look for an "@" line -- a quality string may begin with '@' -- it has to know how many
// newlines precede its range.  The file is mapped and cut into fixed slices; threads claim slices in order,
// index the newlines of their slice in parallel (one AVX2 pass, which is also the pass that faults the pages
// in) and pass (newlines before the slice, end of the last whole record) down the chain.  The hand-over is a
// handful of instructions per slice, so reading, framing and parsing all scale with the threads; the blocks
// that come out are views into the mapping (no copy) holding the same whole-record runs a serial getline
// loop would deliver, together with their newline index, which the parser reuses.
#pragma once
#include <atomic>
#include <cstdint>
#include <memory>
#include <string>

#include "host.h"

namespace hasthost {

class PlainSlicer {
public:
    PlainSlicer() = default;
    ~PlainSlicer();
    PlainSlicer(const PlainSlicer&) = delete;
    // "" or an error message.  usable() is false for anything that is not a mappable regular file (pipe,
    // stdin, tty): those go through the sequential FastqSource.
    std::string open(const std::string& path, size_t slice_bytes);
    bool usable() const { return n_slices_ != kNotRegular; }
    uint64_t size() const { return size_; }
    const std::string& path() const { return path_; }
    // Claims the next slice and points blk (view / len / nl) at the whole records that END in it; blk.len may be
    // 0 (no record ends inside the slice).  Returns false when every slice has been claimed.  Thread safe.
    // first is set for the call that claimed slice 0.
    // index (optional) receives the slice number: blocks of consecutive indices are consecutive in the file.
    bool next(TextBlock& blk, bool* first = nullptr, uint64_t* index = nullptr);
private:
    static constexpr uint64_t kNotRegular = ~0ull;
    struct Hand {                         // what slice i needs from slices 0..i-1
        std::atomic<uint32_t> ready{0};
        uint64_t newlines_before = 0;     // '\n' in [0, start of slice i)
        uint64_t cut = 0;                 // offset just past the last newline in [0, start) whose ordinal is a multiple of 4
    };
    std::string path_;
    int fd_ = -1;
    const char* map_ = nullptr;
    uint64_t size_ = 0, slice_ = 0, n_slices_ = kNotRegular;
    std::atomic<uint64_t> next_{0};
    std::unique_ptr<Hand[]> hand_;
};

}  // namespace hasthost
