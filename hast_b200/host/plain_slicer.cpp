// plain_slicer.cpp -- see plain_slicer.h
#include "plain_slicer.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace hasthost {

PlainSlicer::~PlainSlicer() {
    if (map_) munmap(const_cast<char*>(map_), (size_t)size_);
    if (fd_ >= 0) close(fd_);
}

std::string PlainSlicer::open(const std::string& path, size_t slice_bytes) {
    path_ = path;
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) return "cannot open " + path + ": " + strerror(errno);
    struct stat sb;
    if (fstat(fd_, &sb) != 0 || !S_ISREG(sb.st_mode)) return "";       // not sliceable: usable() stays false
    size_ = (uint64_t)sb.st_size;
    if (size_) {
        void* m = mmap(nullptr, (size_t)size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (m == MAP_FAILED) return "";                                // e.g. a file system without mmap: sequential reader
        map_ = static_cast<const char*>(m);
#ifdef MADV_SEQUENTIAL
        madvise(m, (size_t)size_, MADV_SEQUENTIAL);
#endif
    }
    slice_ = std::max<size_t>(slice_bytes, 1);
    n_slices_ = (size_ + slice_ - 1) / slice_;
    hand_.reset(new Hand[n_slices_ + 1]);
    hand_[0].ready.store(1, std::memory_order_release);
    return "";
}

bool PlainSlicer::next(TextBlock& blk, bool* first, uint64_t* index) {
    const uint64_t i = next_.fetch_add(1, std::memory_order_relaxed);
    if (first) *first = i == 0;
    if (index) *index = i;
    if (i >= n_slices_) return false;
    const uint64_t a = i * slice_, b = std::min(size_, a + slice_);
    const size_t n = (size_t)(b - a);
    const char* buf = map_ + a;
    blk.view = nullptr;
    blk.begin = 0;
    blk.len = 0;
    blk.has_nl = false;
    blk.last_of_file = i + 1 == n_slices_;
#ifdef MADV_WILLNEED
    if (i + 2 < n_slices_) madvise(const_cast<char*>(map_ + a + 2 * slice_), (size_t)std::min(slice_, size_ - a - 2 * slice_), MADV_WILLNEED);
#endif
    constexpr size_t kFront = 8;                    // index slots kept free for the newlines of a straddling record
    const uint64_t c = newline_index(buf, n, blk.nl, kFront);
    uint32_t* nl = blk.nl.data() + kFront;
    // ---- hand-over: in slice order, a few instructions per slice ----
    Hand& in = hand_[i];
    for (unsigned spin = 0; !in.ready.load(std::memory_order_acquire); ++spin)
        if (spin > 64) std::this_thread::yield();
    const uint64_t before = in.newlines_before, prev_cut = in.cut;
    const uint64_t total = before + c;
    uint64_t cut = prev_cut, used = 0;              // used: newlines of this slice that lie inside the block
    if (blk.last_of_file) {
        cut = size_;                                // whatever is left, a partial record included (classify.cpp:257)
        used = c;
    } else {
        const uint64_t leftover = total & 3u;       // newlines after the last one that closes a whole record
        if (c > leftover) {
            used = c - leftover;
            cut = a + nl[used - 1] + 1;
        }
    }
    Hand& out = hand_[i + 1];
    out.newlines_before = total;
    out.cut = cut;
    out.ready.store(1, std::memory_order_release);
    if (cut == prev_cut) return true;               // no record ends in this slice: the text rides with a later one
    // ---- the block is [prev_cut, cut) of the mapping; its index = newlines of the straddling head + ours ----
    const uint64_t head = a - prev_cut;
    blk.view = map_ + prev_cut;
    blk.len = (size_t)(cut - prev_cut);
    if (blk.len > 0xFFFFFFFFull) return true;       // parse_block reports it
    if (head == 0) {
        blk.nl_begin = kFront;
        blk.nl_count = (size_t)used;
        blk.has_nl = true;
        return true;
    }
    size_t n_head = 0;
    uint32_t head_nl[kFront];
    if (head <= (1u << 20)) {
        const char* h = map_ + prev_cut;
        for (uint64_t j = 0; j < head && n_head <= kFront; ++j)
            if (h[j] == '\n') { if (n_head < kFront) head_nl[n_head] = (uint32_t)j; ++n_head; }
    } else {
        n_head = kFront + 1;
    }
    if (n_head > kFront) {                          // a run of records longer than a slice: index it afresh
        blk.nl_count = newline_index(blk.view, blk.len, blk.nl, 0);
        blk.nl_begin = 0;
        blk.has_nl = true;
        return true;
    }
    const uint32_t bias = (uint32_t)head;
    for (uint64_t j = 0; j < used; ++j) nl[j] += bias;
    for (size_t j = 0; j < n_head; ++j) blk.nl[kFront - n_head + j] = head_nl[j];
    blk.nl_begin = kFront - n_head;
    blk.nl_count = n_head + (size_t)used;
    blk.has_nl = true;
    return true;
}

}  // namespace hasthost
