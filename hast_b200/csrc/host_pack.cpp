// host_pack.cpp -- the packing step of bin/classify's parser (host/parser.cpp pack_append) as a bulk operation for
// callers of the C ABI who hold ASCII batches: hast_submit_batch with the option "host_pack_threads" packs the batch
// to 2 bits on the host's cores and sends a quarter of the bytes over PCIe.  Since the reads of a batch lie back to
// back, the packed stream of the batch is simply the packed `bases` array: no per-read work except finding the 'N's.
// Bit-identical to capi.pack_bases / the parser (tests/test_host.py::test_bulk_packer_equals_reference_packing).
#include "host_pack.h"

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace hastpack {

namespace {

inline void mark_n(uint64_t g, const uint32_t* read_off, uint32_t n_reads, uint32_t* has_n) {
    // read r with read_off[r] <= g < read_off[r + 1]
    const uint32_t* it = std::upper_bound(read_off, read_off + n_reads + 1, (uint32_t)g);
    if (it == read_off) return;
    const uint32_t r = (uint32_t)(it - read_off) - 1;
    if (r < n_reads) __atomic_fetch_or(&has_n[r >> 5], 1u << (r & 31u), __ATOMIC_RELAXED);
}

inline uint32_t pack16_scalar(const uint8_t* p, size_t n) {      // n <= 16 bases, zero-padded
    uint32_t w = 0;
    for (size_t i = 0; i < n; ++i) w |= (uint32_t)((p[i] >> 1) & 3u) << (30 - 2 * i);
    return w;
}

// words [w_lo, w_hi) of the stream; only the very last word of the batch may be partial
__attribute__((target("avx2"))) void pack_range_avx2(const uint8_t* bases, uint64_t n_bases, uint64_t w_lo, uint64_t w_hi,
                                                     const uint32_t* read_off, uint32_t n_reads, uint32_t* words,
                                                     uint32_t* has_n) {
    const __m256i three = _mm256_set1_epi8(3), w1 = _mm256_set1_epi16(0x0104), w2 = _mm256_set1_epi32(0x00010010);
    const __m256i gather = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i big_n = _mm256_set1_epi8('N');
    uint64_t w = w_lo;
    for (; w + 2 <= w_hi && (w + 2) * 16 <= n_bases; w += 2) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(bases + w * 16));
        const uint32_t nm = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, big_n));
        if (nm) {
            uint32_t m = nm;
            while (m) { mark_n(w * 16 + (uint64_t)__builtin_ctz(m), read_off, n_reads, has_n); m &= m - 1; }
        }
        const __m256i codes = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
        const __m256i quads = _mm256_madd_epi16(_mm256_maddubs_epi16(codes, w1), w2);   // one byte per 4 bases
        const __m256i packed = _mm256_shuffle_epi8(quads, gather);                      // first base on top
        words[w] = (uint32_t)_mm256_cvtsi256_si32(packed);
        words[w + 1] = (uint32_t)_mm256_extract_epi32(packed, 4);
    }
    for (; w < w_hi; ++w) {
        const uint64_t g = w * 16;
        const size_t n = (size_t)std::min<uint64_t>(16, n_bases - g);
        for (size_t i = 0; i < n; ++i)
            if (bases[g + i] == 'N') mark_n(g + i, read_off, n_reads, has_n);
        words[w] = pack16_scalar(bases + g, n);
    }
}

void pack_range_scalar(const uint8_t* bases, uint64_t n_bases, uint64_t w_lo, uint64_t w_hi, const uint32_t* read_off,
                       uint32_t n_reads, uint32_t* words, uint32_t* has_n) {
    for (uint64_t w = w_lo; w < w_hi; ++w) {
        const uint64_t g = w * 16;
        const size_t n = (size_t)std::min<uint64_t>(16, n_bases - g);
        for (size_t i = 0; i < n; ++i)
            if (bases[g + i] == 'N') mark_n(g + i, read_off, n_reads, has_n);
        words[w] = pack16_scalar(bases + g, n);
    }
}

}  // namespace

class Pool {
public:
    explicit Pool(int threads) : n_(std::max(threads, 1)) {
        for (int i = 1; i < n_; ++i) th_.emplace_back([this, i] { run(i); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    // f(part, n_parts) on every thread of the pool, the caller's included; returns when all are done
    template <class F>
    void parallel(F&& f) {
        job_ = [&f](int i, int n) { f(i, n); };
        {
            std::lock_guard<std::mutex> lk(mu_);
            left_ = n_ - 1;
            ++gen_;
        }
        cv_.notify_all();
        f(0, n_);
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return left_ == 0; });
    }
    int size() const { return n_; }
private:
    void run(int i) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            job_(i, n_);
            std::lock_guard<std::mutex> lk(mu_);
            if (--left_ == 0) cv_done_.notify_all();
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, cv_done_;
    std::function<void(int, int)> job_;
    uint64_t gen_ = 0;
    int left_ = 0;
    bool stop_ = false;
};

Pool* pool_create(int threads) { return new Pool(threads); }
void pool_destroy(Pool* p) { delete p; }

void pack_batch(Pool* p, const uint8_t* bases, uint64_t n_bases, const uint32_t* read_off, uint32_t n_reads,
                uint32_t* words_out, uint32_t* has_n_out) {
    const uint64_t n_words = (n_bases + 15) / 16;
    memset(has_n_out, 0, ((size_t)n_reads + 31) / 32 * 4);
    static const bool avx2 = __builtin_cpu_supports("avx2");
    auto part = [&](int i, int n) {
        // ranges of whole words, a multiple of two so that every range but the last runs the 32-base loop throughout
        const uint64_t per = ((n_words + (uint64_t)n - 1) / (uint64_t)n + 1) & ~1ull;
        const uint64_t lo = std::min(n_words, per * (uint64_t)i), hi = std::min(n_words, lo + per);
        if (hi <= lo) return;
        if (avx2) pack_range_avx2(bases, n_bases, lo, hi, read_off, n_reads, words_out, has_n_out);
        else pack_range_scalar(bases, n_bases, lo, hi, read_off, n_reads, words_out, has_n_out);
    };
    if (p && p->size() > 1 && n_words > 4096) p->parallel(part);
    else part(0, 1);
}

}  // namespace hastpack
