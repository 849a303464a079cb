// hapcall.cpp -- the per-barcode haplotype call and the output table.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>

#include "host.h"

namespace hasthost {

// getHap, classify.cpp:66-86.  The reference keeps map<int,int> per barcode whose
// keys 0 / 1 exist only once a positive vote was added (classify.cpp:203-206),
// so "key present" is exactly "count > 0".  The ratio is formed in IEEE double in
// the reference's operation order: divide by the set size, then multiply by the
// weight (:70-73).
int get_hap(const std::string& barcode, int c0, int c1, uint64_t n0, uint64_t n1, double w0, double w1) {
    if (barcode == "0_0_0" || barcode == "0_0" || barcode == "0") return -1;
    const bool has0 = c0 > 0, has1 = c1 > 0;
    if (has0 && has1) {
        double df0 = double(c0) / double(n0);
        double df1 = double(c1) / double(n1);
        df0 *= w0;
        df1 *= w1;
        if (df0 > df1) return 0;
        if (df1 > df0) return 1;
        return -1;
    } else if (has0) {
        return 0;
    } else if (has1) {
        return 1;
    }
    return -1;
}

static inline char* put_int(char* p, long v) {
    char tmp[24];
    int n = 0;
    unsigned long u = v < 0 ? (unsigned long)(-v) : (unsigned long)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

// ---- output order ---------------------------------------------------------------------------
// std::map order of the barcode strings (classify.cpp:50,93-102): bytewise lexicographic.  A human-scale run
// prints 20-50 M barcodes, and one std::sort over string indices takes longer than classifying the reads
// (6 s for 5 M names), so: strip the prefix all names share, take the next 8 bytes of every name as a
// big-endian integer (zero padded: a shorter name sorts first, as it must), split the (key, id) pairs into
// ranges by sampled splitters, sort the ranges on all cores; names are compared only where keys tie.
namespace {
struct KeyId { uint64_t key; uint32_t id; };

inline uint64_t prefix_key(const std::string& s, size_t skip) {
    unsigned char b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (s.size() > skip) memcpy(b, s.data() + skip, std::min<size_t>(8, s.size() - skip));
    uint64_t k = 0;
    for (int i = 0; i < 8; ++i) k = (k << 8) | b[i];
    return k;
}

template <class F>
void parallel_ranges(size_t n, unsigned threads, F f) {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < threads; ++t)
        th.emplace_back([=]() { f(t, n * t / threads, n * (t + 1) / threads); });
    for (auto& x : th) x.join();
}

void sorted_order(const std::vector<std::string>& names, std::vector<uint32_t>& order) {
    const size_t n = names.size();
    order.resize(n);
    unsigned threads = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("HAST_SORT_THREADS")) threads = (unsigned)std::max(1, atoi(e));
    if (n < 200000 || threads == 1) {
        std::iota(order.begin(), order.end(), 0u);
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return names[a] < names[b]; });
        return;
    }
    // bytes that every name starts with carry no order
    size_t lcp = names[0].size();
    for (size_t i = 1; i < n && lcp; ++i) {
        const std::string& s = names[i];
        size_t j = 0;
        const size_t m = std::min(lcp, s.size());
        while (j < m && s[j] == names[0][j]) ++j;
        lcp = j;
    }
    std::vector<KeyId> keys(n), sorted(n);
    parallel_ranges(n, threads, [&](unsigned, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) keys[i] = KeyId{prefix_key(names[i], lcp), (uint32_t)i};
    });
    // splitters from a regular sample; equal keys always land in the same range
    const size_t n_ranges = (size_t)threads * 8;
    std::vector<uint64_t> sample;
    const size_t step = std::max<size_t>(1, n / (n_ranges * 64));
    for (size_t i = 0; i < n; i += step) sample.push_back(keys[i].key);
    std::sort(sample.begin(), sample.end());
    std::vector<uint64_t> split;
    for (size_t r = 1; r < n_ranges; ++r) split.push_back(sample[r * sample.size() / n_ranges]);
    split.erase(std::unique(split.begin(), split.end()), split.end());
    const size_t R = split.size() + 1;
    auto range_of = [&](uint64_t k) { return (size_t)(std::upper_bound(split.begin(), split.end(), k) - split.begin()); };
    std::vector<size_t> cnt((size_t)threads * R, 0);
    parallel_ranges(n, threads, [&](unsigned t, size_t lo, size_t hi) {
        size_t* c = &cnt[(size_t)t * R];
        for (size_t i = lo; i < hi; ++i) ++c[range_of(keys[i].key)];
    });
    std::vector<size_t> start(R + 1, 0), at((size_t)threads * R);
    size_t run = 0;
    for (size_t r = 0; r < R; ++r) {
        start[r] = run;
        for (unsigned t = 0; t < threads; ++t) { at[(size_t)t * R + r] = run; run += cnt[(size_t)t * R + r]; }
    }
    start[R] = run;
    parallel_ranges(n, threads, [&](unsigned t, size_t lo, size_t hi) {
        size_t* a = &at[(size_t)t * R];
        for (size_t i = lo; i < hi; ++i) sorted[a[range_of(keys[i].key)]++] = keys[i];
    });
    std::atomic<size_t> next{0};
    parallel_ranges(threads, threads, [&](unsigned, size_t, size_t) {
        for (size_t r; (r = next.fetch_add(1)) < R;)
            std::sort(sorted.begin() + (ptrdiff_t)start[r], sorted.begin() + (ptrdiff_t)start[r + 1],
                      [&](const KeyId& a, const KeyId& b) { return a.key != b.key ? a.key < b.key : names[a.id] < names[b.id]; });
    });
    parallel_ranges(n, threads, [&](unsigned, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) order[i] = sorted[i].id;
    });
}
}  // namespace

// printBarcodeInfos, classify.cpp:93-102: every barcode seen, in std::map order
// (bytewise lexicographic, e.g. "0_0_0" < "10_1_1" < "1_2_3"), one line
// barcode \t hap \t count0 \t count1.
void print_table(FILE* out, const std::vector<std::string>& names, const int32_t* counts, uint64_t n0,
                 uint64_t n1, double w0, double w1, std::vector<uint32_t>* order_out, std::vector<int8_t>* hap_out) {
    std::vector<uint32_t> order;
    sorted_order(names, order);
    // lines are formatted on all cores, a round of 4 M barcodes at a time, and written in order
    const size_t n = order.size();
    if (hap_out) hap_out->resize(n);
    unsigned threads = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("HAST_SORT_THREADS")) threads = (unsigned)std::max(1, atoi(e));
    if (n < 200000) threads = 1;
    std::vector<std::vector<char>> bufs(threads);
    std::vector<size_t> lens(threads, 0);
    constexpr size_t kRound = (size_t)4 << 20;
    for (size_t r0 = 0; r0 < n; r0 += kRound) {
        const size_t rn = std::min(kRound, n - r0);
        auto format = [&](unsigned t, size_t lo, size_t hi) {
            std::vector<char>& buf = bufs[t];
            size_t need = 0;
            for (size_t i = lo; i < hi; ++i) need += names[order[r0 + i]].size() + 40;
            if (buf.size() < need) buf.resize(need);
            char* p = buf.data();
            for (size_t i = lo; i < hi; ++i) {
                const uint32_t id = order[r0 + i];
                const std::string& nm = names[id];
                const int c0 = counts[2 * (size_t)id], c1 = counts[2 * (size_t)id + 1];
                memcpy(p, nm.data(), nm.size());
                p += nm.size();
                *p++ = '\t';
                const int hap = get_hap(nm, c0, c1, n0, n1, w0, w1);
                if (hap_out) (*hap_out)[r0 + i] = (int8_t)hap;
                p = put_int(p, hap);
                *p++ = '\t';
                p = put_int(p, c0);
                *p++ = '\t';
                p = put_int(p, c1);
                *p++ = '\n';
            }
            lens[t] = (size_t)(p - buf.data());
        };
        if (threads == 1) format(0, 0, rn);
        else parallel_ranges(rn, threads, format);
        for (unsigned t = 0; t < threads; ++t)
            if (lens[t]) fwrite(bufs[t].data(), 1, lens[t], out);
        std::fill(lens.begin(), lens.end(), (size_t)0);
    }
    fflush(out);
    if (order_out) order_out->swap(order);
}

}  // namespace hasthost
