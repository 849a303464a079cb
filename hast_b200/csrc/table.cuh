// table.cuh -- the parent-unique k-mer table (replaces the two
// std::unordered_set<Kmer> of classify.cpp:27).
//
// Layout in HBM: 2^b buckets of 32 bytes (= one DRAM/L2 sector), four 8-byte
// slots per bucket, so a lookup -- hit or miss -- costs ONE 32-byte sector read
// (a single LDG.E.256 on sm_100a) unless its home bucket overflowed.
//
//   h      = bijective mix of the canonical k-mer inside the 2k-bit domain
//   bucket = top b bits of h          (home bucket)
//   rem    = low 2k-b bits of h       (quotient: with the home bucket it
//                                      identifies the k-mer exactly, which is
//                                      what makes room for k = 32)
//   slot   = rem << 7 | disp << 3 | tag << 1 | ovf
//              disp : 0..15, distance (in buckets) from the home bucket
//              tag  : bit0 = member of hap0 set, bit1 = member of hap1 set
//                     (a k-mer may be in both: classify.cpp:195-202 scores the
//                     two sets independently); tag 0 = empty or erased
//              ovf  : meaningful in slot 0 only: some insertion passed through
//                     this full bucket, so a miss must look at the next one
//
// An all-zero slot is empty; it can only compare equal to (rem 0, disp 0) and
// then yields tag 0, which is the correct answer for an absent k-mer.
#pragma once
#include <cstdint>
#include "kmer.cuh"

namespace hast {

constexpr int kSlotsPerBucket = 4;
constexpr int kMaxDisp = 15;
constexpr int kRemShift = 7;
constexpr uint64_t kHashC1 = 0x9E3779B97F4A7C15ull;
constexpr uint64_t kHashC2 = 0xD6E8FEB86659FD93ull;

struct TableView {
    uint64_t* slots;        // n_buckets * 4
    uint64_t kmask;         // 2k low bits
    uint64_t rem_mask;      // (1 << rem_bits) - 1
    uint32_t bucket_mask;   // n_buckets - 1
    int32_t k;
    int32_t rem_bits;       // 2k - b, 0..57
    // Bloom pre-filter over every key ever inserted (fused.cuh): 2^(32-filt_shift)
    // 64-bit words, two bits per key, one in each half of the key's word
    uint64_t* filt;
    uint32_t filt_shift;
    // 0: the filter word is picked by a hash of the k-mer itself; m > 0: by the k-mer's MINIMIZER of
    // length m (below), so that neighbouring positions of a read mostly share one filter word
    uint32_t filt_m;
};

// multiply / xor-shift / multiply, every step a bijection of Z/2^(2k)
__host__ __device__ __forceinline__ uint64_t table_hash(uint64_t key, int k, uint64_t kmask) {
    uint64_t h = (key * kHashC1) & kmask;
    h ^= h >> k;
    h = (h * kHashC2) & kmask;
    return h;
}

// Pre-filter hash of a canonical k-mer: `word` (32 well-mixed bits, the top ones pick
// the 64-bit filter word) and `bits` (an independent multiplicative hash whose top
// 5 + 5 bits pick one bit in the low and one in the high half of that word).  Two
// separate 32-bit states, so that large filters (> 2^22 words) do not run out of
// hash entropy.
struct FilterHash { uint32_t word, bits; };
__host__ __device__ __forceinline__ FilterHash filter_hash(uint64_t canon) {
    const uint32_t lo = (uint32_t)canon, hi = (uint32_t)(canon >> 32);
    uint32_t m = lo * 0x9E3779B1u + hi * 0x85EBCA77u;
    m ^= m >> 15;
    m *= 0xC2B2AE3Du;
    m ^= m >> 13;
    FilterHash f;
    f.word = m;
    f.bits = lo * 0x27D4EB2Fu + hi * 0x165667B1u;
    return f;
}
__host__ __device__ __forceinline__ uint32_t filter_bit_lo(FilterHash f) { return f.bits >> 27; }
__host__ __device__ __forceinline__ uint32_t filter_bit_hi(FilterHash f) { return (f.bits >> 22) & 31u; }
__host__ __device__ __forceinline__ uint64_t filter_bits(FilterHash f) {
    return (1ull << filter_bit_lo(f)) | (1ull << (32u + filter_bit_hi(f)));
}

// ---- minimizer-addressed pre-filter ------------------------------------------------------
// The k-mer's filter WORD is chosen by its minimizer: over the k-m+1 m-mers s inside the k-mer,
// the smallest value of  mini_hash(min(s, revcomp(s))).  The set of canonical m-mers is the same
// for a k-mer and its reverse complement, so the choice is orientation-free, and consecutive
// k-mers of a sequence share their minimizer for (k-m+2)/2 positions on average: a thread that
// walks 16 neighbouring positions needs ~0.3 filter loads per position instead of one (fused.cuh,
// the MINI sweep).  The two bits inside the word still come from the k-mer's own hash
// (filter_hash().bits), so the filter stays a per-k-mer membership test.
// Minimizer lengths: long enough that a human genome's minimizers far outnumber the filter words.
constexpr uint32_t kMiniC = 0x9E3779B1u, kMiniD = 0x7F4A7C15u, kMiniC2 = 0x85EBCA77u;
__host__ __device__ __forceinline__ constexpr int mini_len(int k) {
    return (k == 21 || k == 25 || k == 31) ? 16 : 0;      // 0: no minimizer sweep for this k
}
// Bit selectors (5 + 5 bits) of a k-mer in the minimizer-addressed filter, k >= 16: a hash of the first and
// last 16 bases of both strands that is symmetric in the strands, so the sweep needs neither the canonical
// k-mer nor any 64-bit arithmetic (fwd / rc right-aligned, 2k bits)
constexpr uint32_t kSelC1 = 0x27D4EB2Fu, kSelC2 = 0x165667B1u;
__host__ __device__ __forceinline__ uint32_t mini_sel(uint64_t fwd, uint64_t rc, int k) {
    const uint32_t ff = (uint32_t)(fwd >> (2 * k - 32)), fl = (uint32_t)fwd;
    const uint32_t rf = (uint32_t)(rc >> (2 * k - 32)), rl = (uint32_t)rc;
    return (ff * kSelC1 + fl * kSelC2 + rf * kSelC1 + rl * kSelC2) >> 22;
}
__host__ __device__ __forceinline__ uint64_t mini_bits(uint32_t sel) {
    return (1ull << ((sel >> 5) & 31u)) | (1ull << (32u + (sel & 31u)));
}
__host__ __device__ __forceinline__ uint32_t mini_hash(uint32_t canon_mmer) { return canon_mmer * kMiniC + kMiniD; }
// top bits pick the filter word (the minimum of several hashes is biased towards 0: remix it)
__host__ __device__ __forceinline__ uint32_t mini_word(uint32_t min_hash) { return min_hash * kMiniC2; }
// minimizer hash of one k-mer given both strands right-aligned (fwd, rc = its reverse complement)
__host__ __device__ __forceinline__ uint32_t minimizer_hash(uint64_t fwd, uint64_t rc, int k, int m) {
    const uint32_t mmask = (uint32_t)((1ull << (2 * m)) - 1ull);
    uint32_t best = 0xFFFFFFFFu;
    for (int i = 0; i + m <= k; ++i) {
        // the m-mer at offset i from the right end of fwd; its reverse complement sits at the mirrored offset of rc
        const uint32_t a = (uint32_t)(fwd >> (2 * i)) & mmask;
        const uint32_t b = (uint32_t)(rc >> (2 * (k - m - i))) & mmask;
        const uint32_t h = mini_hash(a < b ? a : b);
        best = h < best ? h : best;
    }
    return best;
}

#ifdef __CUDACC__
struct Bucket { uint64_t s0, s1, s2, s3; };

// one 32-byte sector, read-only path, no L1 allocation (no reuse inside an SM)
__device__ __forceinline__ Bucket load_bucket(const uint64_t* p) {
    Bucket b;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(b.s0), "=l"(b.s1), "=l"(b.s2), "=l"(b.s3) : "l"(p));
    return b;
}
// coherent variant for the build / erase kernels (slots change under them)
__device__ __forceinline__ Bucket load_bucket_volatile(const uint64_t* p) {
    Bucket b;
    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(b.s0), "=l"(b.s1) : "l"(p));
    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(b.s2), "=l"(b.s3) : "l"(p + 2));
    return b;
}

// tag bits of the slot matching `want` (= rem << 4 | disp), 0 if none
__device__ __forceinline__ uint32_t match_bucket(const Bucket& b, uint64_t want, bool& found) {
    const bool m0 = (b.s0 >> 3) == want, m1 = (b.s1 >> 3) == want;
    const bool m2 = (b.s2 >> 3) == want, m3 = (b.s3 >> 3) == want;
    found = m0 | m1 | m2 | m3;
    const uint32_t lo = m0 ? (uint32_t)b.s0 : m1 ? (uint32_t)b.s1 : m2 ? (uint32_t)b.s2 : m3 ? (uint32_t)b.s3 : 0u;
    return (lo >> 1) & 3u;
}

// g_kmers[0].find / g_kmers[1].find (classify.cpp:195-202) in one probe.
// `extra` counts bucket reads beyond the first.
__device__ __forceinline__ uint32_t table_probe(const TableView& t, uint64_t canon, uint32_t& extra) {
    const uint64_t h = table_hash(canon, t.k, t.kmask);
    uint32_t bucket = (uint32_t)(h >> t.rem_bits);
    uint64_t want = (h & t.rem_mask) << 4;
    Bucket b = load_bucket(t.slots + (size_t)bucket * kSlotsPerBucket);
    bool found;
    uint32_t tag = match_bucket(b, want, found);
    if (!found && (b.s0 & 1ull)) {                       // rare: home bucket overflowed
        for (int d = 1; d <= kMaxDisp; ++d) {
            bucket = (bucket + 1) & t.bucket_mask;
            want += 1;
            b = load_bucket(t.slots + (size_t)bucket * kSlotsPerBucket);
            ++extra;
            tag = match_bucket(b, want, found);
            if (found || !(b.s0 & 1ull)) break;
        }
    }
    return tag;
}
#endif

}  // namespace hast
