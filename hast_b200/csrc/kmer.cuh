// kmer.cuh -- 2-bit k-mer arithmetic shared by every kernel.
//
// Semantics follow 01.classify_stlfr_reads/kmer/kmer.h of the reference
// (base code kmer.h:11, MSB-first packing kmer.h:156-160, reverse complement
// kmer.h:196-223, canonical = numerically smaller word kmer.h:161-165); the
// formulation is our own: the reference rolls two 128-bit words per read
// (kmer.h:109-127), here every k-mer position is cut straight out of a 2-bit
// packed, MSB-first bit stream held in shared memory, so positions are
// independent and can be spread over the lanes of a CTA.
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace hast {

// kmer.h:11  (c & 0x06) >> 1 : A0 C1 T2 G3, defined for every byte value
__host__ __device__ __forceinline__ uint32_t base_code(uint32_t c) { return (c >> 1) & 3u; }

// kmer.h:129-148 createFilter, k <= 32
__host__ __device__ __forceinline__ uint64_t kmer_mask(int k) {
    return k < 32 ? ((1ull << (2 * k)) - 1ull) : ~0ull;
}

#ifdef __CUDACC__
// Reverse complement of the k-mer that occupies the TOP 2k bits of x, returned
// right-aligned (low 2k bits).  Complement = ^10b per base (kmer.h:13,199);
// __brevll reverses base order but also swaps the two bits of every base, which
// the 0x5555 swap undoes.
__device__ __forceinline__ uint64_t revcomp_top(uint64_t x, uint64_t mask) {
    uint64_t y = __brevll(x ^ 0xAAAAAAAAAAAAAAAAull);
    y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
    return y & mask;
}

// reverse complement of all 16 bases of one packed word
__device__ __forceinline__ uint32_t revcomp16(uint32_t w) {
    const uint32_t y = __brev(w ^ 0xAAAAAAAAu);
    return ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
}

// 16 ASCII bases (one 128-bit load) -> one 32-bit word, first base in the top
// two bits.  Per 4 bytes: codes = (w >> 1) & 0x03030303 leaves base i in byte i;
// the multiply gathers the four 2-bit fields into the top byte with no carries
// (fields land on pairwise distinct bit positions).
__device__ __forceinline__ uint32_t pack4(uint32_t w) {
    return (((w >> 1) & 0x03030303u) * 0x40100401u) >> 24;
}
__device__ __forceinline__ uint32_t pack16(uint4 v) {
    return (pack4(v.x) << 24) | (pack4(v.y) << 16) | (pack4(v.z) << 8) | pack4(v.w);
}
// does any byte of w equal 'N' (0x4E)?  exact (classify.cpp:182-185)
__device__ __forceinline__ bool any_N4(uint32_t w) {
    uint32_t x = w ^ 0x4E4E4E4Eu;
    return ((x - 0x01010101u) & ~x & 0x80808080u) != 0u;
}

// 4-bit mask of the bytes of w that are NOT one of 'A' 'C' 'G' 'T' (upper case), exact per byte
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) {       // 0x80 in every byte of x that is 0
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}
__device__ __forceinline__ uint32_t not_acgt4(uint32_t w) {
    const uint32_t good = zero_bytes(w ^ 0x41414141u) | zero_bytes(w ^ 0x43434343u) |
                          zero_bytes(w ^ 0x47474747u) | zero_bytes(w ^ 0x54545454u);
    const uint32_t bad = ~good & 0x80808080u;
    return ((bad >> 7) * 0x00204081u >> 21) & 0xFu;                // bits 0,8,16,24 -> bits 0..3
}

// The 64 stream bits that start at base position p (p counted from the first
// base of the packed tile): words are 16 bases each, MSB-first.
__device__ __forceinline__ uint64_t window64(const uint32_t* __restrict__ s_packed, uint32_t p) {
    const uint32_t wi = p >> 4, sh = (p & 15u) * 2u;
    const uint32_t w0 = s_packed[wi], w1 = s_packed[wi + 1], w2 = s_packed[wi + 2];
    const uint32_t hi = __funnelshift_l(w1, w0, sh);
    const uint32_t lo = __funnelshift_l(w2, w1, sh);
    return ((uint64_t)hi << 32) | lo;
}

// canonical k-mer at stream position p (Kmer::chopRead2Kmer, kmer.h:169-194)
__device__ __forceinline__ uint64_t canonical_at(const uint32_t* __restrict__ s_packed, uint32_t p,
                                                 int k, uint64_t mask) {
    const uint64_t x = window64(s_packed, p);
    const uint64_t fwd = x >> (64 - 2 * k);
    const uint64_t rc = revcomp_top(x, mask);
    return fwd < rc ? fwd : rc;
}
#endif

}  // namespace hast
