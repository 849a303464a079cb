// kcount.cuh -- on-device construction of the parent-unique k-mer lists (HAST stage 00).
//
// Replaces 00.build_unshare_kmers_by_jellyfish/build_unshared_kmers.sh:163-291, i.e. the five
// `jellyfish count -C` runs, the dumps between them and the "2 copies of maternal + 1 copy of
// paternal" trick, by ONE count table that holds both parents' counts per canonical k-mer:
//
//     paternal.unique.filter = { x : cntP(x) in [PL, PU]  and  cntM(x) == 0 }      (:262-291)
//     maternal.unique.filter = { x : cntM(x) in [ML, MU]  and  cntP(x) == 0 }
//
// Counting rule (jellyfish 2.3.0 `count -m k -C`, the binary the reference vendors): every
// window of k bases all in ACGTacgt counts once for its canonical form, the lexicographically
// smaller (A<C<G<T) of the window and its reverse complement; any other byte breaks the
// window.  K-mers are therefore held in jellyfish's code A0 C1 G2 T3 here (NOT kmer.h's
// A0 C1 T2 G3): the numerically smaller packed word IS the lexicographically smaller string,
// so what the dump kernels emit is exactly the text jellyfish prints.
//
// Layout in HBM: 2^b slots of 16 bytes { key + 1 (0 = empty), count[paternal], count[maternal] },
// linear probing from the low bits of a 64-bit mix of the key.  A context may own just one
// PARTITION of the key space (the top bits of the same mix): a table for both human parents
// (several 10^9 distinct k-mers with their error k-mers) does not fit one GPU, so the reads
// are streamed once per partition, or once to several GPUs that each keep a different one.
// Bound by random 16-byte atomics on HBM; no tensor cores.
#pragma once
#include <cstdint>
#include "kernels.cuh"

namespace hast {

struct KcSlot { unsigned long long key; uint32_t cnt[2]; };
static_assert(sizeof(KcSlot) == 16, "one slot = 16 bytes");

struct KcView {
    KcSlot* slots;
    uint64_t mask;          // n_slots - 1
    int32_t k;
    uint32_t part, n_parts; // this table keeps keys with part_of(mix) == part
    uint32_t max_probe;
};

struct KcStats {
    unsigned long long windows;      // valid k-mer windows seen (all partitions)
    unsigned long long counted;      // windows that fell into this partition
    unsigned long long table_full;   // insertions that found no slot
    unsigned long long too_long;     // chunks larger than a pass (host must chunk)
};

__host__ __device__ __forceinline__ uint64_t kc_mix(uint64_t x) {      // murmur3 finaliser: bijective
    x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull;
    x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull;
    x ^= x >> 33;
    return x;
}
__host__ __device__ __forceinline__ uint32_t kc_part_of(uint64_t mix, uint32_t n_parts) {
    return (uint32_t)(((mix >> 32) * (uint64_t)n_parts) >> 32);
}

// kmer.h-coded 2-bit fields (A0 C1 T2 G3, what pack16 produces) <-> jellyfish code (A0 C1 G2 T3):
// x ^ (x >> 1) per field, an involution on {0,1,2,3} that swaps 2 and 3
__host__ __device__ __forceinline__ uint32_t recode32(uint32_t w) { return w ^ ((w >> 1) & 0x55555555u); }
__host__ __device__ __forceinline__ uint64_t recode64(uint64_t w) { return w ^ ((w >> 1) & 0x5555555555555555ull); }

#ifdef __CUDACC__
__device__ __forceinline__ void kc_insert(const KcView& t, uint64_t canon, uint64_t mix, uint32_t parent,
                                          uint32_t& full) {
    const unsigned long long key = canon + 1ull;
    uint64_t s = mix & t.mask;
    for (uint32_t probe = 0; probe < t.max_probe; ++probe) {
        KcSlot* slot = t.slots + s;
        unsigned long long cur = *(volatile unsigned long long*)&slot->key;
        if (cur == 0ull) {
            cur = atomicCAS(&slot->key, 0ull, key);
            if (cur == 0ull) cur = key;
        }
        if (cur == key) {
            atomicAdd(&slot->cnt[parent], 1u);
            return;
        }
        s = (s + 1) & t.mask;
    }
    ++full;
}

constexpr int kKcReadsPerTile = 240;
constexpr int kKcCap = 40960;                          // bases per pass
constexpr int kKcWords = kKcCap / 16;

struct __align__(16) KcSmem {
    uint32_t packed[kKcWords + 4];                     // jellyfish-coded, 16 bases per word, MSB first
    uint32_t bad[kKcWords / 2 + 2];                    // 1 bit per position: starts no valid window
    uint32_t off[kKcReadsPerTile + 1];
};

// One CTA walks tiles of sequence chunks ("reads" of a BatchView; barcode_id unused).
//   (a) 128-bit streaming loads -> 2-bit words in shared memory, bytes outside ACGTacgt flagged
//   (b) flags smeared over the k-1 positions before them; the last k-1 positions of a chunk and
//       chunks shorter than k start no window
//   (c) one thread per 16-position word: forward and reverse-complement k-mers roll in registers,
//       canonical = min, mix, partition test, insert-or-increment in the count table
template <int KT>
__global__ void __launch_bounds__(kTileThreads, 4)
kc_count_kernel(KcView t, BatchView b, uint32_t parent, KcStats* __restrict__ stats) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    KcSmem& sm = *reinterpret_cast<KcSmem*>(smem_raw);
    uint32_t* const s_off = sm.off;
    uint32_t* const s_packed = sm.packed;
    uint32_t* const s_bad = sm.bad;
    const uint32_t tid = threadIdx.x;
    const int k = KT ? KT : t.k;
    const uint64_t kmask = kmer_mask(k);
    const int km1 = k - 1;
    const uint64_t mask_km1 = kmer_mask(km1);
    const uint32_t fwd_init_shift = 64u - 2u * (uint32_t)km1;
    const uint32_t rc_shift = 2u * (uint32_t)km1;
    const uint32_t nxt_word = (uint32_t)km1 >> 4, nxt_sh = ((uint32_t)km1 & 15u) * 2u;
    const uint32_t n_tiles = (b.n_reads + kKcReadsPerTile - 1) / kKcReadsPerTile;
    unsigned long long st_windows = 0, st_counted = 0, st_long = 0;
    uint32_t st_full = 0;

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t r0 = tile * kKcReadsPerTile;
        const uint32_t R = min((uint32_t)kKcReadsPerTile, b.n_reads - r0);
        __syncthreads();                               // previous tile's sweep is done with s_off
        for (uint32_t i = tid; i <= R; i += kTileThreads) s_off[i] = b.read_off[r0 + i];
        __syncthreads();
        uint32_t ra = 0;
        while (ra < R) {
            const uint32_t lo = s_off[ra] & ~15u;
            uint32_t rb;
            {
                uint32_t a = ra, c = R;
                while (a < c) {
                    const uint32_t m = (a + c + 1) >> 1;
                    if (s_off[m] - lo <= (uint32_t)kKcCap) a = m; else c = m - 1;
                }
                rb = a;
            }
            if (rb == ra) {                            // a chunk larger than a pass: the host chunks smaller
                if (tid == 0) ++st_long;
                ra += 1;
                continue;
            }
            const uint32_t hi = s_off[rb];
            const uint32_t nseg = (hi - lo + 15u) >> 4;
            __syncthreads();                           // previous pass's sweep is done with s_packed / s_bad
            for (uint32_t i = tid; i < (nseg >> 1) + 2; i += kTileThreads) s_bad[i] = 0;
            __syncthreads();
            // (a)
            for (uint32_t seg = tid; seg < nseg + 4; seg += kTileThreads) {
                uint32_t word = 0;
                if (seg < nseg) {
                    const uint64_t g = (uint64_t)lo + 16ull * seg;
                    uint4 v;
                    if (g + 16 <= b.n_bases) {
                        v = load_stream16(b.bases + g);
                    } else {
                        uint32_t w[4] = {0, 0, 0, 0};
                        for (uint32_t j = 0; j < 16 && g + j < b.n_bases; ++j)
                            w[j >> 2] |= (uint32_t)b.bases[g + j] << (8 * (j & 3));
                        v = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    word = recode32(pack16(v));
                    const uint32_t up = 0xDFDFDFDFu;   // fold lower case onto upper case
                    const uint32_t m16 = not_acgt4(v.x & up) | (not_acgt4(v.y & up) << 4) |
                                         (not_acgt4(v.z & up) << 8) | (not_acgt4(v.w & up) << 12);
                    if (m16) atomicOr(&s_bad[seg >> 1], m16 << ((seg & 1u) * 16u));
                }
                s_packed[seg] = word;
            }
            __syncthreads();
            // (b) an invalid byte spoils the k windows that contain it
            {
                uint32_t smeared[(kKcWords / 2 + 2 + kTileThreads - 1) / kTileThreads];
                const uint32_t nbw = (nseg >> 1) + 1;
                int it = 0;
                for (uint32_t w = tid; w < nbw; w += kTileThreads, ++it) {
                    const uint64_t pair = (uint64_t)s_bad[w] | ((uint64_t)s_bad[w + 1] << 32);
                    uint64_t d = pair;
                    for (int j = 1; j < k; ++j) d |= pair >> j;
                    smeared[it] = (uint32_t)d;
                }
                __syncthreads();
                it = 0;
                for (uint32_t w = tid; w < nbw; w += kTileThreads, ++it) s_bad[w] = smeared[it];
                __syncthreads();
            }
            for (uint32_t r = ra + tid; r < rb; r += kTileThreads) {
                const uint32_t s = s_off[r] - lo, e = s_off[r + 1] - lo, L = e - s;
                if (L < (uint32_t)k) set_bits(s_bad, s, e);
                else set_bits(s_bad, e - (uint32_t)k + 1u, e);
            }
            if (tid == 0) {
                set_bits(s_bad, 0, s_off[ra] - lo);
                set_bits(s_bad, hi - lo, nseg * 16u);
            }
            __syncthreads();
            // (c)
            for (uint32_t wi = tid; wi < nseg; wi += kTileThreads) {
                uint32_t valid = ~(s_bad[wi >> 1] >> ((wi & 1u) * 16u)) & 0xFFFFu;
                if (!valid) continue;
                st_windows += __popc(valid);
                const uint32_t w0 = s_packed[wi], w1 = s_packed[wi + 1];
                const uint32_t nxt = __funnelshift_l(s_packed[wi + nxt_word + 1], s_packed[wi + nxt_word], nxt_sh);
                const uint64_t x = ((uint64_t)w0 << 32) | w1;
                uint64_t fwd = km1 ? (x >> fwd_init_shift) : 0ull;
                // reverse complement of the first k-1 bases (complement = ^3 in this code), shifted up by
                // one base so that the first roll lands it in place
                uint64_t rcv;
                {
                    uint64_t y = __brevll(~x);
                    y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
                    rcv = (y & mask_km1) << 2;
                }
#pragma unroll 4
                for (int j = 0; j < 16; ++j) {
                    const uint32_t c = (nxt >> (30 - 2 * j)) & 3u;
                    fwd = ((fwd << 2) | c) & kmask;
                    rcv = (rcv >> 2) | ((uint64_t)(c ^ 3u) << rc_shift);
                    if ((valid >> j) & 1u) {
                        const uint64_t canon = fwd < rcv ? fwd : rcv;
                        const uint64_t mix = kc_mix(canon);
                        if (t.n_parts == 1u || kc_part_of(mix, t.n_parts) == t.part) {
                            ++st_counted;
                            kc_insert(t, canon, mix, parent, st_full);
                        }
                    }
                }
            }
            ra = rb;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        st_windows += __shfl_xor_sync(0xFFFFFFFFu, st_windows, o);
        st_counted += __shfl_xor_sync(0xFFFFFFFFu, st_counted, o);
        st_long += __shfl_xor_sync(0xFFFFFFFFu, st_long, o);
        st_full += __shfl_xor_sync(0xFFFFFFFFu, st_full, o);
    }
    if ((tid & 31u) == 0) {
        if (st_windows) atomicAdd(&stats->windows, st_windows);
        if (st_counted) atomicAdd(&stats->counted, st_counted);
        if (st_long) atomicAdd(&stats->too_long, st_long);
        if (st_full) atomicAdd(&stats->table_full, (unsigned long long)st_full);
    }
}

// `jellyfish histo` (low 1, high `high`): histo[c] = distinct k-mers of `parent` with count c,
// counts above `high` collected in histo[high + 1]; histo[0] = distinct k-mers of the parent.
__global__ void __launch_bounds__(256)
kc_histo_kernel(const KcSlot* __restrict__ slots, uint64_t n_slots, uint32_t parent, uint32_t high,
                unsigned long long* __restrict__ histo) {
    constexpr uint32_t kLocal = 1024;
    __shared__ uint32_t s_h[kLocal];
    for (uint32_t i = threadIdx.x; i < kLocal; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {
        const uint4 v = *reinterpret_cast<const uint4*>(slots + i);
        if ((v.x | v.y) == 0u) continue;
        const uint32_t c = parent ? v.w : v.z;
        if (!c) continue;
        const uint32_t bin = min(c, high + 1u);
        if (bin < kLocal) atomicAdd(&s_h[bin], 1u);
        else atomicAdd(&histo[bin], 1ull);
        atomicAdd(&s_h[0], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kLocal && i <= high + 1u; i += blockDim.x)
        if (s_h[i]) atomicAdd(&histo[i], (unsigned long long)s_h[i]);
}

// k-mers of `parent` with lower <= count <= upper that the other parent never showed:
// appended (any order) to out[0 .. cap); *n_out counts all of them, also beyond cap.
// hast_code: emit in kmer.h's code (A0 C1 T2 G3) for table_insert_packed_kernel instead of
// jellyfish's.
__global__ void __launch_bounds__(256)
kc_select_kernel(const KcSlot* __restrict__ slots, uint64_t n_slots, uint32_t parent, uint32_t lower,
                 uint32_t upper, bool require_unique, bool hast_code, uint64_t* __restrict__ out, uint64_t cap,
                 unsigned long long* __restrict__ n_out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t base = first - lane; base < n_slots; base += stride) {        // warp-uniform trip count
        const uint64_t i = base + lane;
        bool take = false;
        uint64_t key = 0;
        if (i < n_slots) {
            const uint4 v = *reinterpret_cast<const uint4*>(slots + i);
            key = ((uint64_t)v.y << 32) | v.x;
            const uint32_t mine = parent ? v.w : v.z, other = parent ? v.z : v.w;
            take = key != 0ull && mine >= lower && mine <= upper && (!require_unique || other == 0u);
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, take);
        if (!m) continue;
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(n_out, (unsigned long long)__popc(m));
        at = __shfl_sync(0xFFFFFFFFu, at, 0) + __popc(m & ((1u << lane) - 1u));
        if (take && at < cap) out[at] = hast_code ? recode64(key - 1ull) : key - 1ull;
    }
}

struct KcTotals { unsigned long long distinct[2], both, occupied, occurrences[2]; };
__global__ void __launch_bounds__(256)
kc_totals_kernel(const KcSlot* __restrict__ slots, uint64_t n_slots, KcTotals* __restrict__ out) {
    unsigned long long d0 = 0, d1 = 0, both = 0, occ = 0, o0 = 0, o1 = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {
        const uint4 v = *reinterpret_cast<const uint4*>(slots + i);
        if ((v.x | v.y) == 0u) continue;
        ++occ;
        d0 += v.z != 0u; d1 += v.w != 0u; both += (v.z != 0u && v.w != 0u);
        o0 += v.z; o1 += v.w;
    }
    for (int o = 16; o > 0; o >>= 1) {
        d0 += __shfl_xor_sync(0xFFFFFFFFu, d0, o); d1 += __shfl_xor_sync(0xFFFFFFFFu, d1, o);
        both += __shfl_xor_sync(0xFFFFFFFFu, both, o); occ += __shfl_xor_sync(0xFFFFFFFFu, occ, o);
        o0 += __shfl_xor_sync(0xFFFFFFFFu, o0, o); o1 += __shfl_xor_sync(0xFFFFFFFFu, o1, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out->distinct[0], d0); atomicAdd(&out->distinct[1], d1); atomicAdd(&out->both, both);
        atomicAdd(&out->occupied, occ); atomicAdd(&out->occurrences[0], o0); atomicAdd(&out->occurrences[1], o1);
    }
}
#endif  // __CUDACC__

}  // namespace hast
