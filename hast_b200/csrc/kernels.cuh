// kernels.cuh -- the sm_100a kernels of the read-classification path.
//
//   K1  table_insert_text_kernel / table_insert_packed_kernel / table_erase_kernel
//       load_kmers (classify.cpp:30-46), str2Kmer (kmer.h:153-166),
//       InitAdaptor (classify.cpp:314-339)
//   K2  tile_kernel<MODE_EXTRACT>   chopRead2Kmer (kmer.h:169-194)
//   K3  lookup_kernel               g_kmers[i].find (classify.cpp:195-202)
//   K2+K3+K4 fused  tile_kernel<MODE_CLASSIFY>
//       process_reads (classify.cpp:186-209) + IncrBarcodeHaps (:52-56)
//
// All of this is HBM/L2-bound integer work: no tensor cores.  Design rules
// used: 128-bit coalesced streaming loads of the read bytes, 2-bit packing
// into shared memory, one 256-bit sector load per table probe, several probes
// in flight per thread, shared-memory vote aggregation before any global
// atomic, persistent CTAs sized from the SM count.
#pragma once
#include <cstdint>
#include "kmer.cuh"
#include "table.cuh"

namespace hast {

struct DevStats {
    unsigned long long lookups;
    unsigned long long reads_with_n;
    unsigned long long reads_short;
    unsigned long long extra_probes;
    unsigned long long bad_kmer_lines;
    unsigned long long table_full;
    unsigned long long reads_too_long;
    unsigned long long bad_barcode;
    unsigned long long filter_pass;
    unsigned long long filter_loads;
};

struct BatchView {
    const uint8_t* bases;         // 16-byte aligned ASCII, or null when `packed` is used
    const uint32_t* read_off;     // n_reads + 1 (byte offsets == base offsets)
    const uint32_t* barcode_id;   // n_reads (may be null for MODE_EXTRACT)
    uint64_t n_bases;
    uint32_t n_reads;
    // host-packed form (hast_submit_batch_packed): the same base stream, 16 bases per word,
    // first base in the top two bits, reads back to back; bit i of has_n = read i contains 'N'
    const uint32_t* packed;
    const uint32_t* has_n;
    // reads per tile of classify_kernel (fused.cuh), chosen per batch so that a tile's reads fill one pass
    uint32_t reads_per_tile = 0;
};

constexpr int kTileThreads = 256;
constexpr int kReadsPerTile = 256;
constexpr int kTileCapBytes = 40960;                    // bases staged per pass
constexpr int kTileWords = kTileCapBytes / 16;          // packed words (16 bases each)
constexpr int kProbeUnroll = 4;

enum { MODE_CLASSIFY = 0, MODE_EXTRACT = 1 };

// streaming 128-bit load of read bytes: read-only path, do not keep in L1
__device__ __forceinline__ uint4 load_stream16(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void set_bits(uint32_t* bits, uint32_t from, uint32_t to) {   // [from, to)
    while (from < to) {
        const uint32_t w = from >> 5, lo = from & 31u;
        const uint32_t n = min(32u - lo, to - from);
        const uint32_t m = (n == 32u ? 0xFFFFFFFFu : ((1u << n) - 1u)) << lo;
        atomicOr(&bits[w], m);
        from += n;
    }
}
__device__ __forceinline__ void clear_bits(uint32_t* bits, uint32_t from, uint32_t to) {   // [from, to)
    while (from < to) {
        const uint32_t w = from >> 5, lo = from & 31u;
        const uint32_t n = min(32u - lo, to - from);
        const uint32_t m = (n == 32u ? 0xFFFFFFFFu : ((1u << n) - 1u)) << lo;
        atomicAnd(&bits[w], ~m);
        from += n;
    }
}
__device__ __forceinline__ bool any_bits(const uint32_t* bits, uint32_t from, uint32_t to) {
    while (from < to) {
        const uint32_t w = from >> 5, lo = from & 31u;
        const uint32_t n = min(32u - lo, to - from);
        const uint32_t m = (n == 32u ? 0xFFFFFFFFu : ((1u << n) - 1u)) << lo;
        if (bits[w] & m) return true;
        from += n;
    }
    return false;
}

// One CTA works through tiles of kReadsPerTile reads.  Per pass over at most
// kTileCapBytes of read bytes:
//   (a) 128-bit coalesced loads of the byte stream -> 2-bit MSB-first words in
//       shared memory; bytes equal to 'N' are flagged in a bit mask
//   (b) one thread per read: containN over its own range, mark every position
//       that starts no k-mer (the last k-1 positions; the whole read if it has
//       an N or is shorter than k)
//   (c) one thread per k-mer POSITION (not per read): cut the 64-bit window out
//       of the packed stream, reverse-complement with brev, take the smaller,
//       hash, one 32-byte bucket load; kProbeUnroll probes in flight per thread;
//       the rare hit adds its tag bits to the read's shared-memory vote word
//   (d) one thread per read: votes -> global per-barcode counters, aggregated
//       across the warp by barcode first
template <int MODE>
__global__ void __launch_bounds__(kTileThreads)
tile_kernel(TableView t, BatchView b, int32_t* __restrict__ counts, uint32_t n_barcodes,
            DevStats* __restrict__ stats, uint64_t* __restrict__ kmers_out,
            uint8_t* __restrict__ has_n_out) {
    __shared__ uint32_t s_off[kReadsPerTile + 1];
    __shared__ uint32_t s_votes[kReadsPerTile];
    __shared__ uint32_t s_packed[kTileWords + 4];
    __shared__ uint32_t s_bad[kTileWords / 2 + 1];

    const uint32_t tid = threadIdx.x;
    const int k = t.k;
    const uint64_t kmask = t.kmask;
    const uint32_t n_tiles = (b.n_reads + kReadsPerTile - 1) / kReadsPerTile;

    unsigned long long st_lookups = 0, st_n = 0, st_short = 0, st_long = 0, st_badbc = 0;
    uint32_t st_extra = 0;

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t r0 = tile * kReadsPerTile;
        const uint32_t R = min((uint32_t)kReadsPerTile, b.n_reads - r0);
        for (uint32_t i = tid; i <= R; i += kTileThreads) s_off[i] = b.read_off[r0 + i];
        for (uint32_t i = tid; i < R; i += kTileThreads) s_votes[i] = 0;
        __syncthreads();

        uint32_t ra = 0;
        while (ra < R) {
            // reads [ra, rb) of the tile whose bytes fit one pass
            const uint32_t lo = s_off[ra] & ~15u;
            uint32_t rb;
            {
                uint32_t a = ra, c = R;                    // largest rb with s_off[rb] - lo <= cap
                while (a < c) {
                    const uint32_t m = (a + c + 1) >> 1;
                    if (s_off[m] - lo <= (uint32_t)kTileCapBytes) a = m; else c = m - 1;
                }
                rb = a;
            }
            if (rb == ra) {                                // a single read larger than a pass
                if (tid == 0) ++st_long;
                ra += 1;
                continue;
            }
            const uint32_t hi = s_off[rb];
            const uint32_t nseg = (hi - lo + 15u) >> 4;

            for (uint32_t i = tid; i < (nseg >> 1) + 1; i += kTileThreads) s_bad[i] = 0;
            __syncthreads();

            // (a) pack
            for (uint32_t seg = tid; seg < nseg + 4; seg += kTileThreads) {
                uint32_t word = 0;
                if (seg < nseg) {
                    const uint64_t g = (uint64_t)lo + 16ull * seg;
                    uint4 v;
                    if (g + 16 <= b.n_bases) {
                        v = load_stream16(b.bases + g);
                    } else {                               // last, partial segment of the batch
                        uint32_t w[4] = {0, 0, 0, 0};
                        for (uint32_t j = 0; j < 16 && g + j < b.n_bases; ++j)
                            w[j >> 2] |= (uint32_t)b.bases[g + j] << (8 * (j & 3));
                        v = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    word = pack16(v);
                    if (any_N4(v.x) | any_N4(v.y) | any_N4(v.z) | any_N4(v.w)) {
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                        for (uint32_t j = 0; j < 16; ++j)
                            if (((w[j >> 2] >> (8 * (j & 3))) & 0xFFu) == 'N') {
                                const uint32_t pos = 16u * seg + j;
                                atomicOr(&s_bad[pos >> 5], 1u << (pos & 31u));
                            }
                    }
                }
                s_packed[seg] = word;
            }
            __syncthreads();

            // (b) per read
            for (uint32_t r = ra + tid; r < rb; r += kTileThreads) {
                const uint32_t s = s_off[r] - lo, e = s_off[r + 1] - lo, L = e - s;
                const bool has_n = any_bits(s_bad, s, e);
                const bool too_short = L < (uint32_t)k;
                if (MODE == MODE_EXTRACT) {
                    if (has_n_out) has_n_out[r0 + r] = has_n ? 1 : 0;
                    // chopRead2Kmer itself packs an 'N' like any byte (kmer.h:11): drop the N marks
                    if (has_n) clear_bits(s_bad, s, e);
                    if (too_short) set_bits(s_bad, s, e);
                    else set_bits(s_bad, e - (uint32_t)k + 1u, e);
                } else {
                    if (has_n) {                           // classify.cpp:190-193: no votes at all
                        ++st_n;
                        set_bits(s_bad, s, e);
                    } else if (too_short) {                // kmer.h:171 assert in the reference
                        ++st_short;
                        set_bits(s_bad, s, e);
                    } else {
                        st_lookups += L - (uint32_t)k + 1u;
                        set_bits(s_bad, e - (uint32_t)k + 1u, e);
                    }
                }
            }
            __syncthreads();

            // (c) per k-mer position
            const uint32_t pbeg = s_off[ra] - lo, pend = hi - lo;
            if (MODE == MODE_EXTRACT) {
                // K2 alone is a streaming kernel (8 bytes out per position): four consecutive positions per thread,
                // the first cut out of the packed stream, the other three ROLLED from it (kmer.h:109-127 does the
                // same on its 128-bit words), written as two 16-byte stores, so a warp writes 1 KiB in one piece.
                // Positions that start no k-mer get all ones.
                const int km = k;
                const uint32_t rc_shift = 2u * (uint32_t)(km - 1);
                for (uint32_t q = (pbeg & ~3u) + 4u * tid; q < pend; q += 4u * kTileThreads) {
                    const uint64_t x = window64(s_packed, q);
                    uint64_t fwd = x >> (64 - 2 * km);
                    uint64_t rcv = revcomp_top(x, kmask);
                    // the three bases that enter next: positions q+k .. q+k+2
                    const uint32_t nb = q + (uint32_t)km;
                    const uint32_t wn = __funnelshift_l(s_packed[(nb >> 4) + 1], s_packed[nb >> 4], (nb & 15u) * 2u);
                    const uint32_t bad4 = (s_bad[q >> 5] >> (q & 31u)) & 15u;     // q is a multiple of 4: one word
                    uint64_t v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j) {
                            const uint64_t c = (wn >> (32 - 2 * j)) & 3u;
                            fwd = ((fwd << 2) | c) & kmask;
                            rcv = (rcv >> 2) | ((c ^ 2ull) << rc_shift);
                        }
                        const uint64_t canon = fwd < rcv ? fwd : rcv;
                        v[j] = ((bad4 >> j) & 1u) ? ~0ull : canon;
                    }
                    uint64_t* dst = kmers_out + (uint64_t)lo + q;
                    if (q >= pbeg && q + 4 <= pend && (((uintptr_t)dst) & 15u) == 0) {
                        asm volatile("st.global.v2.u64 [%0], {%1,%2};" :: "l"(dst), "l"(v[0]), "l"(v[1]) : "memory");
                        asm volatile("st.global.v2.u64 [%0], {%1,%2};" :: "l"(dst + 2), "l"(v[2]), "l"(v[3]) : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (q + j >= pbeg && q + j < pend) dst[j] = v[j];
                    }
                }
            } else
            for (uint32_t base = pbeg; base < pend; base += kTileThreads * kProbeUnroll) {
                uint64_t canon[kProbeUnroll];
                uint64_t want[kProbeUnroll];
                Bucket bk[kProbeUnroll];
                bool valid[kProbeUnroll];
#pragma unroll
                for (int u = 0; u < kProbeUnroll; ++u) {
                    const uint32_t p = base + u * kTileThreads + tid;
                    valid[u] = p < pend && !((s_bad[p >> 5] >> (p & 31u)) & 1u);
                    canon[u] = valid[u] ? canonical_at(s_packed, p, k, kmask) : 0ull;
                    if (MODE == MODE_CLASSIFY) {
                        const uint64_t h = table_hash(canon[u], k, kmask);
                        const uint32_t bucket = (uint32_t)(h >> t.rem_bits);
                        want[u] = (h & t.rem_mask) << 4;
                        bk[u].s0 = bk[u].s1 = bk[u].s2 = bk[u].s3 = 0ull;
                        if (valid[u]) bk[u] = load_bucket(t.slots + (size_t)bucket * kSlotsPerBucket);
                    }
                }
#pragma unroll
                for (int u = 0; u < kProbeUnroll; ++u) {
                    const uint32_t p = base + u * kTileThreads + tid;
                    if (MODE == MODE_EXTRACT) {
                        if (valid[u]) kmers_out[(uint64_t)lo + p] = canon[u];
                    } else {
                        bool found;
                        uint32_t tag = match_bucket(bk[u], want[u], found);
                        if (valid[u] && !found && (bk[u].s0 & 1ull)) {      // overflowed home bucket
                            const uint64_t h = table_hash(canon[u], k, kmask);
                            uint32_t bucket = (uint32_t)(h >> t.rem_bits);
                            uint64_t w = want[u];
                            for (int d = 1; d <= kMaxDisp; ++d) {
                                bucket = (bucket + 1) & t.bucket_mask;
                                w += 1;
                                const Bucket nb = load_bucket(t.slots + (size_t)bucket * kSlotsPerBucket);
                                ++st_extra;
                                tag = match_bucket(nb, w, found);
                                if (found || !(nb.s0 & 1ull)) break;
                            }
                        }
                        if (valid[u] && tag) {
                            // which read owns position p: s_off[r] <= lo + p < s_off[r+1]
                            const uint32_t gp = lo + p;
                            uint32_t a = ra, c = rb - 1;
                            while (a < c) {
                                const uint32_t m = (a + c + 1) >> 1;
                                if (s_off[m] <= gp) a = m; else c = m - 1;
                            }
                            atomicAdd(&s_votes[a], (tag & 1u) | ((tag >> 1) << 16));
                        }
                    }
                }
            }
            __syncthreads();
            ra = rb;
        }

        // (d) votes -> per-barcode counters (IncrBarcodeHaps, classify.cpp:203-206)
        if (MODE == MODE_CLASSIFY) {
            for (uint32_t rbase = 0; rbase < R; rbase += kTileThreads) {
                const uint32_t r = rbase + tid;
                const uint32_t v = r < R ? s_votes[r] : 0u;
                const unsigned voters = __ballot_sync(0xFFFFFFFFu, v != 0u);
                if (v) {
                    const uint32_t bc = b.barcode_id[r0 + r];
                    int v0 = (int)(v & 0xFFFFu), v1 = (int)(v >> 16);
                    const unsigned peers = __match_any_sync(voters, bc);
                    const int leader = __ffs(peers) - 1;
                    int s0 = 0, s1 = 0;
                    for (unsigned m = peers; m; m &= m - 1) {
                        const int src = __ffs(m) - 1;
                        s0 += __shfl_sync(peers, v0, src);
                        s1 += __shfl_sync(peers, v1, src);
                    }
                    if ((int)(tid & 31u) == leader) {
                        if (bc < n_barcodes) {
                            if (s0) atomicAdd(&counts[2ull * bc], s0);
                            if (s1) atomicAdd(&counts[2ull * bc + 1], s1);
                        } else {
                            ++st_badbc;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    // per-warp flush of the statistics
    for (int o = 16; o > 0; o >>= 1) {
        st_lookups += __shfl_xor_sync(0xFFFFFFFFu, st_lookups, o);
        st_n += __shfl_xor_sync(0xFFFFFFFFu, st_n, o);
        st_short += __shfl_xor_sync(0xFFFFFFFFu, st_short, o);
        st_long += __shfl_xor_sync(0xFFFFFFFFu, st_long, o);
        st_badbc += __shfl_xor_sync(0xFFFFFFFFu, st_badbc, o);
        st_extra += __shfl_xor_sync(0xFFFFFFFFu, st_extra, o);
    }
    if ((tid & 31u) == 0) {
        if (st_lookups) atomicAdd(&stats->lookups, st_lookups);
        if (st_n) atomicAdd(&stats->reads_with_n, st_n);
        if (st_short) atomicAdd(&stats->reads_short, st_short);
        if (st_long) atomicAdd(&stats->reads_too_long, st_long);
        if (st_badbc) atomicAdd(&stats->bad_barcode, st_badbc);
        if (st_extra) atomicAdd(&stats->extra_probes, (unsigned long long)st_extra);
    }
}

// K3 standalone: tag bits of n canonical k-mers
__global__ void __launch_bounds__(256)
lookup_kernel(TableView t, const uint64_t* __restrict__ canon, uint64_t n, uint8_t* __restrict__ tags,
              DevStats* __restrict__ stats) {
    uint32_t extra = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        tags[i] = (uint8_t)table_probe(t, canon[i], extra);
    if (extra) atomicAdd(&stats->extra_probes, (unsigned long long)extra);
}

// ---- K1 ------------------------------------------------------------------
__device__ __forceinline__ void table_insert(const TableView& t, uint64_t canon, uint32_t parent,
                                             DevStats* stats) {
    const uint64_t h = table_hash(canon, t.k, t.kmask);
    uint32_t bucket = (uint32_t)(h >> t.rem_bits);
    const uint64_t rem4 = (h & t.rem_mask) << 4;
    const uint64_t tagbits = (uint64_t)(1u << parent) << 1;
    {   // pre-filter first: a key must never be in the table without its filter bits
        const FilterHash fh = filter_hash(canon);
        uint32_t word = fh.word;
        unsigned long long bits = filter_bits(fh);
        if (t.filt_m) {                                    // minimizer-addressed filter (table.cuh)
            const uint64_t rc = revcomp_top(t.k < 32 ? (canon << (64 - 2 * t.k)) : canon, t.kmask);
            word = mini_word(minimizer_hash(canon, rc, t.k, (int)t.filt_m));
            bits = mini_bits(mini_sel(canon, rc, t.k));
        }
        unsigned long long* fw = (unsigned long long*)(t.filt + (word >> t.filt_shift));
        if ((*(volatile unsigned long long*)fw & bits) != bits) atomicOr(fw, bits);
    }
    for (int d = 0; d <= kMaxDisp; ++d) {
        const uint64_t want = rem4 | (uint64_t)d;
        unsigned long long* base = (unsigned long long*)(t.slots + (size_t)bucket * kSlotsPerBucket);
        for (int s = 0; s < kSlotsPerBucket; ++s) {
            for (;;) {
                const unsigned long long cur = *(volatile unsigned long long*)(base + s);
                if ((cur >> 1) == 0ull) {                  // empty: claim it
                    const unsigned long long nv = (want << 3) | tagbits | (cur & 1ull);
                    if (atomicCAS(base + s, cur, nv) == cur) return;
                    continue;                              // lost the race: look at this slot again
                }
                if ((cur >> 3) == want) {                  // same k-mer (duplicate / other parent)
                    if ((cur & tagbits) == 0ull) atomicOr(base + s, tagbits);
                    return;
                }
                break;
            }
        }
        if ((*(volatile unsigned long long*)base & 1ull) == 0ull) atomicOr(base, 1ull);
        bucket = (bucket + 1) & t.bucket_mask;
    }
    atomicAdd(&stats->table_full, 1ull);
}

// text: n_lines lines of exactly k letters + '\n'
__global__ void __launch_bounds__(256)
table_insert_text_kernel(TableView t, const char* __restrict__ text, uint64_t n_lines, uint32_t parent,
                         DevStats* __restrict__ stats) {
    const int k = t.k;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += stride) {
        const char* line = text + i * (uint64_t)(k + 1);
        uint64_t w = 0;
        bool bad = line[k] != '\n';
        for (int j = 0; j < k; ++j) {
            const uint32_t c = (uint8_t)line[j];
            bad |= (c == '\n');
            w = (w << 2) | base_code(c);
        }
        if (bad) { atomicAdd(&stats->bad_kmer_lines, 1ull); continue; }
        const uint64_t rc = revcomp_top(k < 32 ? (w << (64 - 2 * k)) : w, t.kmask);
        table_insert(t, w < rc ? w : rc, parent, stats);
    }
}

__global__ void __launch_bounds__(256)
table_insert_packed_kernel(TableView t, const uint64_t* __restrict__ kmers, uint64_t n, uint32_t parent,
                           DevStats* __restrict__ stats) {
    const int k = t.k;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t w = kmers[i] & t.kmask;
        const uint64_t rc = revcomp_top(k < 32 ? (w << (64 - 2 * k)) : w, t.kmask);
        table_insert(t, w < rc ? w : rc, parent, stats);
    }
}

// InitAdaptor: one thread walks the adaptor in order, so that the erase log
// keeps the reference's order and a repeated k-mer is reported once.
__global__ void table_erase_kernel(TableView t, const char* __restrict__ seq, uint32_t len,
                                   uint64_t* __restrict__ erased, uint8_t* __restrict__ tags,
                                   uint32_t cap, uint32_t* __restrict__ n_erased) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int k = t.k;
    uint32_t n = 0;
    uint64_t w = 0;
    for (uint32_t i = 0; i < len; ++i) {
        w = ((w << 2) | base_code((uint8_t)seq[i])) & t.kmask;
        if (i + 1 < (uint32_t)k) continue;
        const uint64_t rc = revcomp_top(k < 32 ? (w << (64 - 2 * k)) : w, t.kmask);
        const uint64_t canon = w < rc ? w : rc;
        const uint64_t h = table_hash(canon, k, t.kmask);
        uint32_t bucket = (uint32_t)(h >> t.rem_bits);
        uint64_t want = (h & t.rem_mask) << 4;
        for (int d = 0; d <= kMaxDisp; ++d) {
            unsigned long long* base = (unsigned long long*)(t.slots + (size_t)bucket * kSlotsPerBucket);
            bool done = false;
            for (int s = 0; s < kSlotsPerBucket; ++s) {
                const unsigned long long cur = base[s];
                if ((cur >> 1) != 0ull && (cur >> 3) == want) {
                    const uint32_t old = (uint32_t)(cur >> 1) & 3u;
                    if (old) {
                        base[s] = cur & ~6ull;
                        if (n < cap) { erased[n] = canon; tags[n] = (uint8_t)old; }
                        ++n;
                    }
                    done = true;
                    break;
                }
            }
            if (done || !(base[0] & 1ull)) break;
            bucket = (bucket + 1) & t.bucket_mask;
            want += 1;
        }
    }
    *n_erased = n;
}

struct TableCounts {
    unsigned long long entries, displaced, overflow_buckets, size0, size1;
};

__global__ void __launch_bounds__(256)
table_count_kernel(const uint64_t* __restrict__ slots, uint64_t n_buckets, TableCounts* __restrict__ out) {
    unsigned long long e = 0, dsp = 0, ov = 0, s0 = 0, s1 = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_buckets; i += stride) {
        const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(slots + i * 4);
        const ulonglong2 c = *reinterpret_cast<const ulonglong2*>(slots + i * 4 + 2);
        const unsigned long long v[4] = {a.x, a.y, c.x, c.y};
        ov += v[0] & 1ull;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            // an entry is anything ever claimed: tag != 0, or erased (rem/disp bits left behind)
            if ((v[s] >> 1) != 0ull) {
                ++e;
                dsp += ((v[s] >> 3) & 15ull) ? 1 : 0;
                s0 += (v[s] >> 1) & 1ull;
                s1 += (v[s] >> 2) & 1ull;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        e += __shfl_xor_sync(0xFFFFFFFFu, e, o);
        dsp += __shfl_xor_sync(0xFFFFFFFFu, dsp, o);
        ov += __shfl_xor_sync(0xFFFFFFFFu, ov, o);
        s0 += __shfl_xor_sync(0xFFFFFFFFu, s0, o);
        s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out->entries, e);
        atomicAdd(&out->displaced, dsp);
        atomicAdd(&out->overflow_buckets, ov);
        atomicAdd(&out->size0, s0);
        atomicAdd(&out->size1, s1);
    }
}

// Random 32-byte sector gather: the measured random-access roofline that the
// lookup kernels are compared with.
__global__ void __launch_bounds__(256)
gather_kernel(const uint64_t* __restrict__ buf, uint64_t n_sectors_mask, uint64_t n_probes,
              unsigned long long* __restrict__ sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t acc = 0;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n_probes; i += 4 * stride) {
        Bucket b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint64_t x = (i + u * stride) * 0x9E3779B97F4A7C15ull;
            x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
            b[u] = load_bucket(buf + (x & n_sectors_mask) * 4);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc ^= b[u].s0 ^ b[u].s1 ^ b[u].s2 ^ b[u].s3;
    }
    for (; i < n_probes; i += stride) {
        uint64_t x = i * 0x9E3779B97F4A7C15ull;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        const Bucket b = load_bucket(buf + (x & n_sectors_mask) * 4);
        acc ^= b.s0 ^ b.s1 ^ b.s2 ^ b.s3;
    }
    if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}

}  // namespace hast
