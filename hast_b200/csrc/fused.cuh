// fused.cuh -- the fused read-classification kernel (K2 + K3 + K4), second generation.
//
// Replaces MultiThread::process_reads (classify.cpp:186-209): containN, every
// canonical k-mer of the read (Kmer::chopRead2Kmer, kmer.h:169-194), the two
// unordered_set finds per k-mer (classify.cpp:195-202) and IncrBarcodeHaps
// (classify.cpp:52-56,203-206).
//
// What bounds it (measured on B200, profiles/microbench_r01.csv): an SM retires
// ONE divergent 128-byte-line request per clock, i.e. 290 G random probes/s for
// the whole chip when the probed structure is L2-resident, but only 38 G/s when
// every probe is an HBM transaction.  ~99 % of the k-mer positions of a read are
// in neither parent's set, so the kernel answers them from an L2-resident Bloom
// pre-filter (one 8-byte probe per position) and sends only the few positions
// that pass (members + ~1-5 % false positives) to the exact table in HBM
// (table.cuh; one 32-byte sector per probe).  The filter only ever says "maybe":
// every vote still comes from an exact match in the table, so results stay
// bit-identical to the reference.
//
// Per tile of kReadsPerTile reads, one CTA:
//   (a) 128-bit streaming loads of the read bytes (L2 evict-first) -> 2-bit
//       MSB-first words in shared memory; 'N' bytes flagged in a bit mask
//   (b) one thread per read: containN, mark the positions that start no k-mer
//   (c) one thread per 16-position chunk: the forward and reverse-complement
//       k-mers ROLL in registers (kmer.h:109-127 does the same on 128-bit
//       words), min -> hash -> one 8-byte filter load per valid position, eight
//       loads in flight per thread; passing positions are appended to a
//       shared-memory queue (warp-aggregated)
//   (d) the queue is drained by all threads: exact probe of the table, hits add
//       their tag bits to the owning read's vote word
//   (e) one thread per read: votes -> per-barcode counters, aggregated by
//       barcode across the warp before any global atomic
#pragma once
#include <cstdint>
#include "kernels.cuh"

namespace hast {

// Two ways of getting a pass's read bytes on chip (template parameter TMA):
//   false  every thread issues 128-bit streaming loads (L2 evict-first) and packs them straight
//          into shared memory -- no staging buffer;
//   true   one thread streams the NEXT pass into a raw staging buffer with a TMA bulk copy
//          (cp.async.bulk + mbarrier) while the current pass is swept; packing then reads shared memory.
// Measured on B200 (cfg2, G lookups/s): direct 218; TMA with a 8 / 16 / 24 / 32 KiB staging buffer
// 172 / 199 / 205 / 142.  The staging buffer costs what this kernel needs most: L1.  The random filter
// probes in flight (~0.75 sectors/clk/SM x ~800 clk) each hold a 128-byte L1 line until they return,
// and shared memory is carved out of the same 256 KiB array, so a CTA footprint beyond ~45 KiB (x4)
// throttles the sweep, while a small buffer means more passes and their fixed cost.  The direct
// path is the default; the TMA path stays selectable (hast_set_option "kernel" = 2).
#ifndef HAST_PASS_CAP_TMA
#define HAST_PASS_CAP_TMA 24576
#endif
#ifndef HAST_PASS_CAP
#define HAST_PASS_CAP 40960
#endif
// Most reads a tile may hold (sizes the per-read arrays in shared memory).  The tile size itself is chosen per
// batch (fused_reads_per_tile below): as many reads as fill one pass of HAST_PASS_CAP bytes, because the fixed
// cost of a tile -- six CTA barriers, the stragglers each of them waits for -- is then paid once per 40 KB of
// reads instead of once per 24 KB.  Measured on B200 (profiles/bench_r02_c_ab_*.json, 100-base reads): 240 reads
// per tile 267.0 / 221.7 G lookups/s (128 MiB / 1 GiB table), 320: 266.8 / 226.2, 409 (a full pass): 277.5 / 234.3.
// (r02_d, the same with the tile size chosen per batch: 276-279 / 245-249 with room for 640 reads and passes of 40-56 KB;
// 416 keeps the shared-memory arrays as small as the fixed 409 did, which measured best on the 1 GiB table: 250.9.)
#ifndef HAST_READS_PER_TILE
#define HAST_READS_PER_TILE 416
#endif
#ifndef HAST_FUSED_THREADS
#define HAST_FUSED_THREADS 256
#endif
#ifndef HAST_FUSED_CTAS_PER_SM
#define HAST_FUSED_CTAS_PER_SM 4
#endif
constexpr int kFusedThreads = HAST_FUSED_THREADS;        // threads per CTA of classify_kernel
constexpr int kFusedReadsPerTile = HAST_READS_PER_TILE;   // upper bound; see fused_reads_per_tile
constexpr int kQueueCap = 768 * (HAST_FUSED_THREADS / 32);                          // passing positions buffered per CTA
constexpr int kChunk = 16;                               // positions per thread per sweep (= bases per packed word)
constexpr int kDrainUnroll = 4;                          // exact probes in flight per thread while draining
constexpr int kWarpQueueCap = kQueueCap / (kFusedThreads / 32);             // every warp of the CTA queues and drains on its own
constexpr uint32_t kDrainChunk = 32u * kDrainUnroll;     // one full round of probes for a warp
static_assert(kWarpQueueCap >= (int)kDrainChunk - 1 + 16 * 32, "eight warps; a sweep appends up to 16 positions per lane");

// reads per tile for a batch: what fills one pass (cap - 16 bytes) at the batch's mean read length
__host__ __device__ inline uint32_t fused_reads_per_tile(uint64_t n_bases, uint32_t n_reads, uint32_t pass_cap) {
    const uint64_t mean = n_reads ? (n_bases + n_reads - 1) / n_reads : 1;
    uint64_t r = (pass_cap - 16u) / (mean ? mean : 1);
    if (r > (uint64_t)kFusedReadsPerTile) r = kFusedReadsPerTile;
    if (r < 32) r = 32;
    return (uint32_t)r;
}

template <bool TMA>
struct __align__(128) FusedSmem {
    static constexpr int kCap = TMA ? HAST_PASS_CAP_TMA : HAST_PASS_CAP;   // read bytes per pass (longest read: cap - 16)
    static constexpr int kWords = kCap / 16;             // packed words (16 bases each)
    uint8_t raw[TMA ? kCap : 16];                        // TMA destination: the pass's ASCII bytes
    uint32_t packed[kWords + 4];                         // 2-bit MSB-first, 16 bases per word
    uint32_t bad[kWords / 2 + 2];                        // 1 bit per position: starts no k-mer
    uint16_t queue[kQueueCap];                           // positions that passed the pre-filter
    uint32_t off[kFusedReadsPerTile + 1];
    uint32_t votes[kFusedReadsPerTile];
    unsigned long long mbar;                             // completion barrier of the bulk copy into raw
    uint32_t pad_;
};
static_assert(HAST_FUSED_CTAS_PER_SM * (sizeof(FusedSmem<true>) + 1024) <= 227 * 1024, "the CTAs of an SM must fit");
static_assert(HAST_FUSED_CTAS_PER_SM * (sizeof(FusedSmem<false>) + 1024) <= 227 * 1024, "the CTAs of an SM must fit");

// hit at global base offset gp: add the tag bits to the vote word of the read that owns it,
// s_off[r] <= gp < s_off[r+1] with r in [ra, rb)
__device__ __forceinline__ void vote(uint32_t* s_votes, const uint32_t* s_off, uint32_t ra, uint32_t rb,
                                     uint32_t gp, uint32_t tag) {
    uint32_t a = ra, c = rb - 1;
    while (a < c) {
        const uint32_t m = (a + c + 1) >> 1;
        if (s_off[m] <= gp) a = m; else c = m - 1;
    }
    atomicAdd(&s_votes[a], (tag & 1u) | ((tag >> 1) << 16));
}

__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
// one filter word; keep it in L2 ahead of the streaming read bytes
__device__ __forceinline__ uint64_t load_filter(const uint64_t* p, uint64_t pol) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}

// the same under a predicate that the compiler cannot see through: with a plain `if (p) v = load`, ptxas
// folds the sweep's "keep the previous word" select into the load's destination register, which makes
// every fetch wait for the one before it (measured: the sweep then runs at one L2 latency per fetch)
__device__ __forceinline__ uint64_t load_filter_if(const uint64_t* p, uint64_t pol, uint32_t on) {
    uint64_t v = 0ull;
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %3, 0;\n"
                 "@q ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;\n}"
                 : "+l"(v) : "l"(p), "l"(pol), "r"(on));
    return v;
}

// 16 read bytes, used once: do not let them push the filter / table out of L2
__device__ __forceinline__ uint4 load_stream16_ef(const uint8_t* p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// One thread: stream `bytes` (multiple of 16, > 0) from global into shared memory, evict-first in
// L2 (read once: must not push the filter / table out), completion signalled on `bar`.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar,
                                            uint64_t policy) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// KT > 0: k is the compile-time constant KT (shift amounts fold into immediates); KT == 0: any k in 1..32.
// PACKED: the batch arrives 2-bit packed from the host parser (a quarter of the PCIe bytes); phase (a)
// is then a plain copy of words and containN comes from the per-read flag the parser set.
// SEQ: the per-sequence classifier of stage 03 (03.mkoutput_by_fabulous2.0/src_main/classify.cpp:203-218).
// That program matches k-mer STRINGS, so a window votes only if all its bytes are upper-case A/C/G/T
// (the lists are jellyfish dumps: upper-case ACGT); a window over anything else ('N', lower case, '\r')
// simply does not match -- it does not silence the rest of the sequence -- and a sequence shorter than
// k is not an error.  "Reads" are chunks of a sequence, "barcodes" are sequence ids.
// MINI (KT > 0 with mini_len(KT) > 0 only): the filter word of a position is chosen by the k-mer's
// minimizer (table.cuh), so a thread walking its 16 neighbouring positions fetches a new filter word
// only where the minimizer changes (~2/(k-m+2) of the positions) instead of at every position.  The
// sweep was bound by the one-divergent-request-per-clock limit of the L1 (profiles/r01_c); this trades
// ~15 integer instructions per position for ~70 % of those requests.
template <int KT, bool TMA, bool PACKED = false, bool SEQ = false, bool MINI = false>
__global__ void __launch_bounds__(kFusedThreads, HAST_FUSED_CTAS_PER_SM)
classify_kernel(TableView t, BatchView b, int32_t* __restrict__ counts, uint32_t n_barcodes,
                DevStats* __restrict__ stats) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    using Smem = FusedSmem<TMA>;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    constexpr uint32_t kPassCapBytes = Smem::kCap;
    const uint32_t kReadsPerTile = b.reads_per_tile;       // <= kFusedReadsPerTile (set by the launcher)
    uint32_t* const s_off = sm.off;
    uint32_t* const s_votes = sm.votes;
    uint32_t* const s_packed = sm.packed;
    uint32_t* const s_bad = sm.bad;
    uint16_t* const s_queue = sm.queue;
    uint16_t* const wqueue = s_queue + (threadIdx.x >> 5) * kWarpQueueCap;

    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const int k = KT ? KT : t.k;
    const uint64_t kmask = KT ? kmer_mask(KT) : t.kmask;
    const uint32_t n_tiles = (b.n_reads + kReadsPerTile - 1) / kReadsPerTile;
    const uint64_t pol_last = policy_evict_last(), pol_first = policy_evict_first();

    // rolling-window constants: the chunk at base position p starts from the k-1
    // bases p .. p+k-2 and then takes in base p+k-1+j for its j-th k-mer
    const int km1 = k - 1;
    const uint64_t mask_km1 = kmer_mask(km1);
    const uint32_t fwd_init_shift = 64u - 2u * (uint32_t)km1;   // 64 when k == 1 (handled below)
    const uint32_t rc_shift = 2u * (uint32_t)km1;
    const uint32_t nxt_word = (uint32_t)km1 >> 4, nxt_sh = ((uint32_t)km1 & 15u) * 2u;

    unsigned long long st_lookups = 0, st_n = 0, st_short = 0, st_long = 0, st_badbc = 0;
    uint32_t st_extra = 0, st_pass = 0, st_loads = 0;

    // The read bytes of pass i+1 are streamed into sm.raw by one bulk copy while pass i is being
    // swept (raw is only needed until pass i has been packed).  Thread 0 issues, everybody waits
    // on the mbarrier; exactly one copy (or a bare arrive for an empty range) per pass, in order.
    auto prefetch = [&](uint32_t lo16, uint32_t end) {           // thread 0 only; bytes [lo16, end) of the batch
        if (!TMA) return;
        const uint32_t bytes = min((uint32_t)kPassCapBytes, (end - lo16 + 15u) & ~15u);
        if (bytes) tma_load_1d(sm.raw, b.bases + lo16, bytes, &sm.mbar, pol_first);
        else mbar_arrive(&sm.mbar);
    };
    uint32_t parity = 0;
    if (TMA && tid == 0) {
        mbar_init(&sm.mbar, 1);
        const uint32_t r0 = blockIdx.x * kReadsPerTile;           // grid <= n_tiles: the first tile exists
        const uint32_t R = min((uint32_t)kReadsPerTile, b.n_reads - r0);
        prefetch(b.read_off[r0] & ~15u, b.read_off[r0 + R]);
    }
    __syncthreads();

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t r0 = tile * kReadsPerTile;
        const uint32_t R = min((uint32_t)kReadsPerTile, b.n_reads - r0);
        for (uint32_t i = tid; i <= R; i += kFusedThreads) s_off[i] = b.read_off[r0 + i];
        for (uint32_t i = tid; i < R; i += kFusedThreads) s_votes[i] = 0;
        // first pass of this CTA's next tile (thread 0 keeps its byte range in registers)
        const uint32_t ntile = tile + gridDim.x;
        uint32_t nt_lo = 0, nt_end = 0;
        if (TMA && tid == 0 && ntile < n_tiles) {
            const uint32_t nr0 = ntile * kReadsPerTile;
            const uint32_t nR = min((uint32_t)kReadsPerTile, b.n_reads - nr0);
            nt_lo = b.read_off[nr0] & ~15u;
            nt_end = b.read_off[nr0 + nR];
        }
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
                    if (s_off[m] - lo <= (uint32_t)kPassCapBytes) a = m; else c = m - 1;
                }
                rb = a;
            }
            const bool too_long = rb == ra;                // a single read larger than a pass: skipped, reported
            if (too_long) rb = ra + 1;
            const uint32_t hi = too_long ? lo : s_off[rb];
            const uint32_t nseg = (hi - lo + 15u) >> 4;

            for (uint32_t i = tid; i < (nseg >> 1) + 2; i += kFusedThreads) s_bad[i] = 0;
            if (TMA) {
                mbar_wait(&sm.mbar, parity);               // this pass's bytes have landed in sm.raw
                parity ^= 1u;
            }
            __syncthreads();

            // (a) pack: 16 ASCII bytes -> one 2-bit word; 'N' bytes flagged.  kPackUnroll segments per thread
            // are fetched before the first is packed, so a pass pays the DRAM latency of its ~10 segments
            // per thread three times instead of ten
            constexpr int kPackUnroll = 4;
            for (uint32_t seg0 = tid; seg0 < nseg + 4; seg0 += kFusedThreads * kPackUnroll) {
                uint4 v[kPackUnroll];
                bool full[kPackUnroll];
#pragma unroll
                for (int u = 0; u < kPackUnroll; ++u) {
                    const uint32_t seg = seg0 + (uint32_t)u * kFusedThreads;
                    v[u] = make_uint4(0u, 0u, 0u, 0u);
                    full[u] = false;
                    if (seg >= nseg) continue;
                    if (PACKED) {
                        const uint32_t gw = (lo >> 4) + seg;
                        if ((uint64_t)gw * 16u < b.n_bases)
                            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;"
                                         : "=r"(v[u].x) : "l"(b.packed + gw), "l"(pol_first));
                    } else if (TMA) {
                        v[u] = *reinterpret_cast<const uint4*>(sm.raw + 16u * seg);
                        full[u] = true;
                    } else {
                        const uint64_t g = (uint64_t)lo + 16ull * seg;
                        full[u] = g + 16 <= b.n_bases;
                        if (full[u]) v[u] = load_stream16_ef(b.bases + g, pol_first);
                    }
                }
#pragma unroll
                for (int u = 0; u < kPackUnroll; ++u) {
                    const uint32_t seg = seg0 + (uint32_t)u * kFusedThreads;
                    if (seg >= nseg + 4) continue;
                    uint32_t word = 0;
                    if (PACKED) {
                        word = v[u].x;
                    } else if (seg < nseg) {
                        uint4 x = v[u];
                        if (!full[u]) {                        // last, partial segment of the batch
                            const uint64_t g = (uint64_t)lo + 16ull * seg;
                            uint32_t w[4] = {0, 0, 0, 0};
                            for (uint32_t j = 0; j < 16 && g + j < b.n_bases; ++j)
                                w[j >> 2] |= (uint32_t)b.bases[g + j] << (8 * (j & 3));
                            x = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                        word = pack16(x);
                        if (SEQ) {
                            const uint32_t m16 = not_acgt4(x.x) | (not_acgt4(x.y) << 4) | (not_acgt4(x.z) << 8) |
                                                 (not_acgt4(x.w) << 12);
                            if (m16) atomicOr(&s_bad[seg >> 1], m16 << ((seg & 1u) * 16u));
                        } else if (any_N4(x.x) | any_N4(x.y) | any_N4(x.z) | any_N4(x.w)) {
                            const uint32_t w[4] = {x.x, x.y, x.z, x.w};
                            uint32_t m16 = 0;
                            for (uint32_t j = 0; j < 16; ++j)
                                if (((w[j >> 2] >> (8 * (j & 3))) & 0xFFu) == 'N') m16 |= 1u << j;
                            atomicOr(&s_bad[seg >> 1], m16 << ((seg & 1u) * 16u));
                        }
                    }
                    s_packed[seg] = word;
                }
            }
            __syncthreads();
            // sm.raw is free again: stream in the next pass (of this tile, else of the next tile)
            if (TMA && tid == 0) {
                if (rb < R) prefetch(s_off[rb] & ~15u, s_off[R]);
                else if (ntile < n_tiles) prefetch(nt_lo, nt_end);
            }
            if (too_long) {
                if (tid == 0) ++st_long;
                ra = rb;
                continue;
            }

            if (SEQ) {
                // a byte that is not ACGT spoils the k windows that contain it: smear every flag over
                // the k-1 positions before it.  Read boundaries need no care, the last k-1 positions
                // of a read start no k-mer anyway.
                uint32_t smeared[(FusedSmem<TMA>::kWords / 2 + 2 + kFusedThreads - 1) / kFusedThreads];
                const uint32_t nbw = (nseg >> 1) + 1;
                int it = 0;
                for (uint32_t w = tid; w < nbw; w += kFusedThreads, ++it) {
                    const uint64_t pair = (uint64_t)s_bad[w] | ((uint64_t)s_bad[w + 1] << 32);
                    uint64_t d = pair;
                    for (int j = 1; j < k; ++j) d |= pair >> j;
                    smeared[it] = (uint32_t)d;
                }
                __syncthreads();
                it = 0;
                for (uint32_t w = tid; w < nbw; w += kFusedThreads, ++it) s_bad[w] = smeared[it];
                __syncthreads();
            }
            // (b) per read: containN (classify.cpp:182-185), positions that start no k-mer
            for (uint32_t r = ra + tid; r < rb; r += kFusedThreads) {
                const uint32_t s = s_off[r] - lo, e = s_off[r + 1] - lo, L = e - s;
                const bool has_n = SEQ ? false
                                 : PACKED ? ((b.has_n[(r0 + r) >> 5] >> ((r0 + r) & 31u)) & 1u) != 0u
                                          : any_bits(s_bad, s, e);
                if (SEQ && L < (uint32_t)k) {              // stage 03: the loop over windows is simply empty
                    set_bits(s_bad, s, e);
                } else if (has_n) {                        // classify.cpp:190-193: no votes at all
                    ++st_n;
                    set_bits(s_bad, s, e);
                } else if (L < (uint32_t)k) {              // kmer.h:171 assert in the reference
                    ++st_short;
                    set_bits(s_bad, s, e);
                } else {
                    st_lookups += L - (uint32_t)k + 1u;
                    set_bits(s_bad, e - (uint32_t)k + 1u, e);
                }
            }
            // positions before the first read of the pass and after the last one
            if (tid == 0) {
                set_bits(s_bad, 0, s_off[ra] - lo);
                set_bits(s_bad, hi - lo, nseg * 16u);
            }
            __syncthreads();

            // (d) exact probe of queued positions [base, base + n) of this warp's queue, n <= kDrainChunk,
            // kDrainUnroll probes in flight per lane
            auto drain = [&](uint32_t base, uint32_t n) {
                uint32_t p[kDrainUnroll];
                uint64_t want[kDrainUnroll];
                Bucket bk[kDrainUnroll];
                uint32_t bucket[kDrainUnroll];
                bool on[kDrainUnroll];
#pragma unroll
                for (int u = 0; u < kDrainUnroll; ++u) {
                    const uint32_t i = u * 32u + lane;
                    on[u] = i < n;
                    p[u] = on[u] ? wqueue[base + i] : 0u;
                }
                __syncwarp();                              // the next append may overwrite these entries
#pragma unroll
                for (int u = 0; u < kDrainUnroll; ++u) {
                    const uint64_t canon = canonical_at(s_packed, p[u], k, kmask);
                    const uint64_t h = table_hash(canon, k, kmask);
                    bucket[u] = (uint32_t)(h >> t.rem_bits);
                    want[u] = (h & t.rem_mask) << 4;
                    bk[u].s0 = bk[u].s1 = bk[u].s2 = bk[u].s3 = 0ull;
                    if (on[u]) bk[u] = load_bucket(t.slots + (size_t)bucket[u] * kSlotsPerBucket);
                }
#pragma unroll
                for (int u = 0; u < kDrainUnroll; ++u) {
                    bool found;
                    uint32_t tag = match_bucket(bk[u], want[u], found);
                    if (on[u] && !found && (bk[u].s0 & 1ull)) {           // overflowed home bucket
                        uint32_t bkt = bucket[u];
                        uint64_t w = want[u];
                        for (int d = 1; d <= kMaxDisp; ++d) {
                            bkt = (bkt + 1) & t.bucket_mask;
                            w += 1;
                            const Bucket nb = load_bucket(t.slots + (size_t)bkt * kSlotsPerBucket);
                            ++st_extra;
                            tag = match_bucket(nb, w, found);
                            if (found || !(nb.s0 & 1ull)) break;
                        }
                    }
                    if (on[u] && tag) vote(s_votes, s_off, ra, rb, lo + p[u], tag);
                }
            };
            uint32_t wq = 0;                               // entries in this warp's queue (same value in every lane)

            // (c) filter sweep: thread <-> packed word (16 positions)
            for (uint32_t wbase = 0; wbase < nseg; wbase += kFusedThreads) {
                const uint32_t wi = wbase + tid;
                uint32_t pass = 0;
                uint32_t valid = 0;
                if (wi < nseg) valid = ~(s_bad[wi >> 1] >> ((wi & 1u) * 16u)) & 0xFFFFu;
                if (valid) {
                    const uint32_t w0 = s_packed[wi], w1 = s_packed[wi + 1];
                    const uint32_t nxt = __funnelshift_l(s_packed[wi + nxt_word + 1], s_packed[wi + nxt_word], nxt_sh);
                    const uint64_t x = ((uint64_t)w0 << 32) | w1;
                    uint64_t fwd = km1 ? (x >> fwd_init_shift) : 0ull;
                    uint64_t rcv = revcomp_top(x, mask_km1) << 2;
                    // MINI: filter word index of each of the 16 positions, and the positions that must fetch
                    constexpr int kM = MINI ? mini_len(KT) : 1, kW = MINI ? KT - kM + 1 : 1, kT = 16 + kW - 1;
                    static_assert(!MINI || (kM == 16 && kT <= 32), "the MINI sweep works on 16-base windows of a 48-base stretch");
                    uint32_t idx[MINI ? 16 : 1];
                    uint32_t need = valid;
                    uint64_t cur = 0ull;
                    if (MINI) {
                        // Everything comes from the 32-bit windows f[t] = bases t .. t+15 of this thread's 48-base
                        // stretch and g[t] = their reverse complements (cut out of the stretch's reverse complement,
                        // base i <-> 47-i), t = 0 .. kT-1:
                        //   m-mer (m = 16) at t, canonical:  min(f[t], g[t])           -> mini_hash -> sliding minimum
                        //   k-mer at j: first / last 16 bases f[j], f[j+kW-1]; of its reverse complement g[j+kW-1], g[j]
                        //               -> mini_sel, a symmetric function of the two strands (no canonical compare)
                        const uint32_t w2 = s_packed[wi + 2];
                        const uint32_t r0 = revcomp16(w2), r1 = revcomp16(w1), r2 = revcomp16(w0);
                        uint32_t hm[kT], sa[16], selv[16];
#pragma unroll
                        for (int tt = 0; tt < kT; ++tt) {
                            const uint32_t f = tt == 0 ? w0 : tt < 16 ? __funnelshift_l(w1, w0, 2 * tt)
                                             : tt == 16 ? w1 : __funnelshift_l(w2, w1, 2 * (tt - 16));
                            const int q = 32 - tt;                 // first base of the mirrored window
                            const uint32_t g = q >= 32 ? r2 : q >= 16 ? __funnelshift_l(r2, r1, 2 * (q - 16))
                                                                      : __funnelshift_l(r1, r0, 2 * q);
                            hm[tt] = mini_hash(min(f, g));
                            if (tt < 16) sa[tt] = f * kSelC1 + g * kSelC2;
                            if (tt >= kW - 1) selv[tt - (kW - 1)] = (sa[tt - (kW - 1)] + f * kSelC2 + g * kSelC1) >> 22;
                        }
                        // sliding minimum over kW hashes: windows of 3, then of 9, then two or three of those
                        constexpr int kWd1 = kW >= 3 ? 3 : 1, kWd = kW >= 9 ? 9 : kWd1;
                        if (kW >= 3) {
#pragma unroll
                            for (int i = 0; i + 2 < kT; ++i) hm[i] = min(min(hm[i], hm[i + 1]), hm[i + 2]);
                        }
                        if (kW >= 9) {
#pragma unroll
                            for (int i = 0; i + 8 < kT; ++i) hm[i] = min(min(hm[i], hm[i + 3]), hm[i + 6]);
                        }
                        uint32_t chg = 1u;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const uint32_t mn = kWd == kW ? hm[j]
                                              : 2 * kWd >= kW ? min(hm[j], hm[j + kW - kWd])
                                                              : min(min(hm[j], hm[j + kWd]), hm[j + kW - kWd]);
                            idx[j] = mini_word(mn) >> t.filt_shift;
                            if (j) chg |= (uint32_t)(idx[j] != idx[j - 1]) << j;
                        }
                        need = valid & (chg | ~(valid << 1));      // word changed, or the position before holds none
                        st_loads += __popc(need);
                        // all fetches first; the selectors are packed two per register while they are in flight
                        uint64_t fw[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            fw[j] = load_filter_if(t.filt + idx[j], pol_last, (need >> j) & 1u);
                        uint32_t hbp[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) hbp[j] = selv[2 * j] | (selv[2 * j + 1] << 16);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if ((need >> j) & 1u) cur = fw[j];     // positions in between keep the last fetched word
                            const uint32_t sel = hbp[j >> 1] >> ((j & 1) * 16);
                            const uint32_t hit = ((uint32_t)cur >> ((sel >> 5) & 31u)) & ((uint32_t)(cur >> 32) >> (sel & 31u)) &
                                                 (valid >> j) & 1u;
                            pass |= hit << j;
                        }
                    } else {
                    st_loads += __popc(valid);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint64_t fw[8];
                        uint32_t hb[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int j = half * 8 + u;
                            const uint32_t c = (nxt >> (30 - 2 * j)) & 3u;
                            fwd = ((fwd << 2) | c) & kmask;
                            rcv = (rcv >> 2) | ((uint64_t)(c ^ 2u) << rc_shift);
                            const uint64_t canon = fwd < rcv ? fwd : rcv;
                            const FilterHash fh = filter_hash(canon);
                            hb[u] = fh.bits;
                            fw[u] = 0ull;
                            if ((valid >> j) & 1u) fw[u] = load_filter(t.filt + (fh.word >> t.filt_shift), pol_last);
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const uint32_t hit = ((uint32_t)fw[u] >> (hb[u] >> 27)) &
                                                 ((uint32_t)(fw[u] >> 32) >> ((hb[u] >> 22) & 31u)) & 1u;
                            pass |= hit << (half * 8 + u);
                        }
                    }
                    }
                }
                // append the passing positions to this WARP's queue (count kept in a register, no atomics), and
                // as soon as it holds a full round of exact probes (kDrainUnroll per lane) drain that round:
                // the warps of a CTA never wait for each other between sweeping and probing, and one warp's
                // HBM latency is covered by the sweeps of the others
                const uint32_t cnt = __popc(pass);
                uint32_t incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= (uint32_t)o) incl += v;
                }
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                if (total) {                               // warp-uniform
                    uint32_t q = wq + incl - cnt;
                    st_pass += cnt;
                    while (pass) {
                        const uint32_t j = __ffs(pass) - 1;
                        pass &= pass - 1;
                        wqueue[q++] = (uint16_t)(wi * 16u + j);
                    }
                    wq += total;                           // < kDrainChunk + 512 <= kWarpQueueCap
                    __syncwarp();
                    while (wq >= kDrainChunk) {
                        wq -= kDrainChunk;
                        drain(wq, kDrainChunk);
                    }
                }
            }
            if (wq) {                                      // warp-uniform
                drain(0u, wq);
                wq = 0;
            }
            __syncthreads();
            ra = rb;
        }

        // (e) votes -> per-barcode counters (IncrBarcodeHaps, classify.cpp:203-206)
        for (uint32_t rbase = 0; rbase < R; rbase += kFusedThreads) {
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
                if ((int)lane == leader) {
                    if (bc < n_barcodes) {
                        if (s0) atomicAdd(&counts[2ull * bc], s0);
                        if (s1) atomicAdd(&counts[2ull * bc + 1], s1);
                    } else {
                        ++st_badbc;
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
        st_pass += __shfl_xor_sync(0xFFFFFFFFu, st_pass, o);
        st_loads += __shfl_xor_sync(0xFFFFFFFFu, st_loads, o);
    }
    if (lane == 0) {
        if (st_lookups) atomicAdd(&stats->lookups, st_lookups);
        if (st_n) atomicAdd(&stats->reads_with_n, st_n);
        if (st_short) atomicAdd(&stats->reads_short, st_short);
        if (st_long) atomicAdd(&stats->reads_too_long, st_long);
        if (st_badbc) atomicAdd(&stats->bad_barcode, st_badbc);
        if (st_extra) atomicAdd(&stats->extra_probes, (unsigned long long)st_extra);
        if (st_pass) atomicAdd(&stats->filter_pass, (unsigned long long)st_pass);
        if (st_loads) atomicAdd(&stats->filter_loads, (unsigned long long)st_loads);
    }
}

}  // namespace hast
