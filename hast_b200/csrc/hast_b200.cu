// hast_b200.cu -- C ABI (include/hast_b200.h) over the sm_100a kernels.
//
// Host-side plumbing only: contexts, streams, the double-buffered H2D ring,
// launches, the NCCL reduce of the per-barcode counters.  No compute happens
// on the host and there is no CPU fallback: without a CUDA device every entry
// point that needs one fails with HAST_E_CUDA.
#include "../../include/hast_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "fused.cuh"
#include "kcount.cuh"
#include "host_pack.h"
#include <cub/device/device_radix_sort.cuh>

using namespace hast;

namespace {

constexpr int kSlots = 3;                 // device-side staging ring (double buffering + 1)

struct Slot {
    uint8_t* d_bases = nullptr;  size_t cap_bases = 0;
    uint32_t* d_off = nullptr;   size_t cap_off = 0;
    uint32_t* d_bc = nullptr;    size_t cap_bc = 0;
    uint32_t* d_hasn = nullptr;  size_t cap_hasn = 0;
    cudaEvent_t copied = nullptr, done = nullptr;
    // pinned staging of a batch packed on the host (option host_pack_threads)
    uint32_t* h_packed = nullptr; size_t cap_h_packed = 0;
    uint32_t* h_hasn = nullptr;   size_t cap_h_hasn = 0;
};

// ---- NCCL through dlopen: a single-GPU run never needs the library ---------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t,
                           cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
std::string g_global_err;

bool nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.ok) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return false;
#define HAST_SYM(field, name) \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name); \
    if (!g_nccl.field) return false;
    HAST_SYM(GetUniqueId, "ncclGetUniqueId")
    HAST_SYM(CommInitRank, "ncclCommInitRank")
    HAST_SYM(CommInitAll, "ncclCommInitAll")
    HAST_SYM(CommDestroy, "ncclCommDestroy")
    HAST_SYM(Reduce, "ncclReduce")
    HAST_SYM(GroupStart, "ncclGroupStart")
    HAST_SYM(GroupEnd, "ncclGroupEnd")
    HAST_SYM(GetErrorString, "ncclGetErrorString")
#undef HAST_SYM
    g_nccl.ok = true;
    return true;
}

}  // namespace

struct hast_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t cs = nullptr;            // compute stream
    cudaStream_t hs = nullptr;            // host-to-device copy stream
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    cudaEvent_t f0 = nullptr, f1 = nullptr;   // hast_finish: reduce / read-back timing

    TableView tv{};
    uint64_t n_buckets = 0;
    bool table_ready = false;

    int32_t* d_counts = nullptr;
    uint64_t n_barcodes = 0, cap_barcodes = 0;
    int32_t* d_reduced = nullptr;
    uint64_t cap_reduced = 0;

    DevStats* d_stats = nullptr;
    Slot slot[kSlots];
    uint64_t seq = 0;
    hast_stats st{};

    void* d_scratch = nullptr;
    size_t cap_scratch = 0;

    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;

    int tile_blocks = 0;                  // persistent grid of the tile kernels
    int extract_blocks = 0;               // ... of tile_kernel<MODE_EXTRACT> (fewer registers: more CTAs per SM)
    int fused_blocks = 0, fused_blocks_tma = 0;   // persistent grids of classify_kernel<*, false / true>
    uint64_t filt_words = 0;
    // options (hast_set_option)
    int64_t opt_kernel = 3;               // 3 = classify_kernel with the minimizer-addressed pre-filter (k = 21/25/31, the k with
                                          // mini_len(k) != 0 in table.cuh; every other k -- 17 included -- runs as 1), 4 = as 3 with TMA-staged reads,
                                          // 1 = classify_kernel (per-k-mer filter word), 2 = same with TMA-staged reads, 0 = tile_kernel<MODE_CLASSIFY>
    int64_t opt_seq_mode = 0;             // 1 = stage-03 window rule (classify_kernel<.., SEQ>)
    hastpack::Pool* pack_pool = nullptr;  // option host_pack_threads > 0: hast_submit_batch packs ASCII batches to 2 bits on the host
    int64_t l2_persist = -1;              // bytes of L2 set aside for persisting accesses (option l2_persist_bytes); -1 = default
    int64_t opt_reads_per_tile = 0;       // 0 = per batch, what fills one pass (fused_reads_per_tile); else fixed (tuning / tests)
    int64_t opt_filter_bits_per_key = 16;
    int64_t opt_filter_max_bytes = (int64_t)64 << 20;
    // stage 00 (kcount.cuh)
    KcView kc{};
    uint64_t kc_slots = 0;
    KcStats* d_kc_stats = nullptr;
    std::string err;
};

namespace {

int fail(hast_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_global_err = msg;
    return code;
}

#define CU(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(ctx, HAST_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));   \
    } while (0)

#define NC(call)                                                                                 \
    do {                                                                                         \
        ncclResult_t r_ = (call);                                                                \
        if (r_ != ncclSuccess)                                                                   \
            return fail(ctx, HAST_E_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
    } while (0)

int grid_for(const hast_ctx* c, uint64_t n, int threads, int per_sm = 8) {
    uint64_t want = (n + threads - 1) / threads;
    uint64_t cap = (uint64_t)c->sm_count * per_sm;
    return (int)std::max<uint64_t>(1, std::min(want, cap));
}

int ensure(hast_ctx* ctx, void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *p) return HAST_OK;
    if (*p) { CU(cudaStreamSynchronize(ctx->cs)); CU(cudaStreamSynchronize(ctx->hs)); CU(cudaFree(*p)); *p = nullptr; }
    size_t want = std::max<size_t>(bytes + bytes / 4, 1 << 16);
    CU(cudaMalloc(p, want));
    *cap = want;
    return HAST_OK;
}

int read_stats(hast_ctx* ctx, DevStats* out) {
    CU(cudaMemcpyAsync(out, ctx->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, ctx->cs));
    CU(cudaStreamSynchronize(ctx->cs));
    return HAST_OK;
}

int launch_tile(hast_ctx* ctx, int mode, const BatchView& bv_in, uint64_t* d_kmers, uint8_t* d_has_n) {
    BatchView bv = bv_in;
    const uint32_t n_tiles = (bv.n_reads + kReadsPerTile - 1) / kReadsPerTile;
    if (!n_tiles) return HAST_OK;
    const int grid = (int)std::min<uint32_t>(n_tiles, (uint32_t)ctx->tile_blocks);
    if (mode == MODE_CLASSIFY && ctx->opt_kernel >= 1) {
        const bool tma = (ctx->opt_kernel == 2 || ctx->opt_kernel == 4) && !bv.packed;
        bv.reads_per_tile = ctx->opt_reads_per_tile > 0
            ? (uint32_t)std::min<int64_t>(ctx->opt_reads_per_tile, kFusedReadsPerTile)
            : fused_reads_per_tile(bv.n_bases, bv.n_reads, tma ? FusedSmem<true>::kCap : FusedSmem<false>::kCap);
        const uint32_t n_ftiles = (bv.n_reads + bv.reads_per_tile - 1) / bv.reads_per_tile;
        const int fgrid = (int)std::min<uint32_t>(n_ftiles, (uint32_t)(tma ? ctx->fused_blocks_tma : ctx->fused_blocks));
        const uint32_t nbc = (uint32_t)std::min<uint64_t>(ctx->n_barcodes, 0xFFFFFFFFull);
#define HAST_LAUNCH_K(KT)                                                                                         \
    do {                                                                                                          \
        if (mini_len(KT) && ctx->tv.filt_m) {                                                                     \
            if (bv.packed) classify_kernel<KT, false, true, false, mini_len(KT) != 0><<<fgrid, kFusedThreads, sizeof(FusedSmem<false>), ctx->cs>>>( \
                    ctx->tv, bv, ctx->d_counts, nbc, ctx->d_stats);                                               \
            else if (tma) classify_kernel<KT, true, false, false, mini_len(KT) != 0><<<fgrid, kFusedThreads, sizeof(FusedSmem<true>), ctx->cs>>>( \
                    ctx->tv, bv, ctx->d_counts, nbc, ctx->d_stats);                                               \
            else classify_kernel<KT, false, false, false, mini_len(KT) != 0><<<fgrid, kFusedThreads, sizeof(FusedSmem<false>), ctx->cs>>>( \
                    ctx->tv, bv, ctx->d_counts, nbc, ctx->d_stats);                                               \
        } else                                                                                                    \
        if (bv.packed) classify_kernel<KT, false, true><<<fgrid, kFusedThreads, sizeof(FusedSmem<false>), ctx->cs>>>( \
                ctx->tv, bv, ctx->d_counts, nbc, ctx->d_stats);                                                   \
        else if (tma) classify_kernel<KT, true><<<fgrid, kFusedThreads, sizeof(FusedSmem<true>), ctx->cs>>>(       \
                ctx->tv, bv, ctx->d_counts, nbc, ctx->d_stats);                                                   \
        else classify_kernel<KT, false><<<fgrid, kFusedThreads, sizeof(FusedSmem<false>), ctx->cs>>>(              \
                ctx->tv, bv, ctx->d_counts, nbc, ctx->d_stats);                                                   \
    } while (0)
        if (ctx->opt_seq_mode) {
            if (bv.packed) return fail(ctx, HAST_E_STATE, "seq_mode takes ASCII batches");
            classify_kernel<0, false, false, true><<<fgrid, kFusedThreads, sizeof(FusedSmem<false>), ctx->cs>>>(
                ctx->tv, bv, ctx->d_counts, nbc, ctx->d_stats);
        } else
        switch (ctx->tv.k) {                      // specialised for HAST's default k and the benchmarked sweep
            case 17: HAST_LAUNCH_K(17); break;
            case 21: HAST_LAUNCH_K(21); break;
            case 25: HAST_LAUNCH_K(25); break;
            case 31: HAST_LAUNCH_K(31); break;
            default: HAST_LAUNCH_K(0); break;
        }
#undef HAST_LAUNCH_K
    } else if (mode == MODE_CLASSIFY)
        tile_kernel<MODE_CLASSIFY><<<grid, kTileThreads, 0, ctx->cs>>>(
            ctx->tv, bv, ctx->d_counts, (uint32_t)std::min<uint64_t>(ctx->n_barcodes, 0xFFFFFFFFull),
            ctx->d_stats, nullptr, nullptr);
    else
        tile_kernel<MODE_EXTRACT><<<(int)std::min<uint32_t>(n_tiles, (uint32_t)ctx->extract_blocks), kTileThreads, 0, ctx->cs>>>(ctx->tv, bv, nullptr, 0, ctx->d_stats,
                                                                     d_kmers, d_has_n);
    CU(cudaGetLastError());
    ctx->st.kernel_launches++;
    return HAST_OK;
}

// Both loaders stream their input through TWO halves of the scratch buffer: the insert kernel of one chunk runs while
// the next chunk is being copied (a 62 M-key list is 1.4 GB of text; round 1 synchronised after every chunk).
template <class Launch>
int stream_chunks(hast_ctx* ctx, const char* src, uint64_t n_items, uint64_t item_bytes, uint64_t chunk_items, Launch launch) {
    const size_t half = (size_t)(chunk_items * item_bytes + 255) & ~(size_t)255;
    int rc = ensure(ctx, &ctx->d_scratch, &ctx->cap_scratch, 2 * half);
    if (rc) return rc;
    cudaEvent_t done[2] = {nullptr, nullptr};
    for (auto& e : done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int h = 0;
    cudaError_t err = cudaSuccess;
    for (uint64_t at = 0; at < n_items && err == cudaSuccess; at += chunk_items, h ^= 1) {
        const uint64_t n = std::min(chunk_items, n_items - at);
        char* d = (char*)ctx->d_scratch + (size_t)h * half;
        err = cudaEventSynchronize(done[h]);                       // the kernel that last read this half (no-op the first time)
        if (err == cudaSuccess) err = cudaMemcpyAsync(d, src + at * item_bytes, n * item_bytes, cudaMemcpyHostToDevice, ctx->hs);
        if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->hs);   // pageable source: the caller may free it after we return
        if (err != cudaSuccess) break;
        ctx->st.h2d_bytes += n * item_bytes;
        launch(d, n);
        err = cudaGetLastError();
        ctx->st.kernel_launches++;
        if (err == cudaSuccess) err = cudaEventRecord(done[h], ctx->cs);
    }
    if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->cs);
    for (auto& e : done) cudaEventDestroy(e);
    if (err != cudaSuccess) return fail(ctx, HAST_E_CUDA, std::string("table load: ") + cudaGetErrorString(err));
    return HAST_OK;
}
}  // namespace

extern "C" {

int hast_abi_version(void) { return HAST_ABI_VERSION; }

int hast_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* hast_last_error(const hast_ctx* ctx) { return ctx ? ctx->err.c_str() : g_global_err.c_str(); }
int hast_device(const hast_ctx* ctx) { return ctx ? ctx->device : -1; }

int hast_create(int device, hast_ctx** out) {
    hast_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, HAST_E_ARG, "hast_create: out is NULL");
    *out = nullptr;
    int n = hast_device_count();
    if (n <= 0) return fail(nullptr, HAST_E_CUDA, "no CUDA device: this library has no CPU path");
    if (device < 0 || device >= n) return fail(nullptr, HAST_E_ARG, "hast_create: bad device index");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(nullptr, HAST_E_CUDA, "device is not sm_100-class (built for sm_100a only)");
    ctx = new hast_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    cudaError_t e;
#define CU_NEW(call) if ((e = (call)) != cudaSuccess) { g_global_err = std::string(#call) + ": " + cudaGetErrorString(e); delete ctx; return HAST_E_CUDA; }
    CU_NEW(cudaStreamCreateWithFlags(&ctx->cs, cudaStreamNonBlocking));
    CU_NEW(cudaStreamCreateWithFlags(&ctx->hs, cudaStreamNonBlocking));
    CU_NEW(cudaEventCreate(&ctx->t0));
    CU_NEW(cudaEventCreate(&ctx->t1));
    CU_NEW(cudaEventCreate(&ctx->f0));
    CU_NEW(cudaEventCreate(&ctx->f1));
    for (auto& s : ctx->slot) {
        CU_NEW(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
        CU_NEW(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    CU_NEW(cudaMalloc(&ctx->d_stats, sizeof(DevStats)));
    CU_NEW(cudaMemset(ctx->d_stats, 0, sizeof(DevStats)));
    int per_sm = 0;
    CU_NEW(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_kernel<MODE_CLASSIFY>, kTileThreads, 0));
    int per_sm_x = 0;
    CU_NEW(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_x, tile_kernel<MODE_EXTRACT>, kTileThreads, 0));
    // classify_kernel keeps its pass in dynamic shared memory (opt-in above 48 KiB for the TMA variant)
#define HAST_ATTR(KT)                                                                                              \
    CU_NEW(cudaFuncSetAttribute(classify_kernel<KT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                (int)sizeof(FusedSmem<false>)));                                                   \
    CU_NEW(cudaFuncSetAttribute(classify_kernel<KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                (int)sizeof(FusedSmem<true>)));                                                    \
    CU_NEW(cudaFuncSetAttribute(classify_kernel<KT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                (int)sizeof(FusedSmem<false>)));
    HAST_ATTR(0) HAST_ATTR(17) HAST_ATTR(21) HAST_ATTR(25) HAST_ATTR(31)
#undef HAST_ATTR
#define HAST_ATTR(KT)                                                                                              \
    CU_NEW(cudaFuncSetAttribute(classify_kernel<KT, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)sizeof(FusedSmem<false>)));                                                   \
    CU_NEW(cudaFuncSetAttribute(classify_kernel<KT, false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                (int)sizeof(FusedSmem<false>)));                                                   \
    CU_NEW(cudaFuncSetAttribute(classify_kernel<KT, true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                (int)sizeof(FusedSmem<true>)));
    HAST_ATTR(21) HAST_ATTR(25) HAST_ATTR(31)
#undef HAST_ATTR
    CU_NEW(cudaFuncSetAttribute(classify_kernel<0, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(FusedSmem<false>)));
    int per_sm_f = 0, per_sm_t = 0;
    CU_NEW(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_f, classify_kernel<0, false>, kFusedThreads,
                                                         sizeof(FusedSmem<false>)));
    CU_NEW(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_t, classify_kernel<0, true>, kFusedThreads,
                                                         sizeof(FusedSmem<true>)));
    ctx->fused_blocks_tma = std::max(1, per_sm_t) * prop.multiProcessorCount;
#undef CU_NEW
    ctx->tile_blocks = std::max(1, per_sm) * ctx->sm_count;
    ctx->extract_blocks = std::max(1, per_sm_x) * ctx->sm_count;
    ctx->fused_blocks = std::max(1, per_sm_f) * ctx->sm_count;
    *out = ctx;
    return HAST_OK;
}

void hast_destroy(hast_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->comm && g_nccl.ok) g_nccl.CommDestroy(ctx->comm);
    if (ctx->pack_pool) hastpack::pool_destroy(ctx->pack_pool);
    for (auto& s : ctx->slot) {
        cudaFree(s.d_bases); cudaFree(s.d_off); cudaFree(s.d_bc); cudaFree(s.d_hasn);
        if (s.h_packed) cudaFreeHost(s.h_packed);
        if (s.h_hasn) cudaFreeHost(s.h_hasn);
        if (s.copied) cudaEventDestroy(s.copied);
        if (s.done) cudaEventDestroy(s.done);
    }
    cudaFree(ctx->tv.slots);
    cudaFree(ctx->tv.filt);
    cudaFree(ctx->d_counts);
    cudaFree(ctx->d_reduced);
    cudaFree(ctx->d_stats);
    cudaFree(ctx->d_scratch);
    cudaFree(ctx->kc.slots);
    cudaFree(ctx->d_kc_stats);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    if (ctx->f0) cudaEventDestroy(ctx->f0);
    if (ctx->f1) cudaEventDestroy(ctx->f1);
    if (ctx->cs) cudaStreamDestroy(ctx->cs);
    if (ctx->hs) cudaStreamDestroy(ctx->hs);
    delete ctx;
}

int hast_set_option(hast_ctx* ctx, const char* name, int64_t value) {
    if (!ctx || !name) return fail(ctx, HAST_E_ARG, "NULL argument");
    const std::string n(name);
    if (n == "kernel") {
        if (value < 0 || value > 4)
            return fail(ctx, HAST_E_ARG, "kernel: 0 (direct probe), 1 (pre-filter), 2 (pre-filter, TMA-staged reads), "
                                         "3 (pre-filter addressed by minimizer), 4 (as 3, TMA-staged reads)");
        if (value == 0 && ctx->opt_seq_mode) return fail(ctx, HAST_E_STATE, "seq_mode needs the pre-filtered kernel");
        // the pre-filter of an existing table keeps its addressing scheme (tv.filt_m, which is what the launch
        // goes by): 1/2 <-> 3 takes effect at the next hast_table_begin
        ctx->opt_kernel = value;
    } else if (n == "seq_mode") {
        if (value != 0 && value != 1) return fail(ctx, HAST_E_ARG, "seq_mode: 0 or 1");
        if (value && ctx->opt_kernel == 0) return fail(ctx, HAST_E_STATE, "seq_mode needs the pre-filtered kernel");
        if (value && ctx->table_ready && ctx->tv.filt_m)
            return fail(ctx, HAST_E_STATE, "seq_mode: set before hast_table_begin (the pre-filter layout differs)");
        ctx->opt_seq_mode = value;
    } else if (n == "reads_per_tile") {
        if (value < 0 || value > kFusedReadsPerTile) return fail(ctx, HAST_E_ARG, "reads_per_tile: 0 (automatic) .. " + std::to_string(kFusedReadsPerTile));
        ctx->opt_reads_per_tile = value;
    } else if (n == "filter_bits_per_key") {
        if (value < 1 || value > 64) return fail(ctx, HAST_E_ARG, "filter_bits_per_key: 1..64");
        ctx->opt_filter_bits_per_key = value;
    } else if (n == "filter_max_bytes") {
        if (value < 128) return fail(ctx, HAST_E_ARG, "filter_max_bytes: >= 128");
        ctx->opt_filter_max_bytes = value;
    } else if (n == "l2_fetch_granularity") {
        // device-wide hint (cudaLimitMaxL2FetchGranularity): bytes fetched from HBM on an L2 miss.  The table
        // probes and count-table updates touch one 32-byte sector per access.
        if (value != 32 && value != 64 && value != 128) return fail(ctx, HAST_E_ARG, "l2_fetch_granularity: 32, 64 or 128");
        CU(cudaSetDevice(ctx->device));
        CU(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
    } else if (n == "host_pack_threads") {
        // hast_submit_batch (ASCII host buffers): pack the batch to 2 bits on `value` host threads and send a quarter of
        // the bytes over PCIe (what bin/classify's parser does while parsing).  0 = send the ASCII bytes (default).
        if (value < 0 || value > 256) return fail(ctx, HAST_E_ARG, "host_pack_threads: 0..256");
        if (ctx->pack_pool) { hastpack::pool_destroy(ctx->pack_pool); ctx->pack_pool = nullptr; }
        if (value > 0) ctx->pack_pool = hastpack::pool_create((int)value);
    } else if (n == "l2_persist_bytes") {
        // device-wide: L2 set aside for persisting accesses (cudaLimitPersistingL2CacheSize).  The pre-filter words are
        // loaded with an L2 evict_last policy; without a set-aside the hardware has nowhere to keep them apart from the
        // table sectors and read bytes that stream through.  Clamped to what the device allows; 0 = none.
        if (value < 0) return fail(ctx, HAST_E_ARG, "l2_persist_bytes: >= 0");
        CU(cudaSetDevice(ctx->device));
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, ctx->device));
        const size_t want = std::min<size_t>((size_t)value, (size_t)prop.persistingL2CacheMaxSize);
        CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
        ctx->l2_persist = (int64_t)want;
    } else {
        return fail(ctx, HAST_E_ARG, "unknown option: " + n);
    }
    return HAST_OK;
}

int hast_host_alloc(void** ptr, size_t bytes) {
    hast_ctx* ctx = nullptr;
    if (!ptr) return fail(nullptr, HAST_E_ARG, "hast_host_alloc: ptr is NULL");
    CU(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return HAST_OK;
}
int hast_host_free(void* ptr) {
    hast_ctx* ctx = nullptr;
    if (ptr) CU(cudaFreeHost(ptr));
    return HAST_OK;
}

// ---- K1 ------------------------------------------------------------------
int hast_table_begin(hast_ctx* ctx, int k, uint64_t expected_keys) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (k < 1 || k > 32) return fail(ctx, HAST_E_K, "k must be in 1..32 (reference is only correct for k <= 32, kmer.h:225-238)");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->cs));
    if (ctx->tv.slots) { CU(cudaFree(ctx->tv.slots)); ctx->tv.slots = nullptr; }
    if (ctx->tv.filt) { CU(cudaFree(ctx->tv.filt)); ctx->tv.filt = nullptr; }
    ctx->table_ready = false;
    // load factor <= 0.5 over 4-slot buckets => at least expected/2 buckets, power of two
    int b = 4;
    while (((uint64_t)1 << b) * 2 < expected_keys && b < 40) ++b;
    const int bmin = std::max(0, 2 * k - 57);          // rem must fit the slot
    const int bmax = 2 * k;                            // bucket index comes out of the 2k-bit hash
    b = std::max(b, bmin);
    b = std::min(b, bmax);
    if (b > 32) return fail(ctx, HAST_E_ARG, "table too large");
    ctx->n_buckets = (uint64_t)1 << b;
    const size_t bytes = ctx->n_buckets * kSlotsPerBucket * sizeof(uint64_t);
    CU(cudaMalloc(&ctx->tv.slots, bytes));
    CU(cudaMemsetAsync(ctx->tv.slots, 0, bytes, ctx->cs));
    ctx->tv.k = k;
    ctx->tv.kmask = kmer_mask(k);
    ctx->tv.rem_bits = 2 * k - b;
    ctx->tv.rem_mask = ctx->tv.rem_bits ? (((uint64_t)1 << ctx->tv.rem_bits) - 1) : 0;
    ctx->tv.bucket_mask = (uint32_t)(ctx->n_buckets - 1);
    // pre-filter: opt_filter_bits_per_key bits per expected key, power of two words, capped so that it
    // stays L2-resident (fused.cuh); 2^4 .. 2^26 words
    int fb = 4;
    while (fb < 26 && ((uint64_t)64 << fb) < expected_keys * (uint64_t)ctx->opt_filter_bits_per_key) ++fb;
    while (fb > 4 && ((uint64_t)8 << fb) > (uint64_t)ctx->opt_filter_max_bytes) --fb;
    ctx->filt_words = (uint64_t)1 << fb;
    CU(cudaMalloc(&ctx->tv.filt, ctx->filt_words * 8));
    CU(cudaMemsetAsync(ctx->tv.filt, 0, ctx->filt_words * 8, ctx->cs));
    ctx->tv.filt_shift = 32 - fb;
    // The filter words are loaded with an L2 evict_last policy (fused.cuh).  That only has teeth when part of L2 is set
    // aside for persisting accesses: with 64 MiB set aside the 64 MiB filter of a human-scale table keeps its place next
    // to 1 GiB of table sectors and 160 MB of counters (B200, profiles/r02_f_l2_persist.txt: 250.1 -> 262.0 G lookups/s;
    // the 16 MiB filter of a 128 MiB table fits L2 anyway: 276.8 -> 276.7).  Device-wide limit; the option overrides it.
    if (ctx->l2_persist < 0) {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, ctx->device));
        const size_t want = std::min<size_t>((size_t)64 << 20, (size_t)prop.persistingL2CacheMaxSize);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) cudaGetLastError();
    }
    // minimizer-addressed filter for the k with a MINI sweep (fused.cuh; kernel 3 and its TMA-staged form 4);
    // the stage-03 window rule, kernels 1/2 and every other k use the per-k-mer filter word
    ctx->tv.filt_m = (ctx->opt_kernel >= 3 && !ctx->opt_seq_mode) ? (uint32_t)mini_len(k) : 0u;
    CU(cudaMemsetAsync(ctx->d_stats, 0, sizeof(DevStats), ctx->cs));
    if (ctx->d_counts) CU(cudaMemsetAsync(ctx->d_counts, 0, ctx->cap_barcodes * 2 * sizeof(int32_t), ctx->cs));
    ctx->table_ready = true;
    return HAST_OK;
}

static int table_check(hast_ctx* ctx) {
    DevStats ds;
    int rc = read_stats(ctx, &ds);
    if (rc) return rc;
    if (ds.bad_kmer_lines)
        return fail(ctx, HAST_E_KMER_LINE, std::to_string(ds.bad_kmer_lines) +
                    " k-mer line(s) whose length differs from k=" + std::to_string(ctx->tv.k));
    if (ds.table_full)
        return fail(ctx, HAST_E_TABLE_FULL, std::to_string(ds.table_full) +
                    " k-mer(s) could not be placed: rebuild with a larger expected_keys");
    return HAST_OK;
}

int hast_table_add_text(hast_ctx* ctx, const char* text, uint64_t n_lines, int parent) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (!ctx->table_ready) return fail(ctx, HAST_E_STATE, "hast_table_begin first");
    if (parent < 0 || parent > 1) return fail(ctx, HAST_E_ARG, "parent must be 0 or 1");
    if (!n_lines) return HAST_OK;
    if (!text) return fail(ctx, HAST_E_ARG, "text is NULL");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->cs));                            // hast_table_begin's memsets
    const uint64_t stride = (uint64_t)ctx->tv.k + 1;
    const uint64_t chunk_lines = std::max<uint64_t>(1, ((uint64_t)128 << 20) / stride);
    int rc = stream_chunks(ctx, text, n_lines, stride, chunk_lines, [&](char* d, uint64_t n) {
        table_insert_text_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->cs>>>(ctx->tv, d, n, (uint32_t)parent, ctx->d_stats);
    });
    if (rc) return rc;
    return table_check(ctx);
}

int hast_table_add_packed(hast_ctx* ctx, const uint64_t* kmers, uint64_t n, int parent) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (!ctx->table_ready) return fail(ctx, HAST_E_STATE, "hast_table_begin first");
    if (parent < 0 || parent > 1) return fail(ctx, HAST_E_ARG, "parent must be 0 or 1");
    if (!n) return HAST_OK;
    if (!kmers) return fail(ctx, HAST_E_ARG, "kmers is NULL");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->cs));
    int rc = stream_chunks(ctx, (const char*)kmers, n, 8, (uint64_t)16 << 20, [&](char* d, uint64_t m) {
        table_insert_packed_kernel<<<grid_for(ctx, m, 256), 256, 0, ctx->cs>>>(ctx->tv, (const uint64_t*)d, m, (uint32_t)parent, ctx->d_stats);
    });
    if (rc) return rc;
    return table_check(ctx);
}

int hast_table_erase_seq(hast_ctx* ctx, const char* seq, uint32_t len, uint64_t* erased_out,
                         uint8_t* tags_out, uint32_t cap, uint32_t* n_erased) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (!ctx->table_ready) return fail(ctx, HAST_E_STATE, "hast_table_begin first");
    if (!seq) return fail(ctx, HAST_E_ARG, "seq is NULL");
    if (len < (uint32_t)ctx->tv.k)                      // chopRead2Kmer assert, kmer.h:171
        return fail(ctx, HAST_E_SHORT_READ, "adaptor shorter than k");
    CU(cudaSetDevice(ctx->device));
    const uint32_t nk = len - (uint32_t)ctx->tv.k + 1;
    const size_t o_seq = 0, o_er = (len + 15) & ~(size_t)15, o_tag = o_er + (size_t)nk * 8,
                 o_n = (o_tag + nk + 15) & ~(size_t)15, total = o_n + 16;
    int rc = ensure(ctx, &ctx->d_scratch, &ctx->cap_scratch, total);
    if (rc) return rc;
    char* d = (char*)ctx->d_scratch;
    CU(cudaMemcpyAsync(d + o_seq, seq, len, cudaMemcpyHostToDevice, ctx->cs));
    table_erase_kernel<<<1, 32, 0, ctx->cs>>>(ctx->tv, d + o_seq, len, (uint64_t*)(d + o_er),
                                              (uint8_t*)(d + o_tag), nk, (uint32_t*)(d + o_n));
    CU(cudaGetLastError());
    ctx->st.kernel_launches++;
    uint32_t n = 0;
    CU(cudaMemcpyAsync(&n, d + o_n, 4, cudaMemcpyDeviceToHost, ctx->cs));
    CU(cudaStreamSynchronize(ctx->cs));
    const uint32_t m = std::min(n, cap);
    if (m && erased_out) CU(cudaMemcpy(erased_out, d + o_er, (size_t)m * 8, cudaMemcpyDeviceToHost));
    if (m && tags_out) CU(cudaMemcpy(tags_out, d + o_tag, m, cudaMemcpyDeviceToHost));
    if (n_erased) *n_erased = n;
    return HAST_OK;
}

int hast_table_info_get(hast_ctx* ctx, hast_table_info* out) {
    if (!ctx || !out) return fail(ctx, HAST_E_ARG, "NULL argument");
    if (!ctx->table_ready) return fail(ctx, HAST_E_STATE, "hast_table_begin first");
    CU(cudaSetDevice(ctx->device));
    int rc = ensure(ctx, &ctx->d_scratch, &ctx->cap_scratch, sizeof(TableCounts));
    if (rc) return rc;
    CU(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(TableCounts), ctx->cs));
    table_count_kernel<<<grid_for(ctx, ctx->n_buckets, 256), 256, 0, ctx->cs>>>(
        ctx->tv.slots, ctx->n_buckets, (TableCounts*)ctx->d_scratch);
    CU(cudaGetLastError());
    ctx->st.kernel_launches++;
    TableCounts tc;
    CU(cudaMemcpyAsync(&tc, ctx->d_scratch, sizeof(tc), cudaMemcpyDeviceToHost, ctx->cs));
    CU(cudaStreamSynchronize(ctx->cs));
    out->k = ctx->tv.k;
    out->log2_buckets = 2 * ctx->tv.k - ctx->tv.rem_bits;
    out->n_buckets = ctx->n_buckets;
    out->bytes = ctx->n_buckets * kSlotsPerBucket * 8;
    out->n_entries = tc.entries;
    out->n_displaced = tc.displaced;
    out->n_overflow_buckets = tc.overflow_buckets;
    out->size[0] = tc.size0;
    out->size[1] = tc.size1;
    out->filter_bytes = ctx->filt_words * 8;
    return HAST_OK;
}

int hast_table_clone(hast_ctx* dst, hast_ctx* src) {
    hast_ctx* ctx = dst;
    if (!dst || !src) return fail(dst, HAST_E_ARG, "NULL context");
    if (!src->table_ready) return fail(dst, HAST_E_STATE, "source table not built");
    // the filter layout must match the kernel the destination will launch (hast_set_option guards a
    // context's own table the same way)
    if (dst->opt_seq_mode && src->tv.filt_m)
        return fail(dst, HAST_E_STATE, "hast_table_clone: the source pre-filter is minimizer-addressed, the destination is in seq_mode");
    if (dst->opt_kernel == 0 && src->opt_seq_mode)
        return fail(dst, HAST_E_STATE, "hast_table_clone: a seq_mode table needs the pre-filtered kernel");
    CU(cudaSetDevice(src->device));
    CU(cudaStreamSynchronize(src->cs));
    CU(cudaSetDevice(dst->device));
    CU(cudaStreamSynchronize(dst->cs));
    if (dst->tv.slots) { CU(cudaFree(dst->tv.slots)); dst->tv.slots = nullptr; }
    if (dst->tv.filt) { CU(cudaFree(dst->tv.filt)); dst->tv.filt = nullptr; }
    const size_t bytes = src->n_buckets * kSlotsPerBucket * 8;
    uint64_t *p = nullptr, *f = nullptr;
    CU(cudaMalloc(&p, bytes));
    CU(cudaMalloc(&f, src->filt_words * 8));
    CU(cudaMemcpyPeer(p, dst->device, src->tv.slots, src->device, bytes));
    CU(cudaMemcpyPeer(f, dst->device, src->tv.filt, src->device, src->filt_words * 8));
    dst->tv = src->tv;
    dst->tv.slots = p;
    dst->tv.filt = f;
    dst->filt_words = src->filt_words;
    dst->n_buckets = src->n_buckets;
    dst->table_ready = true;
    return HAST_OK;
}

// ---- K4 state ------------------------------------------------------------
int hast_reserve_barcodes(hast_ctx* ctx, uint64_t n) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (n > 0xFFFFFFFFull) return fail(ctx, HAST_E_ARG, "more than 2^32 barcodes");
    CU(cudaSetDevice(ctx->device));
    if (n > ctx->cap_barcodes) {
        const uint64_t cap = std::max<uint64_t>(n + n / 2, 1024);
        int32_t* p = nullptr;
        CU(cudaMalloc(&p, cap * 2 * sizeof(int32_t)));
        CU(cudaMemsetAsync(p, 0, cap * 2 * sizeof(int32_t), ctx->cs));
        if (ctx->d_counts) {
            CU(cudaMemcpyAsync(p, ctx->d_counts, ctx->cap_barcodes * 2 * sizeof(int32_t),
                               cudaMemcpyDeviceToDevice, ctx->cs));
            CU(cudaStreamSynchronize(ctx->cs));
            CU(cudaFree(ctx->d_counts));
        }
        ctx->d_counts = p;
        ctx->cap_barcodes = cap;
    }
    ctx->n_barcodes = std::max(ctx->n_barcodes, n);
    return HAST_OK;
}

int hast_reset_counts(hast_ctx* ctx) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    CU(cudaSetDevice(ctx->device));
    if (ctx->d_counts) CU(cudaMemsetAsync(ctx->d_counts, 0, ctx->cap_barcodes * 2 * sizeof(int32_t), ctx->cs));
    CU(cudaMemsetAsync(ctx->d_stats, 0, sizeof(DevStats), ctx->cs));
    memset(&ctx->st, 0, sizeof(ctx->st));
    return HAST_OK;
}

// ---- batches ---------------------------------------------------------------
static int batch_args_ok(hast_ctx* ctx, const void* bases, const void* off, const void* bc, uint64_t n_bases,
                         uint32_t n_reads) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (!ctx->table_ready) return fail(ctx, HAST_E_STATE, "no table: hast_table_begin/add first");
    if (n_reads && (!off || !bc || (!bases && n_bases))) return fail(ctx, HAST_E_ARG, "NULL batch array");
    if (n_bases >= 0xFFFFFFF0ull) return fail(ctx, HAST_E_ARG, "batch larger than 4 GiB of bases");
    if (!ctx->d_counts) return fail(ctx, HAST_E_STATE, "hast_reserve_barcodes first");
    return HAST_OK;
}

int hast_submit_batch(hast_ctx* ctx, const uint8_t* bases, uint64_t n_bases, const uint32_t* read_off,
                      const uint32_t* barcode_id, uint32_t n_reads, uint64_t* ticket) {
    int rc = batch_args_ok(ctx, bases, read_off, barcode_id, n_bases, n_reads);
    if (rc) return rc;
    if (ticket) *ticket = ctx->seq;
    if (!n_reads) return HAST_OK;
    CU(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[ctx->seq % kSlots];
    CU(cudaEventSynchronize(s.done));                   // the kernel that last used this slot
    if (ctx->pack_pool && ctx->opt_kernel >= 1 && !ctx->opt_seq_mode) {
        // pack on the host, copy a quarter of the bytes: the caller's buffer is consumed before this call returns
        const size_t n_words = (size_t)((n_bases + 15) / 16), n_flag = ((size_t)n_reads + 31) / 32;
        if (s.cap_h_packed < n_words * 4 + 16) {
            if (s.h_packed) CU(cudaFreeHost(s.h_packed));
            s.cap_h_packed = (n_words * 4 + 16) * 5 / 4;
            CU(cudaHostAlloc((void**)&s.h_packed, s.cap_h_packed, cudaHostAllocPortable));
        }
        if (s.cap_h_hasn < n_flag * 4 + 16) {
            if (s.h_hasn) CU(cudaFreeHost(s.h_hasn));
            s.cap_h_hasn = (n_flag * 4 + 16) * 5 / 4;
            CU(cudaHostAlloc((void**)&s.h_hasn, s.cap_h_hasn, cudaHostAllocPortable));
        }
        hastpack::pack_batch(ctx->pack_pool, bases, n_bases, read_off, n_reads, s.h_packed, s.h_hasn);
        if ((rc = ensure(ctx, (void**)&s.d_bases, &s.cap_bases, n_words * 4 + 16))) return rc;
        if ((rc = ensure(ctx, (void**)&s.d_off, &s.cap_off, ((size_t)n_reads + 1) * 4))) return rc;
        if ((rc = ensure(ctx, (void**)&s.d_bc, &s.cap_bc, (size_t)n_reads * 4))) return rc;
        if ((rc = ensure(ctx, (void**)&s.d_hasn, &s.cap_hasn, n_flag * 4))) return rc;
        CU(cudaMemcpyAsync(s.d_bases, s.h_packed, n_words * 4, cudaMemcpyHostToDevice, ctx->hs));
        CU(cudaMemcpyAsync(s.d_off, read_off, ((size_t)n_reads + 1) * 4, cudaMemcpyHostToDevice, ctx->hs));
        CU(cudaMemcpyAsync(s.d_bc, barcode_id, (size_t)n_reads * 4, cudaMemcpyHostToDevice, ctx->hs));
        CU(cudaMemcpyAsync(s.d_hasn, s.h_hasn, n_flag * 4, cudaMemcpyHostToDevice, ctx->hs));
        CU(cudaEventRecord(s.copied, ctx->hs));
        CU(cudaStreamWaitEvent(ctx->cs, s.copied, 0));
        BatchView pv{nullptr, s.d_off, s.d_bc, n_bases, n_reads, (const uint32_t*)s.d_bases, s.d_hasn};
        if ((rc = launch_tile(ctx, MODE_CLASSIFY, pv, nullptr, nullptr))) return rc;
        CU(cudaEventRecord(s.done, ctx->cs));
        ctx->st.h2d_bytes += n_words * 4 + ((size_t)n_reads + 1) * 4 + (size_t)n_reads * 4 + n_flag * 4;
        ctx->st.batches++;
        ctx->st.reads += n_reads;
        ctx->st.bases += n_bases;
        ctx->seq++;
        return HAST_OK;
    }
    if ((rc = ensure(ctx, (void**)&s.d_bases, &s.cap_bases, n_bases + 16))) return rc;
    if ((rc = ensure(ctx, (void**)&s.d_off, &s.cap_off, ((size_t)n_reads + 1) * 4))) return rc;
    if ((rc = ensure(ctx, (void**)&s.d_bc, &s.cap_bc, (size_t)n_reads * 4))) return rc;
    CU(cudaMemcpyAsync(s.d_bases, bases, n_bases, cudaMemcpyHostToDevice, ctx->hs));
    CU(cudaMemcpyAsync(s.d_off, read_off, ((size_t)n_reads + 1) * 4, cudaMemcpyHostToDevice, ctx->hs));
    CU(cudaMemcpyAsync(s.d_bc, barcode_id, (size_t)n_reads * 4, cudaMemcpyHostToDevice, ctx->hs));
    CU(cudaEventRecord(s.copied, ctx->hs));
    CU(cudaStreamWaitEvent(ctx->cs, s.copied, 0));
    BatchView bv{s.d_bases, s.d_off, s.d_bc, n_bases, n_reads, nullptr, nullptr};
    if ((rc = launch_tile(ctx, MODE_CLASSIFY, bv, nullptr, nullptr))) return rc;
    CU(cudaEventRecord(s.done, ctx->cs));
    ctx->st.h2d_bytes += n_bases + ((size_t)n_reads + 1) * 4 + (size_t)n_reads * 4;
    ctx->st.batches++;
    ctx->st.reads += n_reads;
    ctx->st.bases += n_bases;
    ctx->seq++;
    return HAST_OK;
}

int hast_wait_copied(hast_ctx* ctx, uint64_t ticket) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (ticket >= ctx->seq) return HAST_OK;             // empty batch or nothing submitted
    if (ctx->seq - ticket > kSlots) return HAST_OK;     // slot already recycled => copy long done
    CU(cudaEventSynchronize(ctx->slot[ticket % kSlots].copied));
    return HAST_OK;
}

int hast_submit_batch_device(hast_ctx* ctx, const uint8_t* d_bases, uint64_t n_bases,
                             const uint32_t* d_read_off, const uint32_t* d_barcode_id, uint32_t n_reads) {
    int rc = batch_args_ok(ctx, d_bases, d_read_off, d_barcode_id, n_bases, n_reads);
    if (rc) return rc;
    if (!n_reads) return HAST_OK;
    if ((uintptr_t)d_bases & 15) return fail(ctx, HAST_E_ARG, "d_bases must be 16-byte aligned");
    CU(cudaSetDevice(ctx->device));
    BatchView bv{d_bases, d_read_off, d_barcode_id, n_bases, n_reads, nullptr, nullptr};
    if ((rc = launch_tile(ctx, MODE_CLASSIFY, bv, nullptr, nullptr))) return rc;
    ctx->st.batches++;
    ctx->st.reads += n_reads;
    ctx->st.bases += n_bases;
    return HAST_OK;
}

// ---- host-packed batches -----------------------------------------------------
int hast_submit_batch_packed(hast_ctx* ctx, const uint32_t* packed, uint64_t n_bases, const uint32_t* read_off,
                             const uint32_t* barcode_id, const uint32_t* has_n, uint32_t n_reads, uint64_t* ticket) {
    int rc = batch_args_ok(ctx, packed, read_off, barcode_id, n_bases, n_reads);
    if (rc) return rc;
    if (n_reads && !has_n) return fail(ctx, HAST_E_ARG, "NULL batch array");
    if (ctx->opt_kernel == 0) return fail(ctx, HAST_E_STATE, "packed batches need the pre-filtered kernel (option kernel >= 1)");
    if (ticket) *ticket = ctx->seq;
    if (!n_reads) return HAST_OK;
    CU(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[ctx->seq % kSlots];
    CU(cudaEventSynchronize(s.done));
    const size_t n_words = (size_t)((n_bases + 15) / 16), n_flag = ((size_t)n_reads + 31) / 32;
    if ((rc = ensure(ctx, (void**)&s.d_bases, &s.cap_bases, n_words * 4 + 16))) return rc;
    if ((rc = ensure(ctx, (void**)&s.d_off, &s.cap_off, ((size_t)n_reads + 1) * 4))) return rc;
    if ((rc = ensure(ctx, (void**)&s.d_bc, &s.cap_bc, (size_t)n_reads * 4))) return rc;
    if ((rc = ensure(ctx, (void**)&s.d_hasn, &s.cap_hasn, n_flag * 4))) return rc;
    CU(cudaMemcpyAsync(s.d_bases, packed, n_words * 4, cudaMemcpyHostToDevice, ctx->hs));
    CU(cudaMemcpyAsync(s.d_off, read_off, ((size_t)n_reads + 1) * 4, cudaMemcpyHostToDevice, ctx->hs));
    CU(cudaMemcpyAsync(s.d_bc, barcode_id, (size_t)n_reads * 4, cudaMemcpyHostToDevice, ctx->hs));
    CU(cudaMemcpyAsync(s.d_hasn, has_n, n_flag * 4, cudaMemcpyHostToDevice, ctx->hs));
    CU(cudaEventRecord(s.copied, ctx->hs));
    CU(cudaStreamWaitEvent(ctx->cs, s.copied, 0));
    BatchView bv{nullptr, s.d_off, s.d_bc, n_bases, n_reads, (const uint32_t*)s.d_bases, s.d_hasn};
    if ((rc = launch_tile(ctx, MODE_CLASSIFY, bv, nullptr, nullptr))) return rc;
    CU(cudaEventRecord(s.done, ctx->cs));
    ctx->st.h2d_bytes += n_words * 4 + ((size_t)n_reads + 1) * 4 + (size_t)n_reads * 4 + n_flag * 4;
    ctx->st.batches++;
    ctx->st.reads += n_reads;
    ctx->st.bases += n_bases;
    ctx->seq++;
    return HAST_OK;
}

int hast_submit_batch_packed_device(hast_ctx* ctx, const uint32_t* d_packed, uint64_t n_bases,
                                    const uint32_t* d_read_off, const uint32_t* d_barcode_id,
                                    const uint32_t* d_has_n, uint32_t n_reads) {
    int rc = batch_args_ok(ctx, d_packed, d_read_off, d_barcode_id, n_bases, n_reads);
    if (rc) return rc;
    if (n_reads && !d_has_n) return fail(ctx, HAST_E_ARG, "NULL batch array");
    if (ctx->opt_kernel == 0) return fail(ctx, HAST_E_STATE, "packed batches need the pre-filtered kernel (option kernel >= 1)");
    if (!n_reads) return HAST_OK;
    CU(cudaSetDevice(ctx->device));
    BatchView bv{nullptr, d_read_off, d_barcode_id, n_bases, n_reads, d_packed, d_has_n};
    if ((rc = launch_tile(ctx, MODE_CLASSIFY, bv, nullptr, nullptr))) return rc;
    ctx->st.batches++;
    ctx->st.reads += n_reads;
    ctx->st.bases += n_bases;
    return HAST_OK;
}

// stateless form of the host packer (tests; callers who want to pack ahead of hast_submit_batch_packed)
int hast_pack_bases(const uint8_t* bases, uint64_t n_bases, const uint32_t* read_off, uint32_t n_reads,
                    uint32_t* words_out, uint32_t* has_n_out, int threads) {
    hast_ctx* ctx = nullptr;
    if ((!bases && n_bases) || !read_off || !words_out || !has_n_out) return fail(ctx, HAST_E_ARG, "NULL argument");
    if (n_bases >= 0xFFFFFFF0ull) return fail(ctx, HAST_E_ARG, "batch larger than 4 GiB of bases");
    hastpack::Pool* p = threads > 1 ? hastpack::pool_create(threads) : nullptr;
    hastpack::pack_batch(p, bases, n_bases, read_off, n_reads, words_out, has_n_out);
    if (p) hastpack::pool_destroy(p);
    return HAST_OK;
}

int hast_sync(hast_ctx* ctx) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->hs));
    CU(cudaStreamSynchronize(ctx->cs));
    return HAST_OK;
}

int hast_stats_get(hast_ctx* ctx, hast_stats* out) {
    if (!ctx || !out) return fail(ctx, HAST_E_ARG, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    DevStats ds;
    int rc = read_stats(ctx, &ds);
    if (rc) return rc;
    *out = ctx->st;
    out->lookups = ds.lookups;
    out->reads_with_n = ds.reads_with_n;
    out->reads_short = ds.reads_short;
    out->extra_probes = ds.extra_probes;
    out->filter_pass = ds.filter_pass;
    out->filter_loads = ds.filter_loads;
    return HAST_OK;
}

int hast_counts_device_ptr(hast_ctx* ctx, void** ptr, uint64_t* n_barcodes) {
    if (!ctx || !ptr) return fail(ctx, HAST_E_ARG, "NULL argument");
    *ptr = ctx->d_counts;
    if (n_barcodes) *n_barcodes = ctx->n_barcodes;
    return HAST_OK;
}

int hast_timer_start(hast_ctx* ctx) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->t0, ctx->cs));
    return HAST_OK;
}
int hast_timer_stop(hast_ctx* ctx, float* ms) {
    if (!ctx || !ms) return fail(ctx, HAST_E_ARG, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->t1, ctx->cs));
    CU(cudaEventSynchronize(ctx->t1));
    CU(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return HAST_OK;
}

int hast_finish(hast_ctx* ctx, int32_t* counts_out, uint64_t n_barcodes) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->hs));
    CU(cudaStreamSynchronize(ctx->cs));
    DevStats ds;
    int rc = read_stats(ctx, &ds);
    if (rc) return rc;
    if (ds.reads_short)
        return fail(ctx, HAST_E_SHORT_READ, std::to_string(ds.reads_short) + " read(s) shorter than k=" +
                    std::to_string(ctx->tv.k) + " (the reference aborts on these, kmer.h:171)");
    if (ds.reads_too_long)
        return fail(ctx, HAST_E_ARG, std::to_string(ds.reads_too_long) + " read(s) longer than " +
                    std::to_string(((ctx->opt_kernel == 2 || ctx->opt_kernel == 4) ? FusedSmem<true>::kCap : ctx->opt_kernel >= 1 ? FusedSmem<false>::kCap : kTileCapBytes) - 16) + " bases");
    if (ds.bad_barcode)
        return fail(ctx, HAST_E_ARG, std::to_string(ds.bad_barcode) + " read(s) with barcode id >= reserved barcodes");
    if (n_barcodes > ctx->n_barcodes) return fail(ctx, HAST_E_ARG, "n_barcodes exceeds reserved barcodes");
    const int32_t* src = ctx->d_counts;
    float ms = 0.f;
    if (ctx->comm) {
        if (n_barcodes > ctx->cap_reduced) {
            if (ctx->d_reduced) CU(cudaFree(ctx->d_reduced));
            CU(cudaMalloc(&ctx->d_reduced, std::max<uint64_t>(n_barcodes, 1) * 2 * sizeof(int32_t)));
            ctx->cap_reduced = n_barcodes;
        }
        // BarcodeCache::Add (classify.cpp:57-63) across GPUs: one int32 sum to rank 0
        CU(cudaEventRecord(ctx->f0, ctx->cs));
        NC(g_nccl.Reduce(ctx->d_counts, ctx->d_reduced, n_barcodes * 2, ncclInt32, ncclSum, 0, ctx->comm, ctx->cs));
        CU(cudaEventRecord(ctx->f1, ctx->cs));
        CU(cudaStreamSynchronize(ctx->cs));
        CU(cudaEventElapsedTime(&ms, ctx->f0, ctx->f1));
        ctx->st.finish_reduce_us += (uint64_t)(ms * 1000.f);
        src = ctx->d_reduced;
    }
    if (counts_out && n_barcodes && ctx->rank == 0) {
        CU(cudaEventRecord(ctx->f0, ctx->cs));
        CU(cudaMemcpyAsync(counts_out, src, n_barcodes * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->cs));
        CU(cudaEventRecord(ctx->f1, ctx->cs));
        CU(cudaStreamSynchronize(ctx->cs));
        CU(cudaEventElapsedTime(&ms, ctx->f0, ctx->f1));
        ctx->st.finish_d2h_us += (uint64_t)(ms * 1000.f);
        ctx->st.d2h_bytes += n_barcodes * 2 * sizeof(int32_t);
    }
    return HAST_OK;
}

// ---- multi-GPU -------------------------------------------------------------
int hast_comm_unique_id(void* id128) {
    hast_ctx* ctx = nullptr;
    if (!id128) return fail(nullptr, HAST_E_ARG, "NULL id");
    if (!nccl_load()) return fail(nullptr, HAST_E_NCCL, "libnccl.so.2 not loadable");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return HAST_OK;
}

int hast_comm_init_rank(hast_ctx* ctx, int nranks, int rank, const void* id128) {
    if (!ctx || !id128) return fail(ctx, HAST_E_ARG, "NULL argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, HAST_E_ARG, "bad rank / nranks");
    if (!nccl_load()) return fail(ctx, HAST_E_NCCL, "libnccl.so.2 not loadable");
    CU(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NC(g_nccl.CommInitRank(&ctx->comm, nranks, id, rank));
    ctx->nranks = nranks;
    ctx->rank = rank;
    return HAST_OK;
}

int hast_comm_init_all(hast_ctx** ctxs, int n) {
    hast_ctx* ctx = (ctxs && n > 0) ? ctxs[0] : nullptr;
    if (!ctx) return fail(nullptr, HAST_E_ARG, "no contexts");
    if (n == 1) return HAST_OK;
    if (!nccl_load()) return fail(ctx, HAST_E_NCCL, "libnccl.so.2 not loadable");
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
    NC(g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) { ctxs[i]->comm = comms[i]; ctxs[i]->nranks = n; ctxs[i]->rank = i; }
    return HAST_OK;
}

// ---- standalone K2 / K3 ----------------------------------------------------
int hast_extract_kmers_device(hast_ctx* ctx, const uint8_t* d_bases, uint64_t n_bases,
                              const uint32_t* d_read_off, uint32_t n_reads, uint64_t* d_kmers_out,
                              uint8_t* d_has_n_out) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (!ctx->table_ready) return fail(ctx, HAST_E_STATE, "hast_table_begin first (fixes k)");
    if (!n_reads) return HAST_OK;
    if (!d_bases || !d_read_off || !d_kmers_out) return fail(ctx, HAST_E_ARG, "NULL array");
    if ((uintptr_t)d_bases & 15) return fail(ctx, HAST_E_ARG, "d_bases must be 16-byte aligned");
    if (n_bases >= 0xFFFFFFF0ull) return fail(ctx, HAST_E_ARG, "batch larger than 4 GiB of bases");
    CU(cudaSetDevice(ctx->device));
    BatchView bv{d_bases, d_read_off, nullptr, n_bases, n_reads, nullptr, nullptr};
    return launch_tile(ctx, MODE_EXTRACT, bv, d_kmers_out, d_has_n_out);
}

int hast_extract_kmers(hast_ctx* ctx, const uint8_t* bases, uint64_t n_bases, const uint32_t* read_off,
                       uint32_t n_reads, uint64_t* kmers_out, uint64_t n_kmers_out, uint8_t* has_n_out) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (!ctx->table_ready) return fail(ctx, HAST_E_STATE, "hast_table_begin first (fixes k)");
    if (!n_reads) return HAST_OK;
    if (!bases || !read_off || !kmers_out) return fail(ctx, HAST_E_ARG, "NULL array");
    if (n_kmers_out < n_bases) return fail(ctx, HAST_E_ARG, "kmers_out needs one slot per base position");
    CU(cudaSetDevice(ctx->device));
    const size_t o_b = 0, o_off = (n_bases + 16 + 15) & ~(size_t)15, o_n = o_off + (((size_t)n_reads + 1) * 4 + 15 & ~(size_t)15),
                 o_k = (o_n + n_reads + 15) & ~(size_t)15, total = o_k + n_bases * 8;
    int rc = ensure(ctx, &ctx->d_scratch, &ctx->cap_scratch, total);
    if (rc) return rc;
    char* d = (char*)ctx->d_scratch;
    CU(cudaMemcpyAsync(d + o_b, bases, n_bases, cudaMemcpyHostToDevice, ctx->cs));
    CU(cudaMemcpyAsync(d + o_off, read_off, ((size_t)n_reads + 1) * 4, cudaMemcpyHostToDevice, ctx->cs));
    CU(cudaMemsetAsync(d + o_k, 0xFF, n_bases * 8, ctx->cs));   // positions that start no k-mer
    rc = hast_extract_kmers_device(ctx, (const uint8_t*)(d + o_b), n_bases, (const uint32_t*)(d + o_off), n_reads,
                                   (uint64_t*)(d + o_k), (uint8_t*)(d + o_n));
    if (rc) return rc;
    CU(cudaMemcpyAsync(kmers_out, d + o_k, n_bases * 8, cudaMemcpyDeviceToHost, ctx->cs));
    if (has_n_out) CU(cudaMemcpyAsync(has_n_out, d + o_n, n_reads, cudaMemcpyDeviceToHost, ctx->cs));
    CU(cudaStreamSynchronize(ctx->cs));
    return HAST_OK;
}

int hast_lookup_device(hast_ctx* ctx, const uint64_t* d_canonical, uint64_t n, uint8_t* d_tags_out) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (!ctx->table_ready) return fail(ctx, HAST_E_STATE, "no table");
    if (!n) return HAST_OK;
    if (!d_canonical || !d_tags_out) return fail(ctx, HAST_E_ARG, "NULL array");
    CU(cudaSetDevice(ctx->device));
    lookup_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->cs>>>(ctx->tv, d_canonical, n, d_tags_out, ctx->d_stats);
    CU(cudaGetLastError());
    ctx->st.kernel_launches++;
    return HAST_OK;
}

int hast_lookup(hast_ctx* ctx, const uint64_t* canonical, uint64_t n, uint8_t* tags_out) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (!ctx->table_ready) return fail(ctx, HAST_E_STATE, "no table");
    if (!n) return HAST_OK;
    if (!canonical || !tags_out) return fail(ctx, HAST_E_ARG, "NULL array");
    CU(cudaSetDevice(ctx->device));
    const size_t o_t = n * 8;
    int rc = ensure(ctx, &ctx->d_scratch, &ctx->cap_scratch, o_t + n);
    if (rc) return rc;
    char* d = (char*)ctx->d_scratch;
    CU(cudaMemcpyAsync(d, canonical, n * 8, cudaMemcpyHostToDevice, ctx->cs));
    rc = hast_lookup_device(ctx, (const uint64_t*)d, n, (uint8_t*)(d + o_t));
    if (rc) return rc;
    CU(cudaMemcpyAsync(tags_out, d + o_t, n, cudaMemcpyDeviceToHost, ctx->cs));
    CU(cudaStreamSynchronize(ctx->cs));
    return HAST_OK;
}

int hast_gather_roofline(hast_ctx* ctx, uint64_t n_probes, uint64_t span_bytes, float* gbps) {
    if (!ctx || !gbps) return fail(ctx, HAST_E_ARG, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    uint64_t sectors = 1;
    while (sectors * 2 * 32 <= span_bytes) sectors *= 2;
    const size_t bytes = sectors * 32;
    int rc = ensure(ctx, &ctx->d_scratch, &ctx->cap_scratch, bytes + 64);
    if (rc) return rc;
    CU(cudaMemsetAsync(ctx->d_scratch, 0x5A, bytes, ctx->cs));
    unsigned long long* sink = (unsigned long long*)((char*)ctx->d_scratch + bytes);
    const int grid = ctx->sm_count * 8;
    gather_kernel<<<grid, 256, 0, ctx->cs>>>((const uint64_t*)ctx->d_scratch, sectors - 1, n_probes / 8, sink);  // warm-up
    CU(cudaEventRecord(ctx->t0, ctx->cs));
    gather_kernel<<<grid, 256, 0, ctx->cs>>>((const uint64_t*)ctx->d_scratch, sectors - 1, n_probes, sink);
    CU(cudaEventRecord(ctx->t1, ctx->cs));
    CU(cudaGetLastError());
    CU(cudaEventSynchronize(ctx->t1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->t0, ctx->t1));
    ctx->st.kernel_launches += 2;
    *gbps = ms > 0 ? (float)((double)n_probes * 32.0 / (ms * 1e-3) / 1e9) : 0.f;
    return HAST_OK;
}

// ---- stage 00: k-mer counting (kcount.cuh) -------------------------------------
static int kc_ready(hast_ctx* ctx) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (!ctx->kc.slots) return fail(ctx, HAST_E_STATE, "hast_kc_begin first");
    return HAST_OK;
}

int hast_kc_end(hast_ctx* ctx) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->hs));
    CU(cudaStreamSynchronize(ctx->cs));
    if (ctx->kc.slots) { CU(cudaFree(ctx->kc.slots)); ctx->kc.slots = nullptr; }
    ctx->kc_slots = 0;
    return HAST_OK;
}

int hast_kc_begin(hast_ctx* ctx, int k, uint64_t expected_distinct, uint32_t part, uint32_t n_parts) {
    if (!ctx) return fail(nullptr, HAST_E_ARG, "NULL context");
    if (k < 1 || k > 32) return fail(ctx, HAST_E_K, "k must be in 1..32");
    if (n_parts < 1 || part >= n_parts) return fail(ctx, HAST_E_ARG, "need part < n_parts");
    int rc = hast_kc_end(ctx);
    if (rc) return rc;
    int b = 10;
    while (((uint64_t)1 << b) < 2 * expected_distinct && b < 36) ++b;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    if (((uint64_t)16 << b) > free_b)
        return fail(ctx, HAST_E_TABLE_FULL, "count table of " + std::to_string(((uint64_t)16 << b) >> 20) +
                    " MiB does not fit the free device memory: use more partitions");
    ctx->kc_slots = (uint64_t)1 << b;
    CU(cudaMalloc(&ctx->kc.slots, ctx->kc_slots * sizeof(KcSlot)));
    CU(cudaMemsetAsync(ctx->kc.slots, 0, ctx->kc_slots * sizeof(KcSlot), ctx->cs));
    if (!ctx->d_kc_stats) CU(cudaMalloc(&ctx->d_kc_stats, sizeof(KcStats)));
    CU(cudaMemsetAsync(ctx->d_kc_stats, 0, sizeof(KcStats), ctx->cs));
    ctx->kc.mask = ctx->kc_slots - 1;
    ctx->kc.k = k;
    ctx->kc.part = part;
    ctx->kc.n_parts = n_parts;
    ctx->kc.max_probe = (uint32_t)std::min<uint64_t>(ctx->kc_slots, 8192);
    return HAST_OK;
}

static int kc_launch(hast_ctx* ctx, const BatchView& bv, int parent) {
    const uint32_t n_tiles = (bv.n_reads + kKcReadsPerTile - 1) / kKcReadsPerTile;
    if (!n_tiles) return HAST_OK;
    const int grid = (int)std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count * 4);
    switch (ctx->kc.k) {
        case 21: kc_count_kernel<21><<<grid, kTileThreads, sizeof(KcSmem), ctx->cs>>>(ctx->kc, bv, (uint32_t)parent, ctx->d_kc_stats); break;
        case 31: kc_count_kernel<31><<<grid, kTileThreads, sizeof(KcSmem), ctx->cs>>>(ctx->kc, bv, (uint32_t)parent, ctx->d_kc_stats); break;
        default: kc_count_kernel<0><<<grid, kTileThreads, sizeof(KcSmem), ctx->cs>>>(ctx->kc, bv, (uint32_t)parent, ctx->d_kc_stats); break;
    }
    CU(cudaGetLastError());
    ctx->st.kernel_launches++;
    return HAST_OK;
}

int hast_kc_add_device(hast_ctx* ctx, const uint8_t* d_bases, uint64_t n_bases, const uint32_t* d_seq_off,
                       uint32_t n_seqs, int parent) {
    int rc = kc_ready(ctx);
    if (rc) return rc;
    if (parent < 0 || parent > 1) return fail(ctx, HAST_E_ARG, "parent must be 0 or 1");
    if (!n_seqs) return HAST_OK;
    if (!d_bases || !d_seq_off) return fail(ctx, HAST_E_ARG, "NULL batch array");
    if ((uintptr_t)d_bases & 15) return fail(ctx, HAST_E_ARG, "d_bases must be 16-byte aligned");
    if (n_bases >= 0xFFFFFFF0ull) return fail(ctx, HAST_E_ARG, "batch larger than 4 GiB of bases");
    CU(cudaSetDevice(ctx->device));
    BatchView bv{d_bases, d_seq_off, nullptr, n_bases, n_seqs, nullptr, nullptr};
    return kc_launch(ctx, bv, parent);
}

int hast_kc_add(hast_ctx* ctx, const uint8_t* bases, uint64_t n_bases, const uint32_t* seq_off, uint32_t n_seqs,
                int parent, uint64_t* ticket) {
    int rc = kc_ready(ctx);
    if (rc) return rc;
    if (parent < 0 || parent > 1) return fail(ctx, HAST_E_ARG, "parent must be 0 or 1");
    if (ticket) *ticket = ctx->seq;
    if (!n_seqs) return HAST_OK;
    if (!seq_off || (!bases && n_bases)) return fail(ctx, HAST_E_ARG, "NULL batch array");
    if (n_bases >= 0xFFFFFFF0ull) return fail(ctx, HAST_E_ARG, "batch larger than 4 GiB of bases");
    CU(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[ctx->seq % kSlots];
    CU(cudaEventSynchronize(s.done));
    if ((rc = ensure(ctx, (void**)&s.d_bases, &s.cap_bases, n_bases + 16))) return rc;
    if ((rc = ensure(ctx, (void**)&s.d_off, &s.cap_off, ((size_t)n_seqs + 1) * 4))) return rc;
    CU(cudaMemcpyAsync(s.d_bases, bases, n_bases, cudaMemcpyHostToDevice, ctx->hs));
    CU(cudaMemcpyAsync(s.d_off, seq_off, ((size_t)n_seqs + 1) * 4, cudaMemcpyHostToDevice, ctx->hs));
    CU(cudaEventRecord(s.copied, ctx->hs));
    CU(cudaStreamWaitEvent(ctx->cs, s.copied, 0));
    BatchView bv{s.d_bases, s.d_off, nullptr, n_bases, n_seqs, nullptr, nullptr};
    if ((rc = kc_launch(ctx, bv, parent))) return rc;
    CU(cudaEventRecord(s.done, ctx->cs));
    ctx->st.h2d_bytes += n_bases + ((size_t)n_seqs + 1) * 4;
    ctx->st.batches++;
    ctx->st.reads += n_seqs;
    ctx->st.bases += n_bases;
    ctx->seq++;
    return HAST_OK;
}

static int kc_read_stats(hast_ctx* ctx, KcStats* ks) {
    CU(cudaStreamSynchronize(ctx->hs));
    CU(cudaMemcpyAsync(ks, ctx->d_kc_stats, sizeof(KcStats), cudaMemcpyDeviceToHost, ctx->cs));
    CU(cudaStreamSynchronize(ctx->cs));
    if (ks->too_long)
        return fail(ctx, HAST_E_ARG, std::to_string(ks->too_long) + " sequence chunk(s) longer than " +
                    std::to_string(kKcCap - 16) + " bytes: cut them with k-1 bytes of overlap");
    return HAST_OK;
}

int hast_kc_info_get(hast_ctx* ctx, hast_kc_info* out) {
    int rc = kc_ready(ctx);
    if (rc) return rc;
    if (!out) return fail(ctx, HAST_E_ARG, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    KcStats ks;
    if ((rc = kc_read_stats(ctx, &ks))) return rc;
    if ((rc = ensure(ctx, &ctx->d_scratch, &ctx->cap_scratch, sizeof(KcTotals)))) return rc;
    CU(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(KcTotals), ctx->cs));
    kc_totals_kernel<<<grid_for(ctx, ctx->kc_slots, 256), 256, 0, ctx->cs>>>(ctx->kc.slots, ctx->kc_slots,
                                                                          (KcTotals*)ctx->d_scratch);
    CU(cudaGetLastError());
    ctx->st.kernel_launches++;
    KcTotals t;
    CU(cudaMemcpyAsync(&t, ctx->d_scratch, sizeof(t), cudaMemcpyDeviceToHost, ctx->cs));
    CU(cudaStreamSynchronize(ctx->cs));
    out->k = ctx->kc.k;
    out->part = ctx->kc.part;
    out->n_parts = ctx->kc.n_parts;
    out->n_slots = ctx->kc_slots;
    out->bytes = ctx->kc_slots * sizeof(KcSlot);
    out->occupied = t.occupied;
    out->distinct[0] = t.distinct[0];
    out->distinct[1] = t.distinct[1];
    out->both = t.both;
    out->occurrences[0] = t.occurrences[0];
    out->occurrences[1] = t.occurrences[1];
    out->windows = ks.windows;
    out->table_full = ks.table_full;
    return HAST_OK;
}

static int kc_check_full(hast_ctx* ctx) {
    KcStats ks;
    int rc = kc_read_stats(ctx, &ks);
    if (rc) return rc;
    if (ks.table_full)
        return fail(ctx, HAST_E_TABLE_FULL, std::to_string(ks.table_full) +
                    " k-mer occurrence(s) found no slot in the count table: use a larger expected_distinct or more partitions");
    return HAST_OK;
}

int hast_kc_histo(hast_ctx* ctx, int parent, uint32_t high, uint64_t* histo) {
    int rc = kc_ready(ctx);
    if (rc) return rc;
    if (parent < 0 || parent > 1 || !histo || high < 1 || high > (1u << 24)) return fail(ctx, HAST_E_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    if ((rc = kc_check_full(ctx))) return rc;
    const size_t bytes = ((size_t)high + 2) * 8;
    if ((rc = ensure(ctx, &ctx->d_scratch, &ctx->cap_scratch, bytes))) return rc;
    CU(cudaMemsetAsync(ctx->d_scratch, 0, bytes, ctx->cs));
    kc_histo_kernel<<<grid_for(ctx, ctx->kc_slots, 256, 4), 256, 0, ctx->cs>>>(
        ctx->kc.slots, ctx->kc_slots, (uint32_t)parent, high, (unsigned long long*)ctx->d_scratch);
    CU(cudaGetLastError());
    ctx->st.kernel_launches++;
    CU(cudaMemcpyAsync(histo, ctx->d_scratch, bytes, cudaMemcpyDeviceToHost, ctx->cs));
    CU(cudaStreamSynchronize(ctx->cs));
    return HAST_OK;
}

// selection of one parent into d_scratch: [counter (16 B)] [keys, cap of them] ; sorted when `sort`
static int kc_select_device(hast_ctx* ctx, int parent, uint32_t lower, uint32_t upper, bool require_unique,
                            bool hast_code, bool sort, uint64_t** d_keys, uint64_t* n_out) {
    int rc;
    // pass 1: how many
    if ((rc = ensure(ctx, &ctx->d_scratch, &ctx->cap_scratch, 16))) return rc;
    CU(cudaMemsetAsync(ctx->d_scratch, 0, 16, ctx->cs));
    const int grid = grid_for(ctx, ctx->kc_slots, 256, 8);
    kc_select_kernel<<<grid, 256, 0, ctx->cs>>>(ctx->kc.slots, ctx->kc_slots, (uint32_t)parent, lower, upper,
                                                require_unique, hast_code, nullptr, 0,
                                                (unsigned long long*)ctx->d_scratch);
    CU(cudaGetLastError());
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(&n, ctx->d_scratch, 8, cudaMemcpyDeviceToHost, ctx->cs));
    CU(cudaStreamSynchronize(ctx->cs));
    ctx->st.kernel_launches++;
    *n_out = n;
    *d_keys = nullptr;
    if (!n) return HAST_OK;
    size_t tmp_bytes = 0;
    if (sort) cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (size_t)n,
                                             0, 2 * ctx->kc.k, ctx->cs);
    const size_t o_a = 16, o_b = o_a + n * 8, o_tmp = (o_b + (sort ? n * 8 : 0) + 255) & ~(size_t)255;
    if ((rc = ensure(ctx, &ctx->d_scratch, &ctx->cap_scratch, o_tmp + tmp_bytes))) return rc;
    char* d = (char*)ctx->d_scratch;
    CU(cudaMemsetAsync(d, 0, 16, ctx->cs));
    kc_select_kernel<<<grid, 256, 0, ctx->cs>>>(ctx->kc.slots, ctx->kc_slots, (uint32_t)parent, lower, upper,
                                                require_unique, hast_code, (uint64_t*)(d + o_a), n,
                                                (unsigned long long*)d);
    CU(cudaGetLastError());
    ctx->st.kernel_launches++;
    if (sort) {
        CU(cub::DeviceRadixSort::SortKeys(d + o_tmp, tmp_bytes, (const uint64_t*)(d + o_a), (uint64_t*)(d + o_b), (size_t)n,
                                          0, 2 * ctx->kc.k, ctx->cs));
        *d_keys = (uint64_t*)(d + o_b);
    } else {
        *d_keys = (uint64_t*)(d + o_a);
    }
    return HAST_OK;
}

int hast_kc_select(hast_ctx* ctx, int parent, uint32_t lower, uint32_t upper, int require_unique, uint64_t* out,
                   uint64_t cap, uint64_t* n) {
    int rc = kc_ready(ctx);
    if (rc) return rc;
    if (parent < 0 || parent > 1 || !n || (cap && !out)) return fail(ctx, HAST_E_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    if ((rc = kc_check_full(ctx))) return rc;
    uint64_t* d_keys = nullptr;
    if ((rc = kc_select_device(ctx, parent, lower, upper, require_unique != 0, false, true, &d_keys, n))) return rc;
    const uint64_t m = std::min(*n, cap);
    if (m) {
        CU(cudaMemcpyAsync(out, d_keys, m * 8, cudaMemcpyDeviceToHost, ctx->cs));
        CU(cudaStreamSynchronize(ctx->cs));
        ctx->st.d2h_bytes += m * 8;
    }
    return HAST_OK;
}

int hast_kc_to_table(hast_ctx* dst, hast_ctx* src, uint32_t pl, uint32_t pu, uint32_t ml, uint32_t mu) {
    hast_ctx* ctx = src;
    int rc = kc_ready(src);
    if (rc) return rc;
    if (!dst) return fail(src, HAST_E_ARG, "NULL context");
    if (dst->device != src->device) return fail(src, HAST_E_ARG, "dst and src must be on the same device (clone the table afterwards)");
    if (src->kc.n_parts != 1) return fail(src, HAST_E_STATE, "the count table holds one partition only");
    CU(cudaSetDevice(src->device));
    if ((rc = kc_check_full(src))) return rc;
    // sizes first (the table is sized from them), then one selection + insertion per parent
    uint64_t n[2] = {0, 0};
    uint64_t* d_keys = nullptr;
    const uint32_t lo[2] = {pl, ml}, hi[2] = {pu, mu};
    for (int p = 0; p < 2; ++p) {
        if ((rc = ensure(src, &src->d_scratch, &src->cap_scratch, 16))) return rc;
        CU(cudaMemsetAsync(src->d_scratch, 0, 16, src->cs));
        kc_select_kernel<<<grid_for(src, src->kc_slots, 256, 8), 256, 0, src->cs>>>(
            src->kc.slots, src->kc_slots, (uint32_t)p, lo[p], hi[p], true, true, nullptr, 0,
            (unsigned long long*)src->d_scratch);
        CU(cudaGetLastError());
        unsigned long long c = 0;
        CU(cudaMemcpyAsync(&c, src->d_scratch, 8, cudaMemcpyDeviceToHost, src->cs));
        CU(cudaStreamSynchronize(src->cs));
        src->st.kernel_launches++;
        n[p] = c;
    }
    if ((rc = hast_table_begin(dst, src->kc.k, std::max<uint64_t>(n[0] + n[1], 16)))) {
        if (dst != src) src->err = dst->err;
        return rc;
    }
    for (int p = 0; p < 2; ++p) {
        uint64_t m = 0;
        if ((rc = kc_select_device(src, p, lo[p], hi[p], true, true, false, &d_keys, &m))) return rc;
        if (!m) continue;
        CU(cudaStreamSynchronize(src->cs));
        table_insert_packed_kernel<<<grid_for(dst, m, 256), 256, 0, dst->cs>>>(dst->tv, d_keys, m, (uint32_t)p, dst->d_stats);
        CU(cudaGetLastError());
        dst->st.kernel_launches++;
        CU(cudaStreamSynchronize(dst->cs));
    }
    ctx = dst;
    return table_check(dst);
}

}  // extern "C"
