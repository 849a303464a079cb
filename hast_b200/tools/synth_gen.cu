// synth_gen.cu -- device side of the counter-based read-pair generator (see synth_gen.h).
// Synthetic-data tooling, not on the classification path: it fills device buffers that bench.py then hands
// to the engine exactly like reads that arrived over PCIe.
#include <cuda_runtime.h>
#include <stdint.h>

#include "synth_gen.h"

namespace {

// one warp per pair: the pair's parameters are computed once per warp (uniform), the lanes then write
// the 2 * ceil(L/4) four-base groups of r1 and r2.
__global__ void sg_pairs_kernel(sg_params p, const uint64_t* __restrict__ cdf, const uint8_t* __restrict__ hap0,
                                const uint8_t* __restrict__ hap1, const uint64_t* __restrict__ idx, uint64_t lo,
                                uint64_t n, uint8_t* __restrict__ bases, uint32_t* __restrict__ bc) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t L = p.read_len, Q = (L + 3) / 4;
    for (uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += warps) {
        const uint64_t i = idx ? idx[w] : lo + w;
        const sg_pair pr = sg_pair_of(&p, cdf, i);
        const sg_edits e0 = sg_edits_of(&p, i, 0), e1 = sg_edits_of(&p, i, 1);
        if (lane == 0) { bc[w] = pr.barcode; bc[n + w] = pr.barcode; }
        for (uint32_t g = lane; g < 2 * Q; g += 32) {
            const uint32_t mate = g >= Q, q = mate ? g - Q : g;
            const sg_edits* e = mate ? &e1 : &e0;
            uint8_t* out = bases + (mate ? n + w : w) * (uint64_t)L + 4ull * q;
            uint32_t word = 0;
            const uint32_t m = min(4u, L - 4 * q);
            for (uint32_t t = 0; t < m; ++t) word |= (uint32_t)sg_base(&p, hap0, hap1, &pr, e, mate, 4 * q + t) << (8 * t);
            if (m == 4 && (((uintptr_t)out) & 3) == 0) *reinterpret_cast<uint32_t*>(out) = word;
            else for (uint32_t t = 0; t < m; ++t) out[t] = (uint8_t)(word >> (8 * t));
        }
    }
}

__global__ void sg_barcodes_kernel(sg_params p, const uint64_t* __restrict__ cdf, uint64_t lo, uint64_t n,
                                   uint32_t* __restrict__ bc) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += stride)
        bc[w] = sg_pair_of(&p, cdf, lo + w).barcode;
}

int grid_for(uint64_t items, int threads) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t want = (items + threads - 1) / threads;
    const uint64_t cap = (uint64_t)sms * 16;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

extern "C" {

// rows [0, n) of bases/bc = r1 of the pairs, rows [n, 2n) = r2.  idx == NULL: pairs lo .. lo+n-1.
// All pointers are device pointers; runs on the legacy default stream (what torch uses) and returns
// after the launch -- the caller synchronises.
int sg_gen_pairs_device(const sg_params* p, const uint64_t* d_cdf, const uint8_t* d_hap0, const uint8_t* d_hap1,
                        const uint64_t* d_idx, uint64_t lo, uint64_t n, uint8_t* d_bases, uint32_t* d_bc) {
    if (!n) return 0;
    sg_pairs_kernel<<<grid_for(n * 32, 256), 256>>>(*p, d_cdf, d_hap0, d_hap1, d_idx, lo, n, d_bases, d_bc);
    return (int)cudaGetLastError();
}

int sg_barcodes_device(const sg_params* p, const uint64_t* d_cdf, uint64_t lo, uint64_t n, uint32_t* d_bc) {
    if (!n) return 0;
    sg_barcodes_kernel<<<grid_for(n, 256), 256>>>(*p, d_cdf, lo, n, d_bc);
    return (int)cudaGetLastError();
}

}  // extern "C"
