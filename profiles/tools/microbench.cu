// microbench.cu -- random-access roofline probes for the lookup path.
//
// Measures, on the box the bench runs on, how many independent random probes
// per second the memory system sustains as a function of
//   * the span of the probed structure (L2-resident ... HBM-resident),
//   * the bytes each probe loads (4 / 8 / 16 / 32),
//   * how many adjacent lanes of a warp share one 128-byte line (locality).
// These are the denominators the fused classify kernel is designed against:
// the Bloom pre-filter does one 8-byte probe per k-mer position into an
// L2-resident array, the exact table one 32-byte probe per filter pass.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/microbench profiles/tools/microbench.cu
//   bin/microbench > gpurun_out/microbench.csv
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x *= 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    return x;
}

template <int W> struct Vec;
template <> struct Vec<4>  { uint32_t a; __device__ uint64_t fold() const { return a; } };
template <> struct Vec<8>  { uint64_t a; __device__ uint64_t fold() const { return a; } };
template <> struct Vec<16> { uint64_t a, b; __device__ uint64_t fold() const { return a ^ b; } };
template <> struct Vec<32> { uint64_t a, b, c, d; __device__ uint64_t fold() const { return a ^ b ^ c ^ d; } };

template <int W> __device__ __forceinline__ Vec<W> ldv(const char* p);
template <> __device__ __forceinline__ Vec<4> ldv<4>(const char* p) {
    Vec<4> v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v.a) : "l"(p)); return v; }
template <> __device__ __forceinline__ Vec<8> ldv<8>(const char* p) {
    Vec<8> v; asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v.a) : "l"(p)); return v; }
template <> __device__ __forceinline__ Vec<16> ldv<16>(const char* p) {
    Vec<16> v; asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(v.a), "=l"(v.b) : "l"(p)); return v; }
template <> __device__ __forceinline__ Vec<32> ldv<32>(const char* p) {
    Vec<32> v; asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                            : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p)); return v; }

// G adjacent lanes share one 128-byte line (each lane its own W-byte piece of it
// when G*W <= 128, else pieces wrap).  G = 1: every lane its own random line.
template <int W, int G, int U>
__global__ void __launch_bounds__(256) gather(const char* __restrict__ buf, uint64_t line_mask, uint64_t n_probes,
                                              unsigned long long* __restrict__ sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t acc = 0;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t sub = (threadIdx.x % G) * W % 128;
    for (; i + (U - 1) * stride < n_probes; i += U * stride) {
        Vec<W> v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t id = (i + u * stride) / G;
            const uint64_t x = mix(id);
            // G == 1: random W-aligned offset inside the line too
            const uint32_t off = G == 1 ? (uint32_t)((x >> 40) % (128 / W)) * W : sub;
            v[u] = ldv<W>(buf + ((x & line_mask) << 7) + off);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u].fold();
    }
    if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}

template <int W, int G, int U>
static void run(const char* d, size_t span, unsigned long long* sink, int sms, uint64_t n_probes, int ctas_per_sm) {
    const uint64_t lines = span / 128;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int grid = sms * ctas_per_sm;
    gather<W, G, U><<<grid, 256>>>(d, lines - 1, n_probes / 4, sink);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        gather<W, G, U><<<grid, 256>>>(d, lines - 1, n_probes, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double gps = (double)n_probes / (best * 1e-3) / 1e9;
    printf("gather,%d,%d,%d,%d,%zu,%.3f,%.2f,%.1f\n", W, G, U, ctas_per_sm, span >> 20, best, gps, gps * W);
    fflush(stdout);
    CK(cudaEventDestroy(a)); CK(cudaEventDestroy(b));
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    fprintf(stderr, "%s: %d SMs, L2 %d MiB, clock %d MHz\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20, p.clockRate / 1000);
    const size_t max_span = (size_t)4 << 30;
    char* d; CK(cudaMalloc(&d, max_span + 256));
    CK(cudaMemset(d, 0x5A, max_span));
    unsigned long long* sink; CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
    const int sms = p.multiProcessorCount;
    const uint64_t N = (uint64_t)1 << 30;
    printf("kind,bytes_per_probe,lanes_per_line,unroll,ctas_per_sm,span_MiB,ms,Gprobes_per_s,GBps_useful\n");
    const size_t spans[] = {(size_t)8 << 20, (size_t)16 << 20, (size_t)32 << 20, (size_t)64 << 20,
                            (size_t)128 << 20, (size_t)256 << 20, (size_t)1 << 30, (size_t)4 << 30};
    for (size_t s : spans) {
        run<8, 1, 8>(d, s, sink, sms, N, 8);
        run<32, 1, 4>(d, s, sink, sms, N, 8);
    }
    // bytes per probe, L2-resident and HBM-resident
    for (size_t s : {(size_t)16 << 20, (size_t)4 << 30}) {
        run<4, 1, 8>(d, s, sink, sms, N, 8);
        run<16, 1, 8>(d, s, sink, sms, N, 8);
        run<8, 1, 16>(d, s, sink, sms, N, 8);
        run<8, 1, 4>(d, s, sink, sms, N, 8);
        run<8, 1, 8>(d, s, sink, sms, N, 4);
        run<32, 1, 8>(d, s, sink, sms, N, 8);
    }
    // locality: G adjacent lanes in one line
    for (size_t s : {(size_t)16 << 20, (size_t)64 << 20, (size_t)4 << 30}) {
        run<8, 2, 8>(d, s, sink, sms, N, 8);
        run<8, 4, 8>(d, s, sink, sms, N, 8);
        run<8, 8, 8>(d, s, sink, sms, N, 8);
        run<8, 16, 8>(d, s, sink, sms, N, 8);
        run<32, 4, 4>(d, s, sink, sms, N, 8);
    }
    return 0;
}
