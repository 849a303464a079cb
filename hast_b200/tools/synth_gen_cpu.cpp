// synth_gen_cpu.cpp -- host side of the counter-based read-pair generator (see synth_gen.h): the same
// functions as synth_gen.cu, for the CPU-only tests and for writing FASTQ files.  Synthetic-data tooling.
#include <stdint.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "synth_gen.h"

namespace {
template <class F>
void parallel_for(uint64_t n, F f) {
    const unsigned t = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(std::thread::hardware_concurrency(), n / 4096 + 1));
    if (t == 1) { f(0, n); return; }
    std::vector<std::thread> th;
    for (unsigned i = 0; i < t; ++i) th.emplace_back([=] { f(n * i / t, n * (i + 1) / t); });
    for (auto& x : th) x.join();
}
}  // namespace

extern "C" {

// rows [0, n) of bases/bc = r1 of the pairs, rows [n, 2n) = r2.  idx == NULL: pairs lo .. lo+n-1.
int sg_gen_pairs_host(const sg_params* p, const uint64_t* cdf, const uint8_t* hap0, const uint8_t* hap1,
                      const uint64_t* idx, uint64_t lo, uint64_t n, uint8_t* bases, uint32_t* bc) {
    const uint32_t L = p->read_len;
    parallel_for(n, [=](uint64_t a, uint64_t b) {
        for (uint64_t w = a; w < b; ++w) {
            const uint64_t i = idx ? idx[w] : lo + w;
            const sg_pair pr = sg_pair_of(p, cdf, i);
            bc[w] = pr.barcode;
            bc[n + w] = pr.barcode;
            for (uint32_t mate = 0; mate < 2; ++mate) {
                const sg_edits e = sg_edits_of(p, i, mate);
                uint8_t* out = bases + (mate ? n + w : w) * (uint64_t)L;
                for (uint32_t j = 0; j < L; ++j) out[j] = sg_base(p, hap0, hap1, &pr, &e, mate, j);
            }
        }
    });
    return 0;
}

int sg_barcodes_host(const sg_params* p, const uint64_t* cdf, uint64_t lo, uint64_t n, uint32_t* bc) {
    parallel_for(n, [=](uint64_t a, uint64_t b) {
        for (uint64_t w = a; w < b; ++w) bc[w] = sg_pair_of(p, cdf, lo + w).barcode;
    });
    return 0;
}

}  // extern "C"

// Names of barcode ids 0..B (id B = "0_0_0"): distinct a_b_c triples, a, b, c in 1..1536.  id -> triple code
// is multiplication by a constant coprime to 1536^3 = 2^27 3^3 modulo 1536^3, a bijection of the code space.
// blob receives NUL-terminated names, off[id] their starts (off has B + 2 entries).  Returns bytes used.
#include <cstdio>
#include <cstring>
extern "C" uint64_t sg_barcode_names(uint64_t B, char* blob, uint64_t* off, uint64_t, void*, void*) {
    const uint64_t M = 1536ull * 1536ull * 1536ull;
    uint64_t used = 0;
    for (uint64_t id = 0; id <= B; ++id) {
        off[id] = used;
        if (id == B) { memcpy(blob + used, "0_0_0", 6); used += 6; break; }
        const uint64_t code = (id % M) * 1000003ull % M;
        used += (uint64_t)sprintf(blob + used, "%llu_%llu_%llu", (unsigned long long)(code / (1536 * 1536) + 1),
                                  (unsigned long long)((code / 1536) % 1536 + 1), (unsigned long long)(code % 1536 + 1)) + 1;
    }
    off[B + 1] = used;
    return used;
}
