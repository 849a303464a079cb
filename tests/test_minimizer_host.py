"""CPU: the host side of the minimizer-addressed pre-filter (hast_b200/csrc/table.cuh).

table_insert (device) and the MINI sweep of classify_kernel must derive the same filter word and the same
two bit selectors for a k-mer, from either strand.  The functions are __host__ __device__, so the properties
the sweep relies on are checked here with a small g++ harness: strand symmetry of minimizer_hash and mini_sel,
agreement of the window formulation the sweep uses (first / last 16 bases of both strands) with the packed
k-mer formulation the insert uses, and the locality that is the point of it (consecutive k-mers of a sequence
mostly share their minimizer).
"""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

HARNESS = r"""
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include "table.cuh"
using namespace hast;
static uint64_t revcomp(uint64_t x, int k) {            // kmer.h:13,196-223: complement = ^2, reversed
    uint64_t r = 0;
    for (int i = 0; i < k; ++i) { r = (r << 2) | ((x & 3) ^ 2); x >>= 2; }
    return r;
}
int main() {
    uint64_t seed = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return seed; };
    const int ks[3] = {21, 25, 31};
    long bad = 0, shared = 0, total = 0;
    for (int ki = 0; ki < 3; ++ki) {
        const int k = ks[ki], m = mini_len(k);
        if (m != 16) { printf("mini_len(%d) = %d\n", k, m); return 1; }
        const uint64_t mask = kmer_mask(k);
        for (int it = 0; it < 20000; ++it) {
            // a random 64-base sequence, 2 bits per base, base i in seq[i]
            uint8_t seq[64];
            for (int i = 0; i < 64; ++i) seq[i] = (it % 7 == 0 && i % 3) ? seq[i - 1] : (uint8_t)(rnd() >> 61 & 3);
            uint32_t prev = 0;
            for (int p = 0; p + k <= 64; ++p) {
                uint64_t f = 0;
                for (int i = 0; i < k; ++i) f = (f << 2) | seq[p + i];
                const uint64_t r = revcomp(f, k);
                const uint64_t canon = f < r ? f : r, other = f < r ? r : f;
                const uint32_t h1 = minimizer_hash(canon, other, k, m), h2 = minimizer_hash(other, canon, k, m);
                const uint32_t s1 = mini_sel(canon, other, k), s2 = mini_sel(other, canon, k);
                bad += (h1 != h2) + (s1 != s2) + (s1 >= 1024u);
                // the sweep's formulation: 16-base windows of the forward strand and their reverse complements
                uint32_t best = 0xFFFFFFFFu;
                uint32_t fw[16], gw[16];
                for (int t = 0; t + 16 <= k; ++t) {
                    uint32_t w = 0;
                    for (int i = 0; i < 16; ++i) w = (w << 2) | seq[p + t + i];
                    const uint32_t g = (uint32_t)revcomp(w, 16);
                    fw[t] = w; gw[t] = g;
                    const uint32_t h = mini_hash(w < g ? w : g);
                    best = h < best ? h : best;
                }
                const int W = k - 16 + 1;
                const uint32_t sel = (fw[0] * kSelC1 + fw[W - 1] * kSelC2 + gw[W - 1] * kSelC1 + gw[0] * kSelC2) >> 22;
                bad += (best != h1) + (sel != s1);
                if (p) { shared += (h1 == prev); ++total; }
                prev = h1;
                (void)mask;
            }
        }
    }
    printf("bad %ld shared %.4f\n", bad, (double)shared / (double)total);
    return bad != 0;
}
"""


def test_minimizer_functions_on_the_host(tmp_path):
    src = tmp_path / "mini.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "mini"
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-I", str(ROOT / "hast_b200" / "csrc"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    bad, shared = r.stdout.split()[1], float(r.stdout.split()[3])
    assert bad == "0"
    # a new minimizer about every (w + 1) / 2 positions (w = k - 15 m-mers per k-mer): well over half of the
    # neighbouring k-mers share theirs even at k = 21
    assert shared > 0.6, shared
