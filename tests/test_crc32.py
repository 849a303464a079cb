"""CPU: hast_crc32 (hast_b200/host/crc32_clmul.h, PCLMULQDQ folding) against zlib's crc32 -- random offsets, lengths
around the 16 / 64-byte lane boundaries, non-zero starting values, incremental use."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

HARNESS = r"""
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "crc32_clmul.h"
using namespace hasthost;
int main() {
    std::vector<uint8_t> b((size_t)64 << 20);
    uint64_t x = 1234567;
    for (auto& c : b) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; c = (uint8_t)x; }
    long bad = 0;
    for (int it = 0; it < 20000; ++it) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const size_t off = x % 4096, len = (x >> 20) % (it % 50 == 0 ? 300000 : 700);
        const uint32_t seed = it % 3 ? (uint32_t)(x >> 33) : 0u;
        if (hast_crc32(seed, b.data() + off, len) != (uint32_t)crc32(seed, b.data() + off, (uInt)len)) ++bad;
    }
    // incremental use: the CRC of two pieces equals the CRC of the whole
    uint32_t a = hast_crc32(0, b.data(), 1000003);
    a = hast_crc32(a, b.data() + 1000003, 777);
    if (a != (uint32_t)crc32(0, b.data(), 1000003 + 777)) ++bad;
    auto t0 = std::chrono::steady_clock::now();
    uint32_t c1 = hast_crc32(0, b.data(), b.size());
    double d1 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    t0 = std::chrono::steady_clock::now();
    uint32_t c2 = (uint32_t)crc32(0, b.data(), (uInt)b.size());
    double d2 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("bad %ld  clmul %.2f GB/s  zlib %.2f GB/s  %s\n", bad, b.size() / d1 / 1e9, b.size() / d2 / 1e9, c1 == c2 ? "equal" : "DIFFER");
    return bad != 0 || c1 != c2;
}
"""


def test_crc32_clmul_equals_zlib(tmp_path):
    src = tmp_path / "crc.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "crc"
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-I", str(ROOT / "hast_b200" / "host"), str(src), "-lz", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("bad 0") and "equal" in r.stdout, r.stdout + r.stderr
