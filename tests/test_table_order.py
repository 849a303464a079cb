"""CPU: output order and formatting of the barcode table at scale (hast_b200/host/hapcall.cpp).

print_table orders 10^5..10^7 barcodes with a parallel sample sort on 8-byte keys and formats the lines on all
cores; the result has to be the std::map order of the reference (bytewise lexicographic, classify.cpp:50,93-102)
and must not depend on the thread count.  Names are built to break a key-based sort: one prefix shared by all,
groups that tie on the first 8 bytes after it, a name that is a proper prefix of another, the empty name, bytes
>= 0x80, names beyond the small-string size.
"""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

HARNESS = r"""
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <string>
#include <vector>
#include "host.h"
using namespace hasthost;
int main(int argc, char** argv) {
    const size_t n = 300000;
    const bool shared_prefix = argc > 1 && argv[1][0] == 'p';
    std::vector<std::string> names;
    uint64_t x = 88172645463325252ull;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    const std::string pre = shared_prefix ? "lib7_" : "";
    for (size_t i = 0; names.size() < n; ++i) {
        char b[96];
        const unsigned a = (unsigned)(rnd() % 1536 + 1), c = (unsigned)(rnd() % 1536 + 1);
        switch (i % 5) {
            case 0: snprintf(b, sizeof b, "%u_%u_%zu", a, c, i); break;
            case 1: snprintf(b, sizeof b, "SAMEKEY8%zu", i); break;                 // ties on the 8-byte key
            case 2: snprintf(b, sizeof b, "%zu", i); break;
            case 3: snprintf(b, sizeof b, "\xC3\xA9%u\xFF%zu", a, i); break;          // bytes >= 0x80
            default: snprintf(b, sizeof b, "a_barcode_name_well_beyond_the_small_string_size_%zu_%u", i, a); break;
        }
        names.push_back(pre + b);
    }
    names[7] = pre;                        // the shared prefix itself (the empty name when there is none)
    names[8] = pre + "SAMEKEY8";           // a proper prefix of the tie group
    names[9] = pre + "0_0_0";
    std::vector<int32_t> counts(2 * n);
    for (auto& c : counts) c = (int32_t)(rnd() % 5);
    std::vector<uint32_t> want(n);
    std::iota(want.begin(), want.end(), 0u);
    std::sort(want.begin(), want.end(), [&](uint32_t a, uint32_t b) { return names[a] < names[b]; });
    std::vector<uint32_t> order;
    std::vector<int8_t> haps;
    FILE* out = fopen(argv[2], "wb");
    print_table(out, names, counts.data(), 1000, 1100, 1.04, 1.0, &order, &haps);
    fclose(out);
    if (order != want) { printf("order differs\n"); return 1; }
    if (haps.size() != n) { printf("haps size\n"); return 1; }
    for (size_t i = 0; i < n; ++i)
        if (haps[i] != get_hap(names[order[i]], counts[2 * order[i]], counts[2 * order[i] + 1], 1000, 1100, 1.04, 1.0)) { printf("hap differs\n"); return 1; }
    printf("ok\n");
    return 0;
}
"""


def test_table_order_and_thread_independence(tmp_path):
    src = tmp_path / "order.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "order"
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", str(ROOT / "include"), "-I", str(ROOT / "hast_b200" / "host"),
                        str(src), str(ROOT / "hast_b200" / "host" / "hapcall.cpp"), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    import hashlib
    import os
    for mode in ("p", "n"):
        sums = set()
        for threads in ("1", "3", "8"):
            out = tmp_path / f"t_{mode}_{threads}.tsv"
            r = subprocess.run([str(exe), mode, str(out)], capture_output=True, text=True,
                               env=dict(os.environ, HAST_SORT_THREADS=threads))
            assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr
            sums.add(hashlib.md5(out.read_bytes()).hexdigest())
            assert out.read_bytes().count(b"\n") == 300000
        assert len(sums) == 1
