// hapcall.cpp -- the per-barcode haplotype call and the output table.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>

#include "host.h"

namespace hasthost {

// getHap, classify.cpp:66-86.  The reference keeps map<int,int> per barcode whose
// keys 0 / 1 exist only once a positive vote was added (classify.cpp:203-206),
// so "key present" is exactly "count > 0".  The ratio is formed in IEEE double in
// the reference's operation order: divide by the set size, then multiply by the
// weight (:70-73).
int get_hap(const std::string& barcode, int c0, int c1, uint64_t n0, uint64_t n1, double w0, double w1) {
    if (barcode == "0_0_0" || barcode == "0_0" || barcode == "0") return -1;
    const bool has0 = c0 > 0, has1 = c1 > 0;
    if (has0 && has1) {
        double df0 = double(c0) / double(n0);
        double df1 = double(c1) / double(n1);
        df0 *= w0;
        df1 *= w1;
        if (df0 > df1) return 0;
        if (df1 > df0) return 1;
        return -1;
    } else if (has0) {
        return 0;
    } else if (has1) {
        return 1;
    }
    return -1;
}

static inline char* put_int(char* p, long v) {
    char tmp[24];
    int n = 0;
    unsigned long u = v < 0 ? (unsigned long)(-v) : (unsigned long)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

// printBarcodeInfos, classify.cpp:93-102: every barcode seen, in std::map order
// (bytewise lexicographic, e.g. "0_0_0" < "10_1_1" < "1_2_3"), one line
// barcode \t hap \t count0 \t count1.
void print_table(FILE* out, const std::vector<std::string>& names, const int32_t* counts, uint64_t n0,
                 uint64_t n1, double w0, double w1, std::vector<uint32_t>* order_out, std::vector<int8_t>* hap_out) {
    std::vector<uint32_t> order(names.size());
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return names[a] < names[b]; });
    std::vector<char> buf(1u << 22);
    size_t len = 0;
    for (uint32_t id : order) {
        const std::string& nm = names[id];
        if (len + nm.size() + 64 > buf.size()) {
            fwrite(buf.data(), 1, len, out);
            len = 0;
            if (nm.size() + 64 > buf.size()) buf.resize(nm.size() + 64);
        }
        const int c0 = counts[2 * (size_t)id], c1 = counts[2 * (size_t)id + 1];
        memcpy(buf.data() + len, nm.data(), nm.size());
        char* p = buf.data() + len + nm.size();
        *p++ = '\t';
        const int hap = get_hap(nm, c0, c1, n0, n1, w0, w1);
        if (hap_out) hap_out->push_back((int8_t)hap);
        p = put_int(p, hap);
        *p++ = '\t';
        p = put_int(p, c0);
        *p++ = '\t';
        p = put_int(p, c1);
        *p++ = '\n';
        len = (size_t)(p - buf.data());
    }
    fwrite(buf.data(), 1, len, out);
    fflush(out);
    if (order_out) order_out->swap(order);
}

}  // namespace hasthost
