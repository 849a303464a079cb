// merge_result_main.cpp -- `mergeResult`, drop-in for 01.classify_stlfr_reads/mergeResult.cpp.
//
// Sums several `classify` tables and re-calls the haplotypes.  The reference adds
// BOTH count columns into key 0 (mergeResult.cpp:28-29), so key 1 never exists, the
// ratio branch (:36-43) and both weights are dead, and each line comes out as
//     barcode \t (sum > 0 ? 0 : -1) \t count0+count1 \t 0
// That shipped behaviour is reproduced by default (bit-exact parity).  The evidently
// intended behaviour -- sum the columns separately, then the ratio rule with the
// documented weights (:77-79) -- is available behind --intended.
// Pure host program: text in, text out, nothing to accelerate.
#include <getopt.h>

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

static float g_w0 = 1.0f, g_w1 = 1.0f;      // mergeResult.cpp:8-9 (float, promoted in the product)

struct Tally { bool has[2] = {false, false}; int cnt[2] = {0, 0}; };

static void incr(std::map<std::string, Tally>& m, const std::string& bc, int hap, int v) {   // :16-20
    Tally& t = m[bc];
    if (!t.has[hap]) { t.has[hap] = true; t.cnt[hap] = 0; }
    t.cnt[hap] += v;
}

static int get_hap(const std::string& bc, const Tally& t) {                                  // :33-53
    if (bc == "0_0_0" || bc == "0_0" || bc == "0") return -1;
    if (t.has[0] && t.has[1]) {
        double d0 = double(t.cnt[0]);
        double d1 = double(t.cnt[1]);
        d0 *= g_w0;
        d1 *= g_w1;
        if (d0 > d1) return 0;
        if (d1 > d0) return 1;
        return -1;
    } else if (t.has[0]) {
        return t.cnt[0] > 0 ? 0 : -1;
    } else if (t.has[1]) {
        return t.cnt[1] > 0 ? 1 : -1;
    }
    return -1;
}

static void usage() {                                                                         // :70-81
    std::cerr << "Usage :\n"
                 "    mergeResult --input i1.txt [--input i2.txt --input ...] [OPTIONS]\n\n"
                 "Options:\n"
                 "    --input         result file to be merged.\n"
                 "    --weight0       weight of hap0\n"
                 "                    notice: weight0=(#hap0's unshared kmers)*(weight0 of classify)\n"
                 "    --weight1       weight of hap1\n"
                 "                    notice: weight1=(#hap1's unshared kmers)*(weight1 of classify)\n"
                 "    --intended      (B200 build) sum count0 and count1 separately and apply the weights\n";
}

int main(int argc, char** argv) {
    static struct option long_options[] = {{"input", required_argument, nullptr, 'i'},
                                           {"weight0", required_argument, nullptr, 'w'},
                                           {"weight1", required_argument, nullptr, 'u'},
                                           {"help", no_argument, nullptr, 'h'},
                                           {"intended", no_argument, nullptr, 1000},
                                           {nullptr, 0, nullptr, 0}};
    static const char optstring[] = "p:w:u:h";      // :91 -- no 'i': only the long --input works
    std::vector<std::string> inputs;
    bool intended = false;
    for (;;) {
        int c = getopt_long(argc, argv, optstring, long_options, nullptr);
        if (c < 0) break;
        switch (c) {
            case 'i': inputs.emplace_back(optarg); break;
            case 'u': g_w1 = (float)atof(optarg); break;
            case 'w': g_w0 = (float)atof(optarg); break;
            case 1000: intended = true; break;
            case 'h':
            default: usage(); return -1;
        }
    }
    if (inputs.empty()) { usage(); return -1; }
    std::map<std::string, Tally> cache;
    for (const std::string& path : inputs) {
        std::ifstream in(path);
        if (!in.is_open()) {
            std::cerr << "ERROR : failed to open " << path << " to read !!! exit ... " << std::endl;
            return -1;
        }
        std::string line;
        while (!std::getline(in, line).eof()) {                                               // :124
            std::string bc;
            int type = 0, h0 = 0, h1 = 0;
            std::istringstream is(line);
            is >> bc >> type >> h0 >> h1;                                                     // :26-27
            incr(cache, bc, 0, h0);
            incr(cache, bc, intended ? 1 : 0, h1);                                            // :28-29
        }
    }
    std::ios::sync_with_stdio(false);
    for (const auto& kv : cache)                                                              // :60-69
        std::cout << kv.first << '\t' << get_hap(kv.first, kv.second) << '\t'
                  << (kv.second.has[0] ? kv.second.cnt[0] : 0) << '\t'
                  << (kv.second.has[1] ? kv.second.cnt[1] : 0) << '\n';
    return 0;
}
