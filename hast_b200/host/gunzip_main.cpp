// gunzip_main.cpp -- `hast_gunzip FILE.gz`: GzipInflater (inflate.h) to stdout.  A test and timing tool for the
// decoder the FASTQ readers use; `--null` decodes without writing, `--zlib` runs zlib's gzread instead.
#include <zlib.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "inflate.h"
#include "inflate_par.h"

int main(int argc, char** argv) {
    bool null_out = false, use_zlib = false;
    int threads = 1;
    size_t chunk = 1u << 20;
    std::string path;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--null")) null_out = true;
        else if (!strcmp(argv[i], "--zlib")) use_zlib = true;
        else if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--chunk") && i + 1 < argc) chunk = (size_t)atol(argv[++i]);
        else path = argv[i];
    }
    if (path.empty()) { fputs("usage: hast_gunzip [--null] [--zlib] [--threads N [--chunk BYTES]] FILE.gz\n", stderr); return 2; }
    const auto t0 = std::chrono::steady_clock::now();
    unsigned long long total = 0;
    if (use_zlib) {
        gzFile gz = gzopen(path.c_str(), "rb");
        if (!gz) { fprintf(stderr, "cannot open %s\n", path.c_str()); return 1; }
        gzbuffer(gz, 1u << 20);
        std::vector<char> buf(4u << 20);
        for (;;) {
            const int r = gzread(gz, buf.data(), (unsigned)buf.size());
            if (r < 0) { int e; fprintf(stderr, "error: %s\n", gzerror(gz, &e)); return 1; }
            if (r == 0) break;
            total += (unsigned long long)r;
            if (!null_out) fwrite(buf.data(), 1, (size_t)r, stdout);
        }
        gzclose(gz);
    } else if (threads > 1) {
        hasthost::ParallelGzip inf(threads, chunk);
        const std::string e = inf.open(path);
        if (!e.empty()) { fprintf(stderr, "%s\n", e.c_str()); return 1; }
        const uint8_t* p;
        size_t n;
        while (inf.next(&p, &n)) {
            total += n;
            if (!null_out) fwrite(p, 1, n, stdout);
        }
        if (!inf.error().empty()) { fflush(stdout); fprintf(stderr, "error: %s\n", inf.error().c_str()); return 1; }
        const auto st = inf.stats();
        fprintf(stderr, "parallel: %llu chunks in %llu batches, %llu block starts found, %llu dropped\n",
                (unsigned long long)st.chunks, (unsigned long long)st.batches, (unsigned long long)st.starts_found,
                (unsigned long long)st.starts_dropped);
    } else {
        hasthost::GzipInflater inf;
        const std::string e = inf.open(path);
        if (!e.empty()) { fprintf(stderr, "%s\n", e.c_str()); return 1; }
        const uint8_t* p;
        size_t n;
        while (inf.next(&p, &n)) {
            total += n;
            if (!null_out) fwrite(p, 1, n, stdout);
        }
        if (!inf.error().empty()) { fflush(stdout); fprintf(stderr, "error: %s\n", inf.error().c_str()); return 1; }
    }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "%llu bytes in %.3f s = %.1f MB/s\n", total, dt, total / dt / 1e6);
    return 0;
}
