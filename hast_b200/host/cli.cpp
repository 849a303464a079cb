// cli.cpp -- option table and usage text of `classify`.
// Mirrors classify.cpp:280-310 (usage) and :375-428 (getopt table, validation,
// exit code): unknown option, -h, the dead "-l x", a missing --hap0/--hap1/--read
// or --thread < 1 all print the usage on stderr and exit with (unsigned char)-1.
#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "host.h"

namespace hasthost {

void print_usage() {
    fputs("\n"
          "Uasge :\n"
          "    classify --hap0 hap0 --hap1 hap1 --read read1.fq [options]\n"
          "\n"
          "Options:\n"
          "        -h/--help                       print this uasge and exit.\n"
          "        -p/--hap0                       unshared kmer set of hap0.\n"
          "        -m/--hap1                       unshared kmer set of hap1.\n"
          "        -r/--read                       filial reads in fastq format. gzip file must be ended by \".gz\".\n"
          "        -t/--thread   (8 default)       host threads that parse FASTQ for the GPUs.\n"
          "        -w/--weight0  (1.0 default)     weight of hap0.\n"
          "        -u/--weight1  (1.0 default)     weight of hap1.\n"
          "        -f/--adaptor_f                  forward adaptor sequence.\n"
          "                                        default \"CTGTCTCTTATACACATCTTAGGAAGACAAGCACTGACGACATGA\"\n"
          "        -q/--adaptor_r                  reverse adaptor sequence.\n"
          "                                        default \"TCTGCTGAGTCGAGAACGTCTCTGTGAGCCAAGGAGTTGCTCTGG\"\n"
          "\n"
          "B200 build only (long form):\n"
          "        --gpus N                        GPUs of this box to use (default 1; env HAST_GPUS).  One GPU classifies\n"
          "                                        ~100x faster than the host can read FASTQ: more only add start-up time.\n"
          "        --stats-json FILE               write throughput statistics as JSON.\n"
          "        --split-barcodes                also write {paternal,maternal,homozygous}.unique.barcodes\n"
          "                                        (the three awk passes of classify_stlfr_reads.sh).\n"
          "        --partition-reads               also write NAME.{paternal,maternal,homozygous,nobarcode}.fastq\n"
          "                                        and filter_reads.log for every --read (quartering_fastq.awk).\n"
          "        --outdir DIR                    where those extra files go (default: current directory).\n"
          "        (reads of up to 24560 bases; a longer read stops the run with its name.)\n"
          "\n"
          "Examples:\n"
          "    ./classify --hap0 p.kmers --hap1 m.kmers --read input.fastq.gz\n"
          "\n"
          "    ./classify --hap0 p.kmers --hap1 m.kmers --read input.L01.fastq.gz --read input.L02.fastq.gz\n"
          "\n"
          "Output format:\n"
          "barcode\thaplotype(0/1/-1)\tkmer_count_hap0\tkmer_count_hap1\n"
          "\n"
          "Usage done.\n",
          stderr);
}

int parse_options(int argc, char** argv, Options& opt) {
    static struct option long_options[] = {
        {"hap0", required_argument, nullptr, 'p'},      {"hap1", required_argument, nullptr, 'm'},
        {"read", required_argument, nullptr, 'r'},      {"thread", required_argument, nullptr, 't'},
        {"weight0", required_argument, nullptr, 'w'},   {"weight1", required_argument, nullptr, 'u'},
        {"adaptor_f", required_argument, nullptr, 'f'}, {"adaptor_r", required_argument, nullptr, 'q'},
        {"help", no_argument, nullptr, 'h'},            {"gpus", required_argument, nullptr, 1000},
        {"stats-json", required_argument, nullptr, 1001}, {"split-barcodes", no_argument, nullptr, 1002},
        {"partition-reads", no_argument, nullptr, 1003}, {"outdir", required_argument, nullptr, 1004},
        {nullptr, 0, nullptr, 0}};
    // classify.cpp:387 -- includes the dead "l:" so that "-l x" falls to the usage branch
    static const char optstring[] = "p:m:l:r:t:w:u:f:q:h";
    if (const char* e = getenv("HAST_GPUS")) opt.gpus = atoi(e);
    if (const char* e = getenv("HAST_STATS_JSON")) opt.stats_json = e;
    if (const char* e = getenv("HAST_PACKED")) opt.packed_h2d = atoi(e) != 0;
    if (const char* e = getenv("HAST_BLOCK_MB")) opt.batch_bytes = (size_t)atol(e) << 20;
    optind = 1;
    for (;;) {
        int c = getopt_long(argc, argv, optstring, long_options, nullptr);
        if (c < 0) break;
        switch (c) {
            case 'f': opt.adaptor_f = optarg; break;
            case 'q': opt.adaptor_r = optarg; break;
            case 'p': opt.hap0 = optarg; break;
            case 'm': opt.hap1 = optarg; break;
            case 'r': opt.reads.emplace_back(optarg); break;
            case 't': opt.threads = atoi(optarg); break;       // classify.cpp:411
            case 'u': opt.weight1 = atof(optarg); break;       // :414
            case 'w': opt.weight0 = atof(optarg); break;       // :417
            case 1000: opt.gpus = atoi(optarg); break;
            case 1001: opt.stats_json = optarg; break;
            case 1002: opt.split_barcodes = true; break;
            case 1003: opt.partition_reads = opt.split_barcodes = true; break;
            case 1004: opt.outdir = optarg; break;
            case 'h':
            default:
                print_usage();
                return 255;                                    // `return -1` from main
        }
    }
    if (opt.hap0.empty() || opt.hap1.empty() || opt.reads.empty() || opt.threads < 1) {   // :425-428
        print_usage();
        return 255;
    }
    return 0;
}

}  // namespace hasthost
