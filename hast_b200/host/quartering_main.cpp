// quartering_main.cpp -- `quartering_fastq`, a drop-in for the awk invocation of
// classify_stlfr_reads.sh:176-185:
//     [gzip -dc $x |] awk -v prefix=$name -F '#|/' -f quartering_fastq.awk
//          paternal.unique.barcodes maternal.unique.barcodes homozygous.unique.barcodes  {$x | -}
// becomes
//     quartering_fastq --prefix $name paternal.unique.barcodes maternal.unique.barcodes
//          homozygous.unique.barcodes {$x | -}
// Same outputs in the current directory ($name.{paternal,maternal,homozygous,nobarcode}.fastq,
// filter_reads.log appended), same stderr complaints about unknown barcodes.  A ".gz" input is
// inflated in-process, so the `gzip -dc |` stage can go; in that case pass --filename - to keep
// the "-" that awk logs as FILENAME when it reads the pipe.
#include <getopt.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "host.h"

static void usage() {
    fputs("Usage: quartering_fastq --prefix NAME [--filename TEXT] [--outdir DIR] "
          "PATERNAL.barcodes MATERNAL.barcodes HOMOZYGOUS.barcodes INPUT.fastq[.gz]|-\n", stderr);
}

int main(int argc, char** argv) {
    static struct option lo[] = {{"prefix", required_argument, nullptr, 'p'},
                                 {"filename", required_argument, nullptr, 'f'},
                                 {"outdir", required_argument, nullptr, 'o'},
                                 {"help", no_argument, nullptr, 'h'},
                                 {nullptr, 0, nullptr, 0}};
    std::string prefix, filename, outdir;
    bool have_prefix = false, have_filename = false;
    for (;;) {
        const int c = getopt_long(argc, argv, "p:f:o:h", lo, nullptr);
        if (c < 0) break;
        switch (c) {
            case 'p': prefix = optarg; have_prefix = true; break;
            case 'f': filename = optarg; have_filename = true; break;
            case 'o': outdir = optarg; break;
            default: usage(); return 255;
        }
    }
    if (argc - optind != 4) { usage(); return 255; }
    const std::string input = argv[optind + 3];
    if (!have_prefix) prefix = hasthost::partition_prefix(input);   // awk: an unset prefix gives ".paternal.fastq"
    if (!have_filename) filename = input;
    hasthost::BarcodeLists lists;
    static const int kType[3] = {hasthost::BarcodeLists::kPaternal, hasthost::BarcodeLists::kMaternal,
                                 hasthost::BarcodeLists::kHomozygous};
    for (int i = 0; i < 3; ++i) {
        const std::string e = lists.load(argv[optind + i], kType[i]);
        if (!e.empty()) { fprintf(stderr, "ERROR : %s\n", e.c_str()); return 2; }
    }
    hasthost::PartitionStats st;
    const std::string e = hasthost::partition_fastq(input, filename, prefix, outdir, lists, st);
    if (!e.empty()) { fprintf(stderr, "ERROR : %s\n", e.c_str()); return 2; }
    return 0;
}
