// classify_main.cpp -- `classify`, the drop-in for 01.classify_stlfr_reads/classify
// (process interface of classify.cpp:373-450: argv, stdout table, stderr log, exit code).
#include <cstdio>

#include "host.h"

int main(int argc, char** argv) {
    hasthost::Options opt;
    int rc = hasthost::parse_options(argc, argv, opt);
    if (rc) return rc;
    hasthost::RunStats st;
    rc = hasthost::run_classify(opt, st);
    if (rc == 0 && !opt.stats_json.empty()) {
        if (FILE* f = fopen(opt.stats_json.c_str(), "w")) {
            const double pairs = st.reads / 2.0;
            fprintf(f,
                    "{\"reads\": %llu, \"bases\": %llu, \"fastq_text_bytes\": %llu, \"barcodes\": %llu, "
                    "\"lookups\": %llu, \"gpus\": %d, \"parser_threads\": %d, \"table_bytes\": %llu, "
                    "\"size0\": %llu, \"size1\": %llu, \"kernel_launches\": %llu, \"t_table_s\": %.6f, "
                    "\"t_reads_s\": %.6f, \"t_finish_s\": %.6f, \"t_print_s\": %.6f, \"t_total_s\": %.6f, "
                    "\"t_split_s\": %.6f, \"t_partition_s\": %.6f, \"partition_text_bytes\": %llu, "
                    "\"pairs_per_s_stream\": %.1f, \"pairs_per_s_total\": %.1f}\n",
                    (unsigned long long)st.reads, (unsigned long long)st.bases, (unsigned long long)st.text_bytes,
                    (unsigned long long)st.barcodes, (unsigned long long)st.lookups, st.gpus, st.parser_threads,
                    (unsigned long long)st.table_bytes, (unsigned long long)st.size0, (unsigned long long)st.size1,
                    (unsigned long long)st.kernel_launches, st.t_table, st.t_reads, st.t_finish, st.t_print,
                    st.t_total, st.t_split, st.t_partition, (unsigned long long)st.partition_text_bytes,
                    st.t_reads > 0 ? pairs / st.t_reads : 0.0, st.t_total > 0 ? pairs / st.t_total : 0.0);
            fclose(f);
        }
    }
    return rc;
}
