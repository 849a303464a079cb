#!/bin/bash
# final CLI numbers of the round (sync before the timed legs): 16 M pairs plain + gzip, with --split-barcodes --partition-reads timing
O=gpurun_out; T=${1:-r02_v}; mkdir -p $O
python profiles/tools/bench_cli.py --pairs 16000000 --skip-zlib --no-reference > $O/${T}_cli16m.json 2> $O/${T}_cli16m.log; echo "cli16m rc=$?"; cat $O/${T}_cli16m.json
python - <<'P' > $O/r02_v_files.log 2>&1
import os, sys; sys.path.insert(0, ".")
from hast_b200 import synth
s = synth.config("cfg2"); s.n_pairs = 8_000_000; s.n_barcodes = 200_000
t = synth.make_trio(s, device="cuda")
t.write_kmer_lists("/tmp/pt"); print(t.write_fastq("/tmp/pt", gz=False)); os.sync()
P
mkdir -p /tmp/pt/out; for rep in 1 2; do
./bin/classify --hap0 /tmp/pt/paternal.unique.filter.mer --hap1 /tmp/pt/maternal.unique.filter.mer --weight0 1.04 --thread 14 --read /tmp/pt/child.r1.fq --read /tmp/pt/child.r2.fq --split-barcodes --partition-reads --outdir /tmp/pt/out --stats-json /tmp/pt/s.json > /tmp/pt/out/phased.barcodes 2>/tmp/pt/err
python -c "
import json; d=json.load(open('/tmp/pt/s.json')); print('8 M pairs plain (3.96 GB) with --partition-reads: total %.2f s; table %.2f, stream %.2f (%.1f M pairs/s), print %.2f, split %.2f, partition %.2f s = %.2f GB/s of text' % (d['t_total_s'], d['t_table_s'], d['t_reads_s'], d['pairs_per_s_stream']/1e6, d['t_print_s'], d['t_split_s'], d['t_partition_s'], d['partition_text_bytes']/d['t_partition_s']/1e9))"
done
