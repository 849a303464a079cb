#!/bin/bash
# round 2, pass G (1 GPU): stage 00 with the minimizer-addressed count table (tests, throughput, one ncu capture);
# plain-reader page population A/B
O=gpurun_out; T=${1:-r02_g}; mkdir -p $O
python -m pytest tests/test_stage00.py tests/test_gpu_parity.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"; grep -E "^FAILED|Error" $O/${T}_pytest.log | head -5
python profiles/tools/bench_stage00.py --steps 3 > $O/${T}_stage00.json 2> $O/${T}_stage00.log; echo "stage00 bench rc=$?"; cut -c1-900 $O/${T}_stage00.json
ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:^kc_count_kernel -s 2 -c 1 -f -o $O/${T}_kc python profiles/tools/bench_stage00.py --steps 1 --no-e2e > $O/${T}_kc_ncu.log 2>&1
ncu -i $O/${T}_kc.ncu-rep --page raw --csv > $O/${T}_kc_raw.csv 2>/dev/null; python profiles/tools/ncu_raw.py $O/${T}_kc_raw.csv
python - <<'P' > $O/r02_g_files.log 2>&1
import sys; sys.path.insert(0, ".")
from hast_b200 import synth
s = synth.config("cfg2"); s.n_pairs = 12_000_000; s.n_barcodes = 300_000
t = synth.make_trio(s, device="cuda")
t.write_kmer_lists("/tmp/gzt"); print(t.write_fastq("/tmp/gzt", gz=False))
P
for rep in 1 2 3; do for MODE in "" "HAST_NO_POPULATE=1"; do
  env $MODE ./bin/classify --hap0 /tmp/gzt/paternal.unique.filter.mer --hap1 /tmp/gzt/maternal.unique.filter.mer --weight0 1.04 --thread 14 --read /tmp/gzt/child.r1.fq --read /tmp/gzt/child.r2.fq --stats-json /tmp/gzt/s.json > /dev/null 2>/tmp/gzt/err
  python -c "
import json; d=json.load(open('/tmp/gzt/s.json')); print('plain, 14 parser threads, [$MODE]: stream %.3f s = %.2f M pairs/s, total %.2f s' % (d['t_reads_s'], d['pairs_per_s_stream']/1e6, d['t_total_s']))"
done; done
