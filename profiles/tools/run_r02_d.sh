#!/bin/bash
# round 2, fourth pass (1 GPU): tile size chosen per batch vs the fixed 409, larger passes, pre-filter size on the
# HBM-resident table, inflate threads of the gzip path
O=gpurun_out; T=${1:-r02_d}; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_stage03.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"
for L in rpt409 auto auto_cap49152 auto_cap57344; do
  HAST_B200_LIB=$PWD/hast_b200/lib/ab/$L.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg3 > $O/${T}_ab_${L}_cfg2.json 2> $O/${T}_ab_${L}_cfg2.log
  HAST_B200_LIB=$PWD/hast_b200/lib/ab/$L.so python bench.py --only-cfg3 --cfg3-pairs 80000000 > $O/${T}_ab_${L}_cfg3.json 2> $O/${T}_ab_${L}_cfg3.log
  python - <<P
import json
try:
    d=json.load(open("$O/${T}_ab_${L}_cfg2.json")); c=json.load(open("$O/${T}_ab_${L}_cfg3.json"))
    print("$L: cfg2 %.1f G lookups/s (%.4f ms/launch) parity %s | cfg3 %.1f G lookups/s (%.4f ms/launch) parity %s" % (d["roofline"]["lookups_per_s"]/1e9, d["roofline"]["ms_per_launch"], d["parity"]["ok"], c["roofline"]["lookups_per_s"]/1e9, c["roofline"]["ms_per_launch"], c["parity"]))
except Exception as e: print("$L failed", e)
P
done
for F in 32 128; do
  python bench.py --only-cfg3 --cfg3-pairs 80000000 --filter-max-mib $F > $O/${T}_filter${F}_cfg3.json 2> $O/${T}_filter${F}_cfg3.log
  python - <<P
import json
try:
    c=json.load(open("$O/${T}_filter${F}_cfg3.json"))
    print("filter cap $F MiB: cfg3 %.1f G lookups/s (%.4f ms/launch), filter %d MiB, parity %s" % (c["roofline"]["lookups_per_s"]/1e9, c["roofline"]["ms_per_launch"], c["filter_bytes"]>>20, c["parity"]))
except Exception as e: print("filter $F failed", e)
P
done
python - <<'P' > $O/${T}_gzfiles.log 2>&1
import sys; sys.path.insert(0, ".")
from hast_b200 import synth
s = synth.config("cfg2"); s.n_pairs = 12_000_000; s.n_barcodes = 300_000
t = synth.make_trio(s, device="cuda")
t.write_kmer_lists("/tmp/gzt"); print(t.write_fastq("/tmp/gzt", gz=6))
P
for IT in 5 6 7 8; do for PT in 6 12; do
  HAST_INFLATE_THREADS=$IT ./bin/classify --hap0 /tmp/gzt/paternal.unique.filter.mer --hap1 /tmp/gzt/maternal.unique.filter.mer --weight0 1.04 --thread $PT --read /tmp/gzt/child.r1.fq.gz --read /tmp/gzt/child.r2.fq.gz --stats-json /tmp/gzt/s.json > /dev/null 2>/tmp/gzt/err
  python -c "
import json; d=json.load(open('/tmp/gzt/s.json')); print('inflate threads/file $IT, parser threads $PT: stream %.3f s = %.2f M pairs/s' % (d['t_reads_s'], d['pairs_per_s_stream']/1e6))"
done; done
