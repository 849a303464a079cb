#!/bin/bash
# round 2, third pass (1 GPU): full GPU suite, tile-size A/B of the fused kernel, DRAM-traffic captures (ncu --set full)
# of one launch on the cfg2 and cfg3 tables, the launch list of a bench run, the CLI tool with the copy-free gzip path
O=gpurun_out; T=${1:-r02_c}; mkdir -p $O
S=$(date +%s); python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s $(tail -1 $O/${T}_pytest.log)"; grep -E "^FAILED|^ERROR" $O/${T}_pytest.log | head
for L in base rpt320 rpt409; do for W in cfg2 cfg3t; do
  HAST_B200_LIB=$PWD/hast_b200/lib/ab/$L.so python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg3 > $O/${T}_ab_${L}_$W.json 2> $O/${T}_ab_${L}_$W.log
  python - <<P
import json
try:
    d=json.load(open("$O/${T}_ab_${L}_$W.json"))
    print("$L $W: value %.3f G pairs/s, %.1f G lookups/s per launch, ms/launch %.4f, parity %s" % (d["value"]/1e9, d["roofline"]["lookups_per_s"]/1e9, d["roofline"]["ms_per_launch"], d["parity"]["ok"]))
except Exception as e: print("$L $W failed", e)
P
done; done
S=$(date +%s); bash profiles/tools/capture_traffic.sh ${T} > $O/${T}_traffic.log 2>&1; echo "traffic rc=$? $(( $(date +%s)-S ))s"; tail -22 $O/${T}_traffic.log
S=$(date +%s); ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k "regex:classify_kernel|tile_kernel|table_|gather_kernel|lookup_kernel|sg_" -c 3000 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --cfg3-pairs 40000000 --cfg3-steps 1 > $O/${T}_launches_bench.log 2>&1; echo "launch list rc=$? $(( $(date +%s)-S ))s"; python profiles/tools/ncu_launches.py $O/${T}_launches.csv 2>/dev/null | head -20
S=$(date +%s); python profiles/tools/bench_cli.py --pairs 16000000 --skip-zlib --no-reference > $O/${T}_cli16m.json 2> $O/${T}_cli16m.log; echo "cli16m rc=$? $(( $(date +%s)-S ))s"; cat $O/${T}_cli16m.json
