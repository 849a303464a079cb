#!/bin/bash
# A/B of library builds: run_ab.sh TAG "lib1.so lib2.so ..." "workloads"   (device-resident bench, kernel default)
O=gpurun_out; T=$1; LIBS=$2; WS=${3:-"cfg2 cfg3t"}; mkdir -p $O
for L in $LIBS; do
  N=$(basename $L .so)
  HAST_B200_LIB=$PWD/$L python -m pytest tests/test_gpu_parity.py -x -q -k "prefilter_mini and (fused or low_complex or saturated or packed)" > $O/${T}_${N}_pytest.log 2>&1; echo "$N pytest rc=$? $(tail -1 $O/${T}_${N}_pytest.log)"
  for W in $WS; do
    HAST_B200_LIB=$PWD/$L python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg3 > $O/${T}_${N}_$W.json 2> $O/${T}_${N}_$W.log
    python - <<P
import json
d=json.load(open("$O/${T}_${N}_$W.json"))
print("$N $W: value %.3f G pairs/s, %.1f G lookups/s, ms/launch %.3f" % (d["value"]/1e9, d["roofline"]["lookups_per_s"]/1e9, d["roofline"]["ms_per_launch"]))
P
  done
done
