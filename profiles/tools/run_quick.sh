#!/bin/bash
# quick GPU pass: parity file + device-resident bench per workload/kernel.  run_quick.sh TAG "kernels" "workloads" [pytest -k expr]
O=gpurun_out; T=$1; KS=${2:-3}; WS=${3:-"cfg2 cfg3t"}; mkdir -p $O
if [ -n "$4" ]; then python -m pytest tests/test_gpu_parity.py -x -q -k "$4" > $O/${T}_pytest.log 2>&1; else python -m pytest tests/test_gpu_parity.py -x -q > $O/${T}_pytest.log 2>&1; fi
echo "pytest rc=$?"; tail -2 $O/${T}_pytest.log
for W in $WS; do for K in $KS; do
  python bench.py --workload $W --kernel $K --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/${T}_bench_${W}_k$K.json 2> $O/${T}_bench_${W}_k$K.log; echo "bench $W k$K rc=$?"
  python - <<P
import json
d=json.load(open("$O/${T}_bench_${W}_k$K.json"))
print("$W kernel $K: value %.3f G pairs/s, %.1f G lookups/s, ms/launch %.3f, pass %.4f, loads/lookup %.3f" % (d["value"]/1e9, d["roofline"]["lookups_per_s"]/1e9, d["roofline"]["ms_per_launch"], d["stats"]["filter_pass_frac"], d["stats"]["filter_loads_per_lookup"]))
P
done; done
