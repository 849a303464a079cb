#!/bin/bash
# minimizer-addressed pre-filter (kernel 3) against the per-k-mer one (kernel 1): parity, then cfg2 / cfg3t device-resident bench
O=gpurun_out; T=${1:-r01_n}; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q > $O/${T}_pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${T}_pytest_parity.log
for W in cfg2 cfg3t; do for K in 3 1; do
  python bench.py --workload $W --kernel $K --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/${T}_bench_${W}_k$K.json 2> $O/${T}_bench_${W}_k$K.log; echo "bench $W k$K rc=$?"
  python - <<P
import json
d=json.load(open("$O/${T}_bench_${W}_k$K.json"))
print("$W kernel $K: value %.3f G pairs/s, %.1f G lookups/s, ms/launch %.3f, pass %.4f, loads/lookup %.3f, clocks %s" % (d["value"]/1e9, d["roofline"]["lookups_per_s"]/1e9, d["roofline"]["ms_per_launch"], d["stats"]["filter_pass_frac"], d["stats"]["filter_loads_per_lookup"], d["clocks"]))
P
done; done
