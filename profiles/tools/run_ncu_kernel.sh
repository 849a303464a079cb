#!/bin/bash
# one `ncu --set full` capture of classify_kernel per kernel option given: run_ncu_kernel.sh TAG "3 1" [workload]
O=gpurun_out; T=${1:-rXX}; KS=${2:-3}; W=${3:-cfg2}; mkdir -p $O
for K in $KS; do
  ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:^classify_kernel -s 32 -c 1 -f \
      -o $O/${T}_k${K} python bench.py --workload $W --kernel $K --no-cpu-baseline --no-e2e --steps 1 --warmup 3 > $O/${T}_k${K}_ncu.log 2>&1
  ncu -i $O/${T}_k${K}.ncu-rep --page raw --csv > $O/${T}_k${K}_raw.csv 2>/dev/null
  ncu -i $O/${T}_k${K}.ncu-rep --page source --csv --print-source sass > $O/${T}_k${K}_sass.csv 2>/dev/null
  python profiles/tools/ncu_raw.py $O/${T}_k${K}_raw.csv
done
ls -la $O | tail
