#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), ncu launch list + one full capture of the hot kernel.
# usage: profiles/tools/run_round.sh TAG      (outputs under gpurun_out/TAG_*)
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
tail -3 $O/${TAG}_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.log; echo "bench rc=$?"
python bench.py --workload cfg3t --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_cfg3t.json 2> $O/${TAG}_bench_cfg3t.log; echo "bench cfg3t rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.log; echo "ref rc=$?"
KREG='regex:^(classify_kernel|tile_kernel|lookup_kernel|gather_kernel|table_)'
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k "$KREG" -c 400 --csv \
    --log-file $O/${TAG}_launches.csv python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 3 > $O/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:^classify_kernel -s 32 -c 2 -f \
    -o $O/${TAG}_classify python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 3 > $O/${TAG}_ncu_full.log 2>&1
ncu -i $O/${TAG}_classify.ncu-rep --page raw --csv > $O/${TAG}_classify_raw.csv 2>/dev/null
ncu -i $O/${TAG}_classify.ncu-rep --page source --csv --print-source sass > $O/${TAG}_classify_sass.csv 2>/dev/null
ls -la $O | tail -20
cat $O/${TAG}_bench.json
