#!/bin/bash
# round 2, pass H (1 GPU): stage 00, home line by minimizer + overflow by k-mer hash: tests, throughput, one ncu capture
O=gpurun_out; T=${1:-r02_h}; mkdir -p $O
python -m pytest tests/test_stage00.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"; grep -E "^FAILED|Error" $O/${T}_pytest.log | head -5
python profiles/tools/bench_stage00.py --steps 3 > $O/${T}_stage00.json 2> $O/${T}_stage00.log; echo "stage00 bench rc=$?"; cut -c1-700 $O/${T}_stage00.json; python -c "
import json; d=json.load(open('$O/${T}_stage00.json')); print('e2e', d.get('e2e',{}).get('wall_s'))"
ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:^kc_count_kernel -s 2 -c 1 -f -o $O/${T}_kc python profiles/tools/bench_stage00.py --steps 1 --no-e2e > $O/${T}_kc_ncu.log 2>&1
ncu -i $O/${T}_kc.ncu-rep --page raw --csv > $O/${T}_kc_raw.csv 2>/dev/null; python profiles/tools/ncu_raw.py $O/${T}_kc_raw.csv
