#!/bin/bash
# round 2, pass F (1 GPU): L2 set-aside for the pre-filter (persisting accesses) on both tables; CLI tests after the table-build thread
O=gpurun_out; T=${1:-r02_f}; mkdir -p $O
python -m pytest tests/test_cli_gpu.py tests/test_stage_script_gpu.py tests/test_stage03.py tests/test_gpu_parity.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"
python - <<'P'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "persisting max", getattr(p, "persistingL2CacheMaxSize", None) or getattr(p, "persisting_l2_cache_max_size", None))
P
for M in -1 0 32 64 80 128; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg3 --l2-persist-mib $M > $O/${T}_p${M}_cfg2.json 2> $O/${T}_p${M}_cfg2.log
  python bench.py --only-cfg3 --cfg3-pairs 80000000 --l2-persist-mib $M > $O/${T}_p${M}_cfg3.json 2> $O/${T}_p${M}_cfg3.log
  python - <<P
import json
try:
    d=json.load(open("$O/${T}_p${M}_cfg2.json")); c=json.load(open("$O/${T}_p${M}_cfg3.json"))
    print("persist $M MiB: cfg2 %.1f G lookups/s (%.4f ms/launch) parity %s | cfg3 %.1f G lookups/s (%.4f ms/launch) parity %s" % (d["roofline"]["lookups_per_s"]/1e9, d["roofline"]["ms_per_launch"], d["parity"]["ok"], c["roofline"]["lookups_per_s"]/1e9, c["roofline"]["ms_per_launch"], c["parity"]))
except Exception as e: print("persist $M failed", e)
P
done
python profiles/tools/bench_cli.py --pairs 4000000 --skip-zlib > $O/${T}_cli4m.json 2> $O/${T}_cli4m.log; echo "cli4m rc=$?"; cat $O/${T}_cli4m.json
