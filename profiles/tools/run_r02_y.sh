#!/bin/bash
# host-side packing inside hast_submit_batch: parity, then the e2e legs of the bench line
O=gpurun_out; T=${1:-r02_y}; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_pack or packed_batches or fused_parity" > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-cfg3 > $O/${T}_bench.json 2> $O/${T}_bench.log; echo "bench rc=$?"; tail -3 $O/${T}_bench.log
python - <<P
import json
d=json.load(open("$O/${T}_bench.json")); e=d["e2e"]
print("e2e ascii %.1f M pairs/s (%.1f ms) | host_packed %.1f M pairs/s (%.1f ms, %d threads) | pre-packed %.1f M pairs/s" % (e["value"]/1e6, e["ms_per_step"], e["host_packed"]["value"]/1e6, e["host_packed"]["ms_per_step"], e["host_packed"]["host_pack_threads"], e["packed"]["value"]/1e6))
P
for HP in 6 10 16; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cfg3 --host-pack-threads $HP 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); e=d['e2e']['host_packed']; print('host_pack_threads $HP: %.1f M pairs/s (%.1f ms/step)' % (e['value']/1e6, e['ms_per_step']))"; done
