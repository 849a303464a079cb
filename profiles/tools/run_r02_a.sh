#!/bin/bash
# round 2, first pass: new generator on the device, full GPU suite, the whole bench line (cfg2 + cli + cfg3) at N=1
O=gpurun_out; T=${1:-r02_a}; mkdir -p $O
nproc > $O/${T}_host.txt; free -g >> $O/${T}_host.txt; df -h /tmp >> $O/${T}_host.txt; lscpu | head -20 >> $O/${T}_host.txt
S=$(date +%s); python -m pytest tests/test_synth_stream.py -m gpu -x -q > $O/${T}_pytest_synth.log 2>&1; echo "pytest synth rc=$? $(( $(date +%s)-S ))s $(tail -1 $O/${T}_pytest_synth.log)"
S=$(date +%s); timeout 1200 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.log; echo "bench rc=$? $(( $(date +%s)-S ))s"; tail -25 $O/${T}_bench.log
S=$(date +%s); python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s $(tail -1 $O/${T}_pytest.log)"
python - <<P
import json
d=json.load(open("$O/${T}_bench.json"))
print("value %.3f G pairs/s e2e %.1f M parity %s" % (d["value"]/1e9, d["e2e"]["value"]/1e6, d["parity"]))
print("cli", json.dumps(d.get("cli"))[:1500])
c=d.get("cfg3"); print("cfg3", json.dumps(c)[:3000])
P
