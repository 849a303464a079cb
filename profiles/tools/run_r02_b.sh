#!/bin/bash
# round 2, second pass (1 GPU): new host pipeline (sliced plain reader, token/replay gzip decoder) through the whole
# bench line, the reference's stage script around bin/classify, the H2D probe, the CLI tool at 16 M pairs
O=gpurun_out; T=${1:-r02_b}; mkdir -p $O
S=$(date +%s); python -m pytest tests/test_stage_script_gpu.py tests/test_cli_gpu.py tests/test_stage00.py -m gpu -x -q > $O/${T}_pytest_cli.log 2>&1; echo "pytest cli rc=$? $(( $(date +%s)-S ))s $(tail -1 $O/${T}_pytest_cli.log)"
./bin/h2d_probe 256 1.0 > $O/${T}_h2d.json 2> $O/${T}_h2d.log; echo "h2d rc=$?"; cat $O/${T}_h2d.json
S=$(date +%s); timeout 1500 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.log; echo "bench rc=$? $(( $(date +%s)-S ))s"; tail -8 $O/${T}_bench.log
python - <<P
import json
d=json.load(open("$O/${T}_bench.json"))
print("value %.3f G pairs/s e2e %.1f M parity %s" % (d["value"]/1e9, d["e2e"]["value"]/1e6, d["parity"]["ok"]))
print("cli", json.dumps(d.get("cli"))[:1800])
c=d.get("cfg3"); print("cfg3 value %.3f G ms %.1f kernel %.1f reduce %.2f d2h %.2f parity %s" % (c["value"]/1e9, c["ms_per_step"], c["kernel_ms"], c["reduce_ms"], c["d2h_ms"], c["parity"]))
P
S=$(date +%s); python profiles/tools/bench_cli.py --pairs 16000000 --skip-zlib --no-reference > $O/${T}_cli16m.json 2> $O/${T}_cli16m.log; echo "cli16m rc=$? $(( $(date +%s)-S ))s"; cat $O/${T}_cli16m.json
for t in 4 8 12 16; do HAST_PARSE_ONLY=1 true; done
S=$(date +%s); python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s $(tail -1 $O/${T}_pytest.log)"
