#!/bin/bash
# round 2, final 1-GPU pass of a build: full GPU suite + smoke, the whole bench line, DRAM-traffic captures of the final kernel,
# launch list, CLI tool (plain / gzip / partition)
O=gpurun_out; T=${1:-r02_e}; mkdir -p $O
S=$(date +%s); python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s $(tail -1 $O/${T}_pytest.log)"; grep -E "^FAILED|^ERROR" $O/${T}_pytest.log | head
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
S=$(date +%s); bash profiles/tools/capture_traffic.sh ${T} > $O/${T}_traffic.log 2>&1; echo "traffic rc=$? $(( $(date +%s)-S ))s"; tail -22 $O/${T}_traffic.log; cp $O/${T}_traffic.json profiles/traffic.json
S=$(date +%s); timeout 1500 python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.log; echo "bench rc=$? $(( $(date +%s)-S ))s"; tail -6 $O/${T}_bench.log
python - <<P
import json
d=json.load(open("$O/${T}_bench.json"))
r=d["roofline"]
print("value %.3f G pairs/s (%.2f ms/step) e2e %.1f M packed %.1f M parity %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["packed"]["value"]/1e6, d["parity"]["ok"]))
print("roofline cfg2: %.4f ms/launch %.1f G lookups/s frac_alg %.3f frac_dram %s | hbm_resident: %.4f ms/launch %.1f G lookups/s frac_alg %.3f frac_dram %s" % (r["ms_per_launch"], r["lookups_per_s"]/1e9, r["frac_algorithmic"], r["frac_dram"], r["hbm_resident"]["ms_per_launch"], r["hbm_resident"]["lookups_per_s"]/1e9, r["hbm_resident"]["frac_algorithmic"], r["hbm_resident"]["frac_dram"]))
print("cli", json.dumps(d.get("cli"))[:1500])
c=d["cfg3"]; print("cfg3 value %.3f G ms %.1f kernel %.1f reduce %.2f d2h %.2f parity %s" % (c["value"]/1e9, c["ms_per_step"], c["kernel_ms"], c["reduce_ms"], c["d2h_ms"], c["parity"]))
P
S=$(date +%s); ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k "regex:classify_kernel|tile_kernel|table_|gather_kernel|lookup_kernel|sg_" -c 3000 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --cfg3-pairs 40000000 --cfg3-steps 1 > $O/${T}_launches_bench.log 2>&1; echo "launch list rc=$? $(( $(date +%s)-S ))s"; python profiles/tools/ncu_launches.py $O/${T}_launches.csv 2>/dev/null | head -12
S=$(date +%s); python profiles/tools/bench_cli.py --pairs 16000000 --skip-zlib --no-reference > $O/${T}_cli16m.json 2> $O/${T}_cli16m.log; echo "cli16m rc=$? $(( $(date +%s)-S ))s"; cat $O/${T}_cli16m.json
