#!/bin/bash
# last pass of the round: the driver's own sequence on the final commit -- GPU suite, smoke, default bench line
O=gpurun_out; T=${1:-r02_z}; mkdir -p $O
S=$(date +%s); python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s $(tail -1 $O/${T}_pytest.log)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
S=$(date +%s); python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.log; echo "bench (default flags) rc=$? $(( $(date +%s)-S ))s"
python - <<P
import json
d=json.load(open("$O/${T}_bench.json")); r=d["roofline"]; c=d["cfg3"]
print("value %.3f G pairs/s e2e %.1f M parity %s | roofline %.1f / %.1f G lookups/s frac_dram %.2f / %.2f | cfg3 %.3f G parity %s | cli wall %s tables %s" % (d["value"]/1e9, d["e2e"]["value"]/1e6, d["parity"]["ok"], r["lookups_per_s"]/1e9, r["hbm_resident"]["lookups_per_s"]/1e9, r["frac_dram"], r["hbm_resident"]["frac_dram"], c["value"]/1e9, c["parity"], d["cli"]["speedup_wall"], d["cli"]["tables_identical"]))
P
