#!/bin/bash
# table loaders with overlapped copy / insert: parity + CLI tests, table build time at 62 M keys
O=gpurun_out; T=${1:-r02_w}; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_cli_gpu.py tests/test_stage00.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"
python bench.py --only-cfg3 --cfg3-pairs 16000000 > $O/${T}_cfg3.json 2> $O/${T}_cfg3.log; python -c "
import json; c=json.load(open('$O/${T}_cfg3.json')); print('cfg3 table build (62.4 M packed keys): %.3f s; parity %s; %.1f G lookups/s' % (c['generate_s']['table'], c['parity'], c['roofline']['lookups_per_s']/1e9))"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
