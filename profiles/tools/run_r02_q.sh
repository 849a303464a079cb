#!/bin/bash
# CLI tool only: 16 M and 4 M pairs, plain + gzip (whole-process wall and phase times)
O=gpurun_out; T=${1:-r02_q}; mkdir -p $O
python -m pytest tests/test_cli_gpu.py tests/test_stage_script_gpu.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_pytest.log)"
python profiles/tools/bench_cli.py --pairs 16000000 --skip-zlib --no-reference > $O/${T}_cli16m.json 2> $O/${T}_cli16m.log; echo "cli16m rc=$?"; cat $O/${T}_cli16m.json
python profiles/tools/bench_cli.py --pairs 4000000 --skip-zlib > $O/${T}_cli4m.json 2> $O/${T}_cli4m.log; echo "cli4m rc=$?"; cat $O/${T}_cli4m.json
