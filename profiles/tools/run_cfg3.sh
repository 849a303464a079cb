O=gpurun_out; T=r01_z; mkdir -p $O
S=$(date +%s); python -m pytest tests/test_full_size_gpu.py -x -q -k barcode_scale > $O/${T}_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s $(tail -1 $O/${T}_pytest.log)"
S=$(date +%s); python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.log; echo "bench cfg3 rc=$? $(( $(date +%s)-S ))s"; tail -3 $O/${T}_bench_cfg3.log; cat $O/${T}_bench_cfg3.json | head -c 2500
