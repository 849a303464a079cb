O=gpurun_out; T=r01_m; mkdir -p $O
S=$(date +%s)
python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s"; tail -3 $O/${T}_pytest_gpu.log
S=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke $(( $(date +%s)-S ))s"
S=$(date +%s)
python bench.py --impl reference > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.log; echo "ref rc=$? $(( $(date +%s)-S ))s"
S=$(date +%s)
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.log; echo "bench rc=$? $(( $(date +%s)-S ))s"
cat $O/${T}_bench.json | head -c 3000
nproc; free -g | head -2
