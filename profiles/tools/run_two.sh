O=gpurun_out; T=r01_ac; mkdir -p $O
python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/${T}_pytest_multi.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_pytest_multi.log)"
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/${T}_bench_n2.json 2> $O/${T}_bench_n2.log; echo "bench n=2 rc=$?"; cut -c1-330 $O/${T}_bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/${T}_ref_n2.json 2> $O/${T}_ref_n2.log; echo "ref n=2 rc=$?"; cut -c1-200 $O/${T}_ref_n2.json
