#!/bin/bash
# N-GPU pass on one box: the >= 2-GPU parity tests and the weak-scaling bench under torchrun.
# usage: profiles/tools/run_multi.sh TAG N
TAG=${1:-rXX}; N=${2:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/${TAG}_gpus.txt 2>&1
nvidia-smi topo -m >> $O/${TAG}_gpus.txt 2>&1
python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/${TAG}_pytest_multi.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_multi.log
tail -3 $O/${TAG}_pytest_multi.log
for n in $(seq 1 $N); do
  case $n in 1|2|4|8) ;; *) continue;; esac
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.log
  else
    NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 5 --warmup 3 > $O/${TAG}_bench_n$n.json 2> $O/${TAG}_bench_n$n.log
  fi
  echo "bench n=$n rc=$?"; cat $O/${TAG}_bench_n$n.json | cut -c1-400
done
