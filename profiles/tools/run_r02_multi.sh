#!/bin/bash
# round 2, the 8-GPU box in ONE call: H2D ceiling probe, strong/weak scaling bench at N=8,4,2 under torchrun,
# configs[4] through bin/classify --gpus 1 / 8, the multi-GPU parity tests.    usage: run_r02_multi.sh TAG [NGPU]
O=gpurun_out; T=${1:-r02_m}; NG=${2:-8}; mkdir -p $O
{ nproc; free -g; df -h /tmp | tail -1; lscpu | grep -E "Model name|^CPU\(s\)|NUMA"; nvidia-smi --query-gpu=index,name,memory.total --format=csv; nvidia-smi topo -m; } > $O/${T}_box.txt 2>&1
S=$(date +%s); ./bin/h2d_probe 256 1.0 > $O/${T}_h2d.json 2> $O/${T}_h2d.log; echo "h2d rc=$? $(( $(date +%s)-S ))s"; cut -c1-600 $O/${T}_h2d.json
for n in $NG 4 2; do
  [ $n -gt $NG ] && continue
  S=$(date +%s)
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 5 --warmup 3 > $O/${T}_bench_n$n.json 2> $O/${T}_bench_n$n.log
  echo "bench n=$n rc=$? $(( $(date +%s)-S ))s"
  python - <<P
import json
try:
    d=json.load(open("$O/${T}_bench_n$n.json"))
    c=d["cfg3"]
    print("n=$n cfg2 weak: value %.3f G pairs/s, e2e %.1f M (%.1f GB/s H2D per GPU), parity %s" % (d["value"]/1e9, d["e2e"]["value"]/1e6, d["e2e"]["h2d_gbs_per_gpu"], d["parity"]))
    print("n=$n cfg3 strong: value %.3f G pairs/s, ms/step %.1f, kernel %.1f, reduce %.2f ms (%.1f%%), d2h %.2f ms, parity %s, pairs by rank %s" % (c["value"]/1e9, c["ms_per_step"], c["kernel_ms"], c["reduce_ms"], 100*c["reduce_share_of_step"], c["d2h_ms"], c["parity"], c["parity_detail"]["oracle_pairs_by_rank"]))
except Exception as e:
    print("n=$n: no line:", e)
P
  tail -3 $O/${T}_bench_n$n.log
done
S=$(date +%s); timeout 1500 python profiles/tools/bench_cfg5.py --gpus-list 1,$NG --gz-level 4 > $O/${T}_cfg5.json 2> $O/${T}_cfg5.log; echo "cfg5 rc=$? $(( $(date +%s)-S ))s"; grep "cfg5" $O/${T}_cfg5.log | tail -12; cut -c1-1500 $O/${T}_cfg5.json
S=$(date +%s); python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/${T}_pytest_multi.log 2>&1; echo "pytest multi rc=$? $(( $(date +%s)-S ))s $(tail -1 $O/${T}_pytest_multi.log)"
