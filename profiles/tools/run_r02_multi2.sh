#!/bin/bash
# round 2, second 8-GPU call (final kernel): weak + strong scaling bench at N = 8, 4, 2 under torchrun (20 steps like the driver)
O=gpurun_out; T=${1:-r02_n}; NG=${2:-8}; mkdir -p $O
for n in $NG 4 2; do
  [ $n -gt $NG ] && continue
  S=$(date +%s)
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
      bench.py --gpus $n --steps 20 --warmup 5 > $O/${T}_bench_n$n.json 2> $O/${T}_bench_n$n.log
  echo "bench n=$n rc=$? $(( $(date +%s)-S ))s"
  python - <<P
import json
try:
    d=json.load(open("$O/${T}_bench_n$n.json"))
    c=d["cfg3"]
    print("n=$n cfg2 weak: value %.3f G pairs/s (%.2f ms/step), e2e %.1f M (%.1f GB/s H2D per GPU, %.2f of ceiling), packed %.1f M, parity %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["h2d_gbs_per_gpu"], d["e2e"].get("frac_of_h2d_ceiling") or 0, d["e2e"]["packed"]["value"]/1e6, d["parity"]["ok"]))
    print("n=$n cfg3 strong: value %.3f G pairs/s, ms/step %.1f, kernel %.1f, reduce %.2f ms (%.1f%%), d2h %.2f ms, parity %s, pairs by rank %s" % (c["value"]/1e9, c["ms_per_step"], c["kernel_ms"], c["reduce_ms"], 100*c["reduce_share_of_step"], c["d2h_ms"], c["parity"], c["parity_detail"]["oracle_pairs_by_rank"]))
except Exception as e:
    print("n=$n: no line:", e)
P
done
