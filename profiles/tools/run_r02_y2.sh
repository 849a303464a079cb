#!/bin/bash
# N=2: e2e legs (ascii / host_packed) under torchrun
O=gpurun_out; T=${1:-r02_y2}; mkdir -p $O; nproc
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cfg3 > $O/${T}_bench_n2.json 2> $O/${T}_bench_n2.log; echo "bench n=2 rc=$?"
python - <<P
import json
d=json.load(open("$O/${T}_bench_n2.json")); e=d["e2e"]
print("n=2 value %.3f G | e2e chosen=%s %.1f M pairs/s | ascii %.1f M (%.1f ms) | host_packed %.1f M (%.1f ms, %d threads/rank) | pre-packed %.1f M | rule: %s" % (d["value"]/1e9, e["chosen"], e["value"]/1e6, e["ascii"]["value"]/1e6, e["ascii"]["ms_per_step"], e["host_packed"]["value"]/1e6, e["host_packed"]["ms_per_step"], e["host_packed"]["host_pack_threads"], e["packed"]["value"]/1e6, e["rule"]))
P
