#!/bin/bash
# 2 GPUs: reduce pipelined with the read-back in hast_finish -- multi-GPU parity tests, then the cfg3 leg under torchrun
O=gpurun_out; T=${1:-r02_x}; mkdir -p $O
python -m pytest tests/test_multi_gpu.py -m gpu -q > $O/${T}_pytest_multi.log 2>&1; echo "pytest multi rc=$? $(tail -1 $O/${T}_pytest_multi.log)"
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --only-cfg3 --cfg3-steps 5 > $O/${T}_cfg3_n2.json 2> $O/${T}_cfg3_n2.log; echo "cfg3 n=2 rc=$?"
python - <<P
import json
c=json.load(open("$O/${T}_cfg3_n2.json"))
print("n=2 cfg3 strong: value %.3f G pairs/s, ms/step %.1f, kernel %.1f, reduce %.2f ms, exposed read-back %.2f ms, parity %s %s" % (c["value"]/1e9, c["ms_per_step"], c["kernel_ms"], c["reduce_ms"], c["d2h_ms"], c["parity"], c["parity_detail"]["oracle_pairs_by_rank"]))
P
