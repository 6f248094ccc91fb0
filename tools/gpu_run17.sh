#!/bin/bash
O=gpurun_out/r2t
mkdir -p $O
SECONDS=0
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 \
    tests/multi_gpu_check.py > $O/multi2.log 2>&1
echo "multi rc=$? total ${SECONDS}s"
grep -E "^[a-z0-9_]+: world|Error|error" $O/multi2.log | tail -14
