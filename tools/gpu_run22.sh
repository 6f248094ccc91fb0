#!/bin/bash
O=gpurun_out/r2y
mkdir -p $O
SECONDS=0
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29658 tools/migrate_probe.py > $O/p2p.log 2>&1
echo "one-sided rc=$?"; grep "^(" $O/p2p.log; grep -i "error\|Traceback" $O/p2p.log | head -5
echo "total ${SECONDS}s"
