#!/bin/bash
# 2-GPU validation of the one-sided halo: parity of 7 cases vs the single-rank oracle, then the S4 bench at N=2 with parity
O=gpurun_out/r2h
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tests/multi_gpu_check.py > $O/multi2.log 2>&1
echo "multi_gpu_check rc=$?" >> $O/multi2.log
tail -12 $O/multi2.log
KAMR_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench2.json 2> $O/bench2.err
echo "bench rc=$?"
tail -c 3000 $O/bench2.json
grep -E "Error|error|Traceback" $O/bench2.err | head
