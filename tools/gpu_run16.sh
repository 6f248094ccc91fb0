#!/bin/bash
# adaptation criteria: the relaxed assertions again + the 2-rank check (ghost w / flag exchange)
O=gpurun_out/r2s
mkdir -p $O
SECONDS=0
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -k "(ps_criterion or vs_criterion) and (s2_ib_fine or s4_ib_l3)" > $O/pytest.log 2>&1
echo "pytest rc=$? in ${SECONDS}s" | tee -a $O/pytest.log
grep -E "passed|failed|FAILED|ERROR" $O/pytest.log | tail -10
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 \
    tests/multi_gpu_check.py > $O/multi2.log 2>&1
echo "multi rc=$? total ${SECONDS}s"
grep -v "^W\|^\[W\|warn" $O/multi2.log | tail -14
