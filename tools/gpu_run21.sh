#!/bin/bash
# device-to-device migration: the single-rank reorder test, then the 2-rank check (sensor + migration phases)
O=gpurun_out/r2x
mkdir -p $O
SECONDS=0
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -k "migrate" > $O/pytest.log 2>&1
echo "pytest rc=$? in ${SECONDS}s" | tee -a $O/pytest.log
grep -E "passed|failed|FAILED|ERROR|Error:" $O/pytest.log | tail -10
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29657 \
    tests/multi_gpu_check.py > $O/multi2.log 2>&1
echo "multi rc=$? total ${SECONDS}s"
grep -E "^[a-z0-9_]+: world|Error|error:" $O/multi2.log | cut -c1-14,330-700 | tail -14
