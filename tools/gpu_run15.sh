#!/bin/bash
# adaptation criteria on the device: parity tests only
O=gpurun_out/r2r
mkdir -p $O
SECONDS=0
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -k "ps_criterion or vs_criterion" > $O/pytest.log 2>&1
echo "pytest rc=$? in ${SECONDS}s" | tee -a $O/pytest.log
grep -E "passed|failed|FAILED|ERROR|Error" $O/pytest.log | tail -40
