#!/bin/bash
# kamr_project_cells + regression of everything that runs the shared Newton projection (CIP cases) + euler3d
O=gpurun_out/r2u
mkdir -p $O
SECONDS=0
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -k "project_cells or cip or s1_small or euler3d" > $O/pytest.log 2>&1
echo "pytest rc=$? in ${SECONDS}s" | tee -a $O/pytest.log
grep -E "passed|failed|FAILED|ERROR" $O/pytest.log | tail -30
