#!/bin/bash
N=$1; O=gpurun_out/r2l_$N
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "not 1000 and not elementwise" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
bash tools/gpu_run8.sh $N
cp gpurun_out/r2k_$N/* $O/
