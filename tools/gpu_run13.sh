#!/bin/bash
O=gpurun_out/r2p
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "not 1000 and not elementwise" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log; tail -4 $O/pytest.log
bash tools/gpu_bench_all.sh $O/sweep S1caidvm S3 S2ib S4
