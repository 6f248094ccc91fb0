#!/bin/bash
# round-2 first GPU pass: hardware numbers, the whole GPU test-suite, per-workload kernel breakdown
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
./build/membw > $O/membw.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=15 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
for w in S2ib S1caidvm S3 S4 S5; do
  KAMR_VERBOSE=1 timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err
done
tail -5 $O/pytest.log
cat $O/membw.txt
