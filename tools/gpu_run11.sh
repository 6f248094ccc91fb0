#!/bin/bash
O=gpurun_out/r2n
mkdir -p $O
/usr/bin/time -v python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench rc=$?"
tail -3 $O/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
echo "reference rc=$?"
cat $O/bench_reference.json | cut -c1-600
