#!/bin/bash
# usage: tools/gpu_bench_all.sh <outdir> [workloads...] — base library and every variant under csrc/variants/
O=$1; shift
W=${@:-S4 S2ib S1caidvm S5 S3}
mkdir -p $O
run() {  # tag lib workload
  if [ -n "$2" ]; then export KAMR_LIB=$2; else unset KAMR_LIB; fi
  timeout 600 python bench.py --workload $3 --steps 10 --warmup 3 --no-cpu --no-parity --no-workloads > $O/$1_$3.json 2> $O/$1_$3.err
  python - $O/$1_$3.json $1 <<'P'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = j["roofline"]["kernels_ms_per_step"]
    print(sys.argv[2].ljust(8), j["config"]["workload"][:14].ljust(14), "ms/step %.4f" % j["ms_per_step"], "frac %.3f" % j["roofline"]["frac"],
          " ".join("%s=%.3f" % (a.replace("_kernel", "").replace("phase","ph").replace("slope","sl").replace("regular","reg"), b) for a, b in k.items()))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
}
for w in $W; do
  run base "" $w
  for so in kitamr.jl_b200/csrc/variants/libkamr_*.so; do
    [ -e "$so" ] || continue
    tag=$(basename "$so" .so); tag=${tag#libkamr_}
    run $tag $PWD/$so $w
  done
done
