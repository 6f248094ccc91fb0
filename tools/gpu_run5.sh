#!/bin/bash
O=gpurun_out/r2g
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "not 1000 and not elementwise" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -6 $O/pytest.log
run() {  # tag lib workload
  KAMR_LIB=$2 timeout 600 python bench.py --workload $3 --steps 10 --warmup 3 --no-cpu --no-parity --no-workloads > $O/$1_$3.json 2> $O/$1_$3.err
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
for w in S4 S2ib S1caidvm S5 S3; do
  run base "" $w
  for so in kitamr.jl_b200/csrc/variants/libkamr_*.so; do
    tag=$(basename "$so" .so); tag=${tag#libkamr_}
    run $tag $PWD/$so $w
  done
done
