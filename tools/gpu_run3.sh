#!/bin/bash
O=gpurun_out/r2f
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "not 1000 and not elementwise" > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
for w in S4 S2ib S5 S3; do
  KAMR_VERBOSE=1 timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-parity --no-workloads > $O/bench_$w.json 2> $O/bench_$w.err
  python - $O/bench_$w.json <<'P'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = j["roofline"]["kernels_ms_per_step"]
    print(j["config"]["workload"], "ms/step %.4f" % j["ms_per_step"], "frac %.3f" % j["roofline"]["frac"],
          " ".join("%s=%.3f" % (a.replace("_kernel", ""), b) for a, b in k.items()))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
