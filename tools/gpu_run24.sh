#!/bin/bash
O=gpurun_out/r2z2
mkdir -p $O
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -k "migrate" > $O/pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" $O/pytest.log | tail -5
timeout 80 python bench.py --workload S2ib --steps 5 --warmup 3 --no-cpu --no-workloads > $O/bench_S2ib.json 2> $O/bench_S2ib.err
echo "bench rc=$?"
python - $O/bench_S2ib.json <<'P'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.4f e2e %.3e windows %s reflatten_all %s parity %s" % (j["ms_per_step"], j["e2e"]["value"],
      ["%.3f" % w for w in j["e2e"]["windows_s"]], ["%.0f" % w for w in j["e2e"]["upload_topology_ms_all"]], j.get("parity")))
P
