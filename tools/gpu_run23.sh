#!/bin/bash
# bench.py sanity after the e2e-window change (median of three windows, collector parked)
O=gpurun_out/r2z
mkdir -p $O
timeout 100 python bench.py --workload S2ib --steps 5 --warmup 3 --no-cpu --no-workloads > $O/bench_S2ib.json 2> $O/bench_S2ib.err
echo "bench rc=$?"
python - $O/bench_S2ib.json <<'P'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.4f value %.3e e2e %.3e windows %s reflatten_all %s parity %s" % (j["ms_per_step"], j["value"], j["e2e"]["value"],
      ["%.3f" % w for w in j["e2e"]["windows_s"]], ["%.0f" % w for w in j["e2e"]["upload_topology_ms_all"]], j.get("parity", {}).get("ok")))
P
tail -3 $O/bench_S2ib.err
