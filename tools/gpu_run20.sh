#!/bin/bash
# is the sporadic re-flatten outlier of the bench's e2e window gone with the collector parked?  (laps on stderr)
O=gpurun_out/r2w
mkdir -p $O
SECONDS=0
for i in 1 2 3 4 5; do
  KAMR_VERBOSE=1 timeout 120 python bench.py --workload S2ib --steps 10 --warmup 3 --no-cpu --no-parity --no-workloads > $O/b$i.json 2> $O/b$i.err
  python - $O/b$i.json <<'P'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("S2ib run: in-window re-flatten %.1f ms, first call %.1f ms, ms/step %.4f" % (j["e2e"]["upload_topology_ms"], j["reflatten"]["first_call_ms"], j["ms_per_step"]))
P
done
echo "total ${SECONDS}s"
