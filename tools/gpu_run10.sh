#!/bin/bash
N=$1; O=gpurun_out/r2m_$N; shift
mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus $N --steps 20 --warmup 3 "$@" > $O/bench.json 2> $O/bench.err
echo "bench rc=$?"
python - $O/bench.json <<'P'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N", j["n_gpus"], "value %.4e" % j["value"], "ms/step %.3f" % j["ms_per_step"], "frac %.3f" % j["roofline"]["frac"], "reflatten ms", j["reflatten"]["upload_topology_ms"], j["reflatten"]["first_call_ms"], "\nparity", j.get("parity"), "\nby rank", j["config"]["kernels_ms_per_step_by_rank"], "\ne2e", j["e2e"])
print({k: round(v, 3) for k, v in j["roofline"]["kernels_ms_per_step"].items()})
P
grep -E "Error|Traceback" $O/bench.err | head -5
