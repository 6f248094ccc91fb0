#!/bin/bash
# N-GPU validation: parity of the small cases at world N, then the S4 bench at N with the parity object
N=$1; O=gpurun_out/r2k_$N; shift
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tests/multi_gpu_check.py > $O/multi.log 2>&1
echo "multi_gpu_check rc=$?" >> $O/multi.log
grep -E "world=|rc=" $O/multi.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus $N --steps 20 --warmup 3 "$@" > $O/bench.json 2> $O/bench.err
echo "bench rc=$?"
python - $O/bench.json <<'P'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N", j["n_gpus"], "value %.4e" % j["value"], "ms/step %.3f" % j["ms_per_step"], "frac %.3f" % j["roofline"]["frac"], "reflatten ms", j["reflatten"], "\nparity", j.get("parity"), "\nby rank", j["config"]["kernels_ms_per_step_by_rank"])
print({k: round(v, 3) for k, v in j["roofline"]["kernels_ms_per_step"].items()})
P
grep -E "Error|Traceback" $O/bench.err | head -5
