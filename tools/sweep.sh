#!/bin/bash
# usage: tools/sweep.sh <outdir> [bench args...]  -- runs bench.py once per library variant under csrc/variants/
out=$1; shift
mkdir -p "$out"
python bench.py --no-cpu --steps 10 --warmup 3 --no-parity --no-workloads "$@" > "$out/base.json" 2> "$out/base.err"
for so in kitamr.jl_b200/csrc/variants/libkamr_*.so; do
  tag=$(basename "$so" .so); tag=${tag#libkamr_}
  KAMR_LIB=$PWD/$so python bench.py --no-cpu --steps 10 --warmup 3 --no-parity --no-workloads "$@" > "$out/$tag.json" 2> "$out/$tag.err"
done
python - "$out" <<'P'
import json, sys, glob, os
for f in sorted(glob.glob(sys.argv[1] + "/*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        k = j["roofline"]["kernels_ms_per_step"]
        print(os.path.basename(f)[:-5].ljust(10), "ms/step %.4f" % j["ms_per_step"], "frac %.3f" % j["roofline"]["frac"],
              " ".join("%s=%.3f" % (a.replace("_kernel", ""), b) for a, b in k.items()))
    except Exception as e:
        print(f, "FAILED", e)
P
