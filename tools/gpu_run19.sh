#!/bin/bash
O=gpurun_out/r2v
mkdir -p $O
SECONDS=0
KAMR_VERBOSE=1 timeout 200 python tools/reflatten_probe.py S2ib 6 > $O/probe_S2ib.log 2>&1
KAMR_VERBOSE=1 timeout 400 python tools/reflatten_probe.py S4 5 > $O/probe_S4.log 2>&1
echo "total ${SECONDS}s"
grep "\[probe\]" $O/probe_S2ib.log $O/probe_S4.log
