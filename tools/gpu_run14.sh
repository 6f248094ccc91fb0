#!/bin/bash
# final validation of the round: whole GPU suite, smoke(), quick benches, full-set ncu capture of the top kernels
O=gpurun_out/r2q
mkdir -p $O
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=6 > $O/pytest.log 2>&1
echo "pytest rc=$? in ${SECONDS}s" | tee -a $O/pytest.log
grep -E "passed|failed|FAILED|ERROR" $O/pytest.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
bash tools/gpu_bench_all.sh $O/bench S4 S5 S2ib
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'phase_kernel|phase_regular' -s 12 -c 4 -o $O/full_S4 \
    python bench.py --workload S4 --steps 1 --warmup 3 --no-cpu --no-parity --no-workloads > $O/ncu_full_S4.log 2>&1
echo "total ${SECONDS}s"
