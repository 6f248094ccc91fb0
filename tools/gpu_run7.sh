#!/bin/bash
O=gpurun_out/r2i
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=8 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
grep -E "passed|failed|FAILED|ERROR|rc=" $O/pytest.log | tail -12
grep -E "world=2" $O/pytest.log | head -12
