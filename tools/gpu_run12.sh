#!/bin/bash
O=gpurun_out/r2o
mkdir -p $O
SECONDS=0
KAMR_VERBOSE=1 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench rc=$? in ${SECONDS}s"
grep "re-flatten" $O/bench_default.err | head -24
# variants on the 3-D workloads
bash tools/gpu_bench_all.sh $O/sweep S4 S5
# DRAM traffic of one full S4 step (all launches of the step)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:'phase|slope|solid|limit' -s 75 -c 25 --csv --log-file $O/metrics_S4.csv \
    python bench.py --workload S4 --steps 1 --warmup 3 --no-cpu --no-parity --no-workloads > $O/ncu_S4.log 2>&1
echo "total ${SECONDS}s"
