#!/bin/bash
# per-kernel DRAM traffic / L2 hit rate / occupancy on the 3-D workloads, plus a full-set capture of the dominant kernels
O=gpurun_out/r2b
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
for w in S4 S5 S2ib; do
  timeout 900 ncu --metrics $M --clock-control none -k regex:'phase|slope|solid' -s 24 -c 8 --csv --log-file $O/metrics_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 3 --no-cpu > $O/ncu_$w.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'phase_regular|phase_kernel' -s 9 -c 3 -o $O/full_S4 \
    python bench.py --workload S4 --steps 1 --warmup 3 --no-cpu > $O/ncu_full_S4.log 2>&1
ls -la $O
