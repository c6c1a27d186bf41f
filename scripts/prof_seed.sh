#!/bin/bash
# quick loop: GPU parity suite, then one ncu pass over the seeding kernels of a warmed c2 step (summary metrics)
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,smsp__thread_inst_executed_per_inst_executed.ratio,lts__t_sector_hit_rate.pct --clock-control none -k regex:seed_ --launch-skip 4 -c 2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>&1 | grep -E "^  void|duration|warps_active|inst_executed|issue_active|dram|hit_rate" | cut -c1-150 | tail -20
