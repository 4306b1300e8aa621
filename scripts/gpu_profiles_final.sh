#!/bin/bash
# Evidence for profiles/ (final build): launch list of the default bench step, DRAM traffic + instruction counts of the
# evaluation kernels at full size (C4, C3), a full-set capture of k_eval_or on the full C4 batch.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2f}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-extra --cpu-sample 16 --cpu-seconds 2 > gpurun_out/${TAG}_launches.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none -k regex:'k_eval|k_heap_replay' -s 3 -c 3 --csv --log-file gpurun_out/${TAG}_traffic_c4.csv \
  python bench.py --steps 1 --warmup 1 --no-extra --no-decode --cpu-sample 16 --cpu-seconds 2 > gpurun_out/${TAG}_traffic_c4.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:'k_eval|k_heap_replay' -s 2 -c 2 --csv --log-file gpurun_out/${TAG}_traffic_c3.csv \
  python bench.py --workload c3 --steps 1 --warmup 1 --no-extra --no-decode --cpu-sample 16 --cpu-seconds 2 > gpurun_out/${TAG}_traffic_c3.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_eval_or$' -s 1 -c 1 -o gpurun_out/${TAG}_eval_or_full \
  python bench.py --steps 1 --warmup 1 --no-decode --no-extra --cpu-sample 16 --cpu-seconds 2 > gpurun_out/${TAG}_ncu_or_full.log 2>&1
echo done > gpurun_out/${TAG}_done
