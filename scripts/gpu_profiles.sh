#!/bin/bash
# Evidence for profiles/: the default bench line, its launch list, DRAM traffic of the evaluation kernels at full
# size, full-set captures of k_eval_and (C3) and k_eval_or_ms (scaled C4) and k_decode_segment.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2p}
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-extra --cpu-sample 16 --cpu-seconds 2 > gpurun_out/${TAG}_launches.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none -k regex:'k_eval|k_heap_replay' -s 3 -c 3 --csv --log-file gpurun_out/${TAG}_traffic_c4.csv \
  python bench.py --steps 1 --warmup 1 --no-extra --no-decode --cpu-sample 16 --cpu-seconds 2 > gpurun_out/${TAG}_traffic_c4.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:'k_eval|k_heap_replay' -s 2 -c 2 --csv --log-file gpurun_out/${TAG}_traffic_c3.csv \
  python bench.py --workload c3 --steps 1 --warmup 1 --no-extra --no-decode --cpu-sample 16 --cpu-seconds 2 > gpurun_out/${TAG}_traffic_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval_and -s 1 -c 1 -o gpurun_out/${TAG}_eval_and_c3 \
  python bench.py --workload c3 --steps 1 --warmup 1 --no-extra --no-decode --cpu-sample 16 --cpu-seconds 2 > gpurun_out/${TAG}_ncu_and.log 2>&1
S="--docs 10000000 --terms 100000 --batch 1024 --cpu-sample 16 --cpu-seconds 2 --no-decode --no-extra"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval_or_ms -s 1 -c 1 -o gpurun_out/${TAG}_eval_or_ms_scaled \
  python bench.py $S --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_ms.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_decode_segment -s 1 -c 1 -o gpurun_out/${TAG}_decode_segment \
  python bench.py --docs 20000000 --terms 200000 --batch 64 --steps 1 --warmup 1 --no-extra --cpu-sample 8 --cpu-seconds 1 > gpurun_out/${TAG}_ncu_dec.log 2>&1
echo done > gpurun_out/${TAG}_done
