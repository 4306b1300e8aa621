#!/bin/bash
# Diagnostics of k_eval_or_ms: event counters at scaled and full size, ncu --set full at scaled size.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2b}
S="--docs 10000000 --terms 100000 --batch 1024 --cpu-sample 64 --cpu-seconds 4 --no-decode --no-extra"
timeout 600 python bench.py $S --steps 3 --warmup 2 --stats > gpurun_out/${TAG}_scaled_stats.json 2> gpurun_out/${TAG}_scaled_stats.err
timeout 600 python bench.py $S --steps 3 --warmup 2 --no-maxscore > gpurun_out/${TAG}_scaled_noms.json 2> gpurun_out/${TAG}_scaled_noms.err
timeout 900 python bench.py --steps 2 --warmup 1 --stats --no-decode --no-extra --cpu-sample 64 --cpu-seconds 5 > gpurun_out/${TAG}_c4_stats.json 2> gpurun_out/${TAG}_c4_stats.err
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_eval_or_ms -s 1 -c 1 -o gpurun_out/${TAG}_ms_scaled \
  python bench.py $S --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu.log 2>&1
echo done > gpurun_out/${TAG}_done
