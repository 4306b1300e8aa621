#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-ncu}
S="--docs 10000000 --terms 100000 --batch 1024 --cpu-sample 64 --cpu-seconds 4 --no-decode --no-extra"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_eval_or_ms -s 1 -c 1 -o gpurun_out/${TAG}_ms_scaled \
  python bench.py $S --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu.log 2>&1
echo done > gpurun_out/${TAG}_done
