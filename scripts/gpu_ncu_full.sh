#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-ncuf}
timeout 1500 ncu --section SpeedOfLight --section SchedulerStats --section WarpStateStats --section SourceCounters --section Occupancy --section MemoryWorkloadAnalysis --section LaunchStats \
  --clock-control none --import-source on -k regex:k_eval_or_ms -s 1 -c 1 -o gpurun_out/${TAG}_ms_full \
  python bench.py --steps 1 --warmup 1 --no-decode --no-extra --cpu-sample 16 --cpu-seconds 2 > gpurun_out/${TAG}_ncu.log 2>&1
echo done > gpurun_out/${TAG}_done
