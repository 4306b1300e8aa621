#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-it}
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.log
for wl in c3 c5 c4; do
  RG_PLAN_TIMING=1 timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-decode --no-extra --cpu-sample 64 --cpu-seconds 4 > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err
done
echo done > gpurun_out/${TAG}_done
