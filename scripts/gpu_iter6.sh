#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-it}
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.log
RG_PLAN_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-decode --no-extra --cpu-sample 128 --cpu-seconds 6 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err
echo done > gpurun_out/${TAG}_done
