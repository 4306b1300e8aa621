#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-sanity}
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-decode --no-extra --cpu-sample 32 --cpu-seconds 2 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err
echo done > gpurun_out/${TAG}_done
