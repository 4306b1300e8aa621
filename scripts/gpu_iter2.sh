#!/bin/bash
# One gpurun call: GPU parity tests, then the C4 bench without the extra legs (parity on the CPU sample included).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-it}
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-decode --no-extra --stats --cpu-sample 128 --cpu-seconds 8 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err
echo done > gpurun_out/${TAG}_done
