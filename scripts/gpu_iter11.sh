#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-it}
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.log
B="--steps 3 --warmup 2 --no-decode --no-extra --cpu-sample 32 --cpu-seconds 2"
timeout 600 python bench.py $B > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err
timeout 600 python bench.py --docs 12500000 $B > gpurun_out/${TAG}_leaf12m.json 2> gpurun_out/${TAG}_leaf12m.err
timeout 600 python bench.py --workload c5 $B > gpurun_out/${TAG}_c5.json 2> gpurun_out/${TAG}_c5.err
echo done > gpurun_out/${TAG}_done
