#!/bin/bash
# Quick iteration: GPU parity tests + event counters / timing at scaled and full C4 size.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-it}
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.log
S="--docs 10000000 --terms 100000 --batch 1024 --cpu-sample 64 --cpu-seconds 4 --no-decode --no-extra"
timeout 600 python bench.py $S --steps 3 --warmup 2 --stats > gpurun_out/${TAG}_scaled_stats.json 2> gpurun_out/${TAG}_scaled_stats.err
timeout 900 python bench.py --steps 3 --warmup 2 --stats --no-decode --no-extra --cpu-sample 256 --cpu-seconds 5 > gpurun_out/${TAG}_c4_stats.json 2> gpurun_out/${TAG}_c4_stats.err
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-decode --no-extra --cpu-sample 256 --cpu-seconds 5 > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
echo done > gpurun_out/${TAG}_done
