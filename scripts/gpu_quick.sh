#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-q}
timeout 900 python bench.py --steps 3 --warmup 2 --no-decode --no-extra --cpu-sample 64 --cpu-seconds 3 > gpurun_out/${TAG}_c4_ms.json 2> gpurun_out/${TAG}_c4_ms.err
timeout 900 python bench.py --steps 3 --warmup 2 --maxscore --no-decode --no-extra --cpu-sample 64 --cpu-seconds 3 > gpurun_out/${TAG}_c4_noms.json 2> gpurun_out/${TAG}_c4_noms.err
echo done > gpurun_out/${TAG}_done
