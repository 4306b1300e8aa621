#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-it}
B="--steps 3 --warmup 2 --no-decode --no-extra --cpu-sample 32 --cpu-seconds 2"
RG_MAX_RANGES=64 timeout 600 python bench.py $B > gpurun_out/${TAG}_c4_cap64.json 2> gpurun_out/${TAG}_c4_cap64.err
RG_MAX_RANGES=16 timeout 600 python bench.py --docs 12500000 $B > gpurun_out/${TAG}_leaf12m_cap16.json 2> gpurun_out/${TAG}_leaf12m_cap16.err
RG_MAX_RANGES=64 timeout 600 python bench.py --docs 12500000 $B > gpurun_out/${TAG}_leaf12m_cap64.json 2> gpurun_out/${TAG}_leaf12m_cap64.err
echo done > gpurun_out/${TAG}_done
