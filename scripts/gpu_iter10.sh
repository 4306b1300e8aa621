#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-it}
( timeout 900 python -m pytest tests/test_gpu_maxscore.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.log
B="--steps 3 --warmup 2 --no-decode --no-extra --cpu-sample 32 --cpu-seconds 2"
timeout 600 python bench.py $B --range-postings 32768 > gpurun_out/${TAG}_c4_rp32k.json 2> gpurun_out/${TAG}_c4_rp32k.err
RG_MAX_RANGES=512 timeout 600 python bench.py $B --range-postings 32768 > gpurun_out/${TAG}_c4_rp32k_cap512.json 2> gpurun_out/${TAG}_c4_rp32k_cap512.err
RG_MAX_RANGES=128 timeout 600 python bench.py $B --range-postings 32768 > gpurun_out/${TAG}_c4_rp32k_cap128.json 2> gpurun_out/${TAG}_c4_rp32k_cap128.err
timeout 600 python bench.py $B > gpurun_out/${TAG}_c4_auto.json 2> gpurun_out/${TAG}_c4_auto.err
for rp in 65536 262144; do
  timeout 600 python bench.py --docs 12500000 $B --range-postings $rp > gpurun_out/${TAG}_leaf12m_rp$rp.json 2> gpurun_out/${TAG}_leaf12m_rp$rp.err
done
echo done > gpurun_out/${TAG}_done
