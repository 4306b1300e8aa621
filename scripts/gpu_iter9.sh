#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-it}
for den in 16 32 64; do
  RG_OR_COL_DEN=$den timeout 900 python bench.py --steps 3 --warmup 2 --no-decode --no-extra --cpu-sample 64 --cpu-seconds 4 > gpurun_out/${TAG}_c4_den${den}.json 2> gpurun_out/${TAG}_c4_den${den}.err
done
# per-rank load of an 8-GPU run on one GPU: one 12.5M-doc leaf, the whole batch
for rp in 32768 131072; do
  timeout 600 python bench.py --docs 12500000 --steps 3 --warmup 2 --range-postings $rp --no-decode --no-extra --cpu-sample 64 --cpu-seconds 3 > gpurun_out/${TAG}_leaf12m_rp$rp.json 2> gpurun_out/${TAG}_leaf12m_rp$rp.err
done
RG_PLAN_TIMING=1 timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-decode --no-extra --cpu-sample 64 --cpu-seconds 3 > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
echo done > gpurun_out/${TAG}_done
