#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-it}
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.log
RG_PLAN_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-decode --no-extra --cpu-sample 128 --cpu-seconds 6 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err
S="--docs 10000000 --terms 100000 --batch 1024 --cpu-sample 16 --cpu-seconds 2 --no-decode --no-extra"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_eval_or$' -s 1 -c 1 -o gpurun_out/${TAG}_eval_or_scaled \
  python bench.py $S --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_or.log 2>&1
echo done > gpurun_out/${TAG}_done
