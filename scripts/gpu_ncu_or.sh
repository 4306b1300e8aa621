#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-ncuor}
S="--docs 10000000 --terms 100000 --batch 1024 --cpu-sample 16 --cpu-seconds 2 --no-decode --no-extra"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_eval_or$' -s 1 -c 1 -o gpurun_out/${TAG}_eval_or_scaled \
  python bench.py $S --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_or.log 2>&1
# per-rank load of an 8-GPU run on one GPU: one 12.5M-doc leaf, the whole batch, three range sizes
for rp in 32768 65536 131072; do
  timeout 600 python bench.py --docs 12500000 --steps 3 --warmup 2 --range-postings $rp --no-decode --no-extra --cpu-sample 64 --cpu-seconds 3 > gpurun_out/${TAG}_leaf12m_rp$rp.json 2> gpurun_out/${TAG}_leaf12m_rp$rp.err
done
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/${TAG}_pytest.log
echo done > gpurun_out/${TAG}_done
