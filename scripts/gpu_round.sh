#!/bin/bash
# One gpurun call: GPU parity tests, scaled A/B benches, the full default bench.  Everything lands in gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2a}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/${TAG}_pytest.log
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
for extra in "" "--maxscore"; do
  name=${TAG}_scaled${extra// /_}
  timeout 600 python bench.py --docs 10000000 --terms 100000 --batch 1024 --steps 3 --warmup 2 --cpu-sample 128 --cpu-seconds 6 --no-decode --no-extra $extra > gpurun_out/${name}.json 2> gpurun_out/${name}.err
done
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
echo done > gpurun_out/${TAG}_done
