#!/bin/bash
# GPU parity tests, then the full default bench line (C4 + A/B legs + C3 + C5 + config 2).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-it}
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.log
timeout 1800 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
echo done > gpurun_out/${TAG}_done
