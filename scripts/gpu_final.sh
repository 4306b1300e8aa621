#!/bin/bash
# The round's closing evidence in one gpurun call: GPU parity tests, smoke(), the default bench line (C4 + A/B legs + C3 +
# C5 + config 2), then the profile captures of scripts/gpu_profiles_final.sh.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2f}
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.log
( timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5 ) > gpurun_out/${TAG}_smoke.log
timeout 1800 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
bash scripts/gpu_profiles_final.sh ${TAG}
echo done > gpurun_out/${TAG}_final_done
