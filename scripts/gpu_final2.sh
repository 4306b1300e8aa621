#!/bin/bash
# Closing evidence of the round, final build: GPU parity tests, smoke(), the default bench line.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2g}
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.log
( timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5 ) > gpurun_out/${TAG}_smoke.log
timeout 1800 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
echo done > gpurun_out/${TAG}_final_done
