#!/bin/bash
# Final multi-GPU check (gpurun --gpus N): the whole GPU test suite (the sharded tests then use NCCL across devices),
# the C4 bench line at N with its parity verdict.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-mg}
N=${2:-2}
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-decode --no-extra --cpu-sample 128 --cpu-seconds 6 > gpurun_out/${TAG}_c4_n$N.json 2> gpurun_out/${TAG}_c4_n$N.err
echo done > gpurun_out/${TAG}_done
