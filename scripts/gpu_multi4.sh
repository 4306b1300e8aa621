#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-mg4}
N=${2:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 200 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-decode --no-extra --cpu-sample 64 --cpu-seconds 3 > gpurun_out/${TAG}_c4_n$N.json 2> gpurun_out/${TAG}_c4_n$N.err
echo done > gpurun_out/${TAG}_done
