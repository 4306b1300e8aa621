"""Sharded mode on real GPUs — BASELINE config 5 in small: one segment per rank, mixed AND/OR batch, one all-gather of
per-segment top-k + device merge, TopDocs identical to the oracle's leaf-ordered search_parallel.  Runs on a one-GPU
box (two ranks share the device) and, where there are several GPUs, over NCCL as well."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def _run_worker(world, env_extra, port):
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(here, "sharded_gpu_worker.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "SHARDED_OK" in p.stdout, p.stdout[-3000:]


def test_sharded_search_matches_oracle():
    """The N>1 path: every rank evaluates the batch on its own leaf (rg_batch_leaf_records), one all-gather, the
    leaf-order merge (rg_merge_leaf_records), TopDocs identical to the oracle's search_parallel.  Always with two
    ranks sharing cuda:0 (gloo transport through host memory: NCCL refuses two ranks on one device) — so a one-GPU
    box exercises it too — and, where >= 2 GPUs are visible, one rank per GPU over NCCL, including the in-library
    rg_batch_run_sharded with a raw ncclComm_t."""
    import torch
    _run_worker(2, {"SHARDED_SAME_DEVICE": "1"}, 29519)
    n = torch.cuda.device_count()
    if n >= 2:
        _run_worker(2 if n < 4 else 4, {}, 29517)
