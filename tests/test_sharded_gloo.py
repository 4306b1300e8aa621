"""world_size-2 gloo test (CPU) of the one-segment-per-rank plumbing: statistics broadcast,
leaf-record layout, all-gather order and the leaf-ordered finish_parallel merge — checked against
the oracle's search_parallel semantics over the same two leaves."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_binding as ob
from rucene_b200 import codec, sharded

K = 10
N_TERMS = 400


def _segments():
    return [codec.synth_segment(0x5EED0005 + r, 30000, N_TERMS, doc_version=1, n_threads=1) for r in range(2)]


def _specs():
    rng = np.random.default_rng(5)
    specs = []
    for i in range(24):
        t = [int(x) for x in rng.choice(60, size=int(rng.integers(2, 5)), replace=False)]
        occ = ob.MUST if i % 2 else ob.SHOULD
        specs.append(("bool", [(occ, x) for x in t], 0))
    specs.append(("term", 3))
    return specs


def _worker(rank, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=2)
    try:
        segs = _segments()
        local = segs[rank]
        # every rank must score with the statistics of leaf 0 (largest / first among equals)
        df, doc_count, sum_ttf = sharded.broadcast_stats(local.terms["doc_freq"], local.doc_count,
                                                         local.sum_total_term_freq, src=0)
        assert np.array_equal(df, segs[0].terms["doc_freq"])
        assert (doc_count, sum_ttf) == (segs[0].doc_count, segs[0].sum_total_term_freq)
        # the local leaf's records, produced here by the oracle (no GPU in this test)
        ix = ob.Index()
        for s in segs:
            ix.add_segment(s)
        q, c = ob.make_queries(_specs())
        rec = ix.leaf_records(rank, q, c, K)
        assert rec.shape[1] == sharded.record_bytes(K)
        allrec = sharded.gather_leaf_records(torch.from_numpy(rec.reshape(-1).copy()))
        allrec = allrec.numpy().reshape(2, len(q), sharded.record_bytes(K))
        assert np.array_equal(allrec[rank], rec)
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), allrec)
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_and_leaf_order_merge(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)  # every rank holds the same [leaf][query] record array
    segs = _segments()
    ix = ob.Index()
    for s in segs:
        ix.add_segment(s)
    q, c = ob.make_queries(_specs())
    want_hits, want_counts, want_total = ix.search_batch(q, c, K, parallel_mode=1)
    for qi in range(len(q)):
        leaves, total = [], 0
        for leaf in range(2):
            rec = a[leaf, qi]
            n = int(rec[:4].view(np.uint32)[0])
            total += int(rec[8:16].view(np.uint64)[0])
            leaves.append(rec[16:16 + 8 * n].view(ob.HIT_DTYPE))
        got = ob.topk_merge(leaves, K)
        assert total == want_total[qi]
        n = int(want_counts[qi])
        assert np.array_equal(got["doc"], want_hits[qi][:n]["doc"])
        assert np.array_equal(got["score"].view(np.uint32), want_hits[qi][:n]["score"].view(np.uint32))
    # leaf 1's docids are global (doc_base = 30000)
    assert any((a[1, qi][16:24].view(np.int32)[0] >= 30000) for qi in range(len(q)) if a[1, qi][:4].view(np.uint32)[0])
