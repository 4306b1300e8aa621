"""torchrun worker: one segment per GPU, NCCL all-gather of leaf records, device merge; rank 0
compares with the oracle's search_parallel (leaf-order) result.  Used by test_gpu_sharded.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import helpers  # noqa: E402
import oracle_binding as ob  # noqa: E402
from rucene_b200 import codec, engine, sharded  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    # SHARDED_SAME_DEVICE=1: every rank uses cuda:0 (NCCL refuses two ranks on one device, so the records
    # travel over gloo through host memory) — the N>1 path on a one-GPU box
    same_device = os.environ.get("SHARDED_SAME_DEVICE") == "1"
    if same_device:
        local_rank = 0
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if same_device:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    max_doc, n_terms, k = 120000, 3000, 100
    segs = [codec.synth_segment(0x5EED0005 + r, max_doc, n_terms, doc_version=1) for r in range(world)]
    local = segs[rank]
    eng = engine.Engine(device=local_rank, range_postings=20000)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)   # ShardedSearcher points the engine at torch's current stream itself
    eng.upload_segment(local, doc_base=rank * max_doc)
    df0, doc_count, sum_ttf = sharded.broadcast_stats(local.terms["doc_freq"], local.doc_count,
                                                      local.sum_total_term_freq, src=0,
                                                      device=None if same_device else dev)
    avgdl = codec.bm25_avg_field_length(sum_ttf, doc_count, max_doc * world)
    eng.set_norm_cache(0, codec.bm25_norm_cache(1.2, 0.75, avgdl))
    rng = np.random.default_rng(0x5EED0005)
    specs = []
    for i, ts in enumerate(helpers.distinct_query_terms(rng, n_terms, 96, 2, 5)):
        occ = ob.MUST if i % 2 == 0 and len(ts) <= 3 else ob.SHOULD   # mixed AND / OR batch (config 5)
        specs.append(("bool", [(occ, t) for t in ts], 0))
    oq, oc = ob.make_queries(specs)
    q = np.zeros(len(oq), engine.QUERY_DTYPE)
    q["clause_begin"], q["n_clauses"], q["flags"] = oq["clause_begin"], oq["n_clauses"], engine.Q_BOOLEAN
    c = np.zeros(len(oc), engine.CLAUSE_DTYPE)
    c["occur"], c["term_id"] = oc["occur"], oc["term_id"]
    c["weight"] = [np.float32(codec.bm25_idf(int(df0[t]), doc_count)) for t in oc["term_id"]]
    s = sharded.ShardedSearcher(eng)
    got = s.search_batch(q, c, k)
    torch.cuda.synchronize()
    got_c = None
    if not same_device:
        # the same step through rg_batch_run_sharded: a raw ncclComm_t (created here with ctypes, as a Rust / C++
        # host would with its own binding) — run, ncclAllGather and the leaf-order merge happen inside the C library
        import ctypes

        class UniqueId(ctypes.Structure):
            _fields_ = [("internal", ctypes.c_byte * 128)]
        nccl = ctypes.CDLL("libnccl.so.2")
        uid = UniqueId()
        if rank == 0:
            assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
        t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).to(dev)
        dist.broadcast(t, 0)
        ctypes.memmove(ctypes.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        comm = ctypes.c_void_p()
        nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
        assert nccl.ncclCommInitRank(ctypes.byref(comm), world, uid, rank) == 0
        b = eng.prepare(q, c, k, mode=engine.MODE_SEARCH_PARALLEL)
        got_c = b.run_sharded(comm, world)
        b.close()
        nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
        nccl.ncclCommDestroy(comm)
    ok = True
    if rank == 0:
        ix = ob.Index()
        for sg in segs:
            ix.add_segment(sg)
        want = ix.search_batch(oq, oc, k, parallel_mode=1, n_threads=8)
        try:
            helpers.assert_same_topdocs(got, want, "sharded world=%d" % world)
            if got_c is not None:
                helpers.assert_same_topdocs(got_c, want, "rg_batch_run_sharded world=%d" % world)
            print("SHARDED_OK world=%d queries=%d" % (world, len(q)))
        except AssertionError as e:
            ok = False
            print("SHARDED_MISMATCH", e)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
