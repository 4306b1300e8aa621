"""One index segment (or a contiguous run of segments) per GPU: search_parallel semantics
across ranks.

Every rank holds its leaves (docid ranges [doc_base, doc_base+max_doc)), evaluates the whole query
batch against them with a per-leaf TopDocs heap (search/collector/top_docs.rs:145-155), then ONE
all-gather moves the fixed-size leaf records {u32 n; u32 pad; u64 total_hits; (doc,score)[k] in
heap-array order} and every rank replays finish_parallel (top_docs.rs:157-172) in leaf order.
Weights are computed once from the statistics of the largest leaf — rank 0's when all leaves are
the same size (searcher.rs:311-351,732-767) — and broadcast, so every rank scores with the same
idf/avgdl.  No other collective touches the data path.
"""
import numpy as np
import torch
import torch.distributed as dist


def record_bytes(k):
    return 16 + 8 * k


class _CudaArray:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 2, "strides": None}


def device_bytes_tensor(ptr, nbytes, device):
    """torch uint8 view of engine-owned device memory."""
    return torch.as_tensor(_CudaArray(ptr, nbytes), device=device)


def broadcast_stats(doc_freq, doc_count, sum_total_term_freq, src=0, device=None, group=None):
    """Statistics of the leaf that supplies them (rank `src`) to every rank.
    doc_freq: int32 array (per engine-wide term id) of the local leaf."""
    df = torch.from_numpy(np.ascontiguousarray(doc_freq, dtype=np.int32))
    st = torch.tensor([int(doc_count), int(sum_total_term_freq)], dtype=torch.int64)
    if device is not None:
        df, st = df.to(device), st.to(device)
    dist.broadcast(df, src, group=group)
    dist.broadcast(st, src, group=group)
    return df.cpu().numpy(), int(st[0]), int(st[1])


def gather_leaf_records(local_records, group=None):
    """local_records: uint8 tensor [n_local_leaves * n_queries * record_bytes] (CPU/gloo or CUDA/nccl).
    Returns a uint8 tensor laid out [leaf][query][record] in rank (= leaf) order."""
    world = dist.get_world_size(group)
    out = torch.empty(world * local_records.numel(), dtype=torch.uint8, device=local_records.device)
    if local_records.is_cuda:
        dist.all_gather_into_tensor(out, local_records, group=group)
    else:
        parts = [torch.empty_like(local_records) for _ in range(world)]
        dist.all_gather(parts, local_records, group=group)
        out = torch.cat(parts)
    return out


class ShardedSearcher:
    """GPU searcher over the local leaves + one all-gather of per-leaf top-k.

    The engine launches on torch's current stream (set in every search_batch), so the evaluation
    kernels, the NCCL all-gather and the merge kernel are ordered on ONE stream without events.
    With a gloo process group (several ranks sharing one GPU, CPU-only transports) the records are
    staged through host memory instead."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.on_device = dist.get_backend(group) == "nccl"

    def search_batch(self, queries, clauses, k, k1=1.2):
        from . import engine as E
        stream = torch.cuda.current_stream(self.device)
        self.engine.set_stream(stream.cuda_stream)
        batch = self.engine.prepare(queries, clauses, k, k1=k1, mode=E.MODE_SEARCH_PARALLEL)
        try:
            batch.run()
            ptr, rb = batch.leaf_records()
            n = len(queries)
            n_local = self.engine.n_segments
            local = device_bytes_tensor(ptr, n_local * rb * n, self.device)
            if self.on_device:
                allrec = gather_leaf_records(local, self.group)
            else:
                stream.synchronize()
                allrec = gather_leaf_records(local.cpu(), self.group).to(self.device)
            return self.engine.merge_leaf_records(allrec.data_ptr(), self.world * n_local, n, k)
        finally:
            batch.close()
