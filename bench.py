#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on B200.

Workload (config.workload): BASELINE config 4 — 5-term SHOULD BooleanQuery (DisjunctionSumScorer)
BM25 top-100, batch 4096, 100M-doc synthetic Zipfian index (1M terms), evaluated by
IndexSearcher::search through the C ABI.  One "step" = one pass of the whole query batch.

  value  : queries/sec, whole job, kernels only (plan + index resident in HBM), max over ranks
  e2e    : queries/sec through rg_search_batch with HOST query arrays in and HOST TopDocs out
           (planning, H2D of the plan, kernels, D2H of results inside the timed region)
  roofline: k_eval_or (dominant kernel): algorithmic bytes per launch / CUDA-event time
  cpu_baseline: the oracle (C++ restatement of the reference's CPU path, kind "port") timed on the
           box's host cores on a bounded sample of the same batch
  forutil_decode: BASELINE config 2 (1M x 128-int blocks) GB/s vs the measured HBM peak

N>1 (torchrun): the index is split into N docid-range segments, one per GPU (strong scaling);
every rank evaluates the whole batch on its segment (search_parallel semantics), ONE NCCL
all-gather moves the per-segment top-k records and every rank replays finish_parallel in leaf
order.  `--impl reference` times the oracle alone (rank 0 only).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_INDEX = 0x5EED0001
SEED_BLOCKS = 0x5EED0002
SEED_QUERIES = 0x5EED0004


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs", type=int, default=int(os.environ.get("RUCENE_BENCH_DOCS", 100_000_000)))
    ap.add_argument("--terms", type=int, default=int(os.environ.get("RUCENE_BENCH_TERMS", 1_000_000)))
    ap.add_argument("--batch", type=int, default=int(os.environ.get("RUCENE_BENCH_BATCH", 4096)))
    ap.add_argument("--qterms", type=int, default=5)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("RUCENE_BENCH_CPU_SAMPLE", 256)))
    ap.add_argument("--range-postings", type=int, default=0)
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-columns", action="store_true",
                    help="RG_CFG_NO_COLUMNS: evaluate every clause from its block stream (A/B runs)")
    ap.add_argument("--workload", default="c4", choices=["c4", "c3"],
                    help="c4 (headline): 5-term SHOULD top-100 batch 4096 on 100M docs; c3: 2-term MUST "
                         "(ConjunctionScorer) top-10 batch 1024 on 10M docs")
    a = ap.parse_args()
    if a.workload == "c3":
        env = os.environ
        if "RUCENE_BENCH_DOCS" not in env and "--docs" not in sys.argv:
            a.docs = 10_000_000
        if "--terms" not in sys.argv:
            a.terms = 100_000
        if "--batch" not in sys.argv:
            a.batch = 1024
        if "--qterms" not in sys.argv:
            a.qterms = 2
        if "--k" not in sys.argv:
            a.k = 10
    return a


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.p, self.t, self.index = [], None, None, index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def gen_queries(n_terms, batch, qterms, seed):
    """ranks r = floor(V^U) (log-uniform), distinct within a query; 0-based ids."""
    rng = np.random.default_rng(seed)
    out = np.zeros((batch, qterms), np.int64)
    for i in range(batch):
        chosen = []
        while len(chosen) < qterms:
            r = int(np.floor(float(n_terms) ** rng.random()))
            t = min(max(r, 1), n_terms) - 1
            if t not in chosen:
                chosen.append(t)
        out[i] = chosen
    return out


def build_query_arrays(qt, weights_of, engine_mod, must=False):
    """-> rg_query[], rg_clause[] for BooleanQuery::build(musts | shoulds of TermQuery)."""
    batch, qterms = qt.shape
    q = np.zeros(batch, engine_mod.QUERY_DTYPE)
    c = np.zeros(batch * qterms, engine_mod.CLAUSE_DTYPE)
    q["clause_begin"] = np.arange(batch) * qterms
    q["n_clauses"] = qterms
    q["min_should_match"] = 0
    q["flags"] = engine_mod.Q_BOOLEAN
    c["occur"] = engine_mod.MUST if must else engine_mod.SHOULD
    c["term_id"] = qt.reshape(-1)
    c["weight"] = weights_of(qt.reshape(-1))
    c["cache_id"] = 0
    return q, c


def oracle_setup(seg, stats_df, stats, total_max_doc):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    ix = ob.Index(1.2, 0.75)
    ix.add_segment(seg)
    return ob, ix


def oracle_queries(ob, qt, must=False):
    occ = ob.MUST if must else ob.SHOULD
    specs = [("bool", [(occ, int(t)) for t in row], 0) for row in qt]
    return ob.make_queries(specs)


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rucene_b200 import codec
    seg = codec.synth_segment(SEED_INDEX, args.docs, args.terms, doc_version=1)
    ob, ix = oracle_setup(seg, None, None, args.docs)
    qt = gen_queries(args.terms, args.batch, args.qterms, SEED_QUERIES)
    cores = os.cpu_count() or 1
    sample = min(args.cpu_sample, args.batch)
    times = []
    for step in range(args.warmup + args.steps):
        lo = (step * sample) % max(1, args.batch - sample + 1)
        q, c = oracle_queries(ob, qt[lo:lo + sample], args.workload == "c3")
        t0 = time.perf_counter()
        ix.search_batch(q, c, args.k, parallel_mode=0, n_threads=cores)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    line = {"impl": "reference", "metric": "queries/sec", "value": value, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32/f32",
            "data": "synthetic", "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": "%d queries of the batch per step, one query per thread" % sample},
            "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, world):
    kind = ("C3: %d-term MUST BooleanQuery (ConjunctionScorer)" if args.workload == "c3"
            else "C4: %d-term SHOULD BooleanQuery (DisjunctionSumScorer)") % args.qterms
    return {"workload": "%s BM25 top-%d, batch %d, %d-doc Zipfian synthetic index, %d terms"
                        % (kind, args.k, args.batch, args.docs, args.terms),
            "batch": args.batch, "k": args.k, "docs": args.docs, "terms": args.terms,
            "segments": world, "parallelism": "1 docid-range segment per GPU" if world > 1 else "single GPU",
            "cache": "index image (GBs) is larger than the 126 MB L2; no explicit flush"}


class _CudaArray:
    """__cuda_array_interface__ view of engine-owned device memory (for torch.as_tensor)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 2, "strides": None}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from rucene_b200 import codec, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- index: one docid-range segment per rank -------------------------------------------
    t_gen0 = time.perf_counter()
    seg_docs = args.docs // world
    seg = codec.synth_segment(SEED_INDEX + rank, seg_docs, args.terms, doc_version=1)
    t_gen = time.perf_counter() - t_gen0
    eng = engine.Engine(device=local_rank, range_postings=args.range_postings,
                        flags=engine.CFG_NO_COLUMNS if args.no_columns else 0)
    # a dedicated (non-default) torch stream: the engine launches on it and torch.cuda.Event
    # timing sees exactly those launches
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    t_up0 = time.perf_counter()
    eng.upload_segment(seg, doc_base=rank * seg_docs)
    t_up = time.perf_counter() - t_up0

    # ---- weights: statistics of the largest segment = segment 0 (searcher.rs:311-351) -------
    df0 = torch.from_numpy(seg.terms["doc_freq"].astype(np.int32)).to(dev)
    st0 = torch.tensor([seg.doc_count, seg.sum_total_term_freq], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(df0, 0)
        dist.broadcast(st0, 0)
    df0 = df0.cpu().numpy()
    doc_count, sum_ttf = int(st0[0]), int(st0[1])
    avgdl = codec.bm25_avg_field_length(sum_ttf, doc_count, args.docs)
    eng.set_norm_cache(0, codec.bm25_norm_cache(1.2, 0.75, avgdl))
    idf_cache = {}

    def weights_of(term_ids):
        out = np.zeros(len(term_ids), np.float32)
        for i, t in enumerate(term_ids):
            t = int(t)
            if t not in idf_cache:
                idf_cache[t] = np.float32(codec.bm25_idf(int(df0[t]), doc_count))  # boost 1.0
            out[i] = idf_cache[t]
        return out

    qt = gen_queries(args.terms, args.batch, args.qterms, SEED_QUERIES)
    must = args.workload == "c3"
    q, c = build_query_arrays(qt, weights_of, engine, must)
    mode = engine.MODE_SEARCH_PARALLEL if world > 1 else engine.MODE_SEARCH

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: kernels only, plan resident --------------------------------------------------
    batch = eng.prepare(q, c, args.k, k1=1.2, mode=mode)
    rec_ptr = rec_bytes = None
    gathered = local_rec = None
    if world > 1:
        rec_ptr, rec_bytes = batch.leaf_records()
        local_rec = torch.as_tensor(_CudaArray(rec_ptr, rec_bytes * args.batch), device=dev)
        gathered = torch.empty(world * rec_bytes * args.batch, dtype=torch.uint8, device=dev)

    def one_step(fetch=False):
        batch.run()
        if world > 1:
            dist.all_gather_into_tensor(gathered, local_rec)
            return eng.merge_leaf_records(gathered.data_ptr(), world, args.batch, args.k)
        return batch.fetch() if fetch else None

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launch_count()
    ev0.record(stream)
    for _ in range(args.steps):
        one_step()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total[0]) / args.steps
    launches = eng.launch_count() - launches0
    result = one_step(fetch=True)
    if world > 1:
        batch.fetch()  # populates the per-kernel CUDA-event timings of the last run
    eval_ms = eng.last_kernel_ms("eval")
    replay_ms = eng.last_kernel_ms("replay")
    bstats = batch.stats()
    n_cols, col_bytes = batch.columns()
    value = args.batch / (ms_step / 1e3)

    # ---- e2e: host arrays in, host TopDocs out ---------------------------------------------
    def e2e_step():
        if world == 1:
            return eng.search_batch(q, c, args.k, k1=1.2, mode=mode)
        b2 = eng.prepare(q, c, args.k, k1=1.2, mode=mode)
        b2.run()
        p2, _ = b2.leaf_records()
        loc = torch.as_tensor(_CudaArray(p2, rec_bytes * args.batch), device=dev)
        dist.all_gather_into_tensor(gathered, loc)
        out = eng.merge_leaf_records(gathered.data_ptr(), world, args.batch, args.k)
        b2.close()
        return out

    e2e_steps = max(1, min(args.steps, 2))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_res = e2e_step()
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms[0])
    # where the e2e time goes (one extra, untimed-for-the-metric step through the split calls)
    e2e_split = None
    if world == 1:
        torch.cuda.synchronize()
        ta = time.perf_counter()
        b3 = eng.prepare(q, c, args.k, k1=1.2, mode=mode)
        tb = time.perf_counter()
        b3.run()
        torch.cuda.synchronize()
        tc = time.perf_counter()
        b3.fetch()
        td = time.perf_counter()
        b3.close()
        te = time.perf_counter()
        e2e_split = {"prepare_ms": (tb - ta) * 1e3, "run_ms": (tc - tb) * 1e3, "fetch_ms": (td - tc) * 1e3,
                     "destroy_ms": (te - td) * 1e3}
    d2h = args.batch * args.k * 8 + args.batch * 4 + args.batch * 8
    h2d = bstats["h2d_bytes"] + q.nbytes + c.nbytes

    # ---- consistency: device-resident path == e2e path ---------------------------------------
    if result is not None:
        assert np.array_equal(result[0]["doc"], e2e_res[0]["doc"]) and np.array_equal(result[2], e2e_res[2])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_eval_or) ----------------------------------------
    peak, peak_src = measured_peaks()
    algo_bytes = bstats["algorithmic_bytes"]
    achieved = algo_bytes / (eval_ms / 1e3) / 1e9 if eval_ms > 0 else 0.0
    traffic = None
    try:  # DRAM bytes per step from the committed ncu --set full captures (same workload only)
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            tr = json.load(f)
        if world == 1 and not must and tr["docs"] == args.docs and tr["batch"] == args.batch:
            traffic = tr["k_eval_or_dram_bytes_per_step"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_eval_and" if must else "k_eval_or", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "note": "instruction/latency-bound, not HBM-bound: ~3.6 warp instructions per posting "
                        "(profiles/README.md); DRAM traffic is far below the algorithmic bytes because hot posting "
                        "blocks and the batch's score columns hit in L2; kernel_ms includes k_build_columns",
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": eval_ms, "replay_ms": replay_ms,
                "postings_per_launch": bstats["postings"], "work_items": bstats["items"],
                "candidate_slots": bstats["candidate_slots"],
                "score_columns": {"n": n_cols, "bytes": col_bytes,
                                  "note": "dense clauses shared by >= 4 disjunctions of the batch are scored once per "
                                          "step (k_build_columns, inside the timed region) and read as f32 columns"}}

    # ---- ForUtil decode microbench (BASELINE config 2) --------------------------------------
    decode = None
    if not args.no_decode:
        nb = 1_000_000
        bs = codec.synth_blocks(SEED_BLOCKS, nb, mode=0, doc_version=1)
        st = eng.stage_blocks(bs.stream, bs.offsets, 1, codec.forutil_table())
        for _ in range(3):
            st.decode()
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        d0.record(stream)
        for _ in range(reps):
            st.decode()
        d1.record(stream)
        torch.cuda.synchronize()
        dms = d0.elapsed_time(d1) / reps
        s = st.stats()
        rw = (s["encoded_bytes"] + s["decoded_bytes"]) / (dms / 1e3) / 1e9
        ro = s["encoded_bytes"] / (dms / 1e3) / 1e9
        decode = {"blocks": nb, "ms": dms, "read_write_GBs": rw, "read_only_GBs": ro, "frac_of_hbm_peak": rw / peak,
                  "bytes_read": s["encoded_bytes"], "bytes_written": s["decoded_bytes"],
                  "note": "staged 16B-aligned blocks, uniform widths 1..32; 777 MB/pass > L2"}
        st.close()

    # ---- cpu baseline: the oracle on the host cores, bounded sample --------------------------
    cpu = None
    if world == 1:
        ob, ix = oracle_setup(seg, None, None, args.docs)
        cores = os.cpu_count() or 1
        sample = min(args.cpu_sample, args.batch)
        oq, oc = oracle_queries(ob, qt[:sample], must)
        t0 = time.perf_counter()
        want = ix.search_batch(oq, oc, args.k, parallel_mode=0, n_threads=cores)
        cdt = time.perf_counter() - t0
        cpu = {"value": sample / cdt, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": "first %d queries of the batch, one query per thread, %.2f s wall" % (sample, cdt)}
        # the timed GPU result must be identical to the reference's on the sample
        got = result
        ok = (np.array_equal(got[2][:sample], want[2]) and np.array_equal(got[1][:sample], want[1])
              and all(np.array_equal(got[0][i][:want[1][i]]["doc"], want[0][i][:want[1][i]]["doc"]) and
                      np.array_equal(got[0][i][:want[1][i]]["score"].view(np.uint32),
                                     want[0][i][:want[1][i]]["score"].view(np.uint32)) for i in range(sample)))
        cpu["parity_on_sample"] = "identical TopDocs" if ok else "MISMATCH"

    line = {"metric": "queries/sec", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32/f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": args.batch / (e2e_ms / 1e3), "unit": "queries/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "split": e2e_split},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "forutil_decode": decode,
            "setup": {"index_gen_s": t_gen, "upload_s": t_up, "index_image_bytes": eng.index_bytes(),
                      "doc_file_bytes": int(seg.doc_file.size), "postings": int(seg.sum_doc_freq),
                      "host_cores": os.cpu_count()}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
