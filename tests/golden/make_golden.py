"""Generates tests/golden/topdocs_v1.npz.

The reference (Rust) cannot run here, so these fixtures are produced by the oracle
(oracle/oracle.cpp — the C++ restatement pinned on the reference's own known-answer vectors,
tests/test_oracle_kat.py) over a deterministic synthetic segment (rc_synth_segment uses only
IEEE +,*,/ so every host regenerates the same bytes).  They guard (a) the oracle itself against
regressions, (b) cross-machine determinism of the generator, (c) the GPU path on the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import helpers  # noqa: E402
import oracle_binding as ob  # noqa: E402
from rucene_b200 import codec  # noqa: E402

SEED, MAX_DOC, N_TERMS, K = 0x5EED0001, 120000, 4000, 10


def specs():
    rng = np.random.default_rng(0x60D)
    out = [("term", 0), ("term", 5), ("term", 333), ("term", 3999)]
    for i, ts in enumerate(helpers.distinct_query_terms(rng, N_TERMS, 36, 2, 5)):
        occ = ob.MUST if i % 2 == 0 else ob.SHOULD
        out.append(("bool", [(occ, t) for t in ts], 0))
    return out


def flat_specs():
    rows = []
    for s in specs():
        if s[0] == "term":
            rows.append([0, s[1], -1, -1, -1, -1])
        else:
            ts = [c[1] for c in s[1]]
            rows.append([1 + s[1][0][0]] + ts + [-1] * (5 - len(ts)))
    return np.array(rows, dtype=np.int64)


def main():
    seg = codec.synth_segment(SEED, MAX_DOC, N_TERMS, doc_version=1, n_threads=1)
    ix = helpers.oracle_index([seg])
    q, c = ob.make_queries(specs())
    hits, counts, total = ix.search_batch(q, c, K)
    np.savez_compressed(os.path.join(HERE, "topdocs_v1.npz"), docs=hits["doc"], scores=hits["score"].view(np.uint32),
                        counts=counts, total=total, specs=flat_specs(),
                        doc_file_crc=np.uint32(zlib.crc32(seg.doc_file.tobytes())),
                        norms_crc=np.uint32(zlib.crc32(seg.norms.tobytes())),
                        params=np.array([SEED, MAX_DOC, N_TERMS, K], dtype=np.int64))
    print("wrote topdocs_v1.npz:", len(q), "queries")


if __name__ == "__main__":
    main()
