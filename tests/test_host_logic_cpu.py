"""Host logic that needs no GPU: the Python mirror of the reference's query surface, and the
rule that the product never touches the oracle."""
import os
import re

import numpy as np
import pytest

from rucene_b200 import engine, search

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tq(t, boost=1.0):
    return search.TermQuery.new(search.Term.new("body", t), boost, None)


def test_boolean_query_build_semantics():
    """search/query/boolean_query.rs:40-87."""
    with pytest.raises(search.IllegalArgument):
        search.BooleanQuery.build([], [], [], [], 0)
    a, b = _tq(1), _tq(2)
    assert search.BooleanQuery.build([a], [], [], [], 0) is a          # single clause collapses
    assert search.BooleanQuery.build([], [b], [], [], 0) is b
    q = search.BooleanQuery.build([], [a, b], [], [], 0)
    assert q.min_should_match == 1                                      # no MUST => defaults to 1
    q = search.BooleanQuery.build([a], [b], [], [], 0)
    assert q.min_should_match == 0
    q = search.BooleanQuery.build([a], [], [], [b], 0)
    assert isinstance(q, search.BooleanQuery) and q.must_not_queries == [b]
    assert search.BooleanQuery.build([], [a, b], [], [], 3).min_should_match == 3


def test_collector_and_term_surface():
    c = search.TopDocsCollector.new(10)
    assert c.needs_scores() and c.top_docs().total_hits() == 0 and c.top_docs().score_docs() == []
    with pytest.raises(search.IllegalArgument):
        search.TopDocsCollector.new(0)
    t = search.Term.new("body", 42)
    assert t.bytes == b"42" and search.IndexReader([], {}).term_id(t) == 42
    assert search.IndexReader([], {("body", b"x"): 7}).term_id(search.Term.new("body", b"x")) == 7
    assert search.IndexReader([], {("body", b"x"): 7}).term_id(search.Term.new("body", b"y")) is None
    assert search.IndexReader([], {}).term_id(search.Term.new("title", 3)) is None


def test_abi_struct_layouts_match_the_header():
    """ctypes/numpy mirrors vs include/rucene_gpu.h."""
    assert engine.CLAUSE_DTYPE.itemsize == 16 and engine.QUERY_DTYPE.itemsize == 16
    assert engine.HIT_DTYPE.itemsize == 8 and engine.TERM_STATE_DTYPE.itemsize == 32
    import ctypes as C
    assert C.sizeof(engine.Config) == 24 and C.sizeof(engine.SearchParams) == 16
    hdr = open(os.path.join(ROOT, "include", "rucene_gpu.h")).read()
    for name, val in [("RG_EINVAL", -1), ("RG_ENODEVICE", -2), ("RG_ECUDA", -3), ("RG_EUNSUPPORTED", -4),
                      ("RG_ENOMEM", -5)]:
        assert re.search(r"#define %s \(%d\)" % (name, val), hdr)
        assert getattr(engine, name) == val


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
    use oracle/: nothing under rucene_b200/ or include/ mentions it."""
    bad = []
    for base in ("rucene_b200", "include"):
        for d, _dirs, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    txt = open(os.path.join(d, f), errors="replace").read()
                    if re.search(r"liboracle|oracle_binding|oracle\.h|orc_[a-z_]+\(", txt):
                        bad.append(os.path.join(d, f))
    assert bad == []
    # and the libraries do not link it
    import subprocess
    from rucene_b200 import _build
    for so in (_build.build_gpu(), _build.build_codec()):
        out = subprocess.run(["ldd", so], stdout=subprocess.PIPE, text=True).stdout
        assert "oracle" not in out
