"""The C++ host mirror (rucene_b200/csrc/host/searcher.hpp) driven like examples/example.rs:
compiled here with g++, run on the GPU box, compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import helpers
import oracle_binding as ob
from rucene_b200 import _build, codec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_example():
    exe = os.path.join(ROOT, "tests", "cpp", "host_mirror_example")
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_example.cpp")
    lib = os.path.dirname(_build.build_gpu())
    _build.build_codec()
    deps = [src, os.path.join(ROOT, "rucene_b200", "csrc", "host", "searcher.hpp")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
                               src, "-o", exe, "-L" + lib, "-lrucene_gpu", "-lrucene_codec",
                               "-Wl,-rpath," + lib])
    return exe


def test_cpp_example_builds():
    _build_example()


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle():
    exe = _build_example()
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    seg = codec.synth_segment(0x5EED0001, 50000, 500, doc_version=1, n_threads=2)
    ix = helpers.oracle_index([seg])
    specs = [("term", 5), ("bool", [(ob.MUST, 3), (ob.MUST, 40)], 0),
             ("bool", [(ob.SHOULD, 1), (ob.SHOULD, 77), (ob.SHOULD, 499999)], 0),
             ("bool", [(ob.MUST, 2), (ob.SHOULD, 9), (ob.SHOULD, 30)], 0)]
    q, c = ob.make_queries(specs)
    hits, counts, total = ix.search_batch(q, c, 10)
    for i in range(4):
        parts = lines[i].split()
        assert int(parts[0]) == total[i]
        got = [tuple(int(x) for x in p.split(":")) for p in parts[1:]]
        want = [(int(h["doc"]), int(np.float32(h["score"]).view(np.uint32))) for h in hits[i][:counts[i]]]
        assert got == want
    assert lines[4] == "filter_same_as_must:1"
    assert lines[5] == "unsupported:1"
