"""GPU parity: ForUtil block decode through the C ABI vs the oracle (bit-exact)."""
import numpy as np
import pytest

import oracle_binding as ob
from rucene_b200 import codec, engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine()
    yield e
    e.close()


@pytest.mark.parametrize("version", [1, 0])
def test_decode_every_width_raw_and_staged(eng, version):
    table = codec.forutil_table()
    for width in range(0, 33):
        bs = codec.synth_blocks(0x5EED0002 + width, 257, mode=1, param=width, doc_version=version)
        want = ob.forutil_decode(bs.stream, bs.offsets, version, table)
        assert np.array_equal(want.reshape(-1), bs.values)  # the oracle inverts the writer
        got = eng.forutil_decode(bs.stream, bs.offsets, version, table)
        assert np.array_equal(got, want), ("raw", width)
        st = eng.stage_blocks(bs.stream, bs.offsets, version, table)
        st.decode()
        assert np.array_equal(st.fetch(), want), ("staged", width)
        st.close()


@pytest.mark.parametrize("version", [1, 0])
def test_decode_mixed_widths(eng, version):
    table = codec.forutil_table()
    bs = codec.synth_blocks(0x5EED0002, 50000, mode=0, doc_version=version)
    want = ob.forutil_decode(bs.stream, bs.offsets, version, table, n_threads=4)
    got = eng.forutil_decode(bs.stream, bs.offsets, version, table)
    assert np.array_equal(got, want)
    st = eng.stage_blocks(bs.stream, bs.offsets, version, table)
    st.decode()
    assert np.array_equal(st.fetch(), want)
    s = st.stats()
    assert s["decoded_bytes"] == 50000 * 512
    assert s["encoded_bytes"] == int(np.diff(np.append(bs.offsets, bs.stream.size - 64)).sum())


def test_decode_empty_and_ragged(eng):
    table = codec.forutil_table()
    bs = codec.synth_blocks(1, 5, mode=0)
    assert eng.forutil_decode(bs.stream, bs.offsets[:0], 1, table).shape == (0, 128)
    # blocks addressed out of order / subset (ragged offsets)
    idx = np.array([4, 0, 2], dtype=np.int64)
    got = eng.forutil_decode(bs.stream, bs.offsets[idx], 1, table)
    assert np.array_equal(got.reshape(-1), bs.values.reshape(5, 128)[idx].reshape(-1))
    with pytest.raises(engine.EngineError):
        eng.forutil_decode(bs.stream, np.array([bs.stream.size + 5], np.uint64), 1, table)


def test_decode_full_size_round_trip(eng):
    """BASELINE config 2 size (1M blocks): size-independent property decode(encode(x)) == x."""
    table = codec.forutil_table()
    bs = codec.synth_blocks(0x5EED0002, 1_000_000, mode=0, doc_version=1)
    st = eng.stage_blocks(bs.stream, bs.offsets, 1, table)
    st.decode()
    got = st.fetch()
    assert np.array_equal(got.reshape(-1), bs.values)
