"""Host side of the Python boundary (splintr_b200/csrc/spl_pyhost.c, the work PyO3 does for
/root/reference/src/python/bindings.rs:337-350): packing list[str] and building list[list[int]] in C against the same
steps in Python.  CPU only."""
import random

import numpy as np
import pytest

from splintr_b200 import _lib
from splintr_b200.tokenizer import Tokenizer
from fuzz_alphabet import random_text


@pytest.fixture(scope="module")
def ph():
    mod = _lib.pyhost()
    assert mod is not None, "spl_pyhost.c did not build"
    return mod


def test_pack_equals_python_packing(ph):
    rng = random.Random(8)
    texts = [random_text(rng, 60) for _ in range(3000)] + ["", "a", "é", "日本語", "🌍" * 5, "x" * 100_000, "\x00nul\x00"]
    rng.shuffle(texts)
    data, off = Tokenizer._pack(texts)
    want_data, want_off = Tokenizer._pack_py(texts)
    assert isinstance(data, np.ndarray) and data.dtype == np.uint8
    assert data.tobytes() == want_data and np.array_equal(off, want_off)
    d0, o0 = Tokenizer._pack([])
    assert len(d0) == 0 and o0.tolist() == [0]
    d1, o1 = Tokenizer._pack(("tuple", "of", "texts"))             # any sequence, as before
    assert d1.tobytes() == b"tupleoftexts" and o1.tolist() == [0, 5, 7, 12]


def test_pack_errors_match_python_packing(ph):
    with pytest.raises(TypeError, match="'int' object cannot be converted to 'PyString'"):
        Tokenizer._pack(["ok", 3])
    with pytest.raises(TypeError, match="'bytes' object cannot be converted to 'PyString'"):
        Tokenizer._pack([b"raw"])
    with pytest.raises(UnicodeEncodeError):
        Tokenizer._pack(["lone \ud800 surrogate"])
    with pytest.raises(UnicodeEncodeError):
        Tokenizer._pack_py(["lone \ud800 surrogate"])


def test_ids_to_lists(ph):
    rng = np.random.default_rng(4)
    counts = rng.integers(0, 40, size=2000)
    counts[:3] = 0
    off = np.zeros(len(counts) + 1, dtype=np.uint64)
    np.cumsum(counts, out=off[1:])
    ids = rng.integers(0, 2 ** 32, size=int(off[-1]), dtype=np.uint64).astype(np.uint32)
    got = ph.ids_to_lists(ids.ctypes.data, off.ctypes.data, len(counts))
    flat, o = ids.tolist(), off.tolist()
    assert got == [flat[o[i]:o[i + 1]] for i in range(len(counts))]
    assert all(type(x) is int for row in got[:50] for x in row)
    assert ph.ids_to_lists(ids.ctypes.data, off.ctypes.data, 0) == []


def test_ids_to_lists_shares_int_objects_and_leaves_gc_alone(ph):
    """token ids below 2^21 come from a per-process table of int objects (made on first use); the result is
    indistinguishable from fresh ints, the garbage collector is back on afterwards, and repeated calls agree"""
    import gc
    import sys
    rng = np.random.default_rng(9)
    ids = np.concatenate([rng.integers(0, 200_000, size=50_000), [0, 1, 255, 256, 257, 2 ** 21 - 1, 2 ** 21, 2 ** 21 + 1, 2 ** 32 - 1]]).astype(np.uint32)
    off = np.array([0, 10, 10, 5000, len(ids)], dtype=np.uint64)
    want = [ids.tolist()[int(off[i]):int(off[i + 1])] for i in range(4)]
    assert gc.isenabled()
    a = ph.ids_to_lists(ids.ctypes.data, off.ctypes.data, 4)
    assert gc.isenabled()
    b = ph.ids_to_lists(ids.ctypes.data, off.ctypes.data, 4)
    assert a == want and b == want
    assert all(type(x) is int for x in a[3][-9:])
    a[0].append(7); a[2][0] = -1                                    # the lists are the caller's own
    assert b == want
    x = int(ids[20])
    before = sys.getrefcount(b[2][10])
    del a
    assert sys.getrefcount(b[2][10]) <= before                       # references were counted: dropping lists releases them
    gc.disable()
    try:
        ph.ids_to_lists(ids.ctypes.data, off.ctypes.data, 4)
        assert not gc.isenabled()                                    # a collector the caller turned off stays off
    finally:
        gc.enable()
