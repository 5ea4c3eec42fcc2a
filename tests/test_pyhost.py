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
