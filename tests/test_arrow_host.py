"""Tokenizer.encode_arrow (host side only, no GPU): an Arrow string column is handed to encode_packed as the column's
own data buffer + widened offsets.  Here encode_packed is replaced by a recorder that "tokenizes" every document into
its bytes, so what reaches the C ABI -- and how the per-chunk results are stitched together -- is checked on the CPU.
What the call replaces: `texts = column.to_pylist()` in front of Tokenizer.encode_batch
(/root/reference/python/splintr/__init__.py documents encode_batch over a list)."""
import random

import numpy as np
import pyarrow as pa
import pytest

from splintr_b200.tokenizer import Tokenizer


class _Recorder(Tokenizer):
    def __init__(self):                                   # no device, no library: only encode_arrow's own logic runs
        self.calls = []

    def encode_packed(self, data, offsets, with_special=False, return_stats=False, _as_lists=False):
        data = np.asarray(data, dtype=np.uint8)
        offsets = np.asarray(offsets, dtype=np.uint64)
        assert offsets[0] == 0 and int(offsets[-1]) == len(data), "the chunk's offsets must start at 0 and span its data"
        assert np.all(np.diff(offsets.astype(np.int64)) >= 0)
        self.calls.append((len(data), len(offsets) - 1, with_special))
        return data.astype(np.uint32), offsets.copy()     # one "id" per byte


def _docs(ids, off):
    return [bytes(ids[int(off[i]):int(off[i + 1])].astype(np.uint8)) for i in range(len(off) - 1)]


def test_encode_arrow_passes_the_columns_buffers():
    rng = random.Random(3)
    texts = [None if rng.random() < 0.15 else "".join(rng.choice("abc é日本🙂\n") for _ in range(rng.randint(0, 40))) for _ in range(3000)]
    want = [b"" if t is None else t.encode() for t in texts]
    for col, exp in ((pa.array(texts, pa.string()), want),
                     (pa.array(texts, pa.large_string()), want),
                     (pa.array(texts, pa.string()).slice(100, 1234), want[100:1334]),
                     (pa.chunked_array([pa.array(texts[:10]), pa.array([], pa.string()), pa.array(texts[10:])]), want),
                     (pa.array([t.encode() if t is not None else None for t in texts], pa.binary()), want),
                     (pa.array([], pa.string()), [])):
        r = _Recorder()
        ids, off = r.encode_arrow(col, with_special=True)
        assert off[0] == 0 and len(off) == len(exp) + 1 and off.dtype == np.uint64 and ids.dtype == np.uint32
        assert _docs(ids, off) == exp
        assert all(ws for _, _, ws in r.calls)
        assert sum(n for _, n, _ in r.calls) == len(exp)


def test_encode_arrow_null_slots_that_hold_bytes_become_empty_documents():
    data = pa.py_buffer(b"helloJUNKworldMORE")
    offs = pa.py_buffer(np.array([0, 5, 9, 14, 18], dtype=np.int32).tobytes())
    valid = pa.py_buffer(bytes([0b0101]))
    odd = pa.Array.from_buffers(pa.string(), 4, [valid, offs, data])
    assert odd.to_pylist() == ["hello", None, "world", None]
    r = _Recorder()
    ids, off = r.encode_arrow(odd)
    assert _docs(ids, off) == [b"hello", b"", b"world", b""]


def test_encode_arrow_refuses_other_types():
    with pytest.raises(TypeError, match="string column"):
        _Recorder().encode_arrow(pa.array([1, 2, 3]))
    with pytest.raises(TypeError, match="string column"):
        _Recorder().encode_arrow(pa.chunked_array([pa.array([1.5])]))
