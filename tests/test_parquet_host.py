"""Parquet ingestion (row N4), CPU side: the footer / page-header reader (splintr_b200/csrc/spl_parquet_meta.cpp) and
the page decoder the device runs (spl_parquet.h: snappy, RLE / bit-packed hybrid, PLAIN and dictionary pages, V1 / V2
data pages), compiled for the host, against pyarrow on files pyarrow wrote.  What is replaced is the user's loop in
front of Tokenizer.encode_batch (/root/reference/python/splintr/__init__.py documents encode_batch over a list):
    texts = pyarrow.parquet.read_table(path, columns=[column])[column].to_pylist()
with a null row read as an empty document."""
import io
import random

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import hostlib

WORDS = ["the", "of", "tokenizer", "Parquet", "naïve", "日本語", "数据", "🙂", "snappy", "x" * 70, "", " ", "\n", "a\tb"]


def make_texts(rng, n, null_rate=0.0, repeat=False, big=False):
    pool = [" ".join(rng.choice(WORDS) for _ in range(rng.randint(0, 30))) for _ in range(40)] if repeat else None
    out = []
    for _ in range(n):
        if rng.random() < null_rate:
            out.append(None)
        elif repeat:
            out.append(rng.choice(pool))
        else:
            k = rng.randint(0, 400 if not big else 20000)
            out.append("".join(rng.choice(WORDS) + rng.choice(" .,\n") for _ in range(k // 6)))
    return out


def write(table, **kw):
    buf = io.BytesIO()
    pq.write_table(table, buf, **kw)
    return buf.getvalue()


def expect(texts):
    return [b"" if t is None else (t if isinstance(t, bytes) else t.encode()) for t in texts]


OPTS = [dict(compression="none", use_dictionary=False),
        dict(compression="snappy", use_dictionary=False),
        dict(compression="none", use_dictionary=True),
        dict(compression="snappy", use_dictionary=True),
        dict(compression="snappy", use_dictionary=True, data_page_version="2.0"),
        dict(compression="none", use_dictionary=False, data_page_version="2.0"),
        dict(compression="snappy", use_dictionary=False, data_page_version="2.0", data_page_size=2000),
        dict(compression="snappy", use_dictionary=True, data_page_size=1500, row_group_size=700),
        dict(compression="snappy", use_dictionary=True, dictionary_pagesize_limit=3000, data_page_size=4000),   # falls back to PLAIN mid-chunk
        dict(compression="none", use_dictionary=True, dictionary_pagesize_limit=3000, data_page_size=4000, data_page_version="2.0")]


@pytest.mark.parametrize("opt", range(len(OPTS)))
@pytest.mark.parametrize("nulls", [0.0, 0.2])
def test_text_column_matches_pyarrow(opt, nulls):
    rng = random.Random(100 * opt + int(nulls * 10))
    for n, repeat in ((0, False), (1, False), (3000, False), (5000, True)):
        texts = make_texts(rng, n, nulls, repeat)
        table = pa.table({"id": pa.array(range(n), pa.int64()), "text": pa.array(texts, pa.string()), "tail": pa.array([1.5] * n)})
        data = write(table, **OPTS[opt])
        want = expect(pq.read_table(io.BytesIO(data), columns=["text"])["text"].to_pylist())
        got, info = hostlib.parquet(data, "text")
        assert got == want, (OPTS[opt], nulls, n, repeat)
        got2, info2 = hostlib.parquet(data, "text", batch_bytes=1, staged=False)          # one batch per row group; the plain snappy decoder
        assert got2 == want
        assert info2["batches"] >= info["batches"]


def test_required_large_string_binary_and_nested_columns():
    rng = random.Random(5)
    texts = [t for t in make_texts(rng, 2000, 0.0)]
    req = pa.field("text", pa.string(), nullable=False)
    t1 = pa.table([pa.array(texts, pa.string())], schema=pa.schema([req]))
    assert hostlib.parquet(write(t1, compression="snappy"), "text")[0] == expect(texts)
    t2 = pa.table({"text": pa.array(texts, pa.large_string())})
    assert hostlib.parquet(write(t2), "text")[0] == expect(texts)
    blobs = [bytes(rng.randrange(256) for _ in range(rng.randint(0, 300))) for _ in range(500)]
    t3 = pa.table({"blob": pa.array(blobs, pa.binary())})
    assert hostlib.parquet(write(t3, compression="snappy", use_dictionary=False), "blob")[0] == blobs
    # a string inside (nullable) structs: definition levels up to 3, a null at any level is an empty document
    inner = [None if rng.random() < 0.1 else {"body": (None if rng.random() < 0.2 else rng.choice(texts)), "n": 1} for _ in range(3000)]
    outer = [None if rng.random() < 0.1 else {"doc": x} for x in inner]
    t4 = pa.table({"meta": pa.array(outer, pa.struct([("doc", pa.struct([("body", pa.string()), ("n", pa.int32())]))])), "k": pa.array(range(3000))})
    want = [b"" if (o is None or o["doc"] is None or o["doc"]["body"] is None) else o["doc"]["body"].encode() for o in outer]
    for kw in (dict(compression="snappy"), dict(compression="none", use_dictionary=False, data_page_version="2.0")):
        assert hostlib.parquet(write(t4, **kw), "meta.doc.body")[0] == want


def test_large_pages_and_long_documents():
    rng = random.Random(9)
    texts = make_texts(rng, 300, 0.05, big=True)
    table = pa.table({"text": pa.array(texts, pa.string())})
    for kw in (dict(compression="snappy", use_dictionary=False), dict(compression="snappy"), dict(compression="none")):
        data = write(table, **kw)
        assert hostlib.parquet(data, "text")[0] == expect(texts)
    # highly repetitive text: long snappy copies, copies that overlap their own output
    rep = ["ab" * 5000, "x" * 70000, "", "abc" * 3 + "z" * 100, ("0123456789" * 7 + "\n") * 900]
    data = write(pa.table({"text": pa.array(rep)}), compression="snappy", use_dictionary=False)
    assert hostlib.parquet(data, "text")[0] == expect(rep)
    assert hostlib.parquet(data, "text", staged=False)[0] == expect(rep)
    # incompressible rows (literals of several KiB: many input slots per element) between compressible ones, 3 MB page
    blobs = [bytes(rng.randrange(256) for _ in range(rng.randint(3000, 9000))) if k % 3 else b"hello world " * rng.randint(1, 900) for k in range(400)]
    data = write(pa.table({"b": pa.array(blobs, pa.binary())}), compression="snappy", use_dictionary=False, data_page_size=3 << 20)
    assert hostlib.parquet(data, "b")[0] == blobs


def _varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def test_snappy_streams_by_hand():
    """Element forms a compressor rarely emits (standard snappy works in 64 KiB blocks): copies from 65 473 .. 65 535
    bytes back (beyond what the decoder's ring serves), 4-byte offsets, literals with 1..4 length bytes, copies that
    overlap their own output -- random streams built here element by element, both decoders against the construction."""
    rng = random.Random(31)
    for trial in range(60):
        out = bytearray()
        stream = bytearray()
        n_el = rng.randint(1, 400)
        for _ in range(n_el):
            kind = rng.random()
            if kind < 0.35 or len(out) == 0:
                l = rng.choice([1, 5, 60, 61, 255, 256, 257, 4095, 4096, 4097, rng.randint(1, 70000), 65536 + rng.randint(0, 9)])
                lit = bytes(rng.randrange(256) for _ in range(l)) if l < 5000 else rng.randbytes(l)
                if l <= 60:
                    stream.append((l - 1) << 2)
                else:
                    nb = 1 if l - 1 < 1 << 8 else 2 if l - 1 < 1 << 16 else 3 if l - 1 < 1 << 24 else 4
                    if rng.random() < 0.2:
                        nb = 4
                    stream.append((59 + nb) << 2)
                    stream += (l - 1).to_bytes(nb, "little")
                stream += lit
                out += lit
            else:
                l = rng.randint(1, 64)
                off = rng.choice([1, 2, 3, 7, 63, 64, 65, rng.randint(1, 70000), 65472, 65473, 65500, 65535, 65536, 65537, 69000])
                off = min(off, len(out))
                form = rng.random()
                if form < 0.4 and 4 <= l <= 11 and off < 2048:
                    stream.append(1 | ((l - 4) << 2) | ((off >> 8) << 5))
                    stream.append(off & 0xFF)
                elif form < 0.8 and off < 65536:
                    stream.append(2 | ((l - 1) << 2))
                    stream += off.to_bytes(2, "little")
                else:
                    stream.append(3 | ((l - 1) << 2))
                    stream += off.to_bytes(4, "little")
                for i in range(l):
                    out.append(out[len(out) - off])
        full = _varint(len(out)) + bytes(stream)
        for staged in (False, True) + ((2,) if len(full) < 40000 or trial % 10 == 0 else ()):
            assert hostlib.snappy(full, len(out), staged) == bytes(out), (trial, staged)
        # damaged: wrong size, cut stream, offset beyond the output
        assert hostlib.snappy(full, len(out) + 1, True) is None and hostlib.snappy(full, len(out) + 1, False) is None
        if len(full) > 3:
            cut = full[:rng.randrange(1, len(full))]
            assert hostlib.snappy(cut, len(out), True) == hostlib.snappy(cut, len(out), False)
            if len(cut) < 20000:
                assert hostlib.snappy(cut, len(out), 2) is None
    bad = _varint(10) + bytes([(5 - 1) << 2]) + b"hello" + bytes([2 | ((5 - 1) << 2)]) + (6).to_bytes(2, "little")
    assert hostlib.snappy(bad, 10, True) is None and hostlib.snappy(bad, 10, False) is None and hostlib.snappy(bad, 10, 2) is None


def test_warp_wide_decoder_on_written_files():
    """the 32-lane snappy decoder (what the device runs), emulated by 32 threads in lock step, on pages pyarrow wrote:
    text (short elements), repetitive text (overlapping copies), incompressible rows (long literals), dictionary pages"""
    rng = random.Random(23)
    texts = make_texts(rng, 700, 0.1)
    rep = ["ab" * 3000, "x" * 9000, "", "abc" * 3 + "z" * 100, ("0123456789" * 7 + "\n") * 200] * 3
    blobs = [bytes(rng.randrange(256) for _ in range(rng.randint(3000, 9000))) if k % 3 else b"hello world " * rng.randint(1, 300) for k in range(40)]
    cases = [(pa.array(texts, pa.string()), dict(compression="snappy", use_dictionary=False)),
             (pa.array(texts, pa.string()), dict(compression="snappy", data_page_version="2.0", data_page_size=20000)),
             (pa.array(make_texts(rng, 2000, 0.2, repeat=True), pa.string()), dict(compression="snappy")),
             (pa.array(rep, pa.string()), dict(compression="snappy", use_dictionary=False)),
             (pa.array(blobs, pa.binary()), dict(compression="snappy", use_dictionary=False))]
    for col, kw in cases:
        data = write(pa.table({"text": col}), **kw)
        want = expect(col.to_pylist())
        assert hostlib.parquet(data, "text", staged=2)[0] == want, kw
    # damaged pages: same verdict as the plain decoder (an error, or the same rows)
    good = write(pa.table({"text": pa.array(texts[:200])}), compression="snappy", use_dictionary=False)
    for k in range(25):
        b = bytearray(good)
        for _ in range(3):
            b[rng.randrange(60, len(b) - 300)] ^= 1 << rng.randrange(8)
        res = []
        for staged in (0, 2):
            try:
                res.append(hostlib.parquet(bytes(b), "text", text_cap=1 << 21, max_rows=1 << 12, staged=staged)[0])
            except hostlib.ParquetError as e:
                res.append(e.code)
        assert res[0] == res[1], k


def test_refusals_say_what_to_do():
    table = pa.table({"text": pa.array(["a", "b"]), "n": pa.array([1, 2]), "l": pa.array([["x"], ["y", "z"]])})
    data = write(table)
    with pytest.raises(hostlib.ParquetError, match="no column named") as e:
        hostlib.parquet(data, "body")
    assert e.value.code == -1
    with pytest.raises(hostlib.ParquetError, match="BYTE_ARRAY") as e:
        hostlib.parquet(data, "n")
    assert e.value.code == -2
    with pytest.raises(hostlib.ParquetError, match="repeated") as e:
        hostlib.parquet(data, "l.list.element")
    assert e.value.code == -2
    for codec in ("zstd", "gzip"):
        with pytest.raises(hostlib.ParquetError, match="not supported.*snappy") as e:
            hostlib.parquet(write(table, compression=codec), "text")
        assert e.value.code == -2
    with pytest.raises(hostlib.ParquetError, match="DELTA") as e:
        hostlib.parquet(write(table, use_dictionary=False, column_encoding={"text": "DELTA_BYTE_ARRAY"}), "text")
    assert e.value.code == -2


def test_damaged_files_never_crash():
    """Truncations and flipped bytes anywhere in the file: an error or (when the damage hits no structure this column
    needs, or only text bytes) some result -- never a crash or an out-of-bounds read (run under the sanitizer build too)."""
    rng = random.Random(17)
    texts = make_texts(rng, 400, 0.1)
    good = write(pa.table({"text": pa.array(texts)}), compression="snappy", data_page_size=3000)
    assert hostlib.parquet(good, "text")[0] == expect(texts)
    outcomes = {"ok": 0, "err": 0}
    for i in range(400):
        b = bytearray(good)
        if i % 4 == 0:
            b = b[:rng.randrange(len(b))]
            if len(b) >= 4 and rng.random() < 0.5:
                b[-4:] = b"PAR1"
        else:
            for _ in range(rng.randint(1, 4)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        try:
            hostlib.parquet(bytes(b), "text", text_cap=1 << 22, max_rows=1 << 16, staged=bool(i & 1))
            outcomes["ok"] += 1
        except hostlib.ParquetError:
            outcomes["err"] += 1
    assert outcomes["err"] > 50
