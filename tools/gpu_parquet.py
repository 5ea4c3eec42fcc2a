#!/usr/bin/env python
"""Parquet ingestion (row N4): cfg2 as a Parquet file -> ids, through spl_encode_parquet (pages decoded on the device)
against the loop it replaces (pyarrow read + to_pylist + encode_batch_packed, and pyarrow read + encode_arrow)."""
import io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pyarrow as pa, pyarrow.parquet as pq
import synth
from splintr_b200 import Tokenizer, presets as P

tok = Tokenizer.from_pretrained("cl100k_base", devices=[0])
vb = P.load_vocab_bytes(P.PRESETS["cl100k_base"].vocab_file)
d, o = synth.cfg2(vb, 100_000)
texts = synth.unpack_texts(d, o)
table = pa.table({"id": pa.array(range(len(texts))), "text": pa.array(texts, pa.string())})
want_ids, want_off = tok.encode_packed(d, o)


def best(f, n=4):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
    return min(ts), r


for name, kw in (("snappy, dictionary (pyarrow defaults)", dict(compression="snappy")),
                 ("snappy, PLAIN", dict(compression="snappy", use_dictionary=False)),
                 ("snappy, PLAIN, 64 KiB pages", dict(compression="snappy", use_dictionary=False, data_page_size=65536)),
                 ("uncompressed, PLAIN", dict(compression="none", use_dictionary=False))):
    if len(sys.argv) > 1 and sys.argv[1] not in name:
        continue
    buf = io.BytesIO(); pq.write_table(table, buf, **kw); data = buf.getvalue()
    arr = np.frombuffer(data, dtype=np.uint8)
    t_dev, (ids, off, st) = best(lambda: tok.encode_parquet(arr, "text", return_stats=True))
    assert np.array_equal(ids, want_ids) and np.array_equal(off, want_off)
    t_list, _ = best(lambda: tok.encode_batch_packed(pq.read_table(io.BytesIO(data), columns=["text"])["text"].to_pylist()), 2)
    t_arrow, r2 = best(lambda: tok.encode_arrow(pq.read_table(io.BytesIO(data), columns=["text"])["text"]), 3)
    assert np.array_equal(r2[0], want_ids)
    t_read, _ = best(lambda: pq.read_table(io.BytesIO(data), columns=["text"]), 3)
    print(f"{name}: file {len(data) / 1e6:.1f} MB, text {len(d) / 1e6:.1f} MB, {st['n_launches']} launches")
    print(f"    encode_parquet (device pages)            {t_dev * 1e3:8.2f} ms = {len(d) / t_dev / 1e9:6.2f} GB/s of text   (device part {st['total_ms']:.2f} ms)")
    print(f"    pyarrow read_table + encode_arrow        {t_arrow * 1e3:8.2f} ms = {len(d) / t_arrow / 1e9:6.2f} GB/s   (read_table alone {t_read * 1e3:.2f} ms)")
    print(f"    pyarrow read + to_pylist + encode_packed {t_list * 1e3:8.2f} ms = {len(d) / t_list / 1e9:6.2f} GB/s")
sys.stdout.flush()
os._exit(0)          # (pyarrow's thread pool and the CUDA context do not agree on who goes first at interpreter exit)
