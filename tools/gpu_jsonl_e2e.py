#!/usr/bin/env python
"""End to end for a JSON Lines dataset shard (row N4): file bytes in pinned host memory -> ids in host memory through
spl_encode_jsonl, next to what the host-side alternative costs (json.loads per line, then spl_encode_batch)."""
import os, sys, time, json, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, synth
from splintr_b200 import Tokenizer, presets as P, _lib
lib = _lib.load()
vb = P.load_vocab_bytes("cl100k_base.tiktoken")
d, o = synth.cfg2(vb, int(os.environ.get("DOCS", "100000")))
texts = synth.unpack_texts(d, o)
blob = ("\n".join(json.dumps({"id": i, "text": t}) for i, t in enumerate(texts)) + "\n").encode()
n = len(blob)
hp = lib.spl_alloc_pinned(n + 64)
ctypes.memmove(hp, blob, n)
tok = Tokenizer.from_pretrained("cl100k_base", devices=[0])
want_ids, want_off = tok.encode_packed(d, o)
ts = []
for it in range(15):
    t0 = time.perf_counter()
    ids, off, st = tok.encode_jsonl((hp, n), return_stats=True)
    ts.append(time.perf_counter() - t0)
assert np.array_equal(ids, want_ids) and np.array_equal(off, want_off)
res = ctypes.c_void_p(); ist = _lib.SplIngestStats(); tc = []
for it in range(15):
    t0 = time.perf_counter()
    rc = lib.spl_encode_jsonl(tok._handle, ctypes.c_void_p(hp), n, b"text", 0, ctypes.byref(res), ctypes.byref(ist))
    tc.append(time.perf_counter() - t0)
    assert rc == 0
    lib.spl_result_free(res)
t0 = time.perf_counter()
sample = blob[:blob.rfind(b"\n", 0, n // 10) + 1]
got = [json.loads(l)["text"] for l in sample.decode().splitlines() if l.strip()]
t_json = (time.perf_counter() - t0) * 10
print(f"jsonl file {n/1e6:.1f} MB, {st['n_docs']} docs: spl_encode_jsonl {min(tc[3:])*1e3:.2f} ms = {n/min(tc[3:])/1e9:.1f} GB/s of file bytes "
      f"(device timeline {st['total_ms']:.2f} ms; through Tokenizer.encode_jsonl incl. numpy copies {min(ts[3:])*1e3:.2f} ms); "
      f"host json.loads of the same file alone: {t_json*1e3:.0f} ms (extrapolated from 10 %)")
