#!/usr/bin/env python
"""Timeline of one pipelined spl_encode_batch call (SPL_TRACE=1) on cfg2."""
import os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ["SPL_TRACE"] = "1"
import numpy as np, synth
from splintr_b200 import Tokenizer, presets as P, _lib
lib = _lib.load()
vb = P.load_vocab_bytes("cl100k_base.tiktoken")
d, o = synth.cfg2(vb, int(os.environ.get("DOCS", "100000")))
o = np.ascontiguousarray(o, dtype=np.uint64)
tok = Tokenizer.from_pretrained("cl100k_base")
n = len(d); nd = len(o) - 1
hp = lib.spl_alloc_pinned(n + 64); ho = lib.spl_alloc_pinned((nd + 1) * 8)
ctypes.memmove(hp, d.ctypes.data, n); ctypes.memmove(ho, o.ctypes.data, (nd + 1) * 8)
for it in range(4):
    print(f"--- call {it}", file=sys.stderr, flush=True)
    res = ctypes.c_void_p()
    t0 = time.perf_counter()
    rc = lib.spl_encode_batch(tok._handle, ctypes.c_void_p(hp), ctypes.c_void_p(ho), nd, 0, ctypes.byref(res))
    dt = time.perf_counter() - t0
    assert rc == 0, _lib.last_error(tok._handle)
    st = _lib.SplStats(); lib.spl_result_stats(res, ctypes.byref(st))
    print(f"call {it}: wall {dt*1e3:.2f} ms  dev total {st.total_ms:.2f}  kernels {st.kernel_ms:.2f}  -> {n/dt/1e9:.1f} GB/s", file=sys.stderr, flush=True)
    lib.spl_result_free(res)
