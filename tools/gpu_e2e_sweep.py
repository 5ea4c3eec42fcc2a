#!/usr/bin/env python
"""End-to-end spl_encode_batch on cfg2 under different pipeline chunk sizes (SPL_CHUNK_BYTES; 0 = automatic ramp)."""
import os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, synth
from splintr_b200 import Tokenizer, presets as P, _lib
lib = _lib.load()
vb = P.load_vocab_bytes("cl100k_base.tiktoken")
d, o = synth.cfg2(vb, int(os.environ.get("DOCS", "100000")))
o = np.ascontiguousarray(o, dtype=np.uint64)
n = len(d); nd = len(o) - 1
hp = lib.spl_alloc_pinned(n + 64); ho = lib.spl_alloc_pinned((nd + 1) * 8)
ctypes.memmove(hp, d.ctypes.data, n); ctypes.memmove(ho, o.ctypes.data, (nd + 1) * 8)
for cb in [int(x) for x in (sys.argv[1:] or ["0", "6250000", "12500000", "25000000", "50000000", "0"])]:
    if cb:
        os.environ["SPL_CHUNK_BYTES"] = str(cb)
    else:
        os.environ.pop("SPL_CHUNK_BYTES", None)
    tok = Tokenizer.from_pretrained("cl100k_base")
    ts, dev = [], []
    for it in range(25):
        res = ctypes.c_void_p()
        t0 = time.perf_counter()
        rc = lib.spl_encode_batch(tok._handle, ctypes.c_void_p(hp), ctypes.c_void_p(ho), nd, 0, ctypes.byref(res))
        dt = time.perf_counter() - t0
        assert rc == 0, _lib.last_error(tok._handle)
        st = _lib.SplStats(); lib.spl_result_stats(res, ctypes.byref(st))
        lib.spl_result_free(res)
        if it >= 5:
            ts.append(dt); dev.append(st.total_ms)
    print(f"chunk_bytes {cb:>9}: wall {np.mean(ts)*1e3:.3f} ms (min {min(ts)*1e3:.3f})  dev {np.mean(dev):.3f} ms  kernels {st.kernel_ms:.2f} -> {n/np.mean(ts)/1e9:.1f} GB/s", flush=True)
    del tok
