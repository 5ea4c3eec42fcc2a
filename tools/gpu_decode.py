#!/usr/bin/env python
"""Device-resident decode timing on cfg2 (development aid): ids of the 100 MB batch -> bytes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import synth
from splintr_b200 import Tokenizer, presets as P
vb = P.load_vocab_bytes("cl100k_base.tiktoken")
d, o = synth.cfg2(vb, int(os.environ.get("DOCS", "100000")))
tok = Tokenizer.from_pretrained("cl100k_base", devices=[0])
ids, off = tok.encode_packed(d, o)
d_ids = torch.from_numpy(ids.astype(np.int64)).to(torch.int32).cuda()
d_off = torch.from_numpy(off.astype(np.int64)).cuda()
best = 1e9
for it in range(8):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out, out_off = tok.decode_device(d_ids, d_off)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
assert torch.equal(out.cpu(), torch.from_numpy(d))
nb = out.numel()
print(f"decode_device: {len(ids)} ids -> {nb} bytes: best {best:.3f} ms = {nb/best/1e6:.1f} GB/s of output bytes "
      f"({(len(ids)*4 + nb)/best/1e6:.1f} GB/s of ids in + bytes out; includes the size query round trip and the output allocation)")
t0 = time.perf_counter(); data, boff = tok.decode_packed(ids, off); dt = time.perf_counter() - t0
print(f"decode_packed (host in / host out): {dt*1e3:.2f} ms = {nb/dt/1e9:.2f} GB/s")
t0 = time.perf_counter(); raw = b"".join(tok.decode_bytes(ids[int(off[i]):int(off[i+1])].tolist()) for i in range(2000)); dt = time.perf_counter() - t0
print(f"host table lookup (Python mirror of decode_bytes), 2000 docs: {len(raw)/dt/1e6:.1f} MB/s")
