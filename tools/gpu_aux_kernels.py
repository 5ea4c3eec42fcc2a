#!/usr/bin/env python
"""One decode call (k_dec_*) and one SentencePiece-mode encode call (k_sp_*) on cfg2-shaped text, for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch, synth
from splintr_b200 import Tokenizer, presets as P
vb = P.load_vocab_bytes("cl100k_base.tiktoken")
d, o = synth.cfg2(vb, int(os.environ.get("DOCS", "100000")))
n = len(d)
buf = torch.zeros(n + ((-n) % 16), dtype=torch.uint8, device="cuda")
buf[:n].copy_(torch.from_numpy(d))
d_off = torch.from_numpy(o.astype(np.int64)).cuda()
tok = Tokenizer.from_pretrained("cl100k_base", devices=[0])
ids, out_off, nt = tok.encode_device(buf[:n], d_off)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out, boff = tok.decode_device(ids[:nt], out_off); e1.record(); torch.cuda.synchronize()
assert torch.equal(out, buf[:n])
print(f"decode: {nt} ids -> {n} bytes in {e0.elapsed_time(e1):.3f} ms ({n/e0.elapsed_time(e1)/1e6:.1f} GB/s of output bytes)")
sp = Tokenizer.from_pretrained("mistral_v2", devices=[0])
sp.set_profiling(True)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sids, soff, snt = sp.encode_device(buf[:n], d_off); e1.record(); torch.cuda.synchronize()
print(f"sentencepiece (mistral_v2): {n} bytes -> {snt} ids in {e0.elapsed_time(e1):.3f} ms ({n/e0.elapsed_time(e1)/1e6:.1f} GB/s), "
      f"encode stage: { {k: round(v * 1000) for k, v in sp.last_kernel_times().items()} } us")
