#!/usr/bin/env python
"""Per-kernel device times of one device-resident pass at pipeline-chunk sizes (3 MB, 12.5 MB) and at the full
100 MB batch: what a chunk costs beyond its share of the streaming work."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch, synth
from splintr_b200 import Tokenizer, presets as P
vb = P.load_vocab_bytes("cl100k_base.tiktoken")
tok = Tokenizer.from_pretrained("cl100k_base", devices=[0])
tok.set_profiling(True)
for docs in (3000, 12500, 100000):
    d, o = synth.cfg2(vb, docs)
    n = len(d)
    buf = torch.zeros(n + ((-n) % 16), dtype=torch.uint8, device="cuda")
    buf[:n].copy_(torch.from_numpy(d))
    d_off = torch.from_numpy(o.astype(np.int64)).cuda()
    acc = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for it in range(8):
        e0.record()
        tok.encode_device(buf[:n], d_off, sync=False)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            tot += e0.elapsed_time(e1)
            for k, v in tok.last_kernel_times().items():
                acc[k] = acc.get(k, 0.0) + v
    print(f"{n/1e6:.1f} MB: step {tot/5*1e3:.0f} us | " + "  ".join(f"{k} {v/5*1e3:.0f}" for k, v in acc.items()), flush=True)
