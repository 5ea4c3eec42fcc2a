#!/usr/bin/env python
"""Experiment: does running the device pass as two half-batches on two streams (two handles = two workspaces) beat
one pass over the whole batch?  Kernels of one half can fill the tails and the small launch-bound kernels of the other."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch, synth
from splintr_b200 import Tokenizer, presets as P
vb = P.load_vocab_bytes("cl100k_base.tiktoken")
d, o = synth.cfg2(vb, 100000)
n = len(d)


def dev_batch(d, o):
    n = len(d)
    buf = torch.zeros(n + ((-n) % 16) + 16, dtype=torch.uint8, device="cuda")
    buf[:n].copy_(torch.from_numpy(np.ascontiguousarray(d)))
    return buf[:n], torch.from_numpy((o - o[0]).astype(np.int64)).cuda()


for parts in (1, 2, 3, 4):
    toks = [Tokenizer.from_pretrained("cl100k_base", devices=[0]) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    nd = len(o) - 1
    cuts = [nd * k // parts for k in range(parts + 1)]
    batches = []
    for k in range(parts):
        a, b = cuts[k], cuts[k + 1]
        lo = int(o[a]) // 16 * 16          # keep 16-byte alignment of the slice start
        # simplest: copy the slice to its own aligned buffer
        batches.append(dev_batch(d[int(o[a]):int(o[b])], o[a:b + 1]))
    outs = [(torch.empty(int(b[0].numel()) + 16, dtype=torch.int32, device="cuda"),
             torch.empty(int(b[1].numel()), dtype=torch.int64, device="cuda")) for b in batches]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    times = []
    for it in range(12):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(parts):
            streams[k].wait_event(e0)
            with torch.cuda.stream(streams[k]):
                toks[k].encode_device(batches[k][0], batches[k][1], ids_out=outs[k][0], out_offsets=outs[k][1], sync=False)
        for k in range(parts):
            ev = torch.cuda.Event(); ev.record(streams[k]); torch.cuda.current_stream().wait_event(ev)
        e1.record()
        torch.cuda.synchronize()
        if it >= 4:
            times.append(e0.elapsed_time(e1))
    print(f"{parts} part(s): {np.mean(times)*1e3:.0f} us (min {min(times)*1e3:.0f}) -> {n/np.mean(times)/1e6:.1f} GB/s", flush=True)
