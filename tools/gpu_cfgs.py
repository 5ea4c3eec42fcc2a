#!/usr/bin/env python
"""Device-resident kernel times of every BASELINE config at a reduced size (development aid).
Usage: python tools/gpu_cfgs.py [cfg ...]   MB=<approx megabytes per config, default 64>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import synth
from splintr_b200 import Tokenizer, presets as P

MB = float(os.environ.get("MB", "64"))
CFG = {"cfg2": ("cl100k_base", lambda vb: synth.cfg2(vb, int(MB * 1000))),
       "cfg3": ("o200k_base", lambda vb: synth.cfg3(vb, int(MB * 500))),
       "cfg4": ("llama3", lambda vb: synth.cfg4(vb, max(int(MB), 1), 1_000_000.0)),
       "cfg5": ("deepseek_v3", lambda vb: synth.cfg5(vb, int(MB * 660)))}

for name in (sys.argv[1:] or list(CFG)):
    vocab, gen = CFG[name]
    vb = P.load_vocab_bytes(P.PRESETS[vocab].vocab_file)
    data, off = gen(vb)
    n = len(data)
    tok = Tokenizer.from_pretrained(vocab, devices=[0])
    buf = torch.zeros(n + ((-n) % 16), dtype=torch.uint8, device="cuda")
    buf[:n].copy_(torch.from_numpy(data))
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    ids = torch.empty(n, dtype=torch.int32, device="cuda")
    out = torch.empty(len(off), dtype=torch.int64, device="cuda")
    tok.set_profiling(True)
    best = None
    for it in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tok.encode_device(buf[:n], d_off, ids_out=ids, out_offsets=out, sync=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        kt = tok.last_kernel_times()
        if best is None or ms < best[0]:
            best = (ms, kt)
    tok.set_profiling(False)                     # the step as the bench times it: one graph launch, k_bpe_long beside k_bpe
    gbest = None
    for it in range(6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tok.encode_device(buf[:n], d_off, ids_out=ids, out_offsets=out, sync=False)
        e1.record()
        torch.cuda.synchronize()
        gbest = e0.elapsed_time(e1) if gbest is None else min(gbest, e0.elapsed_time(e1))
    ntok = int(out[-1].item())
    print(f"{name} {vocab}: {n/1e6:.1f} MB, {len(off)-1} docs, {ntok} ids ({n/max(ntok,1):.2f} B/id): best {best[0]:.3f} ms = {n/best[0]/1e6:.1f} GB/s (kernel by kernel); graph launch {gbest:.3f} ms = {n/gbest/1e6:.1f} GB/s")
    print("   ", {k: round(v * 1000, 1) for k, v in best[1].items()})
    print("   ", tok.debug_counters(), flush=True)
