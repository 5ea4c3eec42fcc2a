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
for _ in range(4):          # (the first calls pay for torch's allocation of the result tensors)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out, boff = tok.decode_device(ids[:nt], out_off); e1.record(); torch.cuda.synchronize()
assert torch.equal(out, buf[:n])
print(f"decode: {nt} ids -> {n} bytes in {e0.elapsed_time(e1):.3f} ms ({n/e0.elapsed_time(e1)/1e6:.1f} GB/s of output bytes)")
sp = Tokenizer.from_pretrained("mistral_v2", devices=[0])
sp.set_profiling(True)
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sids, soff, snt = sp.encode_device(buf[:n], d_off); e1.record(); torch.cuda.synchronize()
print(f"sentencepiece (mistral_v2): {n} bytes -> {snt} ids in {e0.elapsed_time(e1):.3f} ms ({n/e0.elapsed_time(e1)/1e6:.1f} GB/s), "
      f"encode stage: { {k: round(v * 1000) for k, v in sp.last_kernel_times().items()} } us")
# JSON Lines ingestion (row N4): cfg2 written with json.dumps, ingested on the device
import json
texts = synth.unpack_texts(d, o)
blob = ("\n".join(json.dumps({"id": i, "text": t}) for i, t in enumerate(texts)) + "\n").encode()
nj = len(blob)
jb = torch.zeros(nj + ((-nj) % 16) + 16, dtype=torch.uint8, device="cuda")
jb[:nj].copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); text, offs, st = tok.ingest_jsonl_device(jb[:nj]); e1.record(); torch.cuda.synchronize()
assert torch.equal(text, buf[:n])
print(f"ingest_jsonl: {nj} file bytes, {st['n_docs']} docs -> {st['n_text_bytes']} text bytes in {e0.elapsed_time(e1):.3f} ms "
      f"({nj/e0.elapsed_time(e1)/1e6:.1f} GB/s of file bytes, both stream synchronisations and the newline count included)")
