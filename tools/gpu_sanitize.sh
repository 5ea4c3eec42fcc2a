#!/bin/bash
# compute-sanitizer (memcheck, then racecheck + initcheck) over a small mixed workload through the C-ABI.
mkdir -p gpurun_out
cat > /tmp/san_driver.py <<'PY'
import os, sys, random
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools")); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, synth
from splintr_b200 import Tokenizer, presets as P
from fuzz_alphabet import random_text
rng = random.Random(11)
texts = [random_text(rng, 60) for _ in range(300)] + ["", "a", "x" * 3000, " " * 2500, "ab" * 700, "日本語のテキスト" * 40, "Hello <|endoftext|> world"]
texts += ["".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(20, 900))) for _ in range(20)]
for name in ("cl100k_base", "llama3", "deepseek_v3"):
    tok = Tokenizer.from_pretrained(name, devices=[0])
    a = tok.encode_batch(texts)
    b = tok.encode_batch_with_special(texts)
    vb = P.load_vocab_bytes(P.PRESETS[name].vocab_file)
    d, o = synth.cfg2(vb, 300) if name == "cl100k_base" else (synth.cfg4(vb, 1, 300000.0) if name == "llama3" else synth.cfg5(vb, 300))
    ids, off = tok.encode_packed(d, o)
    data, boff = tok.decode_packed(ids, off)                       # decode kernels (row N2)
    assert np.array_equal(data, d) and np.array_equal(boff, o)
    print(name, sum(len(x) for x in a), sum(len(x) for x in b), len(ids), flush=True)
# SentencePiece mode (row N3): transform kernels + encode stage over the transformed text, with specials
ws = [" ", "  ", "\n", "\t", "\x0b", "\xa0", "\u3000"]
sp_texts = texts + ["".join(rng.choice(ws) if rng.random() < 0.4 else rng.choice("ab,Z\u00e9\u4e2d") for _ in range(rng.randint(0, 300))) for _ in range(200)]
sp_texts += [" " * 5000 + "y", "x\x0b" + " " * 5000, "a [INST] b [/INST]  <|think|> "]
tok = Tokenizer.from_pretrained("mistral_v2", devices=[0])
a = tok.encode_batch(sp_texts)
b = tok.encode_batch_with_special(sp_texts)
assert tok.decode_batch(a[:50]) == sp_texts[:50]
print("mistral_v2", sum(len(x) for x in a), sum(len(x) for x in b), flush=True)
# JSON Lines ingestion (row N4)
import json
from jsonl_cases import make_lines, join_lines
lines, want = make_lines(5, 800)
tok = Tokenizer.from_pretrained("cl100k_base", devices=[0])
ids, off = tok.encode_jsonl(join_lines(lines))
assert len(off) == len(want) + 1
print("jsonl", len(want), len(ids), flush=True)
# Parquet ingestion (row N4): snappy (warp-wide decoder), dictionary pages, V2 pages, nulls
import io, pyarrow as pa, pyarrow.parquet as pq
ptexts = [None if rng.random() < 0.1 else t for t in texts[:250]] * 3
for kw in (dict(compression="snappy"), dict(compression="snappy", use_dictionary=False, data_page_version="2.0", data_page_size=4000), dict(compression="none")):
    buf = io.BytesIO(); pq.write_table(pa.table({"text": pa.array(ptexts, pa.string())}), buf, **kw)
    ids, off = tok.encode_parquet(buf.getvalue(), "text")
    flat, o = ids.tolist(), off.tolist()
    assert [flat[o[k]:o[k + 1]] for k in range(len(ptexts))] == tok.encode_batch([t or "" for t in ptexts])
print("parquet", len(ptexts), flush=True)
# tiles dense in ids (k_emit in several phases) and user special-token sets that overlap (candidate walk)
rare = "\u9f98\u9750\u9f49\u7228\u706a\u9ea4\u9c7b\u9955"
dense = ["".join(rng.choice(rare) for _ in range(4000)), "".join(rng.choice("!?;:") + rng.choice("\n\t") for _ in range(6000)),
         "".join(chr(rng.randrange(0x80, 0x250)) for _ in range(4000))]
a = tok.encode_batch(dense)
pc = P.PRESETS["cl100k_base"]
sp = {"<a>": 200001, "<a><b>": 200002, "<b>": 200003, "b><": 200004}
tok2 = Tokenizer.from_bytes(P.load_vocab_bytes(pc.vocab_file), pc.pattern, sp)
alpha = "".join(sp) + " xy\n"
b = tok2.encode_batch_with_special(["", "<a><b>", "x<a><b><a>"] + ["".join(rng.choice(alpha) for _ in range(rng.randint(0, 80))) for _ in range(200)])
print("dense", sum(len(x) for x in a), "overlapping specials", sum(len(x) for x in b), flush=True)
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_driver.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitizer_$tool.log
  tail -6 gpurun_out/sanitizer_$tool.log
done
