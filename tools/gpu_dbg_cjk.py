#!/usr/bin/env python
"""Debug aid: one CJK-heavy batch (tiles that take k_probe's refining pass) against the C oracle."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from splintr_b200 import Tokenizer, presets as P
from oracle.c_oracle import COracle
name = sys.argv[1] if len(sys.argv) > 1 else "cl100k_base"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
rng = random.Random(1)
cjk = "的一是不了人我在有他这为之大来以个中上们到说国和地也子时道出而要于就下得可你年生自会那后能对着事其里所去行过家十用发天如然作方成者多日都三小军二无同么经法当起与好看学进种将还分此心前面又定见只主没公从"
texts = ["".join(rng.choice(cjk + "，。 ab") for _ in range(rng.randint(1, n))) for _ in range(12)]
p = P.PRESETS[name]
tok = Tokenizer.from_pretrained(name, devices=[0])
o = COracle(P.load_vocab_bytes(p.vocab_file), p.pattern, p.special_tokens, p.byte_level)
got = tok.encode_batch(texts)
torch.cuda.synchronize()
print("counters", tok.debug_counters())
want = o.encode_batch(texts)
print("match", got == want, sum(map(len, got)), "ids")
