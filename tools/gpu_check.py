#!/usr/bin/env python
"""Quick on-GPU parity + timing check (development aid; the real tests live in tests/)."""
import os, sys, time, random, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import synth
from splintr_b200 import Tokenizer, presets as P
from oracle.py_oracle import OracleTokenizer
from fuzz_alphabet import random_text

def main():
    names = sys.argv[1:] or ["cl100k_base", "o200k_base", "llama3", "deepseek_v3", "mistral_v3"]
    rng = random.Random(5)
    fuzz = [random_text(rng, 80) for _ in range(4000)] + ["", "a", " ", "\n", "Hello world", "你好世界", "Hello 🌍 World!"]
    fuzz += ["".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(20, 700))) for _ in range(40)]
    fuzz += ["=" * 300, "-" * 77 + "\n", " " * 100 + "x", "#" * 33, "ab" * 100, "a" * 257, "日本語のテキストをトークン化します。" * 9, "x" * 5000, " " * 6000, "ab" * 4000]
    spec_texts = ["<|endoftext|>", "a<|endoftext|>b", "<|im_start|>user\nHi<|im_end|>", "<think>x</think>", "<|endoftext|><|endoftext|>", "no specials here", "<|endoftext", "x<|fim_prefix|> y <|fim_suffix|>"]
    spec_texts += [random_text(rng, 30) + rng.choice(["<|endoftext|>", "<|im_end|>", "<think>"]) + random_text(rng, 30) for _ in range(300)]
    for name in names:
        p = P.PRESETS[name]
        vb = P.load_vocab_bytes(p.vocab_file)
        t0 = time.time()
        tok = Tokenizer.from_pretrained(name)
        t1 = time.time()
        orc = OracleTokenizer.from_bytes(vb, p.pattern, p.special_tokens, p.byte_level)
        got = tok.encode_batch(fuzz)
        bad = 0
        for s, g in zip(fuzz, got):
            e = orc.encode(s)
            if g != e:
                bad += 1
                if bad <= 3: print("  MISMATCH", name, repr(s[:60]), g[:12], e[:12])
        got = tok.encode_batch_with_special(spec_texts)
        bads = 0
        for s, g in zip(spec_texts, got):
            e = orc.encode_with_special(s)
            if g != e:
                bads += 1
                if bads <= 3: print("  SPECIAL MISMATCH", name, repr(s[:60]), g[:12], e[:12])
        print(f"{name}: create {t1-t0:.2f}s fuzz {len(fuzz)} bad {bad}; special {len(spec_texts)} bad {bads}", flush=True)
    # cfg1 parity + cfg2 timing on cl100k
    vb = P.load_vocab_bytes("cl100k_base.tiktoken")
    p = P.PRESETS["cl100k_base"]
    tok = Tokenizer.from_pretrained("cl100k_base")
    orc = OracleTokenizer.from_bytes(vb, p.pattern, p.special_tokens, False)
    d, o = synth.cfg1(vb)
    ids, off = tok.encode_packed(d, o)
    exp = orc.encode_batch(synth.unpack_texts(d, o))
    flat = [x for e in exp for x in e]
    print("cfg1 parity:", ids.tolist() == flat, "offsets:", off.tolist() == np.cumsum([0] + [len(e) for e in exp]).tolist(), flush=True)
    d, o = synth.cfg2(vb, int(os.environ.get("CFG2_DOCS", "100000")))
    for it in range(4):
        t0 = time.time()
        ids, off, st = tok.encode_packed(d, o, return_stats=True)
        dt = time.time() - t0
        print(f"cfg2 iter {it}: wall {dt*1e3:.1f} ms, kernel {st['kernel_ms']:.3f} ms, total {st['total_ms']:.3f} ms, tokens {st['n_tokens']}, "
              f"kernel GB/s {len(d)/st['kernel_ms']/1e6:.1f}, e2e GB/s {len(d)/st['total_ms']/1e6:.2f}", flush=True)
    # parity of cfg2 on a sample of docs
    nd = 300
    exp = orc.encode_batch(synth.unpack_texts(d[:int(o[nd])], o[:nd + 1]))
    flat = [x for e in exp for x in e]
    print("cfg2 sample parity:", ids[:int(off[nd])].tolist() == flat, flush=True)

if __name__ == "__main__":
    main()
