#!/usr/bin/env python
"""Generate tests/golden/tiktoken_xcheck.json -- run in the BUILD container only.

Independent second oracle ("O2"): tiktoken's CoreBPE driven with the
reference's own pattern STRINGS (src/core/tokenizer.rs:39,42,64) and the reference's bundled
vocab files, i.e. the recipe the reference uses for its own correctness check
(benchmarks/benchmark.py:506-554, benchmarks/vocabs/benchmark_llama3.py:253-300,
.github/workflows/ci.yml:110-131).  deepseek_v3 / mistral_v3 use the vocab translated back
to raw bytes (byte-level folding).  The texts are the reference's round-trip
corpora (python/tests/test_cl100k.py:472-570, tests/deepseek_v3.rs:104-170) plus seeded
fuzz strings from tests/fuzz_alphabet.py.  Output: {vocab: [[text, ids], ...]}.
"""
import json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tiktoken
from splintr_b200 import presets as P
from oracle.py_oracle import load_tiktoken_bpe, byte_level_decode_bytes
from fuzz_alphabet import random_text

CORPUS = [
    "", "a", " ", "\n", "Hello world", "Hello, world!", "你好世界", "Hello 🌍 World!", " hello world ",
    "I'm sorry you're hurting—breakups suck, but you'll get through it.",
    "He said, ‘Hello’ and she replied, “Goodbye”.",
    "Check if you're using valid credentials—API key, token—in headers.",
    "word—word", "a—b", "test—", "—start", "one—two—three",
    "Check your brake pads or rotors—they might be worn out.",
    "Grinding while braking? Check your brake pads—they might be worn.",
    "def fibonacci(n):\n    if n <= 1:\n        return n\n    return fibonacci(n-1) + fibonacci(n-2)\n",
    '{"name": "test", "values": [1, 2, 3.14159, true, null], "nested": {"k": "v"}}',
    "The quick brown fox jumps over the lazy dog. 1234567890 times!!!\n\n\nNew   paragraph\t\ttabs  ",
    "CamelCaseXMLParser helloWORLD don't DON'T I'LL we've x''s a't't't",
    "日本語のテキストをトークン化します。한국어 텍스트도 있습니다. Привет мир! مرحبا بالعالم",
    "Hello 你好 World 世界!", "    indented\r\n\r\n  windows line endings\r\n",
    "https://example.com/path/to/resource?query=1&other=two#frag",
    "a" * 300, " " * 77 + "x", "=" * 130, "ab" * 90, "x   \n  y", "$ 100", "a　b",
]


def main():
    rng = random.Random(20261017)
    fuzz = [t for t in (random_text(rng, 48) for _ in range(400)) if "᠎" not in t]
    out = {}
    for name in ["cl100k_base", "o200k_base", "llama3", "deepseek_v3", "mistral_v3"]:
        p = P.PRESETS[name]
        enc = load_tiktoken_bpe(P.load_vocab_bytes(p.vocab_file))
        if p.byte_level:
            raw = {}
            for k, v in enc.items():
                d = byte_level_decode_bytes(k)
                if d:
                    raw[d] = v
            enc = raw
        tk = tiktoken.Encoding(name + "_ref", pat_str=p.pattern, mergeable_ranks=enc, special_tokens={})
        out[name] = [[t, tk.encode_ordinary(t)] for t in CORPUS + fuzz]
    path = os.path.join(ROOT, "tests", "golden", "tiktoken_xcheck.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=True, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
