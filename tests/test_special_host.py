"""Special-token matching for ARBITRARY sets (strings that contain or overlap one another): the per-document walker of
splintr_b200/csrc/spl_special.h on the CPU against the oracle's restatement of aho-corasick's Standard, non-overlapping
find_iter (/root/reference/src/core/tokenizer.rs:429-434, 842-874; oracle/py_oracle.py _special_find_iter)."""
import random

import hostlib
from oracle.py_oracle import OracleTokenizer, CL100K_BASE_PATTERN


def _oracle_matches(text: bytes, specials):
    o = OracleTokenizer({bytes([b]): b for b in range(256)}, {s.decode("latin-1"): 1000 + i for i, s in enumerate(specials)}, CL100K_BASE_PATTERN)
    o._special_bytes = [(s, i) for i, s in enumerate(specials)]
    return [(s, e, k) for s, e, k in o._special_find_iter(text)]


def test_overlapping_sets_match_the_oracle():
    rng = random.Random(42)
    for trial in range(3000):
        alpha = rng.choice(["ab", "abc", "<|>a", "<>|xy"])
        n = rng.randint(1, 6)
        specials = list({bytes(rng.choice(alpha.encode()) for _ in range(rng.randint(1, 5))) for _ in range(n)})
        text = bytes(rng.choice(alpha.encode()) for _ in range(rng.randint(0, 60)))
        assert hostlib.special_walk(text, specials) == _oracle_matches(text, specials), (text, specials)


def test_named_cases():
    # one string contains another; a suffix of one is a prefix of another; a longer match ends later than a shorter one
    cases = [(b"xx<a><b>yy<a>", [b"<a>", b"<a><b>"]),
             (b"ab<|end|><|endoftext|>", [b"<|end|>", b"<|endoftext|>", b"|><"]),
             (b"aaaaa", [b"aa", b"aaa"]),
             (b"abcabcab", [b"abc", b"bca", b"cab", b"ab"]),
             (b"", [b"a"]), (b"zzz", [b"a"])]
    for text, sp in cases:
        assert hostlib.special_walk(text, sp) == _oracle_matches(text, sp), (text, sp)
    assert hostlib.special_walk(b"xx<a><b>yy<a>", [b"<a>", b"<a><b>"]) == [(2, 5, 0), (10, 13, 0)]   # earliest end wins: "<a><b>" never matches
