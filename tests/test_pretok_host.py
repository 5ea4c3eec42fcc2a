"""The device pre-tokenizer rules (splintr_b200/csrc/spl_pretok.h, __host__ __device__) run
on the CPU and fuzzed against the oracle's regex engine -- replaces regex find_iter
(reference src/core/tokenizer.rs:244-257 with the patterns at :39, :42, :64)."""
import random

import numpy as np
import pytest

import hostlib
from conftest import py_oracle
from fuzz_alphabet import random_text

PATS = [("cl100k_base", 0), ("o200k_base", 1), ("mistral_v3", 2)]

SURVEY_EXAMPLES = ["don't", "CamelCaseXMLParser", "helloWORLD", "x   \n  y", "1234567", "foo!!!\n\nbar", "$ 100",
                   "a　b", "x''s", "  hello", "a't't't", "!◌́a", "!!◌́a", "XYZ好ABC", "'ſ", "a'ſb", "\r\n\r\n", "a \n", " \n a",
                   "   ", "a   ", "/path/to//x\n//y", "1/2\n/3"]


def oracle_starts(name, text):
    o = py_oracle(name)
    starts, b = [], 0
    pos = 0
    for s, e in o.find_iter(text):
        assert s == pos, "pattern does not tile the text"
        starts.append(len(text[:s].encode()))
        pos = e
    assert pos == len(text)
    return starts


@pytest.mark.parametrize("name,pid", PATS)
def test_sequential_rules_match_regex(name, pid):
    rng = random.Random(pid + 11)
    texts = SURVEY_EXAMPLES + [random_text(rng, 40) for _ in range(6000)]
    for t in texts:
        if "᠎" in t or not t:
            continue
        assert hostlib.scan_seq(pid, t.encode()) == oracle_starts(name, t), (name, t)


@pytest.mark.parametrize("name,pid", PATS)
@pytest.mark.parametrize("chunk", [1, 3, 16])
def test_chunked_sync_split_matches_regex(name, pid, chunk):
    """One worker per `chunk` bytes starting at its first sync point: every piece start is
    marked exactly once and the union equals the sequential result."""
    rng = random.Random(pid * 7 + chunk)
    for _ in range(2500):
        t = random_text(rng, 50)
        if "᠎" in t:
            continue
        assert hostlib.scan_chunked(pid, t.encode(), chunk) == oracle_starts(name, t), (name, t)


@pytest.mark.parametrize("name,pid", PATS)
def test_kernel_work_split_multi_doc_small_window(name, pid):
    """spl_pretok_chunk (the kernel's per-thread routine) with a tiny tile / halo so provisional
    segment ends are exercised; several documents packed back to back."""
    rng = random.Random(pid + 99)
    for _ in range(600):
        docs = [random_text(rng, 30) for _ in range(rng.randint(1, 6))]
        docs = [d for d in docs if "᠎" not in d] or ["a"]
        if rng.random() < 0.2:
            docs.insert(rng.randint(0, len(docs)), "")
        enc = [d.encode() for d in docs]
        data = b"".join(enc)
        if not data:
            continue
        hard = np.zeros(len(data) + 1, dtype=np.uint8)
        want, off = [], 0
        for d, e in zip(docs, enc):
            hard[off] = 1
            want += [off + s for s in oracle_starts(name, d)] if d else []
            off += len(e)
        hard[len(data)] = 1
        for tile, halo, chunk in ((32, 8, 4), (64, 16, 16), (16, 0, 16)):
            got = hostlib.scan_kernel_emul(pid, data, tile, halo, chunk, hard)
            assert got == sorted(set(want)), (name, docs, tile, halo, chunk)
