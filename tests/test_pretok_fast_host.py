"""The bit-parallel pre-tokenizer (splintr_b200/csrc/spl_pretok_fast.h, __host__ __device__) on the
CPU under an emulation of the kernel's tile / halo structure, against the oracle's regex engine
(reference src/core/tokenizer.rs:244-257, patterns :39 and :42).  Checked per tile: a tile the
fast path accepts must equal the regex; a tile it declines must be completed exactly by the
sequential rules + lead-in worker (what k_pretok_fb runs)."""
import random

import numpy as np
import pytest

import hostlib
from conftest import py_oracle
from fuzz_alphabet import ALPHABET

PATS = [("cl100k_base", 0), ("o200k_base", 1)]
# no marks / non-ASCII digits / contraction pile-ups: mostly decided by the fast path itself
NATURAL = list("aaabbcdeiostrvmlXYZABDST   \t\n\r\n\n''.,!?;:-_=#/(){}[]\"$%0123456789") + [
    " ", "　", " ", " ", "ǅ", "ʰ", "ª", "ſ", "K", "好", "世", "界", "你", "é", "É", "ß", "Ω",
    "\U0001f30d", "’", "—", "，", "。", "\x00", "\x1c", " the", " and", "'s", "'t", "'re", "'ve", "'m", "'ll", "'d",
    "'S", "'LL", "'Re", "'ſ", "\n\n", "  ", "\r\n"]


def oracle_starts(name, text):
    return [len(text[:s].encode()) for s, _ in py_oracle(name).find_iter(text)]


def runs_text(rng):
    out = []
    for _ in range(rng.randint(1, 8)):
        k, n = rng.randint(0, 13), int(2 ** rng.uniform(0, 7))
        out.append([
            lambda: "".join(rng.choice("0123456789") for _ in range(n)),
            lambda: " " * n,
            lambda: "".join(rng.choice("\n\r \t") for _ in range(n)),
            lambda: "".join(rng.choice("好世界你日本語") for _ in range(n)),
            lambda: "".join(rng.choice("ABCXYZÉΩ") for _ in range(n)),
            lambda: "".join(rng.choice("abcxyzéß") for _ in range(n)),
            lambda: "".join(rng.choice("aB好Cd世") for _ in range(n)),
            lambda: "".join(rng.choice("!?.,;=-#—。") for _ in range(n)),
            lambda: rng.choice(["'s", "'t", "'re", "'LL", "'d", "'", "''"]),
            lambda: "\n" * n,
            lambda: "".join(rng.choice("　   ") for _ in range(min(n, 20))),
            lambda: rng.choice(["!\n", "?\r\n\n", ".\n \n", "x\n", "1\n"]),
            lambda: "".join(rng.choice("ABC好") for _ in range(n)),
            lambda: rng.choice([" the", " And", "Hello", "WORLD", "camelCase", "XMLParser", " ", "\t"]),
        ][k]())
    return "".join(out)


def check_batch(name, pid, docs, tilings, specials=None):
    enc = [d.encode() for d in docs]
    data = b"".join(enc)
    if not data:
        return 0, 0
    hard = np.zeros(len(data) + 1, dtype=np.uint8)
    spec = np.zeros(len(data) + 1, dtype=np.uint8) if specials else None
    want, off = [], 0
    for d, e in zip(docs, enc):
        hard[off] = 1
        if specials and d in specials:                     # a whole document that is one special-token span
            spec[off:off + len(e)] = 1
            want.append(off)
        else:
            want += [off + s for s in oracle_starts(name, d)]
        off += len(e)
    hard[len(data)] = 1
    want = sorted(set(want))
    tiles = flagged = 0
    for payload, halo in tilings:
        fast, flags, everything = hostlib.scan_fast(pid, data, payload, halo, hard, spec)
        assert everything == want, (name, payload, halo, docs)
        T = payload * 32
        for ti, f in enumerate(flags):
            tiles += 1
            if f:
                flagged += 1
                continue
            assert [x for x in fast if ti * T <= x < (ti + 1) * T] == [x for x in want if ti * T <= x < (ti + 1) * T], \
                (name, payload, halo, ti, docs)
    return tiles, flagged


@pytest.mark.parametrize("name,pid", PATS)
def test_natural_alphabet_is_decided_by_the_fast_path(name, pid):
    rng = random.Random(7 + pid)
    tiles = flagged = 0
    for _ in range(1500):
        docs = ["".join(rng.choice(NATURAL) for _ in range(rng.randint(1, rng.choice([8, 30, 120])))) for _ in range(rng.randint(1, 4))]
        t, f = check_batch(name, pid, docs, ((1, 1), (2, 2), (4, 8)))
        tiles += t
        flagged += f
    if pid == 0:
        assert flagged == 0                       # CL100K needs the fallback for none of these constructs
    else:
        assert flagged < 0.5 * tiles              # O200K: the contraction pile-ups of this (suffix-dense) alphabet


@pytest.mark.parametrize("name,pid", PATS)
def test_adversarial_alphabet(name, pid):
    rng = random.Random(70 + pid)
    for _ in range(700):
        docs = ["".join(rng.choice(ALPHABET) for _ in range(rng.randint(1, 60))) for _ in range(rng.randint(1, 4))]
        docs = [d for d in docs if "᠎" not in d] or ["a"]
        check_batch(name, pid, docs, ((1, 1), (4, 8)))


@pytest.mark.parametrize("name,pid", PATS)
def test_long_runs_cross_words_and_windows(name, pid):
    rng = random.Random(11 + pid)
    tiles = flagged = 0
    for _ in range(1200):
        docs = [runs_text(rng) for _ in range(rng.randint(1, 3))]
        t, f = check_batch(name, pid, docs, ((1, 1), (2, 3), (4, 8), (8, 16)))
        tiles += t
        flagged += f
    assert flagged < 0.15 * tiles


@pytest.mark.parametrize("name,pid", PATS)
def test_special_spans_are_opaque(name, pid):
    rng = random.Random(5 + pid)
    specials = {"<|endoftext|>", "<|im_start|>", "<think>"}
    for _ in range(500):
        docs = []
        for _ in range(rng.randint(1, 6)):
            docs.append(rng.choice(sorted(specials)) if rng.random() < 0.4 else
                        "".join(rng.choice(NATURAL) for _ in range(rng.randint(1, 40))))
        check_batch(name, pid, docs, ((1, 1), (4, 8)), specials)


def test_survey_examples():
    ex = ["don't", "CamelCaseXMLParser", "helloWORLD", "x   \n  y", "1234567", "foo!!!\n\nbar", "$ 100", "a　b", "x''s",
          "  hello", "XYZ好ABC", "使用GPU加速", "it'solid", "I'd've", "A'SB", "'tis", "a \n\n  \n b", "!\n\n  x"]
    for name, pid in PATS:
        for e in ex:
            check_batch(name, pid, [e], ((1, 1), (8, 16)))


@pytest.mark.parametrize("name,pid", PATS[:2])
def test_fused_kernel_trusts_the_first_halo_words(name, pid):
    """k_pretok_probe (pre-tokenizer + probe in one kernel) reads the piece starts of the first words of its right halo
    from its own computation instead of from global memory.  Under the same rule as for the payload (no carry enters
    the word from outside the window) those bits must equal the reference split, for every tiling."""
    rng = random.Random(31 + pid)
    checked = 0
    for it in range(1500):
        if it % 3 == 0:
            docs = [runs_text(rng) for _ in range(rng.randint(1, 3))]
        elif it % 3 == 1:
            docs = ["".join(rng.choice(NATURAL) for _ in range(rng.randint(1, rng.choice([30, 120, 400])))) for _ in range(rng.randint(1, 4))]
        else:
            docs = ["".join(rng.choice(ALPHABET) for _ in range(rng.randint(1, 90))) for _ in range(rng.randint(1, 4))]
            docs = [d for d in docs if "᠎" not in d] or ["a"]
        enc = [d.encode() for d in docs]
        data = b"".join(enc)
        if not data:
            continue
        hard = np.zeros(len(data) + 1, dtype=np.uint8)
        want, off = set(), 0
        for d, e in zip(docs, enc):
            hard[off] = 1
            want |= {off + s for s in oracle_starts(name, d)}
            off += len(e)
        hard[len(data)] = 1
        for payload, halo, ext in ((1, 2, 1), (2, 3, 2), (2, 8, 5), (4, 16, 5)):
            st, flags = hostlib.scan_fast_ext(pid, data, payload, halo, ext, hard)
            cov = np.flatnonzero(st & 8)
            for i in cov.tolist():
                assert bool(st[i] & 4) == (i in want), (name, payload, halo, ext, i, docs)
            checked += len(cov)
    assert checked > 20000
