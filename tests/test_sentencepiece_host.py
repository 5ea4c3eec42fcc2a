"""SentencePiece mode (reference src/core/tokenizer.rs:737-795) as position-local bitmap rules
(splintr_b200/csrc/spl_sentencepiece.h), run on the CPU through tests/csrc/hosttest.cpp: the transformed text T' and
its piece starts against the pieces the oracle's sequential walk produces, and the ids of the device-equivalent piece
encoder over the first-id-wins tables against the oracle."""
import random

import numpy as np
import pytest

import hostlib
from conftest import SP_VOCABS, py_oracle
from fuzz_alphabet import random_text
from splintr_b200 import presets as P

EDGE = ["", " ", "  ", "a", "a ", " a", "a b", "a  b", "\n", " \n ", "\n\n  x", "a\tb", "\t", "a \t b",
        "　 a", "a 　 b", "a　 b", "x\x0b y", "\x0b", " \x0b ", "a\x0b\x0b b", "a \x85 b", "\xa0 x", "x \xa0",
        "       x", "x       ", "a\r\n b", "a\x0c b", "a\x1c b", "Hello world", " world!", "Hello 🌍 World!",
        "def f():\n    return 1\n\n\nx = 2  # c\n", "a" * 70 + " " * 40 + "b" * 33, " " * 100, "\n" * 50,
        "a b   c", "tab\t\tsep  end ", "日本語 テキスト 　全角　スペース"]


def _pieces(t2: bytes, starts):
    b = starts + [len(t2)]
    return [t2[b[i]:b[i + 1]] for i in range(len(starts))]


def _texts(seed, n, maxlen):
    rng = random.Random(seed)
    ws = [" ", " ", " ", "  ", "\n", "\t", "\x0b", "\x0c", "\r", " ", "　", " ", "\x85", "\x1c"]
    out = []
    for _ in range(n):
        if rng.random() < 0.5:
            t = random_text(rng, maxlen)
        else:
            t = "".join(rng.choice(ws) if rng.random() < 0.45 else rng.choice("ab,Zé中🙂") for _ in range(rng.randint(0, maxlen)))
        if "᠎" not in t:
            out.append(t)
    return out


def test_transform_pieces_equal_the_sequential_walk():
    o = py_oracle("mistral_v1")
    for t in EDGE + _texts(11, 6000, 70):
        data = t.encode("utf-8")
        t2, starts, _ = hostlib.sp_transform(data)
        want = o.sentencepiece_pieces(t)
        assert t2 == b"".join(want), t
        assert _pieces(t2, starts) == want, (t, _pieces(t2, starts), want)


def test_transform_long_text_crosses_words_and_tiles():
    """32-byte word and 4 KiB tile boundaries in every phase: a 40 KB text equals the concatenation rule."""
    o = py_oracle("mistral_v1")
    rng = random.Random(3)
    t = "".join(_texts(17, 900, 60))
    data = t.encode("utf-8")
    assert len(data) > 20000
    t2, starts, _ = hostlib.sp_transform(data)
    assert _pieces(t2, starts) == o.sentencepiece_pieces(t)
    # a whitespace run far longer than a tile, led by a raw character and by a space
    for lead in ("\x0b", " "):
        t = "x" + lead + " " * 9000 + "y"
        t2, starts, _ = hostlib.sp_transform(t.encode())
        assert _pieces(t2, starts) == o.sentencepiece_pieces(t)


def test_segments_restart_the_walk():
    """Document starts and special-span edges are segment boundaries: the pending run is flushed there
    (tokenizer.rs:842-874 calls encode() once per gap)."""
    o = py_oracle("mistral_v2")
    rng = random.Random(9)
    sp = ["[INST]", "[/INST]", "<|think|>"]
    for _ in range(1500):
        parts = []
        for _k in range(rng.randint(1, 5)):
            parts.append(("gap", "".join(rng.choice(" a\n\x0b　b ") for _ in range(rng.randint(0, 12)))))
            if rng.random() < 0.6:
                parts.append(("sp", rng.choice(sp)))
        data = b"".join(p.encode() for _, p in parts)
        hard = np.zeros(len(data) + 1, dtype=np.uint8)
        spec = np.zeros(len(data) + 1, dtype=np.uint8)
        want, pos = [], 0
        for kind, p in parts:
            b = p.encode()
            hard[pos] = 1
            if kind == "sp":
                spec[pos:pos + len(b)] = 1
                want.append(b)
            else:
                want.extend(o.sentencepiece_pieces(p))
            pos += len(b)
        t2, starts, sstarts = hostlib.sp_transform(data, hard, spec)
        assert _pieces(t2, starts) == want, parts


@pytest.mark.parametrize("name", SP_VOCABS)
def test_host_tables_first_id_wins_and_ids_match_oracle(name, ref_vectors):
    p = P.PRESETS[name]
    t = hostlib.HostTables(P.load_vocab_bytes(p.vocab_file), P.SPL_PATTERN_SENTENCEPIECE, False, p.special_tokens,
                           sentencepiece=True)
    assert t.stats()["unambiguous"] == 1
    o = py_oracle(name)
    for text, ids in ref_vectors[name]["encode"]:
        assert t.encode_sp(text.encode()) == ids
    for text in EDGE + _texts(23, 2500, 60):
        assert t.encode_sp(text.encode("utf-8")) == o.encode(text), (name, text)
