"""The oracles against every golden vector the reference's own tests hold for the encode path
(tests/golden/reference_vectors.json cites the reference file:line of each), against the
tiktoken cross-check fixture (tools/make_golden.py), and against each other."""
import random

import pytest

from conftest import VOCABS, ALL_VOCABS, py_oracle, c_oracle
from fuzz_alphabet import random_text


@pytest.mark.parametrize("name", ALL_VOCABS)
def test_py_oracle_reference_vectors(name, ref_vectors):
    o = py_oracle(name)
    for text, ids in ref_vectors[name]["encode"]:
        assert o.encode(text) == ids, (name, text)
    for text, ids in ref_vectors[name]["special"]:
        assert o.encode_with_special(text) == ids, (name, text)


@pytest.mark.parametrize("name", ALL_VOCABS)
def test_c_oracle_reference_vectors(name, ref_vectors):
    o = c_oracle(name)
    for text, ids in ref_vectors[name]["encode"]:
        assert o.encode(text) == ids, (name, text)
    for text, ids in ref_vectors[name]["special"]:
        assert o.encode_with_special(text) == ids, (name, text)


@pytest.mark.parametrize("name", VOCABS)
def test_oracles_match_tiktoken_fixture(name, xcheck_vectors):
    po, co = py_oracle(name), c_oracle(name)
    texts = [t for t, _ in xcheck_vectors[name]]
    want = [ids for _, ids in xcheck_vectors[name]]
    assert co.encode_batch(texts) == want
    for t, ids in zip(texts[:120], want[:120]):
        assert po.encode(t) == ids, (name, t)


@pytest.mark.parametrize("name", ALL_VOCABS)
def test_c_oracle_equals_py_oracle_fuzz(name):
    rng = random.Random(hash(name) & 0xFFFF)
    texts = [t for t in (random_text(rng, 60) for _ in range(1500)) if "᠎" not in t]
    po, co = py_oracle(name), c_oracle(name)
    assert co.encode_batch(texts) == po.encode_batch(texts)
    sp = list(po.special_tokens)[:6]
    st = [random_text(rng, 20) + rng.choice(sp) + random_text(rng, 20) + rng.choice(sp) for _ in range(200)]
    st = [t for t in st if "᠎" not in t]
    assert co.encode_batch(st, with_special=True) == po.encode_batch_with_special(st)


def test_bpe_toy_vocab_cases():
    """bpe.rs:199-251 unit cases on the toy vocabulary."""
    from oracle.py_oracle import byte_pair_encode
    enc = {b"a": 0, b"b": 1, b"c": 2, b"ab": 3, b"bc": 4, b"abc": 5}
    assert byte_pair_encode(b"a", enc) == [0]
    assert byte_pair_encode(b"ab", enc) == [3]
    assert byte_pair_encode(b"abc", enc) == [5]
    assert byte_pair_encode(b"", enc) == []
    assert byte_pair_encode(b"ac", enc) == [0, 2]
    del enc[b"abc"]
    assert byte_pair_encode(b"abc", enc) == [3, 2]          # ab (rank 3) merges before bc (rank 4)
    assert byte_pair_encode(b"zab", enc) == [3]             # unknown byte dropped (bpe.rs:187-191)


def test_byte_level_map():
    """byte_level.rs:166-255."""
    from oracle.py_oracle import byte_level_encode, byte_level_decode_bytes, BYTE_TO_CHAR
    assert len(set(BYTE_TO_CHAR)) == 256
    assert byte_level_encode(b" ").decode() == "Ġ"
    assert byte_level_encode("你好".encode()).decode() == "ä½łå¥½"
    assert byte_level_decode_bytes(byte_level_encode(bytes(range(256)))) == bytes(range(256))


@pytest.mark.parametrize("name", ["mistral_v1", "mistral_v2"])
def test_sentencepiece_mode_properties(name):
    """SentencePiece branch (tokenizer.rs:737-795): round trips of the reference's own corpora
    (python/tests/test_mistral_v1.py:68-109), vocab sizes (tests/mistral_v2.rs:116), the first-byte rule."""
    o = py_oracle(name)
    for text in ["Hello, world!", "The quick brown fox jumps over the lazy dog.", "1234567890", " world!",
                 "Special characters: !@#$%^&*()", "Unicode: こんにちは 世界 🦀", "Mixed: Hello 你好 🌍 World!",
                 "Multi-line\ntext\nwith\nnewlines",
                 'def hello_world():\n    print("Hello, World!")\n\nif __name__ == "__main__":\n    hello_world()\n']:
        assert o.decode(o.encode(text)) == text
    assert o.vocab_size == {"mistral_v1": 32054, "mistral_v2": 32822}[name]
    us = o.encoder["\u2581".encode()]
    assert o.encode(" ") == [us] and o.encode("a ") == o.encode("a") + [us]
    # a whitespace chunk that starts with a non-ASCII space is encoded as a word: its spaces stay 0x20
    assert o.encode("\u3000 a")[-2:] == [o.encoder[b" "], o.encoder[b"a"]]
