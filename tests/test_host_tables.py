"""Host table builder (splintr_b200/csrc/spl_host.cpp) + a CPU emulation of the device piece
encoder over exactly those tables (pair table keyed by symbol ids, whole-piece tables),
against the oracle.  Covers what Tokenizer::with_full_options builds (reference
src/core/tokenizer.rs:410-456) and the bpe.rs:67-197 equivalence of the id-pair formulation."""
import random

import pytest

import hostlib
from conftest import VOCABS, py_oracle
from fuzz_alphabet import random_text
from splintr_b200 import presets as P

PID = {"cl100k_base": 0, "o200k_base": 1, "llama3": 1, "deepseek_v3": 1, "mistral_v3": 2}
# number of (idL, idR) -> id splits
N_PAIRS = {"cl100k_base": 233378, "o200k_base": 446189, "llama3": 280147, "deepseek_v3": 238951}


@pytest.fixture(scope="module", params=VOCABS)
def tables(request):
    name = request.param
    p = P.PRESETS[name]
    return name, hostlib.HostTables(P.load_vocab_bytes(p.vocab_file), PID[name], p.byte_level, p.special_tokens)


def test_table_shapes(tables):
    name, t = tables
    st = t.stats()
    assert st["unambiguous"] == 1
    assert st["max_key_len"] <= 128
    if name in N_PAIRS:
        assert st["n_pairs"] == N_PAIRS[name]


def test_host_emulation_matches_oracle(tables, ref_vectors, xcheck_vectors):
    name, t = tables
    o = py_oracle(name)
    for text, ids in ref_vectors[name]["encode"]:
        assert t.encode(text.encode()) == ids
    for text, ids in xcheck_vectors[name]:
        assert t.encode(text.encode()) == ids, (name, text)
    rng = random.Random(5)
    for _ in range(1500):
        s = random_text(rng, 50)
        if "᠎" in s:
            continue
        assert t.encode(s.encode()) == o.encode(s), (name, s)
    for n in (17, 33, 64, 129, 200):
        s = "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(n))
        assert t.encode(s.encode()) == o.encode(s)


def test_vocab_errors():
    with pytest.raises(ValueError):
        hostlib.HostTables(b"nospace\n", 0, False)
    with pytest.raises(ValueError):
        hostlib.HostTables(b"!!!! 5\n", 0, False)
    with pytest.raises(ValueError):
        hostlib.HostTables(b"YQ== notanumber\n", 0, False)


def test_ambiguous_specials_flagged():
    t = hostlib.HostTables(b"YQ== 0\nYg== 1\n", 0, False, {"<a>": 5, "<a><b>": 6})
    assert t.stats()["unambiguous"] == 0
