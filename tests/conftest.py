import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

VOCABS = ["cl100k_base", "o200k_base", "llama3", "deepseek_v3", "mistral_v3"]
SP_VOCABS = ["mistral_v1", "mistral_v2"]                # SentencePiece mode (tokenizer.rs:737-795)
ALL_VOCABS = VOCABS + SP_VOCABS


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(ROOT, "tests", "golden", name), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_vectors():
    return load_golden("reference_vectors.json")


@pytest.fixture(scope="session")
def xcheck_vectors():
    return load_golden("tiktoken_xcheck.json")


_ORACLES = {}


def py_oracle(name):
    """oracle/py_oracle.py instance for a preset (cached)."""
    from oracle.py_oracle import OracleTokenizer
    from splintr_b200 import presets as P
    if name not in _ORACLES:
        p = P.PRESETS[name]
        _ORACLES[name] = OracleTokenizer.from_bytes(P.load_vocab_bytes(p.vocab_file), p.pattern, p.special_tokens, p.byte_level,
                                                     p.sentencepiece)
    return _ORACLES[name]


_CORACLES = {}


def c_oracle(name):
    from oracle.c_oracle import COracle
    from splintr_b200 import presets as P
    if name not in _CORACLES:
        p = P.PRESETS[name]
        _CORACLES[name] = COracle(P.load_vocab_bytes(p.vocab_file), p.pattern, p.special_tokens, p.byte_level, p.sentencepiece)
    return _CORACLES[name]
