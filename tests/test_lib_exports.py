"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/splintr_b200.h declares; without a device the product fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from splintr_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.load()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "splintr_b200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(spl_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(lib):
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/splintr_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms


def test_version(lib):
    assert b"sm_100a" in lib.spl_version()


def test_invalid_arguments_do_not_crash(lib):
    h = ctypes.c_void_p()
    assert lib.spl_create(None, 0, 0, 0, None, None, 0, None, 0, ctypes.byref(h)) == _lib.SPL_ERR_INVALID_ARG
    assert lib.spl_create(b"YQ== 0\n", 7, 99, 0, None, None, 0, None, 0, ctypes.byref(h)) == _lib.SPL_ERR_INVALID_ARG
    assert not h.value
    lib.spl_destroy(None)
    lib.spl_result_free(None)
    assert lib.spl_result_n_tokens(None) == 0


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    rc = lib.spl_create(b"YQ== 0\n", 7, 0, 0, None, None, 0, None, 0, ctypes.byref(h))
    assert rc == _lib.SPL_ERR_NO_DEVICE and not h.value
    assert "no CPU fallback" in _lib.last_error(None)
    from splintr_b200 import Tokenizer
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Tokenizer.from_pretrained("cl100k_base")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "splintr_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn), encoding="utf-8", errors="replace") as f:
                    src = f.read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle's", "") or fn == "spl_pretok.h", fn


def test_python_surface_matches_reference_class():
    """bindings.rs:57-446: method names of the PyO3 class."""
    from splintr_b200 import Tokenizer
    for m in ["from_pretrained", "from_bytes", "pcre2", "jit", "encode", "encode_rayon", "encode_with_special",
              "decode", "decode_bytes", "decode_lossy", "encode_batch", "encode_batch_with_special", "decode_batch",
              "decode_batch_lossy", "vocab_size", "streaming_decoder", "byte_level_streaming_decoder",
              "clear_cache", "cache_len"]:
        assert hasattr(Tokenizer, m), m
    import splintr_b200
    for c in ["CL100K_BASE_PATTERN", "O200K_BASE_PATTERN", "LLAMA3_PATTERN", "CL100K_AGENT_TOKENS"]:
        assert hasattr(splintr_b200, c)
    with pytest.raises(ValueError, match="Unknown pretrained model"):
        Tokenizer.from_pretrained("nope")
