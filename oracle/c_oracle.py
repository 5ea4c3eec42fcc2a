"""Build recipe + ctypes wrapper for oracle/c_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__ (build / smoke) and bench.py (cpu_baseline, --impl reference)
may import this.  See the header of c_oracle.c for the reference file:line map.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "c_oracle.c")
LIB = os.path.join(_HERE, "liboracle_c.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-ldl"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + p.stdout + p.stderr)
    return LIB


_lib: Optional[ctypes.CDLL] = None


def _load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB)
        vp = ctypes.c_void_p
        lib.orc_create.restype = vp
        lib.orc_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_int,
                                   ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_uint32), ctypes.c_size_t,
                                   ctypes.c_char_p, ctypes.c_size_t]
        lib.orc_destroy.restype = None
        lib.orc_destroy.argtypes = [vp]
        lib.orc_max_threads.restype = ctypes.c_int
        lib.orc_encode_batch.restype = ctypes.c_longlong
        lib.orc_encode_batch.argtypes = [vp, vp, vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, vp, ctypes.c_size_t, vp]
        lib.orc_split.restype = ctypes.c_int
        lib.orc_split.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t, vp]
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(_load().orc_max_threads())


class COracle:
    """C restatement of core::Tokenizer's encode path (PCRE2 backend, linked-list BPE)."""

    def __init__(self, vocab_data: bytes, pattern: str, special_tokens: Optional[Dict[str, int]] = None,
                 byte_level: bool = False, sentencepiece: bool = False):
        lib = _load()
        self._cap_mul = 3 if sentencepiece else 1          # a space becomes the 3 bytes of U+2581
        sp = list((special_tokens or {}).items())
        n = len(sp)
        strs = (ctypes.c_char_p * max(n, 1))(*[s.encode("utf-8") for s, _ in sp])
        ids = (ctypes.c_uint32 * max(n, 1))(*[i for _, i in sp])
        err = ctypes.create_string_buffer(256)
        self._h = lib.orc_create(vocab_data, len(vocab_data), pattern.encode("utf-8"),
                                 int(bool(byte_level)) | (2 if sentencepiece else 0),
                                 strs, ids, n, err, 256)
        if not self._h:
            raise ValueError(err.value.decode("utf-8", "replace"))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            _load().orc_destroy(h)
            self._h = None

    def encode_packed(self, data: np.ndarray, offsets: np.ndarray, with_special: bool = False,
                      n_threads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n_docs = len(offsets) - 1
        cap = self._cap_mul * int(offsets[-1]) + 16 if n_docs >= 0 else 16
        ids = np.empty(cap, dtype=np.uint32)
        out_off = np.zeros(n_docs + 1, dtype=np.uint64)
        n = _load().orc_encode_batch(self._h, data.ctypes.data, offsets.ctypes.data, n_docs, int(with_special),
                                     n_threads, ids.ctypes.data, cap, out_off.ctypes.data)
        if n < 0:
            raise RuntimeError(f"c_oracle encode failed with code {n}")
        return ids[:n].copy(), out_off

    def encode_batch(self, texts: Sequence[str], with_special: bool = False, n_threads: int = 0):
        enc = [t.encode("utf-8") for t in texts]
        offs = np.zeros(len(enc) + 1, dtype=np.uint64)
        if enc:
            np.cumsum(np.fromiter((len(b) for b in enc), dtype=np.uint64, count=len(enc)), out=offs[1:])
        data = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8)
        ids, off = self.encode_packed(data, offs, with_special, n_threads)
        flat = ids.tolist()
        o = off.tolist()
        return [flat[o[i]:o[i + 1]] for i in range(len(texts))]

    def encode(self, text: str):
        return self.encode_batch([text])[0]

    def encode_with_special(self, text: str):
        return self.encode_batch([text], with_special=True)[0]

    def split(self, text: str):
        """Piece start byte offsets of `text` under the pattern."""
        b = text.encode("utf-8")
        starts = np.zeros(len(b) + 1, dtype=np.uint8)
        rc = _load().orc_split(self._h, b, len(b), starts.ctypes.data)
        if rc:
            raise RuntimeError(f"pcre2 error {rc}")
        return np.flatnonzero(starts).tolist()
