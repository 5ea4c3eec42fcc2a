"""Loader (ctypes) and in-tree build of the C-ABI shared library.

The library is the product; there is no Python or CPU fallback behind it.  If it is
missing or cannot be loaded, importing a Tokenizer fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libsplintr_b200.so")
SOURCES = ["spl_api.cu", "spl_kernels.cu", "spl_encode.cu", "spl_decode.cu", "spl_sentencepiece.cu", "spl_ingest.cu", "spl_parquet.cu", "spl_parquet_meta.cpp", "spl_host.cpp"]
HEADERS = ["spl_common.h", "spl_segment.h", "spl_pretok.h", "spl_pretok_fast.h", "spl_sentencepiece.h", "spl_ingest.h", "spl_parquet.h", "spl_parquet_meta.h", "spl_special.h", "spl_host.h", "spl_kernels.cuh", "spl_device.cuh", "spl_fast_dev.cuh", "spl_bpe_bits.h", "unicode_tables.inc",
           os.path.join("..", "..", "include", "splintr_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into splintr_b200/libsplintr_b200.so."""
    if not force and not needs_build():
        return LIB_PATH
    tmp = f"{LIB_PATH}.{os.getpid()}.tmp"                   # ranks / test workers may get here together: build aside, rename
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


# ---- host side of the Python boundary: CPython extension (packing list[str], building list[list[int]]) -------------
PYHOST_SRC = os.path.join(CSRC, "spl_pyhost.c")
_pyhost = None


def _pyhost_path() -> str:
    import sysconfig
    return os.path.join(_HERE, "_pyhost" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_pyhost(force: bool = False) -> str:
    """gcc -shared spl_pyhost.c against this interpreter's headers -> splintr_b200/_pyhost.<abi>.so"""
    import sysconfig
    out = _pyhost_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(PYHOST_SRC):
        return out
    tmp = f"{out}.{os.getpid()}.tmp"                         # ranks of one job may get here together: build aside, rename
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-I" + sysconfig.get_paths()["include"], "-o", tmp, PYHOST_SRC]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, out)
    return out


def pyhost():
    """The extension module, built on first use.  None when it cannot be built (no compiler / headers): the callers
    then run the same steps in Python (tests/test_pyhost.py holds the two against each other) -- host-side packing
    only; the encode path itself has no fallback."""
    global _pyhost
    if _pyhost is None:
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("splintr_b200._pyhost", build_pyhost())
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _pyhost = mod
        except Exception:                                   # noqa: BLE001
            _pyhost = False
    return _pyhost or None


class SplStats(ctypes.Structure):
    _fields_ = [("n_docs", ctypes.c_uint64), ("n_bytes", ctypes.c_uint64), ("n_tokens", ctypes.c_uint64),
                ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64),
                ("kernel_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("n_devices", ctypes.c_int), ("n_launches", ctypes.c_int)]


class SplIngestStats(ctypes.Structure):
    _fields_ = [("n_lines", ctypes.c_uint64), ("n_docs", ctypes.c_uint64), ("n_text_bytes", ctypes.c_uint64),
                ("n_missing", ctypes.c_uint64), ("n_bad", ctypes.c_uint64), ("n_launches", ctypes.c_int)]


# every symbol include/splintr_b200.h declares
EXPORTS = ["spl_create", "spl_destroy", "spl_last_error", "spl_encode_batch", "spl_result_ids",
           "spl_result_offsets", "spl_result_n_docs", "spl_result_n_tokens", "spl_result_stats",
           "spl_result_free", "spl_encode_batch_device", "spl_launches_per_call", "spl_alloc_pinned",
           "spl_free_pinned", "spl_version", "spl_set_profiling", "spl_last_kernel_times",
           "spl_decode_batch", "spl_decode_batch_device", "spl_result_bytes", "spl_result_n_bytes",
           "spl_ingest_jsonl_device", "spl_encode_jsonl", "spl_ingest_parquet", "spl_encode_parquet", "spl_device_status", "spl_debug_counters"]

SPL_OK, SPL_ERR_INVALID_ARG, SPL_ERR_VOCAB, SPL_ERR_CUDA, SPL_ERR_OOM, SPL_ERR_UNSUPPORTED, SPL_ERR_NO_DEVICE = \
    0, -1, -2, -3, -4, -5, -6
SPL_CREATE_BYTE_LEVEL = 1
SPL_CREATE_SENTENCEPIECE = 2
SPL_ENCODE_WITH_SPECIAL = 1

_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "splintr_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, u8p, u32p, u64p = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
    lib.spl_create.restype = ctypes.c_int
    lib.spl_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32,
                               ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_uint32), ctypes.c_size_t,
                               ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(vp)]
    lib.spl_destroy.restype = None
    lib.spl_destroy.argtypes = [vp]
    lib.spl_last_error.restype = ctypes.c_char_p
    lib.spl_last_error.argtypes = [vp]
    lib.spl_encode_batch.restype = ctypes.c_int
    lib.spl_encode_batch.argtypes = [vp, u8p, u64p, ctypes.c_size_t, ctypes.c_uint32, ctypes.POINTER(vp)]
    lib.spl_result_ids.restype = vp
    lib.spl_result_ids.argtypes = [vp]
    lib.spl_result_offsets.restype = vp
    lib.spl_result_offsets.argtypes = [vp]
    lib.spl_result_n_docs.restype = ctypes.c_size_t
    lib.spl_result_n_docs.argtypes = [vp]
    lib.spl_result_n_tokens.restype = ctypes.c_size_t
    lib.spl_result_n_tokens.argtypes = [vp]
    lib.spl_result_stats.restype = None
    lib.spl_result_stats.argtypes = [vp, ctypes.POINTER(SplStats)]
    lib.spl_result_free.restype = None
    lib.spl_result_free.argtypes = [vp]
    lib.spl_encode_batch_device.restype = ctypes.c_int
    lib.spl_encode_batch_device.argtypes = [vp, ctypes.c_int, u8p, ctypes.c_size_t, u64p, ctypes.c_size_t,
                                            ctypes.c_uint32, u32p, ctypes.c_size_t, u64p, vp,
                                            ctypes.POINTER(ctypes.c_uint64)]
    lib.spl_decode_batch.restype = ctypes.c_int
    lib.spl_decode_batch.argtypes = [vp, u32p, u64p, ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.spl_result_bytes.restype = vp
    lib.spl_result_bytes.argtypes = [vp]
    lib.spl_result_n_bytes.restype = ctypes.c_size_t
    lib.spl_result_n_bytes.argtypes = [vp]
    lib.spl_decode_batch_device.restype = ctypes.c_int
    lib.spl_decode_batch_device.argtypes = [vp, ctypes.c_int, u32p, ctypes.c_size_t, u64p, ctypes.c_size_t,
                                            u8p, ctypes.c_size_t, u64p, vp, ctypes.POINTER(ctypes.c_uint64)]
    lib.spl_ingest_jsonl_device.restype = ctypes.c_int
    lib.spl_ingest_jsonl_device.argtypes = [vp, ctypes.c_int, u8p, ctypes.c_size_t, ctypes.c_char_p, u8p, ctypes.c_size_t,
                                            u64p, ctypes.c_size_t, vp, ctypes.POINTER(SplIngestStats)]
    lib.spl_encode_jsonl.restype = ctypes.c_int
    lib.spl_encode_jsonl.argtypes = [vp, u8p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_uint32, ctypes.POINTER(vp),
                                     ctypes.POINTER(SplIngestStats)]
    lib.spl_ingest_parquet.restype = ctypes.c_int
    lib.spl_ingest_parquet.argtypes = [vp, ctypes.c_int, u8p, ctypes.c_size_t, ctypes.c_char_p, u8p, ctypes.c_size_t,
                                       u64p, ctypes.c_size_t, vp, ctypes.POINTER(SplIngestStats)]
    lib.spl_encode_parquet.restype = ctypes.c_int
    lib.spl_encode_parquet.argtypes = [vp, u8p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_uint32, ctypes.POINTER(vp),
                                       ctypes.POINTER(SplIngestStats)]
    lib.spl_launches_per_call.restype = ctypes.c_int
    lib.spl_launches_per_call.argtypes = [vp, ctypes.c_uint32]
    lib.spl_set_profiling.restype = ctypes.c_int
    lib.spl_set_profiling.argtypes = [vp, ctypes.c_int]
    lib.spl_last_kernel_times.restype = ctypes.c_int
    lib.spl_last_kernel_times.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
                                          ctypes.POINTER(ctypes.c_float), ctypes.c_int]
    lib.spl_alloc_pinned.restype = vp
    lib.spl_alloc_pinned.argtypes = [ctypes.c_size_t]
    lib.spl_free_pinned.restype = None
    lib.spl_free_pinned.argtypes = [vp]
    lib.spl_version.restype = ctypes.c_char_p
    lib.spl_version.argtypes = []
    _lib = lib
    return lib


def last_error(handle=None) -> str:
    msg = load().spl_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""
