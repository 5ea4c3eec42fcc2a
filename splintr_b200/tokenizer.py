"""Host-side mirror of the reference's Python `Tokenizer` class
(/root/reference/src/python/bindings.rs:57-446) over the splintr_b200 C-ABI.

Same method names, argument meaning and error behaviour as the PyO3 class; the encode
methods pack the texts into one contiguous UTF-8 buffer + offsets and hand it to
`spl_encode_batch` -- the regex split, special-token scan, BPE merge and batch loop all run
on the GPU.  Decode is a host table lookup (tokenizer.rs:877-911), as in the reference.
"""
from __future__ import annotations

import base64
import ctypes
import os
import threading
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from . import presets as _presets


def _byte_level_maps():
    """byte_level.rs:46-74."""
    direct = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    b2c = {}
    nxt = 256
    for b in range(256):
        if b in direct:
            b2c[b] = chr(b)
        else:
            b2c[b] = chr(nxt)
            nxt += 1
    return b2c, {c: b for b, c in b2c.items()}


_BYTE_TO_CHAR, _CHAR_TO_BYTE = _byte_level_maps()


def _parse_tiktoken(data: bytes) -> Dict[bytes, int]:
    """vocab.rs:57-89 (host copy, used for decode tables and vocab_size only)."""
    enc: Dict[bytes, int] = {}
    for line in data.split(b"\n"):
        if not line:
            continue
        sp = line.rfind(b" ")
        if sp < 0:
            raise ValueError("Invalid line format: Missing space separator")
        try:
            tok = base64.b64decode(line[:sp], validate=True)
        except Exception as e:  # binascii.Error
            raise ValueError(f"Invalid base64 encoding: {e}")
        try:
            rank = int(line[sp + 1:].decode("utf-8").strip())
        except Exception:
            raise ValueError(f"Invalid line format: Invalid rank: {line[sp + 1:]!r}")
        enc[tok] = rank
    return enc


def _parse_tiktoken_decoder(data: bytes) -> Dict[int, bytes]:
    """vocab.rs:101-143, decoder half: EVERY id of a SentencePiece vocabulary decodes (duplicated byte strings
    keep all their ids; the encoder half -- first id wins -- lives in the C++ table builder)."""
    dec: Dict[int, bytes] = {}
    for line in data.split(b"\n"):
        if not line:
            continue
        sp = line.rfind(b" ")
        if sp < 0:
            raise ValueError("Invalid line format: Missing space separator")
        dec[int(line[sp + 1:].decode("utf-8").strip())] = base64.b64decode(line[:sp], validate=True)
    return dec


class Tokenizer:
    """Drop-in for `splintr.Tokenizer` (bindings.rs:57-446)."""

    # -- construction --------------------------------------------------------------------
    def __init__(self, vocab_path: str, pattern: str, special_tokens: Optional[Dict[str, int]] = None):
        # bindings.rs:70-83: file errors surface as IOError
        try:
            with open(vocab_path, "rb") as f:
                data = f.read()
        except OSError as e:
            raise IOError(str(e))
        try:
            self._init(data, pattern, special_tokens or {}, byte_level=False, sentencepiece=False)
        except ValueError as e:
            raise IOError(str(e))

    @classmethod
    def _make(cls, data: bytes, pattern: str, special_tokens: Dict[str, int], byte_level: bool,
              sentencepiece: bool = False, devices: Optional[Sequence[int]] = None) -> "Tokenizer":
        self = cls.__new__(cls)
        self._init(data, pattern, special_tokens, byte_level, sentencepiece, devices)
        return self

    def _init(self, data: bytes, pattern: str, special_tokens: Dict[str, int], byte_level: bool,
              sentencepiece: bool, devices: Optional[Sequence[int]] = None) -> None:
        self._handle = None
        # one call in flight per handle (include/splintr_b200.h): ctypes releases the GIL around the C call, the reference's
        # binding never does (no allow_threads in bindings.rs), so threads that share a Tokenizer are serialised here
        self._lock = threading.RLock()
        if not isinstance(pattern, str):
            raise TypeError("pattern must be str")
        if sentencepiece:
            # tokenizer.rs:589-640 + the encode branch :737-795; the device implements that walk for the split
            # the reference's presets use it with (`[^\s]+|\s+`, tokenizer.rs:56)
            pid = _presets.SPL_PATTERN_SENTENCEPIECE if pattern == _presets.SENTENCEPIECE_PATTERN else None
        else:
            pid = _presets.PATTERN_IDS.get(pattern)
        if pid is None:
            raise ValueError("Regex compilation error: the device pre-tokenizer implements only the "
                             "CL100K_BASE, O200K_BASE/LLAMA3 and MISTRAL_V3 patterns, and SENTENCEPIECE_PATTERN "
                             "in SentencePiece mode")
        for k, v in special_tokens.items():
            if not isinstance(k, str) or not isinstance(v, int):
                raise TypeError("special_tokens must map str -> int")
            if not 0 <= v <= 0xFFFFFFFF:                     # u32 in the reference (PyO3 raises OverflowError)
                raise OverflowError("special token id out of range for u32")
        self._vocab_data = bytes(data)
        self._pattern = pattern
        self._special_tokens = dict(special_tokens)
        self._special_decoder = {v: k for k, v in special_tokens.items()}
        self._byte_level = bool(byte_level)
        self._sentencepiece = bool(sentencepiece)
        self._devices = list(devices) if devices is not None else None
        self._decoder: Optional[Dict[int, bytes]] = None
        self._vocab_size: Optional[int] = None

        lib = _lib.load()
        sp = list(special_tokens.items())
        n = len(sp)
        strs = (ctypes.c_char_p * max(n, 1))(*[s.encode("utf-8") for s, _ in sp])
        ids = (ctypes.c_uint32 * max(n, 1))(*[i for _, i in sp])
        if self._devices:
            devs = (ctypes.c_int * len(self._devices))(*self._devices)
            ndev = len(self._devices)
        else:
            devs, ndev = None, 0
        h = ctypes.c_void_p()
        rc = lib.spl_create(self._vocab_data, len(self._vocab_data), pid,
                            (_lib.SPL_CREATE_BYTE_LEVEL if byte_level else 0) |
                            (_lib.SPL_CREATE_SENTENCEPIECE if sentencepiece else 0),
                            strs, ids, n, devs, ndev, ctypes.byref(h))
        if rc != _lib.SPL_OK:
            msg = _lib.last_error(None)
            if rc == _lib.SPL_ERR_NO_DEVICE:
                raise RuntimeError(f"splintr_b200: {msg}")
            if rc in (_lib.SPL_ERR_CUDA, _lib.SPL_ERR_OOM):
                raise RuntimeError(f"splintr_b200: {msg} (code {rc})")
            raise ValueError(msg or f"spl_create failed with code {rc}")
        self._handle = h

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                _lib.load().spl_destroy(h)
            except Exception:
                pass
            self._handle = None

    @staticmethod
    def from_pretrained(name: str, devices: Optional[Sequence[int]] = None) -> "Tokenizer":
        """bindings.rs:101-166.  `devices` (extension): CUDA device indices to shard over."""
        p = _presets.get_preset(name)
        if p is None:
            raise ValueError(f"Unknown pretrained model: {name}. See from_pretrained docstring for supported models.")
        return Tokenizer._make(_presets.load_vocab_bytes(p.vocab_file), p.pattern, p.special_tokens,
                               p.byte_level, p.sentencepiece, devices)

    @staticmethod
    def from_bytes(vocab_data: bytes, pattern: str, special_tokens: Optional[Dict[str, int]] = None,
                   devices: Optional[Sequence[int]] = None) -> "Tokenizer":
        """bindings.rs:174-187."""
        return Tokenizer._make(bytes(vocab_data), pattern, special_tokens or {}, False, False, devices)

    @staticmethod
    def from_bytes_byte_level(vocab_data: bytes, pattern: str, special_tokens: Optional[Dict[str, int]] = None,
                              devices: Optional[Sequence[int]] = None) -> "Tokenizer":
        """tokenizer.rs:562-569 (Rust-only constructor in the reference; exposed for tests)."""
        return Tokenizer._make(bytes(vocab_data), pattern, special_tokens or {}, True, False, devices)

    @staticmethod
    def from_bytes_sentencepiece(vocab_data: bytes, pattern: str, special_tokens: Optional[Dict[str, int]] = None,
                                 devices: Optional[Sequence[int]] = None) -> "Tokenizer":
        """tokenizer.rs:589-640 (Rust-only constructor in the reference; what from_pretrained("mistral") uses)."""
        return Tokenizer._make(bytes(vocab_data), pattern, special_tokens or {}, False, True, devices)

    def _clone(self) -> "Tokenizer":
        return Tokenizer._make(self._vocab_data, self._pattern, self._special_tokens, self._byte_level,
                               self._sentencepiece, self._devices)

    def pcre2(self, use_pcre2: bool = True) -> "Tokenizer":
        """bindings.rs:207-214.  There is one pre-tokenizer (the device rules); the switch
        is accepted and returns a new instance, like the reference."""
        return self._clone()

    def jit(self, use_jit: bool = True) -> "Tokenizer":
        """bindings.rs:235-242 (accepted and ignored)."""
        return self._clone()

    # -- encode -----------------------------------------------------------------------------
    @staticmethod
    def _pack_py(texts: Sequence[str]) -> Tuple[bytes, np.ndarray]:
        enc = []
        for t in texts:
            if not isinstance(t, str):
                raise TypeError(f"argument 'texts': '{type(t).__name__}' object cannot be converted to 'PyString'")
            enc.append(t.encode("utf-8"))          # lone surrogates raise UnicodeEncodeError, as in PyO3
        offsets = np.zeros(len(enc) + 1, dtype=np.uint64)
        if enc:
            np.cumsum(np.fromiter((len(b) for b in enc), dtype=np.uint64, count=len(enc)), out=offsets[1:])
        return b"".join(enc), offsets

    @staticmethod
    def _pack(texts: Sequence[str]):
        """list[str] -> (concatenated UTF-8, uint64 offsets[n + 1]): what PyO3 does for bindings.rs:337 when it
        extracts Vec<String>.  csrc/spl_pyhost.c copies every text once, straight into one array."""
        ph = _lib.pyhost()
        if ph is None:
            return Tokenizer._pack_py(texts)
        if not isinstance(texts, list):
            texts = list(texts)
        offsets = np.empty(len(texts) + 1, dtype=np.uint64)
        total = ph.pack_sizes(texts, offsets.ctypes.data)
        data = np.empty(total + 16, dtype=np.uint8)
        ph.pack_copy(texts, data.ctypes.data)
        return data[:total], offsets

    def encode_packed(self, data, offsets: np.ndarray, with_special: bool = False, return_stats: bool = False, _as_lists: bool = False):
        """Zero-copy surface: `data` = concatenated UTF-8 (bytes / bytearray / uint8 array /
        integer host address), `offsets` = uint64[n_docs+1].  Returns (ids uint32[n_tokens],
        out_offsets uint64[n_docs+1]) as numpy arrays."""
        lib = _lib.load()
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n_docs = offsets.shape[0] - 1
        if n_docs < 0:
            raise ValueError("offsets must have n_docs + 1 entries")
        n_avail = None
        if isinstance(data, (bytes, bytearray)):
            keep = bytes(data)
            n_avail = len(keep)
            ptr = ctypes.cast(ctypes.c_char_p(keep), ctypes.c_void_p)
        elif isinstance(data, np.ndarray):
            keep = np.ascontiguousarray(data, dtype=np.uint8)
            n_avail = int(keep.size)
            ptr = ctypes.c_void_p(keep.ctypes.data)
        elif isinstance(data, int):
            keep = None
            ptr = ctypes.c_void_p(data)
        else:
            raise TypeError("data must be bytes, bytearray, numpy uint8 array or an address")
        if int(offsets[0]) != 0:
            raise ValueError("offsets must start at 0")
        if n_avail is not None and int(offsets[-1]) > n_avail:
            raise ValueError("offsets[-1] exceeds the length of data")
        res = ctypes.c_void_p()
        with self._lock:
            rc = lib.spl_encode_batch(self._handle, ptr, ctypes.c_void_p(offsets.ctypes.data), n_docs,
                                      _lib.SPL_ENCODE_WITH_SPECIAL if with_special else 0, ctypes.byref(res))
            msg = _lib.last_error(self._handle) if rc != _lib.SPL_OK else ""
        del keep
        if rc != _lib.SPL_OK:
            if rc in (_lib.SPL_ERR_INVALID_ARG, _lib.SPL_ERR_UNSUPPORTED):
                raise ValueError(msg)
            raise RuntimeError(f"splintr_b200: {msg} (code {rc})")
        try:
            n_tok = lib.spl_result_n_tokens(res)
            if _as_lists:                              # list[list[int]] straight from the pinned result (no copy in between)
                ph = _lib.pyhost()
                if ph is not None:
                    return ph.ids_to_lists(int(lib.spl_result_ids(res) or 0), int(lib.spl_result_offsets(res)), n_docs)
            ids = np.empty(n_tok, dtype=np.uint32)
            out_off = np.empty(n_docs + 1, dtype=np.uint64)
            if n_tok:
                ctypes.memmove(ids.ctypes.data, lib.spl_result_ids(res), n_tok * 4)
            ctypes.memmove(out_off.ctypes.data, lib.spl_result_offsets(res), (n_docs + 1) * 8)
            if return_stats:
                st = _lib.SplStats()
                lib.spl_result_stats(res, ctypes.byref(st))
                stats = {f: getattr(st, f) for f, _ in _lib.SplStats._fields_}
                return ids, out_off, stats
            if _as_lists:
                flat, o = ids.tolist(), out_off.tolist()
                return [flat[o[i]:o[i + 1]] for i in range(n_docs)]
            return ids, out_off
        finally:
            lib.spl_result_free(res)

    def encode_device(self, d_bytes, d_offsets, with_special: bool = False, ids_out=None, out_offsets=None,
                      dev_index: int = 0, sync: bool = True):
        """Device-resident surface (spl_encode_batch_device): `d_bytes` = CUDA uint8 tensor whose
        storage is 16-byte aligned and padded to a multiple of 16, `d_offsets` = CUDA int64
        tensor [n_docs+1] (non-negative, so bit-identical to the ABI's uint64).  Work is enqueued
        on torch's current stream.  Returns (ids int32[capacity], out_offsets int64[n_docs+1],
        n_tokens or None when sync=False); ids[:n_tokens] are valid.  With sync=False nothing has been checked yet:
        call device_status() once the stream has run (0 = good)."""
        import torch
        lib = _lib.load()
        n_bytes = int(d_bytes.numel())
        n_docs = int(d_offsets.numel()) - 1
        if d_bytes.dtype != torch.uint8 or d_offsets.dtype != torch.int64 or not d_bytes.is_cuda or not d_offsets.is_cuda:
            raise TypeError("d_bytes must be a CUDA uint8 tensor and d_offsets a CUDA int64 tensor")
        if ids_out is None:                        # SentencePiece mode: a space becomes the three bytes of U+2581
            ids_out = torch.empty(max(n_bytes * (3 if self._sentencepiece else 1), 1), dtype=torch.int32,
                                  device=d_bytes.device)
        if out_offsets is None:
            out_offsets = torch.empty(n_docs + 1, dtype=torch.int64, device=d_bytes.device)
        n_tok = ctypes.c_uint64(0)
        stream = torch.cuda.current_stream(d_bytes.device).cuda_stream
        with self._lock:
            rc = lib.spl_encode_batch_device(self._handle, dev_index, ctypes.c_void_p(d_bytes.data_ptr()), n_bytes,
                                             ctypes.c_void_p(d_offsets.data_ptr()), n_docs,
                                             _lib.SPL_ENCODE_WITH_SPECIAL if with_special else 0,
                                             ctypes.c_void_p(ids_out.data_ptr()), int(ids_out.numel()),
                                             ctypes.c_void_p(out_offsets.data_ptr()), ctypes.c_void_p(stream),
                                             ctypes.byref(n_tok) if sync else None)
            msg = _lib.last_error(self._handle) if rc != _lib.SPL_OK else ""
        if rc != _lib.SPL_OK:
            if rc in (_lib.SPL_ERR_INVALID_ARG, _lib.SPL_ERR_UNSUPPORTED):
                raise ValueError(msg)
            raise RuntimeError(f"splintr_b200: {msg} (code {rc})")
        return ids_out, out_offsets, (int(n_tok.value) if sync else None)

    def device_status(self, dev_index: int = 0) -> int:
        """spl_device_status: error flags of the last encode_device(sync=False) pass on that device (synchronises
        torch's current stream).  0 = good, 1 = the document offsets were rejected, 2 = scratch for pieces beyond
        1 KiB exhausted (run the call again with sync=True)."""
        import torch
        flags = ctypes.c_uint32(0)
        with self._lock:
            stream = torch.cuda.current_stream().cuda_stream
            rc = _lib.load().spl_device_status(self._handle, dev_index, ctypes.c_void_p(stream), ctypes.byref(flags))
        if rc != _lib.SPL_OK:
            raise RuntimeError(f"splintr_b200: spl_device_status failed (code {rc})")
        return int(flags.value)

    def debug_counters(self, dev_index: int = 0) -> Dict[str, int]:
        """spl_debug_counters: what the last device pass did (duplicates skipped, tiles refined, misses per length class)."""
        import torch
        out = (ctypes.c_uint32 * 32)()
        with self._lock:
            stream = torch.cuda.current_stream().cuda_stream
            rc = _lib.load().spl_debug_counters(self._handle, dev_index, ctypes.c_void_p(stream), out)
        if rc != _lib.SPL_OK:
            raise RuntimeError(f"splintr_b200: spl_debug_counters failed (code {rc})")
        return {"error_flags": out[1], "fallback_tiles": out[4], "duplicates": out[5], "refined_tiles": out[6], "settled_multi_id_chars": out[7],
                "misses_by_class": [out[8 + c] for c in range(8)]}

    # -- ingestion: JSON Lines -> packed text + offsets on the device ---------------
    def ingest_jsonl_device(self, d_jsonl, field: str = "text", dev_index: int = 0):
        """spl_ingest_jsonl_device: `d_jsonl` = the bytes of a JSON Lines file as a CUDA uint8 tensor (16-byte aligned
        storage, padded to a multiple of 16).  Returns (text uint8[n_text_bytes] -- a view of a 16-byte padded buffer,
        ready for encode_device --, offsets int64[n_docs + 1], stats dict).  One document per non-blank line: the
        string value of its last top-level member `field`; see include/splintr_b200.h for the full contract."""
        import torch
        lib = _lib.load()
        if d_jsonl.dtype != torch.uint8 or not d_jsonl.is_cuda:
            raise TypeError("d_jsonl must be a CUDA uint8 tensor")
        n = int(d_jsonl.numel())
        n_lines_max = int((d_jsonl == 10).sum().item()) + 2 if n else 2
        text = torch.empty(n + ((-n) % 16) + 16, dtype=torch.uint8, device=d_jsonl.device)
        offs = torch.empty(n_lines_max, dtype=torch.int64, device=d_jsonl.device)
        st = _lib.SplIngestStats()
        stream = torch.cuda.current_stream(d_jsonl.device).cuda_stream
        with self._lock:
            rc = lib.spl_ingest_jsonl_device(self._handle, dev_index, ctypes.c_void_p(d_jsonl.data_ptr()), n, field.encode("utf-8"),
                                             ctypes.c_void_p(text.data_ptr()), n, ctypes.c_void_p(offs.data_ptr()), n_lines_max,
                                             ctypes.c_void_p(stream), ctypes.byref(st))
            msg = _lib.last_error(self._handle) if rc != _lib.SPL_OK else ""
        if rc != _lib.SPL_OK:
            if rc in (_lib.SPL_ERR_INVALID_ARG, _lib.SPL_ERR_UNSUPPORTED):
                raise ValueError(msg)
            raise RuntimeError(f"splintr_b200: {msg} (code {rc})")
        stats = {f: getattr(st, f) for f, _ in _lib.SplIngestStats._fields_}
        return text[:int(st.n_text_bytes)], offs[:int(st.n_docs) + 1], stats

    @staticmethod
    def _host_bytes(data):
        """bytes / bytearray / numpy uint8 / (address, nbytes) / a path -> (object to keep alive, void pointer, size)"""
        if isinstance(data, tuple):
            addr, n = data
            return None, ctypes.c_void_p(addr), int(n)
        if isinstance(data, (str, os.PathLike)):
            data = np.fromfile(data, dtype=np.uint8)
        if isinstance(data, (bytes, bytearray)):
            keep = bytes(data)
            return keep, ctypes.cast(ctypes.c_char_p(keep), ctypes.c_void_p), len(keep)
        keep = np.ascontiguousarray(data, dtype=np.uint8)
        return keep, ctypes.c_void_p(keep.ctypes.data), int(keep.shape[0])

    def _encode_file(self, fn_name: str, data, member: str, with_special: bool, return_stats: bool):
        lib = _lib.load()
        keep, ptr, n = self._host_bytes(data)
        res = ctypes.c_void_p()
        ist = _lib.SplIngestStats()
        with self._lock:
            rc = getattr(lib, fn_name)(self._handle, ptr, n, member.encode("utf-8"),
                                       _lib.SPL_ENCODE_WITH_SPECIAL if with_special else 0, ctypes.byref(res), ctypes.byref(ist))
            msg = _lib.last_error(self._handle) if rc != _lib.SPL_OK else ""
        del keep
        if rc != _lib.SPL_OK:
            if rc in (_lib.SPL_ERR_INVALID_ARG, _lib.SPL_ERR_UNSUPPORTED):
                raise ValueError(msg)
            raise RuntimeError(f"splintr_b200: {msg} (code {rc})")
        try:
            n_tok, n_docs = lib.spl_result_n_tokens(res), lib.spl_result_n_docs(res)
            ids = np.empty(n_tok, dtype=np.uint32)
            out_off = np.empty(n_docs + 1, dtype=np.uint64)
            if n_tok:
                ctypes.memmove(ids.ctypes.data, lib.spl_result_ids(res), n_tok * 4)
            ctypes.memmove(out_off.ctypes.data, lib.spl_result_offsets(res), (n_docs + 1) * 8)
            if return_stats:
                st = _lib.SplStats()
                lib.spl_result_stats(res, ctypes.byref(st))
                stats = {f: getattr(st, f) for f, _ in _lib.SplStats._fields_}
                stats.update({f: getattr(ist, f) for f, _ in _lib.SplIngestStats._fields_ if f != "n_launches"})
                return ids, out_off, stats
            return ids, out_off
        finally:
            lib.spl_result_free(res)

    def encode_jsonl(self, data, field: str = "text", with_special: bool = False, return_stats: bool = False):
        """spl_encode_jsonl: the bytes of a JSON Lines file (bytes / bytearray / numpy uint8 / a path / integer host
        address folded into a (address, nbytes) tuple) -> (ids uint32, offsets uint64): chunks of the raw file
        are copied to the device, member extraction + unescaping and the encode run there, ids come back --
        the `[json.loads(l)[field] for l in f]` loop and the packing of its result never run on the host."""
        return self._encode_file("spl_encode_jsonl", data, field, with_special, return_stats)

    def encode_parquet(self, data, column: str = "text", with_special: bool = False, return_stats: bool = False):
        """spl_encode_parquet: a Parquet file (bytes / numpy uint8 / a path / (address, nbytes)) -> (ids uint32,
        offsets uint64), one document per row of the string column `column` ("text", or a dotted path into structs);
        a null row is an empty document.  The host only reads the footer and the page headers; the column chunks go
        to the device as they lie in the file and are decompressed (snappy) and decoded there -- the
        `pq.read_table(path)[column].to_pylist()` in front of encode_batch never runs.  ValueError for what the
        reader does not take (other codecs, DELTA encodings, list columns), with what to rewrite."""
        return self._encode_file("spl_encode_parquet", data, column, with_special, return_stats)

    def ingest_parquet(self, data, column: str = "text", dev_index: int = 0):
        """spl_ingest_parquet: the same decoding without the encode -> (text uint8 CUDA tensor -- a view of a 16-byte
        padded buffer, ready for encode_device --, offsets int64[rows + 1] CUDA tensor, stats dict)."""
        import torch
        lib = _lib.load()
        keep, ptr, n = self._host_bytes(data)
        dev = torch.device("cuda", self._devices[dev_index] if self._devices else torch.cuda.current_device())
        stream = torch.cuda.current_stream(dev).cuda_stream
        st = _lib.SplIngestStats()

        def call(text, offs):
            with self._lock:
                rc = lib.spl_ingest_parquet(self._handle, dev_index, ptr, n, column.encode("utf-8"),
                                            ctypes.c_void_p(text.data_ptr()) if text is not None else None, int(text.numel()) - 32 if text is not None else 0,
                                            ctypes.c_void_p(offs.data_ptr()) if offs is not None else None, int(offs.numel()) if offs is not None else 0,
                                            ctypes.c_void_p(stream), ctypes.byref(st))
                msg = _lib.last_error(self._handle) if rc != _lib.SPL_OK else ""
            return rc, msg

        rc, msg = call(None, None)                       # sizes first (a dictionary-encoded column can expand)
        text = offs = None
        if rc == _lib.SPL_OK:
            nt, nr = int(st.n_text_bytes), int(st.n_docs)
            text = torch.empty(nt + ((-nt) % 16) + 32, dtype=torch.uint8, device=dev)
            offs = torch.empty(nr + 1, dtype=torch.int64, device=dev)
            rc, msg = call(text, offs)
        del keep
        if rc != _lib.SPL_OK:
            if rc in (_lib.SPL_ERR_INVALID_ARG, _lib.SPL_ERR_UNSUPPORTED):
                raise ValueError(msg)
            raise RuntimeError(f"splintr_b200: {msg} (code {rc})")
        stats = {f: getattr(st, f) for f, _ in _lib.SplIngestStats._fields_}
        return text[:int(st.n_text_bytes)], offs, stats

    def encode_arrow(self, column, with_special: bool = False):
        """An Arrow string column (pyarrow StringArray / LargeStringArray / ChunkedArray of them) -> (ids uint32,
        offsets uint64), one document per element, a null an empty document.  An Arrow string array IS the packed
        layout of spl_encode_batch -- one data buffer + offsets -- so nothing is copied or packed on the host beyond
        widening 32-bit offsets: the data buffer goes to the device as it is (encode_batch_packed)."""
        import pyarrow as pa
        chunks = column.chunks if isinstance(column, pa.ChunkedArray) else [column]
        ids_parts, off_parts, base = [], [np.zeros(1, dtype=np.uint64)], 0
        for a in chunks:
            if pa.types.is_string(a.type) or pa.types.is_binary(a.type):
                odt = np.int32
            elif pa.types.is_large_string(a.type) or pa.types.is_large_binary(a.type):
                odt = np.int64
            else:
                raise TypeError(f"encode_arrow needs a string column, got {a.type}")
            if len(a) == 0:
                continue
            bufs = a.buffers()
            o = np.frombuffer(bufs[1], dtype=odt)[a.offset:a.offset + len(a) + 1].astype(np.uint64)
            if a.null_count:                                       # a null slot may hold any bytes: make it empty
                lens = np.diff(o)
                lens[~np.asarray(a.is_valid())] = 0
                if int(lens.sum()) != int(o[-1] - o[0]):          # some null slot was not empty: repack those bytes out
                    a = pa.array(["" if v is None else v for v in a.to_pylist()], type=a.type)
                    bufs = a.buffers()
                    o = np.frombuffer(bufs[1], dtype=odt)[:len(a) + 1].astype(np.uint64)
            lo, hi = int(o[0]), int(o[-1])
            data = np.frombuffer(bufs[2], dtype=np.uint8)[lo:hi] if bufs[2] is not None else np.zeros(0, dtype=np.uint8)
            ids, off = self.encode_packed(data, o - np.uint64(lo), with_special)
            ids_parts.append(ids)
            off_parts.append(off[1:] + np.uint64(base))
            base += int(off[-1])
        ids = np.concatenate(ids_parts) if ids_parts else np.zeros(0, dtype=np.uint32)
        return ids, np.concatenate(off_parts)

    def set_profiling(self, enable: bool = True) -> None:
        _lib.load().spl_set_profiling(self._handle, int(enable))

    def last_kernel_times(self, dev_index: int = 0) -> Dict[str, float]:
        """{kernel name: ms} of the most recent encode_device call (after a stream sync)."""
        names = (ctypes.c_char_p * 16)()
        ms = (ctypes.c_float * 16)()
        n = _lib.load().spl_last_kernel_times(self._handle, dev_index, names, ms, 16)
        return {names[i].decode(): float(ms[i]) for i in range(max(n, 0))}

    def launches_per_call(self, with_special: bool = False) -> int:
        return int(_lib.load().spl_launches_per_call(self._handle, _lib.SPL_ENCODE_WITH_SPECIAL if with_special else 0))

    def encode_batch_packed(self, texts: Sequence[str], with_special: bool = False):
        """list[str] -> (ids uint32, offsets uint64) without building Python lists."""
        data, offsets = self._pack(texts)
        return self.encode_packed(data, offsets, with_special)

    def _encode_many(self, texts: Sequence[str], with_special: bool) -> List[List[int]]:
        data, offsets = self._pack(texts)
        return self.encode_packed(data, offsets, with_special, _as_lists=True)

    def encode(self, text: str) -> List[int]:
        """bindings.rs:254-256 -> tokenizer.rs:729-808 (special strings are plain text)."""
        return self._encode_many([text], False)[0]

    def encode_rayon(self, text: str) -> List[int]:
        """bindings.rs:273-275: same ids as encode (the device path is parallel inside a text)."""
        return self._encode_many([text], False)[0]

    def encode_with_special(self, text: str) -> List[int]:
        """bindings.rs:286-288 -> tokenizer.rs:842-874."""
        return self._encode_many([text], True)[0]

    def encode_batch(self, texts: List[str]) -> List[List[int]]:
        """bindings.rs:337-339 -> tokenizer.rs:932-934."""
        return self._encode_many(list(texts), False)

    def encode_batch_with_special(self, texts: List[str]) -> List[List[int]]:
        """bindings.rs:348-350 -> tokenizer.rs:937-942."""
        return self._encode_many(list(texts), True)

    # -- decode (host table lookup, tokenizer.rs:877-958) ----------------------------------
    def _ensure_decoder(self) -> Dict[int, bytes]:
        if self._decoder is None and self._sentencepiece:
            dec = _parse_tiktoken_decoder(self._vocab_data)          # vocab.rs:135
            for k, v in self._special_tokens.items():                # tokenizer.rs:597-600
                dec[v] = k.encode("utf-8")
            self._decoder = dec
        if self._decoder is None:
            enc = _parse_tiktoken(self._vocab_data)
            dec: Dict[int, bytes] = {}
            for k, v in enc.items():
                if self._byte_level:
                    # tokenizer.rs:882-887: byte-level keys decode to raw bytes, else stay as they are
                    try:
                        raw = bytes(_CHAR_TO_BYTE[c] for c in k.decode("utf-8"))
                    except (UnicodeDecodeError, KeyError):
                        raw = k
                    dec[v] = raw
                else:
                    dec[v] = k
            self._decoder = dec
        return self._decoder

    def decode_bytes(self, tokens: Iterable[int]) -> bytes:
        dec = self._ensure_decoder()
        out = bytearray()
        for t in tokens:
            b = dec.get(t)
            if b is not None:
                out += b
            else:
                s = self._special_decoder.get(t)
                if s is not None:
                    out += s.encode("utf-8")
        return bytes(out)

    def _postprocess(self, text: str) -> str:
        """tokenizer.rs:923-930: SentencePiece mode turns U+2581 back into a space (string level, after UTF-8)."""
        return text.replace("\u2581", " ") if self._sentencepiece else text

    def decode(self, tokens: Iterable[int]) -> str:
        try:
            return self._postprocess(self.decode_bytes(tokens).decode("utf-8"))
        except UnicodeDecodeError:
            raise ValueError("Decoding error: invalid UTF-8")

    def decode_lossy(self, tokens: Iterable[int]) -> str:
        return self._postprocess(self.decode_bytes(tokens).decode("utf-8", errors="replace"))

    def decode_packed(self, ids, offsets, return_stats: bool = False):
        """Batch decode on the device (spl_decode_batch; tokenizer.rs:877-897, 945-958): `ids` = concatenated
        token ids (uint32), `offsets` = uint64[n_docs+1] token offsets.  Returns (bytes uint8[n_bytes],
        byte_offsets uint64[n_docs+1]); UTF-8 validation is the caller's (decode_batch below does it)."""
        lib = _lib.load()
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n_docs = offsets.shape[0] - 1
        if n_docs < 0:
            raise ValueError("offsets must have n_docs + 1 entries")
        if int(offsets[-1]) != ids.shape[0]:
            raise ValueError("offsets[-1] must equal the number of ids")
        res = ctypes.c_void_p()
        with self._lock:
            rc = lib.spl_decode_batch(self._handle, ctypes.c_void_p(ids.ctypes.data), ctypes.c_void_p(offsets.ctypes.data),
                                      n_docs, ctypes.byref(res))
            msg = _lib.last_error(self._handle) if rc != _lib.SPL_OK else ""
        if rc != _lib.SPL_OK:
            if rc in (_lib.SPL_ERR_INVALID_ARG, _lib.SPL_ERR_UNSUPPORTED):
                raise ValueError(msg)
            raise RuntimeError(f"splintr_b200: {msg} (code {rc})")
        try:
            nb = lib.spl_result_n_bytes(res)
            out = np.empty(nb, dtype=np.uint8)
            out_off = np.empty(n_docs + 1, dtype=np.uint64)
            if nb:
                ctypes.memmove(out.ctypes.data, lib.spl_result_bytes(res), nb)
            ctypes.memmove(out_off.ctypes.data, lib.spl_result_offsets(res), (n_docs + 1) * 8)
            if return_stats:
                st = _lib.SplStats()
                lib.spl_result_stats(res, ctypes.byref(st))
                return out, out_off, {f: getattr(st, f) for f, _ in _lib.SplStats._fields_}
            return out, out_off
        finally:
            lib.spl_result_free(res)

    def decode_device(self, d_ids, d_tok_offsets, dev_index: int = 0):
        """Device-resident decode (spl_decode_batch_device): CUDA int32/uint32 ids tensor and CUDA int64 token
        offsets [n_docs+1] -> (bytes uint8 tensor[n_bytes], byte offsets int64[n_docs+1]) on the same device."""
        import torch
        lib = _lib.load()
        n_tok = int(d_ids.numel())
        n_docs = int(d_tok_offsets.numel()) - 1
        if d_ids.dtype not in (torch.int32, torch.uint32) or d_tok_offsets.dtype != torch.int64 or not d_tok_offsets.is_cuda:
            raise TypeError("d_ids must be a CUDA int32 tensor and d_tok_offsets a CUDA int64 tensor")
        out_off = torch.empty(n_docs + 1, dtype=torch.int64, device=d_tok_offsets.device)
        stream = torch.cuda.current_stream(d_tok_offsets.device).cuda_stream
        cap = max(n_tok * 4, 16)
        for _ in range(2):
            out = torch.empty(cap, dtype=torch.uint8, device=d_tok_offsets.device)
            nb = ctypes.c_uint64(0)
            with self._lock:
                rc = lib.spl_decode_batch_device(self._handle, dev_index, ctypes.c_void_p(d_ids.data_ptr()), n_tok,
                                                 ctypes.c_void_p(d_tok_offsets.data_ptr()), n_docs,
                                                 ctypes.c_void_p(out.data_ptr()), cap, ctypes.c_void_p(out_off.data_ptr()),
                                                 ctypes.c_void_p(stream), ctypes.byref(nb))
                msg = _lib.last_error(self._handle) if rc != _lib.SPL_OK else ""
            if rc == _lib.SPL_OK:
                return out[:int(nb.value)], out_off
            if rc == _lib.SPL_ERR_INVALID_ARG and int(nb.value) > cap:
                cap = int(nb.value)
                continue
            raise RuntimeError(f"splintr_b200: {msg} (code {rc})")
        raise RuntimeError("splintr_b200: decode capacity retry failed")

    def _decode_many(self, token_lists, errors: str) -> List[str]:
        n = len(token_lists)
        offsets = np.zeros(n + 1, dtype=np.uint64)
        if n:
            np.cumsum(np.fromiter((len(t) for t in token_lists), dtype=np.uint64, count=n), out=offsets[1:])
        total = int(offsets[-1])
        ids = np.fromiter((x for t in token_lists for x in t), dtype=np.int64, count=total)
        if total and (ids.min() < 0 or ids.max() > 0xFFFFFFFF):
            raise OverflowError("token ids must fit in u32")
        data, off = self.decode_packed(ids.astype(np.uint32), offsets)
        raw = data.tobytes()
        out = []
        for i in range(n):
            b = raw[int(off[i]):int(off[i + 1])]
            if errors == "strict":
                try:
                    out.append(self._postprocess(b.decode("utf-8")))
                except UnicodeDecodeError:
                    raise ValueError("Decoding error: invalid UTF-8")
            else:
                out.append(self._postprocess(b.decode("utf-8", errors="replace")))
        return out

    def decode_batch(self, token_lists: List[List[int]]) -> List[str]:
        """bindings.rs:364-370 -> tokenizer.rs:945-950: on the device (one call for the whole batch)."""
        return self._decode_many(token_lists, "strict")

    def decode_batch_lossy(self, token_lists: List[List[int]]) -> List[str]:
        return self._decode_many(token_lists, "replace")

    # -- misc -------------------------------------------------------------------------------
    @property
    def vocab_size(self) -> int:
        """tokenizer.rs:964-972: max id over vocabulary and special tokens, plus one."""
        if self._vocab_size is None:
            dec = self._ensure_decoder()
            m1 = max(dec.keys(), default=0)
            m2 = max(self._special_tokens.values(), default=0)
            self._vocab_size = max(m1, m2) + 1
        return self._vocab_size

    def streaming_decoder(self):
        from .streaming import StreamingDecoder
        return StreamingDecoder(self._raw_decoder(), dict(self._special_decoder))

    def byte_level_streaming_decoder(self):
        from .streaming import ByteLevelStreamingDecoder
        return ByteLevelStreamingDecoder(self._raw_decoder(), dict(self._special_decoder))

    def _raw_decoder(self) -> Dict[int, bytes]:
        if self._sentencepiece:
            return _parse_tiktoken_decoder(self._vocab_data)
        return {v: k for k, v in _parse_tiktoken(self._vocab_data).items()}

    def clear_cache(self) -> None:
        """bindings.rs:432-434: the device path keeps no chunk cache (it is result-transparent)."""

    @property
    def cache_len(self) -> int:
        return 0

    def __repr__(self) -> str:
        return f"Tokenizer(vocab_size={self.vocab_size})"
