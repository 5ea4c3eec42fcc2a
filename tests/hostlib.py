"""Builds tests/csrc/hosttest.cpp: the product's __host__ __device__ pre-tokenizer rules and
host table builder compiled for the CPU (tests only), so the exact device logic can be
fuzzed against the oracle without a GPU."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "tests", "csrc", "hosttest.cpp"), os.path.join(ROOT, "splintr_b200", "csrc", "spl_host.cpp"),
       os.path.join(ROOT, "splintr_b200", "csrc", "spl_parquet_meta.cpp")]
DEPS = SRC + [os.path.join(ROOT, "splintr_b200", "csrc", f) for f in ("spl_pretok.h", "spl_pretok_fast.h", "spl_sentencepiece.h", "spl_ingest.h", "spl_bpe_bits.h", "spl_segment.h", "spl_special.h", "spl_parquet.h", "spl_parquet_meta.h", "spl_common.h", "spl_host.h", "unicode_tables.inc")]
LIB = os.path.join(ROOT, "tests", "csrc", "libhosttest.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", LIB] + SRC
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode:
            raise RuntimeError(p.stdout + p.stderr)
    lib = ctypes.CDLL(LIB)
    vp = ctypes.c_void_p
    lib.ht_scan_seq.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_uint32, vp]
    lib.ht_scan_chunked.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, vp, vp, vp]
    lib.ht_scan_kernel_emul.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                        ctypes.c_uint32, vp, vp, vp]
    lib.ht_scan_fast.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, vp, vp, vp, vp]
    lib.ht_create.restype = vp
    lib.ht_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32,
                              ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_uint32), ctypes.c_size_t,
                              ctypes.c_char_p, ctypes.c_size_t]
    lib.ht_destroy.argtypes = [vp]
    lib.ht_stats.argtypes = [vp, vp]
    lib.ht_encode.restype = ctypes.c_long
    lib.ht_encode.argtypes = [vp, ctypes.c_char_p, ctypes.c_uint32, vp, ctypes.c_size_t]
    lib.ht_encode_piece.restype = ctypes.c_long
    lib.ht_encode_piece.argtypes = [vp, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_int, vp, ctypes.c_size_t]
    lib.ht_seg_counters.argtypes = [vp, ctypes.c_int]
    lib.ht_set_tile_limit.argtypes = [ctypes.c_uint32]
    lib.ht_special_walk.restype = ctypes.c_long
    lib.ht_special_walk.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, vp, ctypes.c_uint32, vp, ctypes.c_size_t]
    lib.ht_jsonl.restype = ctypes.c_int
    lib.ht_jsonl.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, vp, vp, vp]
    lib.ht_parquet.restype = ctypes.c_long
    lib.ht_parquet.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_uint64, vp, ctypes.c_size_t, vp, ctypes.c_size_t,
                               ctypes.c_char_p, ctypes.c_size_t, vp, ctypes.c_int]
    lib.ht_snappy.restype = ctypes.c_int
    lib.ht_snappy.argtypes = [ctypes.c_char_p, ctypes.c_uint32, vp, ctypes.c_uint32, ctypes.c_int]
    lib.ht_set_fast_ext.argtypes = [ctypes.c_int]
    lib.ht_sp_transform.restype = ctypes.c_long
    lib.ht_sp_transform.argtypes = [ctypes.c_char_p, ctypes.c_uint32, vp, vp, vp, vp, vp]
    lib.ht_encode_sp.restype = ctypes.c_long
    lib.ht_encode_sp.argtypes = [vp, ctypes.c_char_p, ctypes.c_uint32, vp, ctypes.c_size_t]
    lib.ht_bpe_window_m.restype = ctypes.c_int
    lib.ht_bpe_window_m.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, vp]
    _lib = lib
    return lib


def special_walk(text: bytes, specials):
    """spl_special.h over one document: [(start, end, index into specials)] as aho-corasick's Standard non-overlapping
    find_iter reports them."""
    strs = b"".join(specials)
    off = np.cumsum([0] + [len(x) for x in specials]).astype(np.uint32)
    out = np.zeros(3 * (len(text) + 1), dtype=np.uint32)
    n = load().ht_special_walk(text, len(text), strs, off.ctypes.data, len(specials), out.ctypes.data, len(out))
    return [(int(out[3 * i]), int(out[3 * i + 1]), int(out[3 * i + 2])) for i in range(n)]


def set_tile_limit(limit: int = 0xFFFFFFFF):
    """boundaries at or beyond this byte of a piece are not looked for (the device refines a piece only inside its tile)"""
    load().ht_set_tile_limit(limit)


def seg_counters(reset: bool = True):
    """(segments, single-character segments, segments beyond SPL_SEG_MAX, safe boundaries) since the last reset."""
    out = np.zeros(4, dtype=np.uint64)
    load().ht_seg_counters(out.ctypes.data, 1 if reset else 0)
    return [int(x) for x in out]


def bpe_window_m(ranks, G: int, B: int):
    """spl_bpe_bits.h over one group: ranks of the pairs (0x1FFFFF = none) -> (m flag per pair, boundary passes)."""
    k = np.ascontiguousarray(ranks, dtype=np.uint32)
    m = np.zeros(32, dtype=np.uint32)
    rc = load().ht_bpe_window_m(k.ctypes.data, len(k), G, B, m.ctypes.data)
    assert rc > 0, rc
    return [bool((int(m[e // B]) >> (e % B)) & 1) for e in range(len(k))], rc


def jsonl(data: bytes, field: str = "text"):
    """spl_ingest.h line parser over a whole JSON Lines buffer -> (list of document bytes, missing, bad)."""
    n = len(data)
    out = np.zeros(n + 16, dtype=np.uint8)
    off = np.zeros(n + 2, dtype=np.uint64)
    cnt = np.zeros(4, dtype=np.uint64)
    rc = load().ht_jsonl(data, n, field.encode(), out.ctypes.data, off.ctypes.data, cnt.ctypes.data)
    assert rc == 0, rc
    nd = int(cnt[0])
    raw = out[:int(cnt[1])].tobytes()
    o = off[:nd + 1].tolist()
    return [raw[o[i]:o[i + 1]] for i in range(nd)], int(cnt[2]), int(cnt[3])


def snappy(stream: bytes, size: int, staged: bool):
    """spl_parquet.h snappy decoders on a raw stream -> bytes, or None if the stream is refused"""
    out = np.zeros(size + 16, dtype=np.uint8)
    ok = load().ht_snappy(stream, len(stream), out.ctypes.data, size, int(staged))
    assert ok >= 0, "the lanes of the warp decoder disagree on the outcome"
    return out[:size].tobytes() if ok else None


class ParquetError(Exception):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def parquet(data: bytes, column: str = "text", batch_bytes: int = 0, text_cap: int = 0, max_rows: int = 0, staged: bool = True):
    """spl_parquet_meta.cpp + the page decoder of spl_parquet.h over a whole Parquet file -> (list of row bytes, info).
    text_cap / max_rows: capacities of the output (defaults are generous guesses; dictionary pages can expand).
    staged: True / 1 the one-lane snappy decoder that works out of a 64 KiB ring + input slots, False / 0 the plain
    one, 2 the warp-wide decoder the device runs (32 host threads in lock step: slow, keep the files small)."""
    n = len(data)
    text_cap = text_cap or max(64 * n, 1 << 20)
    max_rows = max_rows or max(8 * n, 1 << 16)
    out = np.zeros(text_cap + 16, dtype=np.uint8)
    off = np.zeros(max_rows + 2, dtype=np.uint64)
    info = np.zeros(4, dtype=np.uint64)
    err = ctypes.create_string_buffer(512)
    rc = load().ht_parquet(data, n, column.encode(), batch_bytes, out.ctypes.data, text_cap, off.ctypes.data, max_rows + 1, err, 512, info.ctypes.data, int(staged))
    if rc < 0:
        raise ParquetError(rc, err.value.decode() or f"page error bits {int(info[0])}")
    o = off[:rc + 1].tolist()
    raw = out[:o[-1]].tobytes()
    return [raw[o[i]:o[i + 1]] for i in range(rc)], {"batches": int(info[1]), "pages": int(info[2])}


def sp_transform(data: bytes, hard=None, spec=None):
    """SentencePiece-mode transform (spl_sentencepiece.h rules): returns (T' bytes, piece starts in T', special-span
    starts in T')."""
    n = len(data)
    h = np.zeros(n + 1, dtype=np.uint8) if hard is None else np.asarray(hard, dtype=np.uint8).copy()
    h[0] = 1
    h[n] = 1
    sp = None if spec is None else np.asarray(spec, dtype=np.uint8)
    t2 = np.zeros(3 * n + 4, dtype=np.uint8)
    ps = np.zeros(3 * n + 1, dtype=np.uint8)
    s2 = np.zeros(3 * n + 1, dtype=np.uint8)
    n2 = load().ht_sp_transform(data, n, h.ctypes.data, None if sp is None else sp.ctypes.data,
                                t2.ctypes.data, ps.ctypes.data, s2.ctypes.data)
    return t2[:n2].tobytes(), np.flatnonzero(ps[:n2]).tolist(), np.flatnonzero(s2[:n2]).tolist()


def scan_seq(pattern_id, data: bytes):
    st = np.zeros(len(data) + 1, dtype=np.uint8)
    rc = load().ht_scan_seq(pattern_id, data, len(data), st.ctypes.data)
    assert rc == 0, rc
    return np.flatnonzero(st[:len(data)]).tolist()


def scan_chunked(pattern_id, data: bytes, chunk, hard=None):
    n = len(data)
    h = np.zeros(n + 1, dtype=np.uint8) if hard is None else np.asarray(hard, dtype=np.uint8)
    h[0] = 1
    h[n] = 1
    st = np.zeros(n + 1, dtype=np.uint8)
    ns = ctypes.c_uint32(0)
    rc = load().ht_scan_chunked(pattern_id, data, n, chunk, h.ctypes.data, st.ctypes.data, ctypes.addressof(ns))
    assert rc == 0, rc
    return np.flatnonzero(st[:n]).tolist()


def scan_kernel_emul(pattern_id, data: bytes, tile, halo, chunk, hard, spec=None):
    n = len(data)
    h = np.asarray(hard, dtype=np.uint8)
    st = np.zeros(n + 1, dtype=np.uint8)
    sp = None if spec is None else np.asarray(spec, dtype=np.uint8)
    rc = load().ht_scan_kernel_emul(pattern_id, data, n, tile, halo, chunk, h.ctypes.data,
                                    None if sp is None else sp.ctypes.data, st.ctypes.data)
    assert rc == 0, rc
    return np.flatnonzero(st[:n]).tolist()


def scan_fast_ext(pattern_id, data: bytes, payload, halo, ext, hard, spec=None):
    """The fused kernel's use of the bit-parallel path: the first `ext` right-halo words of a decided tile are trusted
    too.  Returns (raw flag array: 1 fast start, 2 fallback start, 4 start reported by an ext word, 8 covered by an
    ext word; per-tile fallback flags)."""
    n = len(data)
    h = np.asarray(hard, dtype=np.uint8)
    st = np.zeros(n + 1, dtype=np.uint8)
    n_tiles = max(1, (n + payload * 32 - 1) // (payload * 32))
    flags = np.zeros(n_tiles, dtype=np.uint8)
    sp = None if spec is None else np.asarray(spec, dtype=np.uint8)
    lib = load()
    lib.ht_set_fast_ext(ext)
    try:
        rc = lib.ht_scan_fast(pattern_id, data, n, payload, halo, h.ctypes.data,
                              None if sp is None else sp.ctypes.data, st.ctypes.data, flags.ctypes.data)
    finally:
        lib.ht_set_fast_ext(0)
    assert rc == 0, rc
    return st[:n], flags.tolist()


def scan_fast(pattern_id, data: bytes, payload, halo, hard, spec=None):
    """Bit-parallel path under a (payload, halo)-word tiling.  Returns (starts list, per-tile fallback flags)."""
    n = len(data)
    h = np.asarray(hard, dtype=np.uint8)
    st = np.zeros(n + 1, dtype=np.uint8)
    n_tiles = max(1, (n + payload * 32 - 1) // (payload * 32))
    flags = np.zeros(n_tiles, dtype=np.uint8)
    sp = None if spec is None else np.asarray(spec, dtype=np.uint8)
    rc = load().ht_scan_fast(pattern_id, data, n, payload, halo, h.ctypes.data,
                             None if sp is None else sp.ctypes.data, st.ctypes.data, flags.ctypes.data)
    assert rc == 0, rc
    return np.flatnonzero(st[:n] & 1).tolist(), flags.tolist(), np.flatnonzero(st[:n]).tolist()


class HostTables:
    def __init__(self, vocab: bytes, pattern_id: int, byte_level: bool, specials=None, sentencepiece: bool = False):
        sp = list((specials or {}).items())
        n = len(sp)
        strs = (ctypes.c_char_p * max(n, 1))(*[s.encode() for s, _ in sp])
        ids = (ctypes.c_uint32 * max(n, 1))(*[i for _, i in sp])
        err = ctypes.create_string_buffer(256)
        self.h = load().ht_create(vocab, len(vocab), pattern_id, (1 if byte_level else 0) | (2 if sentencepiece else 0), strs, ids, n, err, 256)
        if not self.h:
            raise ValueError(err.value.decode())

    def stats(self):
        out = np.zeros(14, dtype=np.uint64)
        load().ht_stats(self.h, out.ctypes.data)
        return dict(zip(["n_keys", "n_pairs", "t8_log2", "t16_log2", "tl_log2", "pair_log2", "max_key_len", "unambiguous",
                         "t8_displaced", "pair_displaced", "seg_pairs", "seg_h2_log2", "seg_irr_bits", "char_tok_entries"],
                        out.tolist()))

    def encode_piece(self, piece: bytes, segments: bool = True):
        """One piece through the device path's piece encoder (segments=True: whole-piece probe + segment walker;
        False: whole-piece probe + the plain merge loop of bpe.rs:83-194)."""
        ids = np.zeros(len(piece) + 1, dtype=np.uint32)
        n = load().ht_encode_piece(self.h, piece, len(piece), 1 if segments else 0, ids.ctypes.data, len(ids))
        assert n >= 0, n
        return ids[:n].tolist()

    def encode(self, data: bytes):
        ids = np.zeros(len(data) + 1, dtype=np.uint32)
        n = load().ht_encode(self.h, data, len(data), ids.ctypes.data, len(ids))
        assert n >= 0, n
        return ids[:n].tolist()

    def encode_sp(self, data: bytes):
        ids = np.zeros(3 * len(data) + 1, dtype=np.uint32)
        n = load().ht_encode_sp(self.h, data, len(data), ids.ctypes.data, len(ids))
        assert n >= 0, n
        return ids[:n].tolist()

    def __del__(self):
        if getattr(self, "h", None):
            load().ht_destroy(self.h)
            self.h = None
