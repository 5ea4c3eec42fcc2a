"""Seeded synthetic workloads for the BASELINE.json configs.

Every generator returns (data: np.uint8[N], offsets: np.uint64[n_docs+1]) -- the packed
form the C-ABI takes -- and is fully vectorised (a 100 MB batch takes a few seconds).
The same arrays feed the GPU path, the oracle and the CPU baseline.

  cfg1  cl100k_base   1 000 short English-like texts (~100 B)            seed 101
  cfg2  cl100k_base   100 000 English-like docs (~1 KB)                   seed 102
  cfg3  o200k_base    mixed prose / code / JSON docs (~2 KB)              seed 103
  cfg4  llama3        long docs with deep merge chains (pieces <= 1 KiB)  seed 104
  cfg5  deepseek_v3   CJK-heavy docs (~1.5 KB)                            seed 105
"""
from __future__ import annotations

import base64
from typing import List, Sequence, Tuple

import numpy as np

SEEDS = {"cfg1": 101, "cfg2": 102, "cfg3": 103, "cfg4": 104, "cfg5": 105}


# ---------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------
def _pool(atoms: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    lens = np.fromiter((len(a) for a in atoms), dtype=np.int64, count=len(atoms))
    off = np.zeros(len(atoms) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return np.frombuffer(b"".join(atoms), dtype=np.uint8), off[:-1], lens


def _gather(pool: np.ndarray, poff: np.ndarray, plen: np.ndarray, ids: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenate pool atoms `ids`; returns (bytes, start offset of every atom)."""
    lens = plen[ids]
    out_off = np.zeros(len(ids) + 1, dtype=np.int64)
    np.cumsum(lens, out=out_off[1:])
    total = int(out_off[-1])
    src = np.repeat(poff[ids] - out_off[:-1], lens) + np.arange(total, dtype=np.int64)
    return pool[src], out_off


def _zipf_ids(rng: np.random.Generator, n_items: int, n: int, s: float = 1.1) -> np.ndarray:
    w = 1.0 / np.arange(1, n_items + 1, dtype=np.float64) ** s
    cdf = np.cumsum(w / w.sum())
    return np.minimum(np.searchsorted(cdf, rng.random(n)), n_items - 1)


def vocab_words(vocab_data: bytes, max_rank: int = 20000) -> List[bytes]:
    """ASCII-alphabetic, space-prefixed tokens with rank < max_rank, in rank order.
    One-letter fragments (" t", " s", ...) are dropped except the words " a" and " I": the
    lowest-rank BPE entries are word fragments, and keeping them would make the Zipf head
    un-English (3.1 B/token instead of ~4.5)."""
    out = []
    for line in vocab_data.split(b"\n"):
        if not line:
            continue
        sp = line.rfind(b" ")
        rank = int(line[sp + 1:])
        if rank >= max_rank:
            continue
        tok = base64.b64decode(line[:sp])
        if len(tok) >= 2 and tok[:1] == b" " and tok[1:].isalpha() and tok[1:].isascii():
            if len(tok) == 2 and tok not in (b" a", b" I"):
                continue
            out.append((rank, tok))
    out.sort()
    return [t for _, t in out]


def _split_docs(atom_off: np.ndarray, targets: np.ndarray) -> np.ndarray:
    """Document boundaries at atom boundaries, document k ends once the running byte count
    reaches the k-th cumulative target."""
    cuts = np.searchsorted(atom_off, np.cumsum(targets), side="left")
    cuts = np.minimum(cuts, len(atom_off) - 1)
    cuts = np.maximum.accumulate(cuts)
    offs = np.concatenate([[0], atom_off[cuts]]).astype(np.uint64)
    return offs


# ---------------------------------------------------------------------------------------
# English-like prose (cfg1, cfg2 and the prose third of cfg3)
# ---------------------------------------------------------------------------------------
_SUFFIXES = [b"", b",", b".", b"!", b"?", b";", b"'s", b"'t", b"'re", b"'ll"]


def prose_stream(rng: np.random.Generator, words: List[bytes], n_bytes: int, rich: bool) -> Tuple[np.ndarray, np.ndarray]:
    """~n_bytes of word atoms.  rich=False: words (10 % capitalised, 8 % + punctuation).
    rich=True adds paragraph breaks every ~300 B, 2 % numbers, 1 % contractions."""
    n_words = len(words)
    numbers = [b" " + str(int(rng.integers(0, 10 ** int(rng.integers(1, 8))))).encode() for _ in range(4096)]
    atoms = list(words) + numbers + [b"\n\n"]
    pool, poff, plen = _pool(atoms)
    spool, soff, slen = _pool(_SUFFIXES)
    zw = 1.0 / np.arange(1, n_words + 1, dtype=np.float64) ** 1.1
    mean_len = float((plen[:n_words] * zw).sum() / zw.sum()) + 0.1
    n_atoms = int(n_bytes / mean_len * 1.03) + 64
    ids = _zipf_ids(rng, n_words, n_atoms)
    u = rng.random(n_atoms)
    suf = np.zeros(n_atoms, dtype=np.int64)
    punct = u < 0.08
    suf[punct] = rng.integers(1, 6, size=int(punct.sum()))
    cap = rng.random(n_atoms) < 0.10
    if rich:
        v = rng.random(n_atoms)
        is_num = v < 0.02
        is_par = (v >= 0.02) & (v < 0.02 + 5.3 / 300.0)
        is_con = (v >= 0.05) & (v < 0.06)
        ids[is_num] = n_words + rng.integers(0, len(numbers), size=int(is_num.sum()))
        ids[is_par] = n_words + len(numbers)
        suf[is_con] = rng.integers(6, 10, size=int(is_con.sum()))
        suf[is_par] = 0
        cap &= ~(is_num | is_par)
    # interleave word, suffix
    inter = np.empty(2 * n_atoms, dtype=np.int64)
    all_pool = np.concatenate([pool, spool])
    all_off = np.concatenate([poff, soff + len(pool)])
    all_len = np.concatenate([plen, slen])
    inter[0::2] = ids
    inter[1::2] = suf + len(atoms)
    data, off = _gather(all_pool, all_off, all_len, inter)
    data = data.copy()
    word_start = off[0:-1:2]
    cpos = word_start[cap] + 1                     # first letter after the leading space
    data[cpos] -= 32
    return data, off[0::2]                           # atom (word+suffix) start offsets, incl. end


def gen_prose_docs(rng, words, n_docs: int, mean_len: float, sd: float, rich: bool):
    targets = np.maximum(rng.normal(mean_len, sd, size=n_docs), 8.0)
    data, atom_off = prose_stream(rng, words, int(targets.sum()) + 4096, rich)
    offs = _split_docs(atom_off, targets)
    return data[:int(offs[-1])], offs


def cfg1(vocab_data: bytes, n_docs: int = 1000):
    rng = np.random.default_rng(SEEDS["cfg1"])
    return gen_prose_docs(rng, vocab_words(vocab_data), n_docs, 100.0, 20.0, rich=False)


def cfg2(vocab_data: bytes, n_docs: int = 100_000):
    rng = np.random.default_rng(SEEDS["cfg2"])
    return gen_prose_docs(rng, vocab_words(vocab_data), n_docs, 1000.0, 200.0, rich=True)


# ---------------------------------------------------------------------------------------
# generic helpers for tests
# ---------------------------------------------------------------------------------------
def pack_texts(texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    enc = [t.encode("utf-8") for t in texts]
    offs = np.zeros(len(enc) + 1, dtype=np.uint64)
    if enc:
        np.cumsum(np.fromiter((len(b) for b in enc), dtype=np.uint64, count=len(enc)), out=offs[1:])
    return np.frombuffer(b"".join(enc), dtype=np.uint8), offs


def unpack_texts(data: np.ndarray, offs: np.ndarray) -> List[str]:
    raw = data.tobytes()
    o = offs.tolist()
    return [raw[o[i]:o[i + 1]].decode("utf-8") for i in range(len(o) - 1)]
