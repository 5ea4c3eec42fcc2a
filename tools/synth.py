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


def cfg1(vocab_data: bytes, n_docs: int = 1000, seed_offset: int = 0):
    rng = np.random.default_rng(SEEDS["cfg1"] + seed_offset)
    return gen_prose_docs(rng, vocab_words(vocab_data), n_docs, 100.0, 20.0, rich=False)


def cfg2(vocab_data: bytes, n_docs: int = 100_000, seed_offset: int = 0):
    rng = np.random.default_rng(SEEDS["cfg2"] + seed_offset)
    return gen_prose_docs(rng, vocab_words(vocab_data), n_docs, 1000.0, 200.0, rich=True)


# ---------------------------------------------------------------------------------------
# cfg3: o200k_base, mixed prose / code / JSON (~2 KB docs)
# ---------------------------------------------------------------------------------------
_IDENT_PARTS = ["get", "set", "value", "index", "count", "data", "result", "item", "list", "name", "user", "config",
                "buffer", "size", "node", "key", "error", "request", "response", "handler", "token", "batch", "max",
                "min", "total", "path", "file", "line", "state", "id", "type", "parse", "load", "update", "cache"]
_KEYWORDS = ["def", "return", "if", "else", "elif", "for", "while", "in", "import", "from", "class", "int", "void",
             "const", "static", "struct", "true", "false", "None", "null", "self", "not", "and", "or"]


def _ident(rng) -> str:
    k = int(rng.integers(1, 4))
    parts = [_IDENT_PARTS[int(i)] for i in rng.integers(0, len(_IDENT_PARTS), size=k)]
    if rng.random() < 0.5:
        return "_".join(parts)
    return parts[0] + "".join(p.capitalize() for p in parts[1:])


def _code_lines(rng, n: int) -> List[bytes]:
    out = []
    for _ in range(n):
        ind = " " * (4 * int(rng.integers(0, 4)))
        a, b, c = _ident(rng), _ident(rng), _ident(rng)
        num = str(int(rng.integers(0, 10 ** int(rng.integers(1, 6)))))
        t = int(rng.integers(0, 12))
        line = [f"def {a}({b}, {c}):", f"{a} = {b}({c})", f"if {a} == {num}:", f"return {a}[{b}]",
                f"for {a} in range({num}):", f"# {a} {b} {c}", f"{a}.{b}(\"{c}\", {num})", f"int {a} = {num};",
                f"{a} += {b} * {c} - {num}", f"}} else if ({a} != {b}) {{", f"print(f\"{{{a}}}: {{{b}}}\")",
                f"{a}->{b} = &{c}[{num}];"][t]
        out.append((ind + line + "\n").encode())
    out += [b"\n", b"}\n", b"    pass\n", b"\n\n"]
    return out


def _json_lines(rng, n: int) -> List[bytes]:
    out = []
    for _ in range(n):
        ind = " " * (2 * int(rng.integers(0, 5)))
        k = _ident(rng)
        t = int(rng.integers(0, 8))
        if t == 0:
            v = str(int(rng.integers(-1000, 10 ** int(rng.integers(1, 8)))))
        elif t == 1:
            v = f"{rng.random() * 10 ** int(rng.integers(0, 4)):.{int(rng.integers(1, 6))}f}"
        elif t == 2:
            v = ["true", "false", "null"][int(rng.integers(0, 3))]
        elif t == 3:
            v = "{"
        elif t == 4:
            v = "[" + ", ".join(str(int(x)) for x in rng.integers(0, 100, size=int(rng.integers(1, 6)))) + "]"
        else:
            v = "\"" + " ".join(_IDENT_PARTS[int(i)] for i in rng.integers(0, len(_IDENT_PARTS), size=int(rng.integers(1, 5)))) + "\""
        out.append(f"{ind}\"{k}\": {v}{'' if v == '{' else ','}\n".encode())
    out += [b"}\n", b"},\n", b"{\n", b"]\n"]
    return out


def _line_docs(rng, lines: List[bytes], n_docs: int, mean_len: float, sd: float):
    pool, poff, plen = _pool(lines)
    targets = np.maximum(rng.normal(mean_len, sd, size=n_docs), 16.0)
    n_lines = int(targets.sum() / float(plen.mean()) * 1.05) + 64
    ids = rng.integers(0, len(lines), size=n_lines)
    data, off = _gather(pool, poff, plen, ids)
    offs = _split_docs(off, targets)
    return data[:int(offs[-1])], offs


def _interleave_docs(rng, parts):
    """parts = [(data, offsets), ...] -> one batch with the documents of all parts shuffled."""
    datas = [p[0] for p in parts]
    base = np.cumsum([0] + [len(d) for d in datas[:-1]])
    starts = np.concatenate([p[1][:-1].astype(np.int64) + b for p, b in zip(parts, base)])
    lens = np.concatenate([np.diff(p[1].astype(np.int64)) for p in parts])
    perm = rng.permutation(len(starts))
    data, off = _gather(np.concatenate(datas), starts, lens, perm)
    return data, off.astype(np.uint64)


def cfg3(vocab_data: bytes, n_docs: int = 1_000_000, seed_offset: int = 0):
    rng = np.random.default_rng(SEEDS["cfg3"] + seed_offset)
    n3 = n_docs // 3
    prose = gen_prose_docs(rng, vocab_words(vocab_data), n_docs - 2 * n3, 2000.0, 400.0, rich=True)
    code = _line_docs(rng, _code_lines(rng, 20000), n3, 2000.0, 400.0)
    js = _line_docs(rng, _json_lines(rng, 20000), n3, 2000.0, 400.0)
    return _interleave_docs(rng, [prose, code, js])


# ---------------------------------------------------------------------------------------
# cfg4: llama3, long docs with deep per-piece merge chains (pieces capped at 1 KiB)
# ---------------------------------------------------------------------------------------
def cfg4(vocab_data: bytes, n_docs: int = 10_000, doc_bytes: float = 1_000_000.0, seed_offset: int = 0):
    rng = np.random.default_rng(SEEDS["cfg4"] + seed_offset)
    words = vocab_words(vocab_data)
    n_words = len(words)
    letters = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz", dtype=np.uint8)
    long_words = []
    for _ in range(8192):                      # random lowercase strings, length LogUniform[32, 512]
        ln = int(np.exp(rng.uniform(np.log(32), np.log(512))))
        long_words.append(b" " + letters[rng.integers(0, 26, size=ln)].tobytes())
    punct = [bytes([int(rng.choice(list(b"=-#")))]) * int(rng.integers(2, 257)) + b"\n" for _ in range(2048)]
    ws = [b" " * int(rng.integers(2, 65)) for _ in range(512)] + [b"\n" * int(rng.integers(1, 5)) for _ in range(512)]
    atoms = list(words) + long_words + punct + ws
    pool, poff, plen = _pool(atoms)
    targets = np.maximum(rng.normal(doc_bytes, doc_bytes * 0.1, size=n_docs), 64.0)
    zw = 1.0 / np.arange(1, n_words + 1, dtype=np.float64) ** 1.1
    mean_word = float((plen[:n_words] * zw).sum() / zw.sum())
    mean_atom = 0.93 * mean_word + 0.05 * float(plen[n_words:n_words + 8192].mean()) + 0.01 * float(plen[n_words + 8192:n_words + 8192 + 2048].mean()) + 0.01 * 20
    n_atoms = int(targets.sum() / mean_atom * 1.05) + 64
    ids = _zipf_ids(rng, n_words, n_atoms)
    v = rng.random(n_atoms)
    m = v < 0.05
    ids[m] = n_words + rng.integers(0, 8192, size=int(m.sum()))
    m = (v >= 0.05) & (v < 0.06)
    ids[m] = n_words + 8192 + rng.integers(0, 2048, size=int(m.sum()))
    m = (v >= 0.06) & (v < 0.07)
    ids[m] = n_words + 8192 + 2048 + rng.integers(0, 1024, size=int(m.sum()))
    data, off = _gather(pool, poff, plen, ids)
    offs = _split_docs(off, targets)
    return data[:int(offs[-1])], offs


# ---------------------------------------------------------------------------------------
# cfg5: deepseek_v3, CJK-heavy docs (~1.5 KB); code points whose class is stable across Unicode versions
# ---------------------------------------------------------------------------------------
def cfg5(vocab_data: bytes = b"", n_docs: int = 100_000, seed_offset: int = 0):
    rng = np.random.default_rng(SEEDS["cfg5"] + seed_offset)
    han = rng.permutation(np.arange(0x4E00, 0x9FA6))[:3500]
    kana = np.concatenate([np.arange(0x3041, 0x3097), np.arange(0x30A1, 0x30FB)])
    hangul = rng.permutation(np.arange(0xAC00, 0xD7A4))[:600]
    punct = [ord(c) for c in "，。！？：；“”"]
    ascii_atoms = [w.encode() for w in ["GPU", "API", " the", " data", " model", "2024", "100", "3.14", " AI", "test", " ", "\n"]]
    atoms = [chr(int(c)).encode() for c in han] + [chr(int(c)).encode() for c in kana] + \
            [chr(int(c)).encode() for c in hangul] + [chr(c).encode() for c in punct] + ascii_atoms
    n_han, n_kana, n_hangul, n_p = len(han), len(kana), len(hangul), len(punct)
    pool, poff, plen = _pool(atoms)
    targets = np.maximum(rng.normal(1500.0, 300.0, size=n_docs), 16.0)
    n_atoms = int(targets.sum() / 2.9 * 1.05) + 64
    ids = _zipf_ids(rng, n_han, n_atoms, s=1.0)
    v = rng.random(n_atoms)
    m = (v >= 0.72) & (v < 0.80)
    ids[m] = n_han + n_kana + n_hangul + rng.integers(0, n_p, size=int(m.sum()))
    m = (v >= 0.80) & (v < 0.90)
    ids[m] = n_han + n_kana + n_hangul + n_p + rng.integers(0, len(ascii_atoms), size=int(m.sum()))
    m = (v >= 0.90) & (v < 0.95)
    ids[m] = n_han + rng.integers(0, n_kana, size=int(m.sum()))
    m = v >= 0.95
    ids[m] = n_han + n_kana + rng.integers(0, n_hangul, size=int(m.sum()))
    data, off = _gather(pool, poff, plen, ids)
    offs = _split_docs(off, targets)
    return data[:int(offs[-1])], offs


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
