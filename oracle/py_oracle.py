"""CPU oracle (O1) for the splintr `encode_batch` path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-Python restatement of the reference algorithm.  It is the
checker: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import it.  The product (`splintr_b200/`) never does.

Restated from (all paths under /root/reference):
  * src/core/vocab.rs:57-89        load_tiktoken_bpe      -> load_tiktoken_bpe()
  * src/core/byte_level.rs:46-74   BYTE_TO_CHAR           -> BYTE_TO_CHAR
  * src/core/byte_level.rs:105-107 byte_level_encode      -> byte_level_encode()
  * src/core/bpe.rs:67-197         byte_pair_encode       -> byte_pair_encode()
  * src/core/tokenizer.rs:693-724  encode_chunk_with_position (LRU omitted: it is
                                   result-transparent)    -> OracleTokenizer._encode_chunk()
  * src/core/tokenizer.rs:729-808  encode (non-SentencePiece branch :796-807; SentencePiece
                                   branch :737-795 -> OracleTokenizer._encode_sentencepiece())
  * src/core/vocab.rs:101-143      load_tiktoken_bpe_with_decoder (SentencePiece vocabularies:
                                   the FIRST id of a duplicated byte string encodes, every id decodes)
  * src/core/tokenizer.rs:923-930  postprocess_decode (U+2581 -> space)
  * src/core/tokenizer.rs:842-874  encode_with_special
  * src/core/tokenizer.rs:932-942  encode_batch / encode_batch_with_special
  * src/core/tokenizer.rs:877-911  decode_bytes / decode / decode_lossy

Third-party arithmetic that is NOT under /root/reference:
  * regexr 0.1.0-beta.5 (Cargo.toml:40) decides every piece boundary.  Its source is
    absent; its published contract is leftmost-first (Perl-style) matching of the
    literal pattern strings at tokenizer.rs:39,42,64.  Restated here with the Python
    `regex` module's `finditer` on the SAME pattern strings (the reference itself
    asserts regexr == PCRE2(UTF|UCP) ids, python/tests/test_cl100k.py:436-454).
  * aho-corasick 1.1 (Cargo.toml:36), default MatchKind::Standard, non-overlapping
    `find_iter`: the match reported is the one that ENDS first; scanning restarts at
    its end (tokenizer.rs:851).  Restated in `_special_find_iter`.

Parity pinning: every golden id vector the reference's own tests hold for this path
(tests/{cl100k,o200k,llama3,deepseek_v3}.rs, python/tests/test_*.py) is reproduced by
this file -- see tests/test_oracle_golden.py.  Beyond those strings (Unicode-version
dependent classes) parity with regexr is unpinned and the working definition is
"leftmost-first semantics of the literal pattern".
"""
from __future__ import annotations

import base64
from typing import Dict, Iterable, List, Optional, Tuple

import regex as _regex

# tokenizer.rs:39
CL100K_BASE_PATTERN = r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
# tokenizer.rs:42 (LLAMA3_PATTERN :45 is the same string)
O200K_BASE_PATTERN = r"[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+(?i:'s|'t|'re|'ve|'m|'ll|'d)?|[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*(?i:'s|'t|'re|'ve|'m|'ll|'d)?|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
LLAMA3_PATTERN = O200K_BASE_PATTERN
# tokenizer.rs:64
MISTRAL_V3_PATTERN = r"[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+|[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n/]*|\s*[\r\n]+|\s+(?!\S)|\s+"

U32_MAX = 0xFFFFFFFF


def load_tiktoken_bpe(data: bytes) -> Dict[bytes, int]:
    """vocab.rs:57-89 -- `base64 SP rank LF`; split on the LAST space; later
    duplicates overwrite earlier ones (`insert`, vocab.rs:85)."""
    enc: Dict[bytes, int] = {}
    for line in data.split(b"\n"):
        if not line:
            continue
        sp = line.rfind(b" ")
        if sp < 0:
            raise ValueError("Invalid line format: Missing space separator")
        tok = base64.b64decode(line[:sp], validate=True)
        enc[tok] = int(line[sp + 1:].decode("utf-8").strip())
    return enc


def load_tiktoken_bpe_with_decoder(data: bytes) -> Tuple[Dict[bytes, int], Dict[int, bytes]]:
    """vocab.rs:101-143 -- the encoder keeps the FIRST occurrence of a byte string
    (`entry().or_insert`, :139), the decoder keeps every id (:135)."""
    enc: Dict[bytes, int] = {}
    dec: Dict[int, bytes] = {}
    for line in data.split(b"\n"):
        if not line:
            continue
        sp = line.rfind(b" ")
        if sp < 0:
            raise ValueError("Invalid line format: Missing space separator")
        tok = base64.b64decode(line[:sp], validate=True)
        rank = int(line[sp + 1:].decode("utf-8").strip())
        dec[rank] = tok
        enc.setdefault(tok, rank)
    return enc, dec


SENTENCEPIECE_PATTERN = r"[^\s]+|\s+"          # tokenizer.rs:56
_SP_UNDERSCORE = "\u2581".encode("utf-8")     # E2 96 81
_ASCII_WS = frozenset(b" \t\n\x0c\r")          # u8::is_ascii_whitespace (no VT)


def _build_byte_to_char() -> List[str]:
    """byte_level.rs:46-74."""
    direct = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    m = [""] * 256
    for b in direct:
        m[b] = chr(b)
    nxt = 256
    for b in range(256):
        if b not in direct:
            m[b] = chr(nxt)
            nxt += 1
    return m


BYTE_TO_CHAR = _build_byte_to_char()
CHAR_TO_BYTE = {c: b for b, c in enumerate(BYTE_TO_CHAR)}
_BYTE_TO_UTF8 = [c.encode("utf-8") for c in BYTE_TO_CHAR]


def byte_level_encode(data: bytes) -> bytes:
    """byte_level.rs:105-107 (returned as the UTF-8 bytes of the String)."""
    return b"".join(_BYTE_TO_UTF8[b] for b in data)


def byte_level_decode_bytes(encoded: bytes) -> Optional[bytes]:
    """byte_level.rs:142-146; None when not UTF-8 or a char is outside the alphabet."""
    try:
        s = encoded.decode("utf-8")
    except UnicodeDecodeError:
        return None
    out = bytearray()
    for ch in s:
        b = CHAR_TO_BYTE.get(ch)
        if b is None:
            return None
        out.append(b)
    return bytes(out)


def byte_pair_encode(piece: bytes, encoder: Dict[bytes, int]) -> List[int]:
    """bpe.rs:67-197, node for node.  prev/next/rank/start/len arrays stand in for
    the Node struct (bpe.rs:42-54); -1 stands in for usize::MAX."""
    n = len(piece)
    if n == 0:                                   # :68-70
        return []
    if n == 1:                                   # :73-75
        r = encoder.get(piece)
        return [] if r is None else [r]
    r = encoder.get(piece)                       # :78-80
    if r is not None:
        return [r]

    prev = [i - 1 for i in range(n)]             # :83-97
    nxt = [i + 1 for i in range(n)]
    nxt[n - 1] = -1
    rank = [U32_MAX] * n
    start = list(range(n))
    length = [1] * n

    def get_rank(li: int, ri: int) -> int:      # :99-111
        if li == -1 or ri == -1:
            return U32_MAX
        s = start[li]
        return encoder.get(piece[s:s + length[li] + length[ri]], U32_MAX)

    for i in range(n - 1):                       # :114-116
        rank[i] = get_rank(i, nxt[i])

    while True:                                  # :119-167
        min_rank = U32_MAX
        min_idx = -1
        cur = 0
        while prev[cur] != -1:
            cur = prev[cur]
        while cur != -1:
            if rank[cur] < min_rank:             # strict <  => leftmost wins ties (:133)
                min_rank = rank[cur]
                min_idx = cur
            cur = nxt[cur]
        if min_rank == U32_MAX:
            break
        nx = nxt[min_idx]
        length[min_idx] += length[nx]
        nn = nxt[nx]
        nxt[min_idx] = nn
        if nn != -1:
            prev[nn] = min_idx
        if prev[min_idx] != -1:
            p = prev[min_idx]
            rank[p] = get_rank(p, min_idx)
        rank[min_idx] = get_rank(min_idx, nxt[min_idx])

    out: List[int] = []                          # :170-194
    cur = 0
    while prev[cur] != -1:
        cur = prev[cur]
    while cur != -1:
        sl = piece[start[cur]:start[cur] + length[cur]]
        r = encoder.get(sl)
        if r is not None:
            out.append(r)
        else:
            for b in sl:                         # unknown bytes are silently dropped
                rb = encoder.get(bytes([b]))
                if rb is not None:
                    out.append(rb)
        cur = nxt[cur]
    return out


class OracleTokenizer:
    """Restatement of core::Tokenizer (encode / decode paths, both modes)."""

    def __init__(self, encoder: Dict[bytes, int], special_tokens: Dict[str, int],
                 pattern: str, byte_level: bool = False, sentencepiece: bool = False,
                 decoder: Optional[Dict[int, bytes]] = None):
        self.encoder = encoder
        self.decoder = decoder if decoder is not None else {v: k for k, v in encoder.items()}   # vocab.rs:146-148
        self.sentencepiece = sentencepiece
        if sentencepiece:                                           # tokenizer.rs:597-600
            for s_, i_ in special_tokens.items():
                self.decoder[i_] = s_.encode("utf-8")
        self.special_tokens = dict(special_tokens)
        self.special_tokens_decoder = {v: k for k, v in special_tokens.items()}
        self.pattern = pattern
        self.regex = _regex.compile(pattern)
        self.byte_level = byte_level
        self._special_bytes = [(s.encode("utf-8"), i) for s, i in special_tokens.items()]

    @classmethod
    def from_bytes(cls, vocab_data: bytes, pattern: str,
                   special_tokens: Optional[Dict[str, int]] = None, byte_level: bool = False,
                   sentencepiece: bool = False):
        if sentencepiece:                                           # tokenizer.rs:589-640
            enc, dec = load_tiktoken_bpe_with_decoder(vocab_data)
            return cls(enc, special_tokens or {}, pattern, False, True, dec)
        return cls(load_tiktoken_bpe(vocab_data), special_tokens or {}, pattern, byte_level)

    # -- tokenizer.rs:244-257 ------------------------------------------------------
    def find_iter(self, text: str) -> List[Tuple[int, int]]:
        """(start, end) offsets in CHARACTERS of `text` (callers slice the str)."""
        return [m.span() for m in self.regex.finditer(text)]

    def pieces(self, text: str) -> List[bytes]:
        return [text[s:e].encode("utf-8") for s, e in self.find_iter(text)]

    # -- tokenizer.rs:693-724 (LRU skipped) ------------------------------------------
    def _encode_chunk(self, piece: bytes) -> List[int]:
        b = byte_level_encode(piece) if self.byte_level else piece
        r = self.encoder.get(b)
        if r is not None:
            return [r]
        return byte_pair_encode(b, self.encoder)

    # -- tokenizer.rs:729-808 ----------------------------------------------------------
    def encode(self, text: str) -> List[int]:
        if self.sentencepiece:
            return self._encode_sentencepiece(text)
        out: List[int] = []
        for s, e in self.find_iter(text):
            out.extend(self._encode_chunk(text[s:e].encode("utf-8")))
        return out

    # -- tokenizer.rs:666-690 (LRU skipped) --------------------------------------------
    def _encode_bytes(self, b: bytes) -> List[int]:
        r = self.encoder.get(b)
        if r is not None:
            return [r]
        return byte_pair_encode(b, self.encoder)

    # -- tokenizer.rs:737-795 ----------------------------------------------------------
    def sentencepiece_pieces(self, text: str) -> List[bytes]:
        """The byte strings the SentencePiece branch hands to encode_bytes_with_cache, in order."""
        out: List[bytes] = []
        pending = 0                                   # count of U+2581 to prepend to the next word
        for s, e in self.find_iter(text):
            sl = text[s:e].encode("utf-8")
            if not sl:
                continue
            if sl[0] in _ASCII_WS:                    # :750 -- decided by the FIRST BYTE only
                for b in sl:                          # :752 -- byte by byte
                    if b == 0x20:
                        pending += 1
                    else:
                        if pending:
                            out.append(_SP_UNDERSCORE * pending)
                            pending = 0
                        out.append(bytes([b]))
            else:
                if pending:
                    out.append(_SP_UNDERSCORE * pending + sl)
                    pending = 0
                else:
                    out.append(sl)
        if pending:                                   # :787-790 trailing spaces
            out.append(_SP_UNDERSCORE * pending)
        return out

    def _encode_sentencepiece(self, text: str) -> List[int]:
        out: List[int] = []
        for piece in self.sentencepiece_pieces(text):
            out.extend(self._encode_bytes(piece))
        return out

    # -- aho-corasick Standard, non-overlapping ---------------------------------------
    def _special_find_iter(self, data: bytes) -> Iterable[Tuple[int, int, int]]:
        """Yield (start, end, id): among all occurrences starting at/after the cursor
        the one that ENDS first; ties on the end go to the longest pattern (the
        automaton state's own pattern precedes its fail-link suffixes)."""
        pats = self._special_bytes
        pos = 0
        n = len(data)
        while pos < n:
            best = None
            for p, tid in pats:
                if not p:
                    continue
                i = data.find(p, pos)
                if i < 0:
                    continue
                cand = (i + len(p), i, tid)      # earliest end, then longest (smallest start)
                if best is None or cand[:2] < best[:2]:
                    best = cand
            if best is None:
                return
            yield best[1], best[0], best[2]
            pos = best[0]

    # -- tokenizer.rs:842-874 ----------------------------------------------------------
    def encode_with_special(self, text: str) -> List[int]:
        if not self._special_bytes:
            return self.encode(text)
        data = text.encode("utf-8")
        out: List[int] = []
        last = 0
        for s, e, tid in self._special_find_iter(data):
            if s > last:
                out.extend(self.encode(data[last:s].decode("utf-8")))
            out.append(tid)
            last = e
        if last < len(data):
            out.extend(self.encode(data[last:].decode("utf-8")))
        return out

    # -- tokenizer.rs:932-942 ----------------------------------------------------------
    def encode_batch(self, texts: List[str]) -> List[List[int]]:
        return [self.encode(t) for t in texts]

    def encode_batch_with_special(self, texts: List[str]) -> List[List[int]]:
        return [self.encode_with_special(t) for t in texts]

    # -- tokenizer.rs:877-911 ----------------------------------------------------------
    def decode_bytes(self, tokens: List[int]) -> bytes:
        out = bytearray()
        for t in tokens:
            b = self.decoder.get(t)
            if b is not None:
                if self.byte_level:
                    d = byte_level_decode_bytes(b)
                    out += d if d is not None else b
                else:
                    out += b
            else:
                s = self.special_tokens_decoder.get(t)
                if s is not None:
                    out += s.encode("utf-8")
        return bytes(out)

    def _postprocess(self, text: str) -> str:                       # tokenizer.rs:923-930
        return text.replace("\u2581", " ") if self.sentencepiece else text

    def decode(self, tokens: List[int]) -> str:
        try:
            return self._postprocess(self.decode_bytes(tokens).decode("utf-8"))
        except UnicodeDecodeError:
            raise ValueError("Decoding error: invalid UTF-8")

    def decode_lossy(self, tokens: List[int]) -> str:
        return self._postprocess(self.decode_bytes(tokens).decode("utf-8", errors="replace"))

    @property
    def vocab_size(self) -> int:                 # tokenizer.rs:964-972
        m1 = max(self.decoder.keys(), default=0)
        m2 = max(self.special_tokens.values(), default=0)
        return max(m1, m2) + 1
