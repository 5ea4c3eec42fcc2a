"""Streaming decoders (host, per-token latency path).

Mirror of /root/reference/src/core/streaming.rs:36-210 (StreamingDecoder) and :232-396
(ByteLevelStreamingDecoder) / src/python/bindings.rs:469-834: buffer token bytes and
release the longest prefix that is complete UTF-8.  Outside the GPU hot path; provided so
that users of the reference's Python API find the same surface.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional


def _is_utf8(b: bytes) -> bool:
    try:
        b.decode("utf-8")
        return True
    except UnicodeDecodeError:
        return False


def _could_be_incomplete(tail: bytes) -> bool:
    """streaming.rs:180-200."""
    if not tail:
        return False
    f = tail[0]
    if 0xC0 <= f <= 0xDF:
        return len(tail) < 2
    if 0xE0 <= f <= 0xEF:
        return len(tail) < 3
    if 0xF0 <= f <= 0xF7:
        return len(tail) < 4
    return False


class StreamingDecoder:
    def __init__(self, decoder: Dict[int, bytes], special_decoder: Dict[int, str]):
        self._decoder = decoder
        self._special = special_decoder
        self._buf = bytearray()

    def _token_bytes(self, token_id: int) -> Optional[bytes]:
        b = self._decoder.get(token_id)
        if b is not None:
            return b
        s = self._special.get(token_id)
        return s.encode("utf-8") if s is not None else None

    def add_token(self, token_id: int) -> Optional[str]:
        b = self._token_bytes(token_id)
        if b is None:
            return None
        self._buf += b
        return self._extract()

    def add_tokens(self, token_ids: Iterable[int]) -> Optional[str]:
        for t in token_ids:
            b = self._token_bytes(t)
            if b is not None:
                self._buf += b
        return self._extract()

    def flush(self) -> str:
        out = bytes(self._buf).decode("utf-8", errors="replace")
        self._buf.clear()
        return out

    def reset(self) -> None:
        self._buf.clear()

    @property
    def has_pending(self) -> bool:
        return len(self._buf) > 0

    @property
    def pending_bytes(self) -> int:
        return len(self._buf)

    def _valid_len(self) -> int:
        """streaming.rs:131-177."""
        b = bytes(self._buf)
        n = len(b)
        if n == 0:
            return 0
        if _is_utf8(b):
            return n
        for inc in range(1, min(3, n) + 1):
            chk = n - inc
            if chk == 0:
                continue
            if _is_utf8(b[:chk]) and _could_be_incomplete(b[chk:]):
                return chk
        for i in range(n - 1, -1, -1):
            if _is_utf8(b[:i + 1]):
                return i + 1
        return 0

    def _extract(self) -> Optional[str]:
        if not self._buf:
            return None
        k = self._valid_len()
        if k == 0:
            return None
        out = bytes(self._buf[:k]).decode("utf-8")
        del self._buf[:k]
        return out

    def __repr__(self) -> str:
        return f"StreamingDecoder(pending_bytes={len(self._buf)})"


def _byte_level_inverse() -> Dict[str, int]:
    direct = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    out, nxt = {}, 256
    for b in range(256):
        if b in direct:
            out[chr(b)] = b
        else:
            out[chr(nxt)] = b
            nxt += 1
    return out


_C2B = _byte_level_inverse()


class ByteLevelStreamingDecoder(StreamingDecoder):
    """Vocabulary bytes are byte-level strings: map them back to raw bytes first
    (streaming.rs:232-396); special tokens and undecodable keys pass through as they are."""

    def _token_bytes(self, token_id: int) -> Optional[bytes]:
        b = self._decoder.get(token_id)
        if b is not None:
            try:
                return bytes(_C2B[c] for c in b.decode("utf-8"))
            except (UnicodeDecodeError, KeyError):
                return b
        s = self._special.get(token_id)
        return s.encode("utf-8") if s is not None else None

    def __repr__(self) -> str:
        return f"ByteLevelStreamingDecoder(pending_bytes={len(self._buf)})"
