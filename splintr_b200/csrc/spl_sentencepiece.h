// SentencePiece mode (mistral / mistral_v2): the branch of Tokenizer::encode at
// /root/reference/src/core/tokenizer.rs:737-795, restated as position-local rules over four bitmaps so that
// it runs data-parallel.  Everything here is __host__ __device__: tests/csrc/hosttest.cpp runs the same rules on
// the CPU against the oracle (tests/test_sentencepiece_host.py).
//
// The reference splits the text with `[^\s]+|\s+` (tokenizer.rs:56) and walks the chunks with one counter,
// `pending_underscores`:
//   * a whitespace chunk whose FIRST BYTE is ASCII whitespace (u8::is_ascii_whitespace: 09 0A 0C 0D 20, not 0B) is
//     walked BYTE by byte: a space adds one pending U+2581; any other byte first flushes the pending run as one piece
//     and is then encoded as a one-byte piece (so the bytes of U+3000 inside such a chunk are three pieces);
//   * every other chunk (a word, or a whitespace chunk led by 0B or a non-ASCII space such as U+00A0) is one
//     piece, prefixed with the pending U+2581 run; its own bytes stay as they are (its spaces stay 0x20);
//   * the pending run left at the end of the text is one piece.
// So the ids are those of the non-SentencePiece path over a TRANSFORMED text T' in which every "converted" space is
// E2 96 81, with piece starts given by the rules below.  With, for byte i of a segment (document, or gap between
// special-token spans):
//   W0(i)  byte i belongs to a \s character                      (k_sp_classify)
//   RS(i)  i starts a whitespace chunk whose first byte is not ASCII whitespace ("raw" chunk)
//   A(i)   W0(i) and the chunk is not raw                        (W0 with the raw chunks cleared, k_sp_rawruns)
//   S(i)   segment start (document start, first byte of / first byte after a special span)
//   conv(i)   = A(i) and T[i] == 0x20        -> three bytes of output
//   single(i) = A(i) and T[i] != 0x20        -> a one-byte piece
// a piece starts at the output position of i iff
//   S(i) | single(i) | (conv(i) & !conv(i-1)) | RS(i) | (!W0(i) & W0(i-1) & !conv(i-1))
// (the last term: a word that follows whitespace without a pending run -- after a one-byte piece or a raw chunk).
#pragma once
#include "spl_pretok.h"

SPL_HD bool spl_sp_ascii_ws(uint32_t b) { return b == 0x20u || b == 0x09u || b == 0x0Au || b == 0x0Cu || b == 0x0Du; }

// W0 of byte i: decode the character that covers i (text is valid UTF-8; malformed bytes count as non-space)
template <class T>
SPL_HD bool spl_sp_ws_byte(const T& t, uint32_t i, uint32_t N, const uint8_t* s1, const uint8_t* s2) {
    uint32_t b = t.byte(i);
    if (b < 0x80u) return in_set(spl_ascii_class(b), CSET_WS);
    uint32_t j = i;
    while (j > 0 && i - j < 3u && (t.byte(j) & 0xC0u) == 0x80u) --j;
    SplChar c = spl_decode(t, j, N, s1, s2);
    return j + c.len > i && in_set(c.cls, CSET_WS);
}

// bits of one position and of its left neighbour (the neighbour's bits are 0 at i == 0)
struct SplSpPos { bool w0, a, rs, s; uint32_t b; bool w0_prev, a_prev; uint32_t b_prev; };

SPL_HD bool spl_sp_conv(const SplSpPos& p) { return p.a && p.b == 0x20u; }

SPL_HD bool spl_sp_piece_start(const SplSpPos& p) {
    const bool conv = p.a && p.b == 0x20u, single = p.a && p.b != 0x20u;
    const bool conv_prev = p.a_prev && p.b_prev == 0x20u;
    return p.s || single || (conv && !conv_prev) || p.rs || (!p.w0 && p.w0_prev && !conv_prev);
}
