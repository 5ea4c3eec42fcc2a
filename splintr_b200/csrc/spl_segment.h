// Independent segments of a piece (host + device).
//
// byte_pair_encode (bpe.rs:83-194) only ever joins two adjacent parts whose concatenation is a vocabulary key.  If NO
// vocabulary key can lie across a given byte boundary of the piece, the parts on its two sides never join, every merge
// decision on one side is independent of the other side (the loop always takes the lowest rank it can see, and the
// ranks on one side do not depend on the other), and the id list of the piece is the concatenation of the id lists of
// the two sides.  Such a boundary is "safe"; the stretches between safe boundaries are the piece's SEGMENTS, and the
// merge loop runs per segment.  CJK text falls apart into segments of one to a few characters this way, which is what
// turns its O(n^2) merge loop into a table walk.
//
// A boundary is only ever declared safe between two well-formed UTF-8 characters A | B of which at least one has more
// than one byte, and only if two vocabulary-derived filters (built by spl_host.cpp) both say "no key crosses here":
//   irr   64 Ki bits, indexed by (last byte of A, first byte of B): set if some key has these two bytes next to each
//         other at a place where the key does NOT hold the whole character that ends there or the whole character that
//         starts there (keys that begin or end in the middle of a character);
//   h2    hashed bitmap over (A, B): set for every place inside a key where a whole character A is followed by a whole
//         character B.  A hashed bitmap has no false negatives.
// Proof of soundness: let a key occurrence cover text[s..e] across the boundary.  Either it holds all bytes of A and
// all bytes of B -- then the builder saw this very (A, B) inside the key and set its h2 bit -- or it lacks a byte of A
// or of B -- then, inside the key, the bytes left of the boundary are continuation bytes only, or the character on
// the right is cut off by the key's end, the builder classed the place as irregular and set irr[last(A)][first(B)].
// Both sides use spl_u8_char below, so "well-formed" means the same thing to the builder and to the walker.
#pragma once
#include "spl_common.h"

// Length of the UTF-8 sequence a lead byte announces; 0 for continuation bytes and bytes that never lead.
SPL_HD uint32_t spl_u8_len(uint32_t b0) {
    if (b0 < 0x80u) return 1u;
    if (b0 < 0xC2u) return 0u;
    if (b0 < 0xE0u) return 2u;
    if (b0 < 0xF0u) return 3u;
    if (b0 < 0xF5u) return 4u;
    return 0u;
}

// Character at the start of the little-endian packed word `w` (bytes beyond `avail` are not looked at): its length if
// the lead byte announces 1..4 bytes, that many are available and the rest are continuation bytes; else 0.
// `packed` receives exactly the character's bytes (zero padded).  Overlong forms and surrogates are NOT rejected --
// the filters only need both sides to agree.
SPL_HD uint32_t spl_u8_char(uint32_t w, uint32_t avail, uint32_t& packed) {
    const uint32_t L = spl_u8_len(w & 0xFFu);
    if (L == 0u || L > avail) return 0u;
    const uint32_t keep = L >= 4u ? 0xFFFFFFFFu : ((1u << (8u * L)) - 1u);
    packed = w & keep;
    const uint32_t cont = packed & 0xC0C0C000u & keep, want = 0x80808000u & keep;
    return cont == want ? L : 0u;
}

SPL_HD uint32_t spl_seg_hash(uint32_t a, uint32_t b, uint32_t log2bits) {
    uint32_t x = a * 0x9E3779B1u + (b ^ 0x5BD1E995u) * 0x85EBCA77u;
    x ^= x >> 15;
    x *= 0x2C1B3C6Du;
    x ^= x >> 13;
    x *= 0x297A2D39u;
    return x >> (32u - log2bits);
}

// the two filters as the walker asks them (A, B: packed characters, la / lb their lengths; at least one > 1)
SPL_HD bool spl_boundary_safe(const uint32_t* irr, const uint32_t* h2, uint32_t h2_log2,
                              uint32_t a, uint32_t la, uint32_t b) {
    const uint32_t a_last = (a >> (8u * (la - 1u))) & 0xFFu, b0 = b & 0xFFu;
    const uint32_t ci = (a_last << 8) | b0;
    if ((irr[ci >> 5] >> (ci & 31u)) & 1u) return false;
    const uint32_t hb = spl_seg_hash(a, b, h2_log2);
    return !((h2[hb >> 5] >> (hb & 31u)) & 1u);
}

#define SPL_SEG_MAX 32u            // the per-lane merge loop (k_bpe) holds a piece or segment of up to this many bytes

// char_tok[code point]: SPL_RANK_NONE = ask the merge loop; else bits 30-31 = id count - 1; one id: the id itself in the
// low bits; two or three ids: index of the first one in char_ids
#define SPL_CHAR_COUNT_SHIFT 30u
#define SPL_CHAR_VALUE_MASK  0x3FFFFFFFu

// code point of a well-formed 2- or 3-byte character (index into the single-character table)
SPL_HD uint32_t spl_u8_cp23(uint32_t packed, uint32_t L) {
    const uint32_t b0 = packed & 0xFFu, b1 = (packed >> 8) & 0x3Fu, b2 = (packed >> 16) & 0x3Fu;
    return L == 2u ? ((b0 & 0x1Fu) << 6) | b1 : ((b0 & 0x0Fu) << 12) | (b1 << 6) | b2;
}

#if defined(__CUDA_ARCH__)
#define SPL_CTZ32(x) ((uint32_t)__ffs((int)(x)) - 1u)
#else
#define SPL_CTZ32(x) ((uint32_t)__builtin_ctz(x))
#endif

// Is the boundary in front of byte q safe?  Position-local: B = the character that starts at q, A = the character
// that ends at q - 1, both decoded right here (the proof above needs nothing else -- in particular not that the bytes
// further left or right are UTF-8).  rd.load4(i) = the four bytes from index i on, little endian; indices from `lo` on
// are readable, bytes in front of `lo` count as absent, `avail` = readable bytes from q on.
template <class Reader>
SPL_HD bool spl_boundary_safe_at(const Reader& rd, uint32_t q, uint32_t lo, uint32_t avail,
                                 const uint32_t* irr, const uint32_t* h2, uint32_t h2_log2) {
    uint32_t b = 0;
    const uint32_t lb = spl_u8_char(rd.load4(q), avail, b);
    if (lb == 0u || q <= lo) return false;
    const uint32_t back = q - lo;                                       // bytes in front of q that exist
    const uint32_t wa = back >= 4u ? rd.load4(q - 4u) : rd.load4(lo) << (8u * (4u - back));   // bytes q-4 .. q-1; absent ones 0
    const uint32_t c1 = wa >> 24, c2 = (wa >> 16) & 0xFFu, c3 = (wa >> 8) & 0xFFu, c4 = wa & 0xFFu;
    uint32_t a, la;
    if (c1 < 0x80u) { a = c1; la = 1u; }
    else if ((c1 & 0xC0u) != 0x80u) return false;                      // a lead byte right in front of q
    else if ((c2 & 0xC0u) != 0x80u) { if (spl_u8_len(c2) != 2u || back < 2u) return false; a = c2 | (c1 << 8); la = 2u; }
    else if ((c3 & 0xC0u) != 0x80u) { if (spl_u8_len(c3) != 3u || back < 3u) return false; a = c3 | (c2 << 8) | (c1 << 16); la = 3u; }
    else if ((c4 & 0xC0u) != 0x80u) { if (spl_u8_len(c4) != 4u || back < 4u) return false; a = wa; la = 4u; }
    else return false;
    if (la == 1u && lb == 1u) return false;                            // ASCII | ASCII is never asked
    return spl_boundary_safe(irr, h2, h2_log2, a, la, b);
}

// Every safe boundary of the piece [0, len) at a position below `limit`: f(pos) is called, in increasing order, for
// each byte position pos (0 < pos < min(len, limit)) in front of which spl_boundary_safe_at holds.
template <class Reader, class F>
SPL_HD void spl_safe_boundaries(const Reader& rd, uint32_t len, uint32_t limit,
                                const uint32_t* irr, const uint32_t* h2, uint32_t h2_log2, F f) {
    if (limit > len) limit = len;
    for (uint32_t q = 1; q < limit; ++q) {
        if ((rd.load4(q) & 0xC0u) == 0x80u) continue;                  // a continuation byte starts no character
        if (spl_boundary_safe_at(rd, q, 0u, len - q, irr, h2, h2_log2)) f(q);
    }
}
