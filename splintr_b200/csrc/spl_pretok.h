// Pre-tokenizer: the split patterns of /root/reference/src/core/tokenizer.rs:39 (CL100K),
// :42 (O200K = LLAMA3 :45, also deepseek_v3 via bindings.rs:126) and :64 (MISTRAL_V3),
// restated as deterministic rules over Unicode classes.  This replaces the regex engine
// call `RegexBackend::find_iter` (tokenizer.rs:244-257): leftmost-first, non-overlapping
// matches that tile the text.
//
// Everything here is __host__ __device__ so that the exact device logic can be fuzzed on
// the CPU against the oracle's regex engine (tests/test_pretok_host.py).
//
// Text accessor concept `T`:   uint8_t byte(uint32_t i) const;
// A "segment" is a maximal stretch of text the regex sees as one input: a document, or
// (encode_with_special) a gap between special-token matches.  `E` below is always the
// exclusive end of the current segment; nothing at or beyond E is ever read.
#pragma once
#include "spl_common.h"

struct SplChar { uint32_t cls; uint32_t len; };

SPL_HD uint32_t spl_class_of_cp(uint32_t cp, const uint8_t* s1, const uint8_t* s2) {
    uint32_t blk = s1[cp >> 8];
    uint32_t v = s2[blk * 128u + ((cp & 255u) >> 1)];
    return (cp & 1u) ? (v >> 4) : (v & 15u);
}

// Decode the character starting at byte i (i < E).  Malformed or truncated sequences are
// treated as one-byte CLS_OTHER characters (the reference only ever sees valid UTF-8).
template <class T>
SPL_HD SplChar spl_decode(const T& t, uint32_t i, uint32_t E, const uint8_t* s1, const uint8_t* s2, bool* hit = nullptr) {
    uint32_t b0 = t.byte(i);
    SplChar c;
    if (b0 < 0x80u) { c.cls = spl_ascii_class(b0); c.len = 1; return c; }
    c.cls = CLS_OTHER; c.len = 1;
    if (b0 < 0xC2u || b0 > 0xF4u) return c;
    uint32_t need = b0 < 0xE0u ? 2u : (b0 < 0xF0u ? 3u : 4u);
    if (E - i < need) { if (hit) *hit = true; return c; }
    uint32_t b1 = t.byte(i + 1);
    if ((b1 & 0xC0u) != 0x80u) return c;
    uint32_t cp;
    if (need == 2) {
        cp = ((b0 & 0x1Fu) << 6) | (b1 & 0x3Fu);
    } else {
        uint32_t b2 = t.byte(i + 2);
        if ((b2 & 0xC0u) != 0x80u) return c;
        if (need == 3) {
            cp = ((b0 & 0x0Fu) << 12) | ((b1 & 0x3Fu) << 6) | (b2 & 0x3Fu);
            if (cp < 0x800u || (cp >= 0xD800u && cp <= 0xDFFFu)) return c;
        } else {
            uint32_t b3 = t.byte(i + 3);
            if ((b3 & 0xC0u) != 0x80u) return c;
            cp = ((b0 & 0x07u) << 18) | ((b1 & 0x3Fu) << 12) | ((b2 & 0x3Fu) << 6) | (b3 & 0x3Fu);
            if (cp < 0x10000u || cp > 0x10FFFFu) return c;
        }
    }
    c.cls = spl_class_of_cp(cp, s1, s2);
    c.len = need;
    return c;
}

// Start of the character that ends right before byte i (segment start S < i).
template <class T>
SPL_HD uint32_t spl_prev_char_start(const T& t, uint32_t i, uint32_t S) {
    uint32_t j = i - 1;
    uint32_t lim = (i - S > 4u) ? i - 4u : S;
    while (j > lim && (t.byte(j) & 0xC0u) == 0x80u) --j;
    return j;
}

// (?i:'s|'t|'re|'ve|'m|'ll|'d) at byte a; returns matched length in bytes or 0.
// The only non-ASCII code point that case-folds onto one of the letters is U+017F
// (LATIN SMALL LETTER LONG S, bytes C5 BF) -> 's' (verified over all code points by
// tools/gen_unicode_tables.py).
template <class T>
SPL_HD uint32_t spl_contraction(const T& t, uint32_t a, uint32_t E, bool* hit = nullptr) {
    if (a >= E) { if (hit) *hit = true; return 0; }
    if (t.byte(a) != '\'') return 0;
    if (a + 1 >= E) { if (hit) *hit = true; return 0; }
    uint32_t b1 = t.byte(a + 1);
    uint32_t l1 = b1 | 0x20u;
    if (b1 < 0x80u) {
        if (l1 == 's' || l1 == 't' || l1 == 'm' || l1 == 'd') return 2;
        if (l1 != 'r' && l1 != 'v' && l1 != 'l') return 0;
        if (a + 2 >= E) { if (hit) *hit = true; return 0; }
        {
            uint32_t b2 = t.byte(a + 2);
            uint32_t l2 = b2 | 0x20u;
            if (b2 < 0x80u) {
                if ((l1 == 'r' || l1 == 'v') && l2 == 'e') return 3;
                if (l1 == 'l' && l2 == 'l') return 3;
            }
        }
        return 0;
    }
    if (b1 == 0xC5u) {                                                    // 'ſ
        if (a + 2 >= E) { if (hit) *hit = true; return 0; }
        if (t.byte(a + 2) == 0xBFu) return 3;
    }
    return 0;
}

// TRACK = true: `hit` is set whenever a decision depended on the segment end E (any
// comparison `i < E` that came out false).  A caller that only knows a provisional E
// (its staging window ends before the real segment end) re-runs with a larger E when hit.
template <class T, bool TRACK = false>
struct SplScanner {
    const T& t;
    const uint8_t* s1;
    const uint8_t* s2;
    int pattern;
    mutable bool hit;

    SPL_HD SplScanner(const T& t_, const uint8_t* s1_, const uint8_t* s2_, int pat) : t(t_), s1(s1_), s2(s2_), pattern(pat), hit(false) {}

    SPL_HD bool lt(uint32_t i, uint32_t E) const {
        if (i < E) return true;
        if (TRACK) hit = true;
        return false;
    }
    SPL_HD SplChar dec(uint32_t i, uint32_t E) const { return spl_decode(t, i, E, s1, s2, TRACK ? &hit : nullptr); }
    SPL_HD uint32_t contr(uint32_t a, uint32_t E) const { return spl_contraction(t, a, E, TRACK ? &hit : nullptr); }

    // end of the maximal run of chars whose class is in `set`, starting at i
    SPL_HD uint32_t run_end(uint32_t i, uint32_t E, uint32_t set) const {
        while (lt(i, E)) {
            SplChar c = dec(i, E);
            if (!in_set(c.cls, set)) break;
            i += c.len;
        }
        return i;
    }

    // Letter alternatives of O200K / MISTRAL_V3 starting at s (after the optional prefix):
    //   A1 = U* W+   (backtracks to just after the last Lm/Lo/M char of the U-run)
    //   A2 = U+ W*
    // a1/a2 = end offset or 0 when the alternative does not match at s.
    SPL_HD void letters_o200k(uint32_t s, uint32_t E, uint32_t& a1, uint32_t& a2) const {
        a1 = 0; a2 = 0;
        uint32_t i = s, last_both_end = 0;
        while (lt(i, E)) {
            SplChar c = dec(i, E);
            if (!in_set(c.cls, CSET_U)) break;
            i += c.len;
            if (in_set(c.cls, CSET_BOTH)) last_both_end = i;
        }
        uint32_t e1 = i;
        bool next_lower = false;
        if (lt(e1, E)) { SplChar c = dec(e1, E); next_lower = (c.cls == CLS_LOWER); }
        if (next_lower) { a1 = run_end(e1, E, CSET_W); if (e1 > s) a2 = a1; }
        else {
            if (last_both_end) a1 = last_both_end;
            if (e1 > s) a2 = e1;
        }
    }

    // End of the piece that starts at p (p < E, p on a character boundary).
    SPL_HD uint32_t next_end(uint32_t p, uint32_t E) const {
        SplChar c = dec(p, E);
        uint32_t q = p + c.len;                       // start of the second character
        if (pattern == SPL_PAT_CL100K) {
            uint32_t k = contr(p, E);                                   // alt 1
            if (k) return p + k;
            if (in_set(c.cls, CSET_L)) return run_end(q, E, CSET_L);                 // alt 2
            if (in_set(c.cls, CSET_PREFIX) && lt(q, E)) {
                SplChar n = dec(q, E);
                if (in_set(n.cls, CSET_L)) return run_end(q + n.len, E, CSET_L);
            }
        } else {
            uint32_t a1p = 0, a2p = 0, a10 = 0, a20 = 0;
            bool pref_ok = in_set(c.cls, CSET_PREFIX) && lt(q, E);
            if (pref_ok) letters_o200k(q, E, a1p, a2p);
            if (in_set(c.cls, CSET_U | CSET_W)) letters_o200k(p, E, a10, a20);
            uint32_t e = a1p ? a1p : (a10 ? a10 : (a2p ? a2p : a20));
            if (e) {
                if (pattern == SPL_PAT_O200K) e += contr(e, E);
                return e;
            }
        }
        if (c.cls == CLS_NUM) {                                                       // \p{N}{1,3}
            if (pattern == SPL_PAT_MISTRAL_V3) return q;
            uint32_t i = q;
            for (int k = 1; k < 3 && lt(i, E); ++k) {
                SplChar n = dec(i, E);
                if (n.cls != CLS_NUM) break;
                i += n.len;
            }
            return i;
        }
        {                                                                             //  ?[^\s\p{L}\p{N}]+[\r\n]*
            uint32_t s = p; SplChar sc = c;
            if (c.cls == CLS_SPACE && lt(q, E)) {
                SplChar n = dec(q, E);
                if (in_set(n.cls, CSET_O)) { s = q; sc = n; }
            }
            if (in_set(sc.cls, CSET_O)) {
                uint32_t i = run_end(s + sc.len, E, CSET_O);
                uint32_t tail = (pattern == SPL_PAT_MISTRAL_V3) ? (CM(CLS_CRLF) | CM(CLS_SLASH)) : CM(CLS_CRLF);
                return run_end(i, E, tail);
            }
        }
        // whitespace run:  \s*[\r\n]+  |  \s+(?!\S)  |  \s+
        uint32_t i = p, last_crlf_end = 0, last_start = p, nchars = 0;
        while (lt(i, E)) {
            SplChar w = dec(i, E);
            if (!in_set(w.cls, CSET_WS)) break;
            last_start = i;
            i += w.len;
            ++nchars;
            if (w.cls == CLS_CRLF) last_crlf_end = i;
        }
        if (nchars == 0) return q;               // unreachable for classified text; guarantees progress
        if (last_crlf_end) return last_crlf_end;
        if (i >= E) { if (TRACK) hit = true; return i; }
        if (nchars >= 2) return last_start;
        return i;
    }

    // Context-free piece starts ("sync points"): true only if a piece starts at i no
    // matter where the scan to the left of i began.  S = segment start, E = segment end,
    // S <= i < E, i on a character boundary.
    SPL_HD bool is_sync(uint32_t i, uint32_t S, uint32_t E) const {
        if (i == S) return true;
        SplChar c = dec(i, E);
        uint32_t pj = spl_prev_char_start(t, i, S);
        SplChar pc = dec(pj, E);
        if (pj + pc.len != i) return false;                       // malformed neighbourhood: no claim
        bool cws = in_set(c.cls, CSET_WS);
        if (pc.cls == CLS_CRLF && !cws &&                                          // after a line break
            !(pattern == SPL_PAT_MISTRAL_V3 && c.cls == CLS_SLASH)) return true;   // ([\r\n/]* tail)
        if ((c.cls == CLS_NUM) != (pc.cls == CLS_NUM)) return true;                // digit <-> non-digit
        if (cws && c.cls != CLS_CRLF && lt(i + c.len, E)) {                           // last blank before a word
            SplChar n = dec(i + c.len, E);
            if (!in_set(n.cls, CSET_WS)) return true;
        }
        if (in_set(pc.cls, CSET_L) && in_set(c.cls, CM(CLS_OTHER) | CM(CLS_SLASH))) return true;   // punctuation after a letter
        return false;
    }
};

// ---------------------------------------------------------------------------------------
// Work split of the piece-boundary kernel, written against an environment so the very same
// code runs in the CUDA kernel (shared-memory window + global fallback) and in the host
// fuzz harness.
//
// Env concept:
//   uint8_t  byte(uint32_t i) const            text byte (i < N)
//   bool     hard(uint32_t i) const            segment-boundary bit (doc start, special-span edge; bit N is set)
//   bool     spec(uint32_t i) const            byte i lies inside a special-token span (a span starts where hard && spec)
//   uint32_t next_hard(uint32_t from, uint32_t lim) const   first hard bit in [from, lim), else lim
//   uint32_t win_end() const                   bits below this position are cheap to query
//   void     mark(uint32_t p)                  record "a piece starts at p"
struct SplSegEnd { uint32_t E; bool exact; uint32_t step; };

template <class Env>
SPL_HD void spl_seg_locate(const Env& env, uint32_t p, uint32_t N, SplSegEnd& s) {
    uint32_t from = p + 1;
    uint32_t W = env.win_end();
    uint32_t lim = W < N + 1 ? W : N + 1;
    s.step = 1024;
    if (lim > from) {
        uint32_t x = env.next_hard(from, lim);
        if (x < lim) { s.E = x; s.exact = true; return; }
        from = lim;
    }
    s.E = from; s.exact = false;            // every position in (p, E) is clear; E itself unknown
}

template <class Env>
SPL_HD void spl_seg_extend(const Env& env, uint32_t N, SplSegEnd& s) {
    uint32_t lim = (N + 1 - s.E > s.step) ? s.E + s.step : N + 1;
    uint32_t x = env.next_hard(s.E, lim);
    if (x < lim) { s.E = x; s.exact = true; } else { s.E = lim; }
    if (s.step < (1u << 30)) s.step <<= 1;
}

template <class Env>
SPL_HD uint32_t spl_prev_limit(const Env& env, uint32_t i) {
    for (uint32_t k = 1; k <= 4 && k <= i; ++k)
        if (env.hard(i - k)) return i - k;
    return i >= 4 ? i - 4 : 0;
}

// One worker owns the pieces that start in [a, b): a = first sync point in its chunk
// [c0, c1), b = first sync point at or after c1.  Workers of all chunks together mark
// every piece start exactly once.
template <class Env>
SPL_HD void spl_pretok_chunk(Env& env, uint32_t c0, uint32_t c1, uint32_t N,
                             const uint8_t* s1, const uint8_t* s2, int pattern, bool with_special) {
    SplScanner<Env, true> sc(env, s1, s2, pattern);
    SplSegEnd se; se.E = 0; se.exact = false; se.step = 1024;
    bool have = false, found = false;
    uint32_t a = c0;
    for (; a < c1; ++a) {
        if (env.hard(a)) { found = true; have = false; break; }
        if (with_special && env.spec(a)) continue;                  // inside a special-token span
        if ((env.byte(a) & 0xC0u) == 0x80u) continue;
        if (!have) { spl_seg_locate(env, a, N, se); have = true; }
        uint32_t S = spl_prev_limit(env, a);
        bool r;
        for (;;) {
            sc.hit = false;
            r = sc.is_sync(a, S, se.E);
            if (!sc.hit || se.exact) break;
            spl_seg_extend(env, N, se);
        }
        if (r) { found = true; break; }
    }
    if (!found) return;
    uint32_t p = a;
    if (!have) spl_seg_locate(env, p, N, se);
    for (;;) {
        env.mark(p);
        uint32_t e;
        if (with_special && env.hard(p) && env.spec(p)) {
            while (!se.exact) spl_seg_extend(env, N, se);
            e = se.E;
        } else {
            for (;;) {
                sc.hit = false;
                e = sc.next_end(p, se.E);
                if (!sc.hit || se.exact) break;
                spl_seg_extend(env, N, se);
            }
        }
        p = e;
        if (p >= N) return;
        while (p == se.E && !se.exact) spl_seg_extend(env, N, se);
        if (p == se.E) {                                  // segment ends here: p is a hard boundary
            if (p >= c1) return;
            spl_seg_locate(env, p, N, se);
            continue;
        }
        if (p >= c1) {
            uint32_t S = spl_prev_limit(env, p);
            bool r;
            for (;;) {
                sc.hit = false;
                r = sc.is_sync(p, S, se.E);
                if (!sc.hit || se.exact) break;
                spl_seg_extend(env, N, se);
            }
            if (r) return;
        }
    }
}

// Lead-in for re-doing an isolated region [t0, ...) with the sequential rules: the pieces that start in
// [t0, first sync point at or after t0) belong to the worker of an EARLIER chunk.  This finds the last sync
// point before t0 (scanning backwards; a segment start always is one) and runs that worker.  Marks left of t0
// are repeats of what the owner of that text wrote -- piece starts are a function of the text only.
template <class Env>
SPL_HD void spl_pretok_leadin(Env& env, uint32_t t0, uint32_t N,
                              const uint8_t* s1, const uint8_t* s2, int pattern, bool with_special) {
    if (t0 == 0 || t0 >= N || env.hard(t0)) return;
    SplScanner<Env, true> sc(env, s1, s2, pattern);
    SplSegEnd se;
    spl_seg_locate(env, t0 - 1, N, se);
    uint32_t a = t0 - 1;
    for (;; --a) {
        if (a == 0 || env.hard(a)) break;
        if (with_special && env.spec(a)) continue;
        if ((env.byte(a) & 0xC0u) == 0x80u) continue;
        uint32_t S = spl_prev_limit(env, a);
        bool r;
        for (;;) {
            sc.hit = false;
            r = sc.is_sync(a, S, se.E);
            if (!sc.hit || se.exact) break;
            spl_seg_extend(env, N, se);
        }
        if (r) break;
    }
    spl_pretok_chunk(env, a, t0, N, s1, s2, pattern, with_special);
}
