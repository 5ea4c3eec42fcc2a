// Bit-parallel pre-tokenizer: the CL100K / O200K split patterns
// (/root/reference/src/core/tokenizer.rs:39, :42/:45) evaluated 32 bytes at a time on class
// bitmasks instead of one character at a time.  Replaces the regex engine's find_iter
// (tokenizer.rs:244-257) for the common case; anything this formulation does not model
// (see "fallback" below) is handed to the sequential rules of spl_pretok.h, tile by tile.
//
// One "word" = 32 consecutive text bytes = one 32-bit mask per property (bit i <-> byte i,
// LSB first).  Class masks are SMEARED over the continuation bytes of multi-byte characters
// and `lead` marks first bytes, so "class of the previous character" at a lead byte is simply
// the class bit of the previous byte.
//
// "Is a piece start" is then a position-local predicate over those masks plus five run
// propagations ("fills", computed with the add-carry trick inside a word and a summary walk
// across words):
//   ABS   CR/LF bytes swallowed by the [\r\n]* tail of a preceding punctuation piece   (forward)
//   BLC   whitespace bytes at or before the last CR/LF of their whitespace run         (backward)
//   PW    Lm/Lo bytes whose nearest non-Lm/Lo letter to the left is lower-case (O200K)  (forward)
//   ALLU  upper-case bytes followed only by upper-case letters to the end of the letter run (backward)
//   digit index mod 3 inside \p{N} runs                                                (forward)
// The derivation of the predicate from the regex alternatives is in DESIGN.md section 3.
//
// Fallback (the tile is re-done by the sequential scanner): invalid UTF-8, non-ASCII digits,
// \p{M} characters under O200K, chained contraction suffixes ('t't), a contraction directly
// followed by an Lm/Lo letter (O200K), and propagations that run past the window halo.
//
// Everything is __host__ __device__ so the exact device logic is fuzzed on the CPU against
// the oracle's regex engine (tests/test_pretok_fast_host.py).
#pragma once
#include "spl_pretok.h"

// per-word base masks (structure-of-arrays in the kernel's shared memory)
enum : int {
    FM_LEAD = 0, FM_UP, FM_LO, FM_BO, FM_NUM, FM_SP, FM_WSO, FM_CR, FM_OTH, FM_CT2, FM_CT3, FM_BAD,
    FM_A2, FM_A3,          // active contraction apostrophes (written by the local phase)
    FM_COUNT
};

struct SplFastWord { uint32_t m[FM_BAD + 1]; };

#define SPL_SWAR_ONES 0x01010101u
// bit 7 of every byte of `a` (all bytes < 0x80) whose value lies in [L, H] / equals C
#define SPL_RNG(a, L, H) ((((a) + (0x80u - (L)) * SPL_SWAR_ONES) & ~((a) + (0x7Fu - (H)) * SPL_SWAR_ONES)) & 0x80808080u)
#define SPL_EQ(a, C) (~((((a) ^ ((C) * SPL_SWAR_ONES))) + 0x7F7F7F7Fu) & 0x80808080u)

SPL_HD uint32_t spl_nib(uint32_t y) { return ((y >> 7) * 0x01020408u) >> 24; }   // bit7 flags of 4 bytes -> 4-bit mask

SPL_HD uint32_t spl_low_run(uint32_t r) { return r & ~(r + 1u); }                 // run of ones containing bit 0
SPL_HD uint32_t spl_brev(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
SPL_HD uint32_t spl_clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz(x);
#else
    return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}

// Forward fill inside one word, carry-in 0:  F(i) = S(i) | (R(i) & ~Bar(i) & F(i-1)),  F subset of R.
SPL_HD uint32_t spl_fill_fwd(uint32_t R, uint32_t S, uint32_t Bar) {
    uint32_t Rn = R & ~Bar;
    uint32_t Sb = S & Bar & R;                      // seeds on a barrier position still count and pass on
    uint32_t S2 = (S & Rn) | ((Sb << 1) & Rn);
    return ((((Rn + S2) ^ Rn) & Rn) | S2) | Sb;
}
// Backward fill: F(i) = S(i) | (R(i) & ~BarN(i) & F(i+1));  BarN(i) = "i must not receive from i+1".
SPL_HD uint32_t spl_fill_bwd(uint32_t R, uint32_t S, uint32_t BarN) {
    return spl_brev(spl_fill_fwd(spl_brev(R), spl_brev(S), spl_brev(BarN)));
}
// smear lead-byte flags over the continuation bytes that follow inside the word
SPL_HD uint32_t spl_smear(uint32_t f, uint32_t cont) {
    f |= (f << 1) & cont; f |= (f << 1) & cont; f |= (f << 1) & cont;
    return f;
}

// 8x8 bit-matrix transpose of 8 bytes held in a u64 (byte r of the result collects bit r of every input byte):
// with bit index 8 * row + column the three masked swaps exchange (row, column) with (column, row).
SPL_HD uint64_t spl_transpose8(uint64_t x) {
    uint64_t t;
    t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;  x = x ^ t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x = x ^ t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x = x ^ t ^ (t << 28);
    return x;
}
// xw[0..7] = 32 bytes as little-endian u32  ->  plane[q] bit i = bit q of byte i
SPL_HD void spl_bitplanes(const uint32_t* xw, uint32_t* plane) {
    uint64_t t[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] = spl_transpose8((uint64_t)xw[2 * k] | ((uint64_t)xw[2 * k + 1] << 32));
#pragma unroll
    for (int q = 0; q < 8; ++q)
        plane[q] = (uint32_t)((t[0] >> (8 * q)) & 0xFFu) | ((uint32_t)((t[1] >> (8 * q)) & 0xFFu) << 8) |
                   ((uint32_t)((t[2] >> (8 * q)) & 0xFFu) << 16) | ((uint32_t)((t[3] >> (8 * q)) & 0xFFu) << 24);
}

// ---- phase A: classify one word -------------------------------------------------------------------
// xw[0..7] = the word's 32 bytes as little-endian u32 (bytes at or beyond N may hold anything: they are masked).
// Text concept: uint8_t byte(uint32_t i) const  (any i < N) -- used only for non-ASCII characters and contractions.
template <class Text>
SPL_HD SplFastWord spl_fast_classify(const Text& t, const uint32_t* xw, uint32_t base, uint32_t N,
                                     const uint8_t* s1, const uint8_t* s2, int pattern) {
    SplFastWord w;
#pragma unroll
    for (int q = 0; q <= FM_BAD; ++q) w.m[q] = 0;
    if (base >= N) return w;
    const uint32_t valid = (N - base >= 32u) ? 0xFFFFFFFFu : ((1u << (N - base)) - 1u);
    // ASCII classes by bit slicing: transpose the 32 bytes into 8 bit planes (plane q, bit i = bit q of byte i), then
    // every class is a few boolean operations on whole planes.
    uint32_t b[8];
    spl_bitplanes(xw, b);
    const uint32_t asc = ~b[7], n6 = asc & ~b[6];
    const uint32_t hn0 = n6 & ~b[5] & ~b[4];                                   // 0x00..0x0F
    const uint32_t hn2 = n6 & b[5] & ~b[4];                                    // 0x20..0x2F
    const uint32_t low5_nz = b[4] | b[3] | b[2] | b[1] | b[0];
    const uint32_t low5_gt26 = b[4] & b[3] & (b[2] | (b[1] & b[0]));
    const uint32_t let = asc & b[6] & low5_nz & ~low5_gt26;                    // A-Z a-z
    const uint32_t m_lo = let & b[5], m_up = let & ~b[5];
    const uint32_t m_num = n6 & b[5] & b[4] & ~(b[3] & (b[2] | b[1]));         // 0-9
    const uint32_t z30 = ~b[3] & ~b[2] & ~b[1] & ~b[0];
    const uint32_t m_sp = hn2 & z30;                                           // 0x20
    const uint32_t m_ap = hn2 & ~b[3] & b[2] & b[1] & b[0];                    // 0x27
    const uint32_t m_ctl = hn0 & b[3] & ((~b[2] & (b[1] | b[0])) | (b[2] & ~b[1]));   // 0x09..0x0D
    const uint32_t m_cr = hn0 & b[3] & ((~b[2] & b[1] & ~b[0]) | (b[2] & ~b[1] & b[0]));   // 0x0A 0x0D
    uint32_t m_hi = b[7];
    m_hi &= valid;
    uint32_t lead = valid & ~m_hi;
    uint32_t up = m_up & valid, lo = m_lo & valid, bo = 0, num = m_num & valid, sp = m_sp & valid;
    uint32_t cr = m_cr & valid, wso = m_ctl & ~m_cr & valid, bad = 0;
    uint32_t oth = valid & ~m_hi & ~(up | lo | num | sp | cr | wso);
    if (m_hi) {
        // non-ASCII characters one by one (bytes are re-read through the cache)
        uint32_t i = 0;
        auto put = [&](uint32_t cls, uint32_t from, uint32_t to) {       // bytes [from, to) of this word
            uint32_t mk = ((to >= 32u) ? 0xFFFFFFFFu : ((1u << to) - 1u)) & ~((1u << from) - 1u) & valid;
            switch (cls) {
                case CLS_UPPER: up |= mk; break;
                case CLS_LOWER: lo |= mk; break;
                case CLS_BOTH:  bo |= mk; break;
                case CLS_NUM:   num |= mk; bad |= mk; break;             // char count != byte count inside \p{N}{1,3}
                case CLS_WS:    wso |= mk; break;
                case CLS_MARK:  oth |= mk; if (pattern != SPL_PAT_CL100K) bad |= mk; break;
                default:        oth |= mk; break;
            }
        };
        if ((m_hi & 1u) && (t.byte(base) & 0xC0u) == 0x80u) {           // a character straddles in from the previous word
            uint32_t j = 1;
            while (j < 3u && j < base && (t.byte(base - j) & 0xC0u) == 0x80u) ++j;
            bool ok = false;
            if (j <= base) {
                SplChar c = spl_decode(t, base - j, N, s1, s2);
                if (c.len > j) { put(c.cls, 0, c.len - j); i = c.len - j; ok = true; }
            }
            if (!ok) { bad |= 1u; oth |= 1u; lead |= 1u; i = 1; }
        }
        for (;;) {
            uint32_t rest = (i >= 32u) ? 0u : (m_hi & ~((1u << i) - 1u));
            if (!rest) break;
#if defined(__CUDA_ARCH__)
            uint32_t p = __ffs(rest) - 1;
#else
            uint32_t p = (uint32_t)__builtin_ctz(rest);
#endif
            SplChar c = spl_decode(t, base + p, N, s1, s2);
            lead |= 1u << p;
            if (c.len == 1) { bad |= 1u << p; oth |= 1u << p; }          // malformed byte
            else put(c.cls, p, p + c.len);
            i = p + c.len;
        }
    }
    uint32_t ct2 = 0, ct3 = 0;
    for (uint32_t m = m_ap & valid; m; m &= m - 1) {
#if defined(__CUDA_ARCH__)
        uint32_t p = __ffs(m) - 1;
#else
        uint32_t p = (uint32_t)__builtin_ctz(m);
#endif
        uint32_t k = spl_contraction(t, base + p, N);
        if (k == 2) ct2 |= 1u << p; else if (k == 3) ct3 |= 1u << p;
    }
    w.m[FM_LEAD] = lead; w.m[FM_UP] = up; w.m[FM_LO] = lo; w.m[FM_BO] = bo; w.m[FM_NUM] = num;
    w.m[FM_SP] = sp; w.m[FM_WSO] = wso; w.m[FM_CR] = cr; w.m[FM_OTH] = oth;
    w.m[FM_CT2] = ct2; w.m[FM_CT3] = ct3; w.m[FM_BAD] = bad;
    return w;
}

// ---- phase B: local masks, fills with carry-in 0, per-word summary ------------------------------------
// summary bits
#define FS_GEN_ABS   (1u << 0)
#define FS_PROP_ABS  (1u << 1)
#define FS_GEN_PW    (1u << 2)
#define FS_PROP_PW   (1u << 3)
#define FS_GEN_BLC   (1u << 4)
#define FS_PROP_BLC  (1u << 5)
#define FS_GEN_ALLU  (1u << 6)
#define FS_PROP_ALLU (1u << 7)
#define FS_OSS_TOP   (1u << 8)     // byte 31 belongs to an o_start character
#define FS_NTAIL_SH  9             // 6 bits: digits at the top of the word connected to byte 31
#define FS_NALL      (1u << 15)    // the whole word continues one digit run
#define FS_BAD       (1u << 16)

struct SplFastLocal {
    uint32_t H, Hn, LD, L, WS, up, lo, bo, num, cr, oth;
    uint32_t pL, pN, pO, pCR, pWS, plo, pbo;            // previous-byte class masks (cleared at segment starts)
    uint32_t OS, OSs, A2, A3, NC;
    uint32_t ABS0, PW0, BLC0, ALLU0;                    // fills with carry-in 0
    uint32_t rn_abs, rn_pw, rn_blc, rn_allu;            // receiving positions of each fill
    uint32_t s3;                                        // lead bytes of "last blank before a non-blank"
    uint32_t valid_bad;
};

// M[q][k] = mask q of window word k (0 <= k < nw); hard[k] / spec[k] = segment-start and special-span bits.
// Out-of-window neighbours read as 0.
template <class Masks>
SPL_HD uint32_t spl_fast_local(const Masks& M, int k, int nw, int pattern, SplFastLocal& o) {
    auto cur = [&](int q) { return M.get(q, k); };
    auto prv = [&](int q) { return k > 0 ? M.get(q, k - 1) : 0u; };
    auto nxt = [&](int q) { return k + 1 < nw ? M.get(q, k + 1) : 0u; };
    const uint32_t H = M.hard(k), Hnext = (k + 1 < nw) ? M.hard(k + 1) : 0u;
    const uint32_t Hn = (H >> 1) | (Hnext << 31), Hn2 = (H >> 2) | (Hnext << 30);
#define PV(q) (((cur(q) << 1) | (prv(q) >> 31)) & ~H)
    o.H = H; o.Hn = Hn;
    o.LD = cur(FM_LEAD);
    o.up = cur(FM_UP); o.lo = cur(FM_LO); o.bo = cur(FM_BO); o.num = cur(FM_NUM); o.cr = cur(FM_CR); o.oth = cur(FM_OTH);
    o.L = o.up | o.lo | o.bo;
    o.WS = cur(FM_SP) | cur(FM_WSO) | o.cr;
    o.plo = PV(FM_LO); o.pbo = PV(FM_BO);
    o.pL = PV(FM_UP) | o.plo | o.pbo;
    o.pN = PV(FM_NUM);
    o.pO = PV(FM_OTH);
    o.pCR = PV(FM_CR);
    const uint32_t pSP = PV(FM_SP);
    o.pWS = pSP | PV(FM_WSO) | o.pCR;
    // o_start: a punctuation character that begins a piece (no punctuation and no ' ' before it)
    o.OS = o.oth & o.LD & ~(o.pO | pSP);
    o.OSs = spl_smear(o.OS, ~o.LD);
    // contraction patterns that lie inside one segment
    const uint32_t CT2 = cur(FM_CT2) & ~Hn, CT3 = cur(FM_CT3) & ~Hn & ~Hn2;
    if (pattern == SPL_PAT_CL100K) { o.A2 = CT2 & o.OS; o.A3 = CT3 & o.OS; }        // alternative 1, tried at a piece start
    else                           { o.A2 = CT2 & o.pL; o.A3 = CT3 & o.pL; }        // suffix of a letter piece
    o.NC = o.num & o.pN;                                                              // digit continuing a digit run
    // fills, carry-in 0
    o.rn_abs = o.cr & ~H;
    o.ABS0 = spl_fill_fwd(o.cr, o.cr & o.pO, H);
    o.rn_pw = o.bo & ~H;
    o.PW0 = spl_fill_fwd(o.bo, o.bo & o.plo, H);
    o.rn_blc = o.WS & ~Hn;
    o.BLC0 = spl_fill_bwd(o.WS, o.cr, Hn);
    // (the last window word cannot see its successor: no seed at byte 31, so a run that reaches it stays "unknown")
    const uint32_t Lnext = (o.L >> 1) | ((k + 1 < nw ? (nxt(FM_UP) | nxt(FM_LO) | nxt(FM_BO)) : 1u) << 31);
    o.rn_allu = o.up & ~Hn;
    o.ALLU0 = spl_fill_bwd(o.up, o.up & ~(Lnext & ~Hn), Hn);
    {
        // S3: the last blank (not CR/LF) of a whitespace run that is followed, inside the segment, by a non-blank.
        // Evaluated at the character's LAST byte, then moved back to its lead byte.
        const uint32_t nWS = nxt(FM_SP) | nxt(FM_WSO) | nxt(FM_CR);
        const uint32_t nV = (k + 1 < nw) ? M.valid(k + 1) : 0u;
        const uint32_t WSnext = (o.WS >> 1) | (nWS << 31);
        const uint32_t Vnext = (M.valid(k) >> 1) | (nV << 31);
        uint32_t f = o.WS & ~o.cr & ~WSnext & ~Hn & Vnext;
        f |= (f & ~o.LD) >> 1; f |= (f & ~o.LD) >> 1; f |= (f & ~o.LD) >> 1;
        f &= o.LD;
        // a blank whose lead is here but whose last byte lies in the next word
        const uint32_t nlead = nxt(FM_LEAD);
        if (k + 1 < nw && !(nlead & 1u) && (nV & 1u)) {
#if defined(__CUDA_ARCH__)
            uint32_t c = nlead ? (uint32_t)(__ffs(nlead) - 1) : 32u;
#else
            uint32_t c = nlead ? (uint32_t)__builtin_ctz(nlead) : 32u;
#endif
            if (c <= 3u && o.LD) {
                uint32_t last = 1u << (c - 1), after = 1u << c;
                bool blank = (nWS & last) && !(nxt(FM_CR) & last);
                bool follows = !(nWS & after) && !(Hnext & after) && (nV & after);
                if (blank && follows) f |= 1u << (31 - spl_clz32(o.LD));
            }
        }
        o.s3 = f;
    }
    o.valid_bad = cur(FM_BAD);
    // digits at the top of the word that connect to byte 31 (byte 0's own link to the previous word is the
    // business of whoever walks the summaries: FS_NALL only says "32 digits, and byte 0 may continue a run")
    uint32_t ntail = 0, nall = 0;
    if (o.num >> 31) {
        ntail = spl_clz32(~o.NC | 1u) + 1;           // bytes p..31, p = highest position that does not continue a run
        nall = (ntail == 32u) && !(H & 1u);
    }
    uint32_t s = 0;
    if (o.ABS0 >> 31) s |= FS_GEN_ABS;
    if (o.rn_abs == 0xFFFFFFFFu) s |= FS_PROP_ABS;
    if (o.PW0 >> 31) s |= FS_GEN_PW;
    if (o.rn_pw == 0xFFFFFFFFu) s |= FS_PROP_PW;
    if (o.BLC0 & 1u) s |= FS_GEN_BLC;
    if (o.rn_blc == 0xFFFFFFFFu) s |= FS_PROP_BLC;
    if (o.ALLU0 & 1u) s |= FS_GEN_ALLU;
    if (o.rn_allu == 0xFFFFFFFFu) s |= FS_PROP_ALLU;
    if (o.OSs >> 31) s |= FS_OSS_TOP;
    s |= ntail << FS_NTAIL_SH;
    if (nall) s |= FS_NALL;
    if (o.valid_bad) s |= FS_BAD;
#undef PV
    return s;
}

// ---- phase C: resolve carries from the summaries, assemble the piece-start mask ---------------------
// Returns the piece-start bits of word k.  Two kinds of "cannot decide here":
//   structural  a construct the predicate does not model was seen in word k; its effect reaches a few bytes (or
//               one Lm/Lo run) to the right, so the caller evaluates this for EVERY window word, halo included
//   unknown     a propagation that word k consumes enters from outside the window (payload words only matter)
// A2/A3 of the neighbouring words are read through M (FM_A2 / FM_A3, written after phase B).
template <class Masks>
SPL_HD uint32_t spl_fast_final(const Masks& M, const SplFastLocal& o, int k, int nw, int pattern, bool with_special,
                               bool& structural, bool& unknown) {
    const uint32_t H = o.H;
    // --- carries -----------------------------------------------------------------------------------
    const bool o200k = pattern != SPL_PAT_CL100K;
    uint32_t cin_abs = 0, cin_pw = 0, cin_blc = 0, cin_allu = 0;
    const uint32_t low_pw = spl_low_run(o.rn_pw);
    // a carry matters only where it is consumed: ABS by the CR/LF run at byte 0, PW by an upper-case letter right
    // after the Lm/Lo run at byte 0 (or at byte 0 itself), BLC / ALLU by the run that ends at byte 31
    bool unk_abs = (o.WS & ~H & 1u) != 0;                     // byte 0 continues the run or asks p(ABS)
    bool unk_pw = o200k && ((((low_pw << 1) | 1u) & o.up) != 0);
    bool unk_blc = (o.rn_blc >> 31) != 0;
    bool unk_allu = o200k && (o.rn_allu >> 31) != 0;
    for (int j = k - 1; j >= 0 && (unk_abs || unk_pw); --j) {
        uint32_t s = M.summary(j);
        if (unk_abs) { if (s & FS_GEN_ABS) { cin_abs = 1; unk_abs = false; } else if (!(s & FS_PROP_ABS)) unk_abs = false; }
        if (unk_pw)  { if (s & FS_GEN_PW)  { cin_pw = 1;  unk_pw = false; }  else if (!(s & FS_PROP_PW))  unk_pw = false; }
    }
    for (int j = k + 1; j < nw && (unk_blc || unk_allu); ++j) {
        uint32_t s = M.summary(j);
        if (unk_blc)  { if (s & FS_GEN_BLC)  { cin_blc = 1;  unk_blc = false; }  else if (!(s & FS_PROP_BLC))  unk_blc = false; }
        if (unk_allu) { if (s & FS_GEN_ALLU) { cin_allu = 1; unk_allu = false; } else if (!(s & FS_PROP_ALLU)) unk_allu = false; }
    }
    if (unk_abs || unk_pw || unk_blc || unk_allu) unknown = true;     // the run leaves the window: undecidable here
    const uint32_t psum = k > 0 ? M.summary(k - 1) : 0u;
    const uint32_t ABS = o.ABS0 | (cin_abs ? spl_low_run(o.rn_abs) : 0u);
    const uint32_t PW = o.PW0 | (cin_pw ? low_pw : 0u);
    const uint32_t BLC = o.BLC0 | (cin_blc ? spl_brev(spl_low_run(spl_brev(o.rn_blc))) : 0u);
    const uint32_t ALLU = o.ALLU0 | (cin_allu ? spl_brev(spl_low_run(spl_brev(o.rn_allu))) : 0u);
    const uint32_t pABS = ((ABS << 1) | cin_abs) & ~H;
    const uint32_t pPW = ((PW << 1) | cin_pw) & ~H;
    // BLC of the previous byte: it is a CR/LF, or a blank that receives from this byte
    const uint32_t pBLC = ((BLC << 1) | ((o.pCR | (o.pWS & BLC)) & 1u)) & ~H;
    // --- contraction ends and interiors ------------------------------------------------------------
    const uint32_t pA2 = k > 0 ? M.get(FM_A2, k - 1) : 0u, pA3 = k > 0 ? M.get(FM_A3, k - 1) : 0u;
    const uint32_t CE = (o.A2 << 2) | (pA2 >> 30) | (o.A3 << 3) | (pA3 >> 29);
    const uint32_t INCT = (((o.A2 | o.A3) << 1) | ((pA2 | pA3) >> 31)) | ((o.A3 << 2) | (pA3 >> 30));
    uint32_t OSx = o.OS;
    if (o200k) {
        if (CE & (o.A2 | o.A3)) structural = true;               // chained suffixes: activity alternates along the chain
        if (CE & o.bo) structural = true;                        // a restart inside an Lm/Lo run changes the case-boundary state
        OSx &= ~(o.A2 | o.A3);                                  // the apostrophe of a suffix belongs to the letter piece
    }
    // --- letters ---------------------------------------------------------------------------------------
    const uint32_t oss_in = (psum & FS_OSS_TOP) ? 1u : 0u;               // an o_start character may straddle in
    const uint32_t OSs = o.OSs | (oss_in ? spl_low_run(~o.LD) : 0u);
    const uint32_t pOSs = ((OSs << 1) | oss_in) & ~H;
    uint32_t LS = H | o.pCR | o.pN | (o.pO & ~pOSs);
    if (!o200k) LS |= o.pL & CE;
    else {
        const uint32_t CB = o.up & (o.plo | pPW);              // lower (through Lm/Lo) then upper: "camelCase"
        const uint32_t TR = o.up & o.pbo & ALLU;               // "...好ABC" at the end of a letter run
        LS |= o.pL & (CB | TR | CE);
    }
    LS &= o.L;
    // --- digits: \p{N}{1,3} ------------------------------------------------------------------------------
    uint32_t Z0 = o.num & ~o.NC, Z1 = 0, Z2 = 0;
    if (o.NC & 1u) {
        uint32_t len = 0; bool known = false;
        for (int j = k - 1; j >= 0; --j) {
            uint32_t s = M.summary(j);
            len += (s >> FS_NTAIL_SH) & 63u;
            if (!(s & FS_NALL)) { known = true; break; }
        }
        if (!known) unknown = true;
        uint32_t r = len % 3u;
        if (r == 0) Z0 |= 1u; else if (r == 1) Z1 |= 1u; else Z2 |= 1u;
    }
    for (;;) {
        uint32_t n1 = (Z0 << 1) & o.NC & ~Z1, n2 = (Z1 << 1) & o.NC & ~Z2, n0 = (Z2 << 1) & o.NC & ~Z0;
        if (!(n0 | n1 | n2)) break;
        Z0 |= n0; Z1 |= n1; Z2 |= n2;
    }
    const uint32_t NS = Z0;
    // --- whitespace:  \s*[\r\n]+ | \s+(?!\S) | \s+ -----------------------------------------------------------
    const uint32_t S1 = o.WS & ~ABS & (~o.pWS | pABS);        // first whitespace not swallowed by a punctuation tail
    const uint32_t S2 = o.WS & ~BLC & pBLC;                    // right after the last CR/LF of the run
    uint32_t start = (LS | NS | OSx | S1 | S2 | o.s3 | H) & o.LD & ~INCT & M.valid(k);
    if (with_special) start &= ~M.spec(k) | H;
    return start;
}
