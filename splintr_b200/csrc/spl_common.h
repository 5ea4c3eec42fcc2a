// Shared host/device definitions for the splintr_b200 encode path:
// character classes, table entry layouts and the hash functions that the host-side
// table builder and the device-side probes must agree on bit for bit.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define SPL_HD __host__ __device__ __forceinline__
#else
#define SPL_HD inline
#endif

// ---- character classes (4 bit; see tools/gen_unicode_tables.py) --------------------
enum : uint8_t {
    CLS_OTHER = 0,   // not whitespace / letter / number / mark
    CLS_CRLF  = 1,   // \r \n
    CLS_SPACE = 2,   // U+0020
    CLS_WS    = 3,   // other White_Space
    CLS_UPPER = 4,   // Lu | Lt
    CLS_LOWER = 5,   // Ll
    CLS_BOTH  = 6,   // Lm | Lo   (member of both the "upper" and "lower" sets of O200K)
    CLS_MARK  = 7,   // M         (same, but not \p{L})
    CLS_NUM   = 8,   // N
    CLS_APOS  = 9,   // '
    CLS_SLASH = 10,  // /         (tail of MISTRAL_V3's punctuation alternative)
};

// class-set bit masks
#define CM(c) (1u << (c))
#define CSET_WS     (CM(CLS_CRLF) | CM(CLS_SPACE) | CM(CLS_WS))
#define CSET_L      (CM(CLS_UPPER) | CM(CLS_LOWER) | CM(CLS_BOTH))
#define CSET_U      (CM(CLS_UPPER) | CM(CLS_BOTH) | CM(CLS_MARK))
#define CSET_W      (CM(CLS_LOWER) | CM(CLS_BOTH) | CM(CLS_MARK))
#define CSET_BOTH   (CM(CLS_BOTH) | CM(CLS_MARK))
#define CSET_O      (CM(CLS_OTHER) | CM(CLS_MARK) | CM(CLS_APOS) | CM(CLS_SLASH))
// [^\r\n\p{L}\p{N}] : the optional one-char prefix of the letter alternatives
#define CSET_PREFIX (CM(CLS_OTHER) | CM(CLS_SPACE) | CM(CLS_WS) | CM(CLS_MARK) | CM(CLS_APOS) | CM(CLS_SLASH))

SPL_HD bool in_set(uint32_t cls, uint32_t set) { return (set >> cls) & 1u; }

// ---- split patterns (C-ABI ids) ------------------------------------------------------
enum : int { SPL_PAT_CL100K = 0, SPL_PAT_O200K = 1, SPL_PAT_MISTRAL_V3 = 2,
             SPL_PAT_SENTENCEPIECE = 3 };   // `[^\s]+|\s+` (tokenizer.rs:56) + the U+2581 walk of tokenizer.rs:737-795

// ---- symbols ---------------------------------------------------------------------------
// A "symbol" is a token id (= merge rank) or, for a single byte that is not in the
// vocabulary, the pseudo id SPL_UNK_BASE + byte (never emitted; reference drops it,
// bpe.rs:187-191).  21 bits each so that (left, right, merged) pack into one u64.
#define SPL_SYM_BITS   21
#define SPL_UNK_BASE   ((1u << SPL_SYM_BITS) - 256u)
#define SPL_RANK_NONE  0xFFFFFFFFu

SPL_HD uint64_t spl_mix64(uint64_t x) {
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}

// pair table: (left symbol, right symbol) -> merged symbol, in buckets of SPL_PAIR_WAYS entries = one 32-byte sector:
// four 32-bit TAGS followed by four 32-bit VALUES, so that the device reads a bucket with ONE 256-bit load
// (LDG.E.256, sm_100) and finds a key with four 32-bit compares:
//     tag   = right:21 | (left & 0x7FF) << 21
//     value = merged:21 | (left >> 11) << 21          (bit 31 clear; an empty slot is tag = value = 0xFFFFFFFF)
// A bucket fills front to back; a key lives in the first bucket from its home bucket on that was not full when it
// was inserted, so a probe ends at the first bucket whose last slot is empty.  `log2size` counts BUCKETS.
#define SPL_PAIR_WAYS 4
#define SPL_PAIR_WORDS (2 * SPL_PAIR_WAYS)
#define SPL_PAIR_EMPTY 0xFFFFFFFFu
#define SPL_SYM_MASK ((1u << SPL_SYM_BITS) - 1u)
SPL_HD uint32_t spl_pair_tag(uint32_t l, uint32_t r) { return r | (l << SPL_SYM_BITS); }
SPL_HD uint32_t spl_pair_hi(uint32_t l) { return (l >> (32 - SPL_SYM_BITS)) << SPL_SYM_BITS; }        // the value's key bits
SPL_HD uint32_t spl_pair_hash(uint32_t l, uint32_t r, uint32_t log2size) {
    uint32_t x = l * 0x9E3779B1u + r * 0x85EBCA77u;
    x ^= x >> 16;
    x *= 0x2C1B3C6Du;
    return x >> (32 - log2size);
}

// whole-piece tables.  Keys are the piece bytes packed little-endian into u64 words,
// zero padded; `len` disambiguates padding from NUL bytes.  len == 0 marks an empty slot.
// T8 is bucketed like the pair table: SPL_T8_WAYS entries (one 32-byte sector) per bucket, t8_log2 counts buckets,
// keys are inserted in id order so that the frequent (low-rank) tokens sit in their home bucket.
struct SplKey8  { uint64_t k0; uint32_t id; uint32_t len; };                       // len 1..8
struct SplKey16 { uint64_t k0; uint64_t k1; uint32_t id; uint32_t len; uint64_t pad; };   // len 9..16
struct SplKeyL  { uint64_t hash; uint32_t id; uint32_t len; };                    // len 17..max, verified against token bytes

#define SPL_T8_WAYS 2
SPL_HD uint32_t spl_hash8(uint32_t lo, uint32_t hi, uint32_t len, uint32_t log2size) {
    uint32_t x = lo * 0x9E3779B1u + hi * 0x85EBCA77u + len * 0xC2B2AE3Du;
    x ^= x >> 15;
    x *= 0x2C1B3C6Du;
    x ^= x >> 13;
    x *= 0x297A2D39u;
    return x >> (32 - log2size);
}
SPL_HD uint32_t spl_hash16(uint64_t k0, uint64_t k1, uint32_t len, uint32_t log2size) {
    uint64_t h = (k0 + len) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    h += k1 * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 31;
    h *= 0xBF58476D1CE4E5B9ull;
    return (uint32_t)(h >> (64 - log2size));
}
// long keys: order-independent sum of position-salted word mixes (so a warp can compute
// it cooperatively), then a final mix.  `w` = i-th little-endian u64 word, zero padded.
SPL_HD uint64_t spl_hashL_word(uint64_t w, uint32_t i) {
    return spl_mix64(w + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1));
}
SPL_HD uint64_t spl_hashL_final(uint64_t sum, uint32_t len) {
    return spl_mix64(sum ^ ((uint64_t)len << 56) ^ len);
}

// ---- device-resident tables ----------------------------------------------------------
struct SplTables {
    // pre-tokenizer
    const uint8_t* ucd_stage1;     // [0x1100]
    const uint8_t* ucd_stage2;     // [nblocks*128], two 4-bit classes per byte
    // whole-piece lookup
    const SplKey8*  t8;   uint32_t t8_log2;
    const SplKey16* t16;  uint32_t t16_log2;
    const SplKeyL*  tl;   uint32_t tl_log2;
    const uint8_t*  tok_bytes;     // concatenated (raw) token bytes
    const uint32_t* tok_off;       // [n_ids+1] offsets into tok_bytes by token id
    uint32_t n_ids;                // max mergeable id + 1
    uint32_t max_key_len;          // longest vocabulary key in bytes
    // BPE
    const uint32_t* pair;  uint32_t pair_log2;     // buckets of SPL_PAIR_WORDS words
    const uint32_t* bpair;         // [65536] dense: rank of the pair (byte b0, byte b1) at [b0 << 8 | b1], SPL_RANK_NONE if none
    uint32_t byte_sym[256];        // symbol of each single byte
    // independent segments (spl_segment.h)
    const uint32_t* seg_irr;       // [2048] words
    const uint32_t* seg_h2;  uint32_t seg_h2_log2;   // hashed bitmap, 2^seg_h2_log2 bits
    const uint32_t* char_tok;      // [65536] what the merge loop makes of the UTF-8 bytes of code point c (SPL_CHAR_* in spl_segment.h)
    const uint32_t* char_ids;      // the id lists of the characters that are two or three ids
    // decode: id -> bytes
    const uint8_t*  dec_bytes;
    const uint32_t* dec_off;       // [n_dec + 1]
    uint32_t n_dec;
    // special tokens (encode_with_special)
    const uint8_t*  sp_bytes;      // concatenated special strings
    const uint32_t* sp_off;        // [n_special+1]
    const uint32_t* sp_id;         // [n_special]
    uint32_t n_special;
    uint32_t sp_first[8];          // 256-bit set: bytes that start some special string
    int pattern;
};

// ASCII class LUT (0..127); built once, identical to the generated table's first block.
SPL_HD uint8_t spl_ascii_class(uint32_t b) {
    if (b >= 'a' && b <= 'z') return CLS_LOWER;
    if (b >= 'A' && b <= 'Z') return CLS_UPPER;
    if (b >= '0' && b <= '9') return CLS_NUM;
    if (b == ' ') return CLS_SPACE;
    if (b == '\n' || b == '\r') return CLS_CRLF;
    if (b == '\t' || b == 0x0B || b == 0x0C) return CLS_WS;
    if (b == '\'') return CLS_APOS;
    if (b == '/') return CLS_SLASH;
    return CLS_OTHER;
}
