// JSON Lines ingestion: the parser of ONE line, __host__ __device__ so that the exact
// device logic is fuzzed on the CPU against Python's json module (tests/test_ingest_host.py).
//
// What it replaces is not in the reference: it is the loop a splintr user runs in front of `encode_batch`,
//     texts = [json.loads(line)[field] for line in open(path) if line.strip()]
// followed by the packing of `texts` into one buffer.  Here the file bytes go to the device as they are and the packed
// UTF-8 text + document offsets that spl_encode_batch_device consumes are produced there.
//
// Contract (per line; lines end at '\n', the last one may lack it):
//   * a line of only blanks ( \t\r) is skipped: no document;
//   * otherwise the line is one JSON object and yields one document: the string value of the LAST top-level member
//     whose (unescaped) name equals `field` -- json.loads semantics for duplicate names -- with the JSON escapes
//     \" \\ \/ \b \f \n \r \t \uXXXX resolved (surrogate pairs -> one 4-byte UTF-8 sequence, a lone surrogate ->
//     U+FFFD); bytes >= 0x80 are copied as they are (the file is UTF-8);
//   * a line without such a member, with a non-string value there, or that does not parse, yields an EMPTY document
//     and is counted (SPL_JL_MISSING / SPL_JL_BAD); nothing is validated beyond what finding the member needs.
// The unescaped value is never longer than its escaped form, so the packed text fits in a buffer of the file's size.
#pragma once
#include "spl_common.h"

#define SPL_JL_FIELD_MAX 64u
enum : uint32_t { SPL_JL_DOC = 1u, SPL_JL_FOUND = 2u, SPL_JL_BAD = 4u };

struct SplJlSpan { uint32_t vs, ve, out_len, flags; };      // raw value bytes [vs, ve) between the quotes

SPL_HD bool spl_jl_blank(uint32_t b) { return b == ' ' || b == '\t' || b == '\r' || b == '\n'; }

SPL_HD int spl_jl_hex(uint32_t c) {
    if (c >= '0' && c <= '9') return (int)(c - '0');
    c |= 0x20u;
    if (c >= 'a' && c <= 'f') return (int)(c - 'a' + 10);
    return -1;
}

// One unescaped character of a JSON string body starting at i (i < e, text[i] is not the closing quote).
// out[0..n) receives its UTF-8 bytes (n = 1..4); returns the position after it.  Malformed escapes are copied raw.
template <class T>
SPL_HD uint32_t spl_jl_char(const T& t, uint32_t i, uint32_t e, uint8_t* out, uint32_t& n) {
    uint32_t b = t.byte(i);
    if (b != '\\' || i + 1 >= e) { out[0] = (uint8_t)b; n = 1; return i + 1; }
    uint32_t c = t.byte(i + 1);
    n = 1;
    switch (c) {
        case 'b': out[0] = 8; return i + 2;
        case 'f': out[0] = 12; return i + 2;
        case 'n': out[0] = 10; return i + 2;
        case 'r': out[0] = 13; return i + 2;
        case 't': out[0] = 9; return i + 2;
        case 'u': break;
        default: out[0] = (uint8_t)c; return i + 2;          // \" \\ \/ (and, leniently, any other character)
    }
    if (i + 6 > e) { out[0] = (uint8_t)b; return i + 1; }
    int h0 = spl_jl_hex(t.byte(i + 2)), h1 = spl_jl_hex(t.byte(i + 3)), h2 = spl_jl_hex(t.byte(i + 4)), h3 = spl_jl_hex(t.byte(i + 5));
    if ((h0 | h1 | h2 | h3) < 0) { out[0] = (uint8_t)b; return i + 1; }
    uint32_t cp = (uint32_t)((h0 << 12) | (h1 << 8) | (h2 << 4) | h3), next = i + 6;
    if (cp >= 0xD800u && cp <= 0xDBFFu) {                    // high surrogate: needs \uDC00..\uDFFF right behind it
        uint32_t lo = 0;
        bool pair = false;
        if (i + 12 <= e && t.byte(i + 6) == '\\' && t.byte(i + 7) == 'u') {
            int g0 = spl_jl_hex(t.byte(i + 8)), g1 = spl_jl_hex(t.byte(i + 9)), g2 = spl_jl_hex(t.byte(i + 10)), g3 = spl_jl_hex(t.byte(i + 11));
            if ((g0 | g1 | g2 | g3) >= 0) {
                lo = (uint32_t)((g0 << 12) | (g1 << 8) | (g2 << 4) | g3);
                pair = lo >= 0xDC00u && lo <= 0xDFFFu;
            }
        }
        if (pair) { cp = 0x10000u + ((cp - 0xD800u) << 10) + (lo - 0xDC00u); next = i + 12; }
        else cp = 0xFFFDu;
    } else if (cp >= 0xDC00u && cp <= 0xDFFFu) {
        cp = 0xFFFDu;                                        // lone low surrogate
    }
    if (cp < 0x80u) { out[0] = (uint8_t)cp; n = 1; }
    else if (cp < 0x800u) { out[0] = (uint8_t)(0xC0u | (cp >> 6)); out[1] = (uint8_t)(0x80u | (cp & 63u)); n = 2; }
    else if (cp < 0x10000u) {
        out[0] = (uint8_t)(0xE0u | (cp >> 12)); out[1] = (uint8_t)(0x80u | ((cp >> 6) & 63u)); out[2] = (uint8_t)(0x80u | (cp & 63u)); n = 3;
    } else {
        out[0] = (uint8_t)(0xF0u | (cp >> 18)); out[1] = (uint8_t)(0x80u | ((cp >> 12) & 63u));
        out[2] = (uint8_t)(0x80u | ((cp >> 6) & 63u)); out[3] = (uint8_t)(0x80u | (cp & 63u)); n = 4;
    }
    return next;
}

// bytes of `w` (eight text bytes, little endian) equal to c, flagged in bit 7 of their byte
SPL_HD uint64_t spl_jl_has(uint64_t w, uint32_t c) {
    const uint64_t x = w ^ (0x0101010101010101ull * c);
    return (x - 0x0101010101010101ull) & ~x & 0x8080808080808080ull;
}

// Parse the line [s, e) (no '\n' inside).  field[0..flen) = the member name wanted.
//
// Written as ONE loop over the line with an explicit state, not as nested scanning loops: on the device one thread
// parses one line, and the 32 lines of a warp are at different places of their objects; with nested loops the lanes
// never reconverge and the warp executes them one after the other (measured: 3.3 of 32 lanes active).  Here every
// lane goes round the same loop and the divergence is confined to the body of one iteration.  Inside strings the
// loop advances eight bytes at a time while they hold no quote and no escape (T::word8(i) = the eight bytes from i
// on; only called with i + 16 <= e).
template <class T>
SPL_HD SplJlSpan spl_jl_parse_line(const T& t, uint32_t s, uint32_t e, const uint8_t* field, uint32_t flen) {
    enum : uint32_t { ST_START, ST_MEMBER, ST_KEY, ST_COLON, ST_VALUE, ST_VSTR, ST_SCALAR, ST_NESTED, ST_NSTR, ST_AFTER, ST_DONE, ST_BAD };
    SplJlSpan r;
    r.vs = r.ve = r.out_len = r.flags = 0;
    uint32_t st = ST_START, i = s, f = 0, depth = 0, vstart = 0;
    bool doc = false, kmatch = true, match = false, found = false;
    while (i < e && st < ST_DONE) {
        const uint32_t b = t.byte(i);
        switch (st) {
            case ST_START:                                   // blanks, then the opening brace
                if (spl_jl_blank(b)) { ++i; break; }
                doc = true;
                if (b == '{') { st = ST_MEMBER; ++i; } else st = ST_BAD;
                break;
            case ST_MEMBER:                                  // a member name, or the closing brace
                if (spl_jl_blank(b)) { ++i; break; }
                if (b == '}') { st = ST_DONE; break; }
                if (b == '"') { st = ST_KEY; f = 0; kmatch = true; ++i; } else st = ST_BAD;
                break;
            case ST_KEY: {                                   // member name: compared unescaped
                if (b == '"') { match = kmatch && f == flen; st = ST_COLON; ++i; break; }
                uint8_t ch[4];
                uint32_t n = 1, ni = i + 1;
                ch[0] = (uint8_t)b;
                if (b == '\\') ni = spl_jl_char(t, i, e, ch, n);
                for (uint32_t q = 0; q < n; ++q) {
                    if (kmatch && f < flen && field[f] == ch[q]) ++f; else kmatch = false;
                }
                i = ni;
                break;
            }
            case ST_COLON:
                if (spl_jl_blank(b)) { ++i; break; }
                if (b == ':') { st = ST_VALUE; ++i; } else st = ST_BAD;
                break;
            case ST_VALUE:                                   // first character of the value decides
                if (spl_jl_blank(b)) { ++i; break; }
                if (b == '"') { st = ST_VSTR; vstart = i + 1; ++i; break; }
                if (match) found = false;                    // the last member of that name wins, and it is not a string
                if (b == '{' || b == '[') { depth = 1; st = ST_NESTED; ++i; } else st = ST_SCALAR;
                break;
            case ST_VSTR:
            case ST_NSTR:                                    // inside a string (a value / somewhere in a nested value)
                if (i + 16u <= e) {
                    const uint64_t w = t.word8(i);
                    if (!(spl_jl_has(w, '"') | spl_jl_has(w, '\\'))) { i += 8u; break; }
                }
                if (b == '"') {
                    if (st == ST_VSTR) { if (match) { r.vs = vstart; r.ve = i; found = true; } st = ST_AFTER; }
                    else st = ST_NESTED;
                    ++i;
                } else {
                    i += (b == '\\') ? 2u : 1u;
                }
                break;
            case ST_SCALAR:                                  // number / true / false / null: up to , } or a blank
                if (b == ',' || b == '}' || spl_jl_blank(b)) st = ST_AFTER; else ++i;
                break;
            case ST_NESTED:                                  // object or array value: skipped with a depth count
                if (b == '"') st = ST_NSTR;
                else if (b == '{' || b == '[') ++depth;
                else if (b == '}' || b == ']') { if (--depth == 0) st = ST_AFTER; }
                ++i;
                break;
            default:                                         // ST_AFTER: a comma or the closing brace
                if (spl_jl_blank(b)) { ++i; break; }
                if (b == ',') { st = ST_MEMBER; ++i; }
                else if (b == '}') st = ST_DONE;
                else st = ST_BAD;
                break;
        }
    }
    if (!doc) return r;                                      // blank line: no document
    r.flags = SPL_JL_DOC;
    if (st != ST_DONE) r.flags |= SPL_JL_BAD;                // ran into the end of the line, or saw something unexpected
    if (found && !(r.flags & SPL_JL_BAD)) {
        r.flags |= SPL_JL_FOUND;
        uint32_t j = r.vs, len = 0;
        while (j < r.ve) {
            if (j + 16u <= r.ve && !spl_jl_has(t.word8(j), '\\')) { j += 8u; len += 8u; continue; }   // no escape: as it is
            uint8_t ch[4]; uint32_t n;
            j = spl_jl_char(t, j, r.ve, ch, n);
            len += n;
        }
        r.out_len = len;
    } else {
        r.vs = r.ve = 0;
    }
    return r;
}
