// Host-side table construction; see spl_host.h.
#include "spl_host.h"
#include "spl_segment.h"
#include <algorithm>
#include <cstring>

namespace {

int b64val(uint8_t c) {
    if (c >= 'A' && c <= 'Z') return c - 'A';
    if (c >= 'a' && c <= 'z') return c - 'a' + 26;
    if (c >= '0' && c <= '9') return c - '0' + 52;
    if (c == '+') return 62;
    if (c == '/') return 63;
    return -1;
}

// Strict RFC 4648 decode with canonical padding (what base64::STANDARD accepts).
bool b64decode(const uint8_t* s, size_t n, std::string& out) {
    out.clear();
    if (n % 4 != 0) return false;
    for (size_t i = 0; i < n; i += 4) {
        int v[4];
        int pad = 0;
        for (int k = 0; k < 4; ++k) {
            uint8_t c = s[i + k];
            if (c == '=') {
                if (i + 4 != n || k < 2) return false;
                v[k] = 0; ++pad;
            } else {
                if (pad) return false;
                v[k] = b64val(c);
                if (v[k] < 0) return false;
            }
        }
        uint32_t w = (v[0] << 18) | (v[1] << 12) | (v[2] << 6) | v[3];
        out.push_back((char)(w >> 16));
        if (pad < 2) out.push_back((char)(w >> 8));
        if (pad < 1) out.push_back((char)w);
        if (pad == 2 && (v[1] & 15)) return false;
        if (pad == 1 && (v[2] & 3)) return false;
    }
    return true;
}

bool is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

uint32_t log2_for(size_t n_entries) {          // table size = 2^k >= 2*n (load factor <= 0.5), k >= 4
    uint32_t k = 4;
    while (((size_t)1 << k) < 2 * n_entries) ++k;
    return k;
}

// byte_level.rs:46-74: byte -> code point
void byte_to_cp_table(uint32_t* cp_of_byte) {
    bool direct[256] = {false};
    for (int b = 33; b <= 126; ++b) direct[b] = true;
    for (int b = 161; b <= 172; ++b) direct[b] = true;
    for (int b = 174; b <= 255; ++b) direct[b] = true;
    uint32_t next = 256;
    for (int b = 0; b < 256; ++b) cp_of_byte[b] = direct[b] ? (uint32_t)b : next++;
}

// decode one UTF-8 char; returns length or 0 if invalid
int utf8_next(const uint8_t* p, size_t n, uint32_t& cp) {
    if (n == 0) return 0;
    uint8_t b0 = p[0];
    if (b0 < 0x80) { cp = b0; return 1; }
    if (b0 < 0xC2 || b0 > 0xF4) return 0;
    int need = b0 < 0xE0 ? 2 : (b0 < 0xF0 ? 3 : 4);
    if ((int)n < need) return 0;
    for (int k = 1; k < need; ++k) if ((p[k] & 0xC0) != 0x80) return 0;
    if (need == 2) cp = ((b0 & 0x1F) << 6) | (p[1] & 0x3F);
    else if (need == 3) {
        cp = ((b0 & 0x0F) << 12) | ((p[1] & 0x3F) << 6) | (p[2] & 0x3F);
        if (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF)) return 0;
    } else {
        cp = ((b0 & 0x07) << 18) | ((p[1] & 0x3F) << 12) | ((p[2] & 0x3F) << 6) | (p[3] & 0x3F);
        if (cp < 0x10000 || cp > 0x10FFFF) return 0;
    }
    return need;
}

uint64_t load_le(const uint8_t* p, uint32_t n) {     // up to 8 bytes, zero padded
    uint64_t v = 0;
    for (uint32_t i = 0; i < n && i < 8; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

}  // namespace

bool spl_parse_tiktoken(const uint8_t* data, size_t len,
                        std::vector<std::pair<std::string, uint32_t>>& out, std::string& err) {
    size_t pos = 0;
    std::string tok;
    while (pos < len) {
        size_t eol = pos;
        while (eol < len && data[eol] != '\n') ++eol;
        size_t n = eol - pos;
        if (n > 0) {
            const uint8_t* line = data + pos;
            size_t sp = n;
            while (sp > 0 && line[sp - 1] != ' ') --sp;
            if (sp == 0) { err = "Invalid line format: Missing space separator"; return false; }
            size_t b64n = sp - 1;
            if (!b64decode(line, b64n, tok)) { err = "Invalid base64 encoding"; return false; }
            size_t a = sp, b = n;
            while (a < b && is_space(line[a])) ++a;
            while (b > a && is_space(line[b - 1])) --b;
            if (a < b && line[a] == '+') ++a;
            if (a == b) { err = "Invalid line format: Invalid rank"; return false; }
            uint64_t r = 0;
            for (size_t i = a; i < b; ++i) {
                if (line[i] < '0' || line[i] > '9') { err = "Invalid line format: Invalid rank"; return false; }
                r = r * 10 + (line[i] - '0');
                if (r > 0xFFFFFFFFull) { err = "Invalid line format: Invalid rank"; return false; }
            }
            out.emplace_back(tok, (uint32_t)r);
        }
        pos = eol + 1;
    }
    return true;
}

uint64_t spl_host_hashL(const uint8_t* p, uint32_t len) {
    uint64_t sum = 0;
    for (uint32_t i = 0; i * 8 < len; ++i) sum += spl_hashL_word(load_le(p + i * 8, len - i * 8), i);
    return spl_hashL_final(sum, len);
}

uint32_t spl_host_lookup_pair(const SplHostTables& t, uint32_t l, uint32_t r) {
    const uint32_t tag = spl_pair_tag(l, r), hi = spl_pair_hi(l);
    uint32_t mask = (1u << t.pair_log2) - 1;
    uint32_t b = spl_pair_hash(l, r, t.pair_log2);
    for (;;) {
        const uint32_t* e = &t.pair[(size_t)b * SPL_PAIR_WORDS];
        for (int k = 0; k < SPL_PAIR_WAYS; ++k)
            if (e[k] == tag && (e[SPL_PAIR_WAYS + k] & ~SPL_SYM_MASK) == hi) return e[SPL_PAIR_WAYS + k] & SPL_SYM_MASK;
        if (e[SPL_PAIR_WORDS - 1] == SPL_PAIR_EMPTY) return SPL_RANK_NONE;
        b = (b + 1) & mask;
    }
}

// The merge loop of bpe.rs:83-194 over the bytes p[0, n) WITHOUT the whole-piece probe (the id list of a segment).
void spl_host_merge_loop(const SplHostTables& t, const uint8_t* p, uint32_t n, std::vector<uint32_t>& out) {
    std::vector<uint32_t> sym(n), rnk(n, SPL_RANK_NONE);
    std::vector<uint8_t> live(n, 1);
    for (uint32_t i = 0; i < n; ++i) sym[i] = t.byte_sym[p[i]];
    for (uint32_t i = 0; i + 1 < n; ++i) rnk[i] = spl_host_lookup_pair(t, sym[i], sym[i + 1]);
    for (;;) {
        uint32_t best = SPL_RANK_NONE, bi = 0;
        for (uint32_t i = 0; i < n; ++i) if (rnk[i] < best) { best = rnk[i]; bi = i; }
        if (best == SPL_RANK_NONE) break;
        uint32_t j = bi + 1; while (!live[j]) ++j;
        sym[bi] = best; live[j] = 0; rnk[j] = SPL_RANK_NONE;
        uint32_t k = j + 1; while (k < n && !live[k]) ++k;
        rnk[bi] = (k < n) ? spl_host_lookup_pair(t, best, sym[k]) : SPL_RANK_NONE;
        if (bi > 0) {
            uint32_t h = bi - 1; while (!live[h]) --h;
            rnk[h] = spl_host_lookup_pair(t, sym[h], best);
        }
    }
    for (uint32_t i = 0; i < n; ++i) if (live[i] && sym[i] < SPL_UNK_BASE) out.push_back(sym[i]);
}

uint32_t spl_host_lookup_piece(const SplHostTables& t, const uint8_t* p, uint32_t len) {
    if (len == 0) return SPL_RANK_NONE;
    if (len <= 8) {
        uint64_t k0 = load_le(p, len);
        uint32_t mask = (1u << t.t8_log2) - 1, b = spl_hash8((uint32_t)k0, (uint32_t)(k0 >> 32), len, t.t8_log2);
        for (;;) {
            const SplKey8* e = &t.t8[(size_t)b * SPL_T8_WAYS];
            for (int k = 0; k < SPL_T8_WAYS; ++k)
                if (e[k].len == len && e[k].k0 == k0) return e[k].id;
            if (e[SPL_T8_WAYS - 1].len == 0) return SPL_RANK_NONE;
            b = (b + 1) & mask;
        }
    }
    if (len <= 16) {
        uint64_t k0 = load_le(p, 8), k1 = load_le(p + 8, len - 8);
        uint32_t mask = (1u << t.t16_log2) - 1, h = spl_hash16(k0, k1, len, t.t16_log2);
        for (;;) {
            const SplKey16& e = t.t16[h];
            if (e.len == 0) return SPL_RANK_NONE;
            if (e.k0 == k0 && e.k1 == k1 && e.len == len) return e.id;
            h = (h + 1) & mask;
        }
    }
    if (len > t.max_key_len) return SPL_RANK_NONE;
    uint64_t hv = spl_host_hashL(p, len);
    uint32_t mask = (1u << t.tl_log2) - 1, h = (uint32_t)(hv >> (64 - t.tl_log2));
    for (;;) {
        const SplKeyL& e = t.tl[h];
        if (e.len == 0) return SPL_RANK_NONE;
        if (e.hash == hv && e.len == len && memcmp(&t.tok_bytes[t.tok_off[e.id]], p, len) == 0) return e.id;
        h = (h + 1) & mask;
    }
}

bool spl_build_tables(SplHostTables& t, const uint8_t* vocab, size_t vocab_len, int pattern, uint32_t flags,
                      const char* const* special_strs, const uint32_t* special_ids, size_t n_special) {
    t.pattern = pattern;
    t.flags = flags;
    if (pattern != SPL_PAT_CL100K && pattern != SPL_PAT_O200K && pattern != SPL_PAT_MISTRAL_V3 &&
        pattern != SPL_PAT_SENTENCEPIECE) {
        t.error = "unsupported split pattern id";
        return false;
    }
    std::vector<std::pair<std::string, uint32_t>> entries;
    if (!spl_parse_tiktoken(vocab, vocab_len, entries, t.error)) return false;

    // ---- encoder map over RAW bytes -------------------------------------------------
    t.encoder.clear();
    t.encoder.reserve(entries.size() * 2);
    if (flags & SPL_FLAG_BYTE_LEVEL) {
        uint32_t cp_of_byte[256];
        byte_to_cp_table(cp_of_byte);
        std::unordered_map<uint32_t, uint8_t> byte_of_cp;
        for (int b = 0; b < 256; ++b) byte_of_cp[cp_of_byte[b]] = (uint8_t)b;
        // later duplicates overwrite (vocab.rs:85) -- apply on the byte-level keys first
        std::unordered_map<std::string, uint32_t> bl;
        bl.reserve(entries.size() * 2);
        for (auto& e : entries) bl[e.first] = e.second;
        uint32_t char_rank[256];
        bool have[256] = {false};
        struct Tr { std::string raw; uint32_t rank; uint32_t nchars; };
        std::vector<Tr> trs;
        trs.reserve(bl.size());
        for (auto& kv : bl) {
            const uint8_t* p = (const uint8_t*)kv.first.data();
            size_t n = kv.first.size(), i = 0;
            std::string raw;
            bool ok = true;
            while (i < n) {
                uint32_t cp;
                int l = utf8_next(p + i, n - i, cp);
                if (l == 0) { t.error = "byte-level vocabulary key is not valid UTF-8"; return false; }
                auto it = byte_of_cp.find(cp);
                if (it == byte_of_cp.end()) { ok = false; break; }   // outside the alphabet: unreachable key
                raw.push_back((char)it->second);
                i += l;
            }
            if (!ok || raw.empty()) continue;
            if (raw.size() == 1) { have[(uint8_t)raw[0]] = true; char_rank[(uint8_t)raw[0]] = kv.second; }
            trs.push_back({raw, kv.second, (uint32_t)raw.size()});
        }
        for (int b = 0; b < 256; ++b)
            if (!have[b]) { t.error = "byte-level vocabulary lacks one of the 256 base characters"; return false; }
        std::unordered_map<uint32_t, int> seen_rank;
        for (auto& tr : trs) {
            if (++seen_rank[tr.rank] > 1) { t.error = "byte-level vocabulary has duplicate ranks"; return false; }
            if (tr.nchars >= 2) {
                // every 2-byte base char inside a longer key must merge before the key can
                // (see DESIGN.md "byte-level folding"); otherwise raw-byte BPE != byte-level BPE
                for (unsigned char c : tr.raw) {
                    if (cp_of_byte[c] >= 0x80 && char_rank[c] >= tr.rank) {
                        t.error = "byte-level vocabulary: a multi-character key outranks one of its base characters";
                        return false;
                    }
                }
            }
            t.encoder[tr.raw] = tr.rank;
        }
    } else if (flags & SPL_FLAG_SENTENCEPIECE) {
        for (auto& e : entries) t.encoder.emplace(e.first, e.second);       // the FIRST occurrence encodes (vocab.rs:139)
    } else {
        for (auto& e : entries) t.encoder[e.first] = e.second;
    }

    // ---- decode table (build_decoder vocab.rs:146-148 + decode_bytes tokenizer.rs:877-897) ----------------
    {
        std::unordered_map<std::string, uint32_t> final_rank;              // later duplicates overwrite (vocab.rs:85)
        final_rank.reserve(entries.size() * 2);
        for (auto& e : entries) final_rank[e.first] = e.second;
        uint32_t cp_of_byte[256];
        byte_to_cp_table(cp_of_byte);
        std::unordered_map<uint32_t, uint8_t> byte_of_cp;
        for (int b = 0; b < 256; ++b) byte_of_cp[cp_of_byte[b]] = (uint8_t)b;
        uint32_t max_dec = 0;
        bool any = false;
        for (auto& e : entries) { max_dec = std::max(max_dec, e.second); any = true; }
        for (size_t i = 0; i < n_special; ++i) { max_dec = std::max(max_dec, special_ids[i]); any = true; }
        if (any && max_dec > 0x3FFFFFFu) { t.error = "token ids above 2^26 are not supported"; return false; }
        std::vector<std::string> dec(any ? (size_t)max_dec + 1 : 0);
        std::vector<uint8_t> has(dec.size(), 0);
        const bool spm = (flags & SPL_FLAG_SENTENCEPIECE) != 0;
        for (auto& e : entries) {
            // SentencePiece vocabularies decode EVERY id (vocab.rs:135); otherwise the decoder is the inverse of the
            // encoder map, which has lost the overwritten duplicates
            if (!spm && final_rank[e.first] != e.second) continue;
            std::string raw = e.first;
            if (flags & SPL_FLAG_BYTE_LEVEL) {
                // byte_level_decode_bytes, falling back to the key itself (tokenizer.rs:883-887)
                const uint8_t* p = (const uint8_t*)e.first.data();
                size_t n = e.first.size(), i = 0;
                std::string tr;
                bool ok = true;
                while (i < n) {
                    uint32_t cp;
                    int l = utf8_next(p + i, n - i, cp);
                    if (l == 0) { ok = false; break; }
                    auto it = byte_of_cp.find(cp);
                    if (it == byte_of_cp.end()) { ok = false; break; }
                    tr.push_back((char)it->second);
                    i += l;
                }
                if (ok) raw = tr;
            }
            dec[e.second] = raw; has[e.second] = 1;
        }
        for (size_t i = 0; i < n_special; ++i)       // from_bytes_sentencepiece inserts the specials INTO the decoder (tokenizer.rs:597-600)
            if (spm || !has[special_ids[i]]) { dec[special_ids[i]] = special_strs[i]; has[special_ids[i]] = 1; }
        t.dec_off.assign(dec.size() + 1, 0);
        t.dec_bytes.clear();
        for (size_t i = 0; i < dec.size(); ++i) {
            t.dec_off[i] = (uint32_t)t.dec_bytes.size();
            t.dec_bytes.insert(t.dec_bytes.end(), dec[i].begin(), dec[i].end());
        }
        t.dec_off[dec.size()] = (uint32_t)t.dec_bytes.size();
        t.dec_bytes.resize(t.dec_bytes.size() + 16, 0);
    }

    // ---- id space ---------------------------------------------------------------------
    uint32_t max_id = 0;
    t.max_key_len = 0;
    for (auto& kv : t.encoder) {
        if (kv.first.empty()) { t.error = "empty vocabulary key"; return false; }
        max_id = std::max(max_id, kv.second);
        t.max_key_len = std::max<uint32_t>(t.max_key_len, (uint32_t)kv.first.size());
    }
    if (max_id >= SPL_UNK_BASE) { t.error = "token ids above 2^21-257 are not supported"; return false; }
    t.n_ids = t.encoder.empty() ? 0 : max_id + 1;

    for (int b = 0; b < 256; ++b) {
        auto it = t.encoder.find(std::string(1, (char)b));
        t.byte_sym[b] = it != t.encoder.end() ? it->second : SPL_UNK_BASE + b;
    }

    // ---- token byte pool (by id; verification of long keys) -----------------------------
    {
        std::vector<const std::string*> by_id(t.n_ids, nullptr);
        for (auto& kv : t.encoder) by_id[kv.second] = &kv.first;    // ids are unique per key after dedup
        t.tok_off.assign(t.n_ids + 1, 0);
        t.tok_bytes.clear();
        for (uint32_t i = 0; i < t.n_ids; ++i) {
            t.tok_off[i] = (uint32_t)t.tok_bytes.size();
            if (by_id[i]) t.tok_bytes.insert(t.tok_bytes.end(), by_id[i]->begin(), by_id[i]->end());
        }
        t.tok_off[t.n_ids] = (uint32_t)t.tok_bytes.size();
        t.tok_bytes.resize(t.tok_bytes.size() + 16, 0);
    }

    // ---- whole-piece tables -----------------------------------------------------------------
    size_t n8 = 0, n16 = 0, nl = 0;
    for (auto& kv : t.encoder) {
        size_t n = kv.first.size();
        if (n <= 8) ++n8; else if (n <= 16) ++n16; else ++nl;
    }
    t.t8_log2 = log2_for(n8) - 1;                       // buckets of two: as many slots as before
    t.t8.assign(((size_t)1 << t.t8_log2) * SPL_T8_WAYS, SplKey8{0, 0, 0});
    t.t16_log2 = log2_for(n16); t.t16.assign((size_t)1 << t.t16_log2, SplKey16{0, 0, 0, 0, 0});
    t.tl_log2 = log2_for(nl);   t.tl.assign((size_t)1 << t.tl_log2, SplKeyL{0, 0, 0});
    // in id order: the low-rank (frequent) tokens take the home buckets
    std::vector<const std::pair<const std::string, uint32_t>*> by_rank;
    by_rank.reserve(t.encoder.size());
    for (auto& kv : t.encoder) by_rank.push_back(&kv);
    std::sort(by_rank.begin(), by_rank.end(), [](auto* a, auto* b) { return a->second != b->second ? a->second < b->second : a->first < b->first; });
    t.t8_displaced = 0;
    for (auto* kvp : by_rank) {
        auto& kv = *kvp;
        const uint8_t* p = (const uint8_t*)kv.first.data();
        uint32_t n = (uint32_t)kv.first.size();
        // two keys can map to one id only if ids collide in the file; tok_off then holds one of them
        if (n <= 8) {
            uint64_t k0 = load_le(p, n);
            uint32_t mask = (1u << t.t8_log2) - 1, b = spl_hash8((uint32_t)k0, (uint32_t)(k0 >> 32), n, t.t8_log2);
            bool home = true;
            for (;; b = (b + 1) & mask, home = false) {
                SplKey8* e = &t.t8[(size_t)b * SPL_T8_WAYS];
                int k = 0;
                while (k < SPL_T8_WAYS && e[k].len) ++k;
                if (k < SPL_T8_WAYS) { e[k] = SplKey8{k0, kv.second, n}; break; }
            }
            if (!home) ++t.t8_displaced;
        } else if (n <= 16) {
            uint64_t k0 = load_le(p, 8), k1 = load_le(p + 8, n - 8);
            uint32_t mask = (1u << t.t16_log2) - 1, h = spl_hash16(k0, k1, n, t.t16_log2);
            while (t.t16[h].len) h = (h + 1) & mask;
            t.t16[h] = SplKey16{k0, k1, kv.second, n, 0};
        } else {
            uint64_t hv = spl_host_hashL(p, n);
            uint32_t mask = (1u << t.tl_log2) - 1, h = (uint32_t)(hv >> (64 - t.tl_log2));
            while (t.tl[h].len) h = (h + 1) & mask;
            // the verifier compares against tok_bytes[tok_off[id]]: make sure that is THIS key
            if (t.tok_off[kv.second + 1] - t.tok_off[kv.second] != n ||
                memcmp(&t.tok_bytes[t.tok_off[kv.second]], p, n) != 0) {
                t.error = "two vocabulary keys share one id";
                return false;
            }
            t.tl[h] = SplKeyL{hv, kv.second, n};
        }
    }

    // ---- pair table: every split of every key into two symbols ------------------------------
    {
        struct Ent { uint32_t l, r, m; };
        std::vector<Ent> ents;
        ents.reserve(t.encoder.size() * 3);
        std::string a, b;
        for (auto& kv : t.encoder) {
            const std::string& k = kv.first;
            size_t n = k.size();
            for (size_t s = 1; s < n; ++s) {
                uint32_t ls, rs;
                if (s == 1) ls = t.byte_sym[(uint8_t)k[0]];
                else {
                    a.assign(k, 0, s);
                    auto it = t.encoder.find(a);
                    if (it == t.encoder.end()) continue;
                    ls = it->second;
                }
                if (n - s == 1) rs = t.byte_sym[(uint8_t)k[n - 1]];
                else {
                    b.assign(k, s, n - s);
                    auto it = t.encoder.find(b);
                    if (it == t.encoder.end()) continue;
                    rs = it->second;
                }
                ents.push_back(Ent{ls, rs, kv.second});
            }
        }
        t.n_pairs = ents.size();
        t.pair_log2 = log2_for(ents.size()) - 1;          // buckets of four at load <= 0.25: a probe rarely leaves its home bucket
        t.pair.assign(((size_t)1 << t.pair_log2) * SPL_PAIR_WORDS, SPL_PAIR_EMPTY);
        uint32_t mask = (1u << t.pair_log2) - 1;
        // low merged rank first: the pairs the merge loop asks for most sit in their home bucket
        std::sort(ents.begin(), ents.end(), [](const Ent& x, const Ent& y) {
            return x.m != y.m ? x.m < y.m : (x.l != y.l ? x.l < y.l : x.r < y.r);
        });
        t.pair_displaced = 0;
        for (const Ent& e : ents) {
            uint32_t bk = spl_pair_hash(e.l, e.r, t.pair_log2);
            bool home = true;
            for (;; bk = (bk + 1) & mask, home = false) {
                uint32_t* s = &t.pair[(size_t)bk * SPL_PAIR_WORDS];
                int k = 0;
                while (k < SPL_PAIR_WAYS && s[SPL_PAIR_WAYS + k] != SPL_PAIR_EMPTY) ++k;
                if (k < SPL_PAIR_WAYS) { s[k] = spl_pair_tag(e.l, e.r); s[SPL_PAIR_WAYS + k] = e.m | spl_pair_hi(e.l); break; }
            }
            if (!home) ++t.pair_displaced;
        }
        // dense copy of the byte x byte corner: the first rank of every part of every piece is one of these
        t.bpair.assign(65536, SPL_RANK_NONE);
        for (uint32_t b0 = 0; b0 < 256; ++b0)
            for (uint32_t b1 = 0; b1 < 256; ++b1)
                t.bpair[(b0 << 8) | b1] = spl_host_lookup_pair(t, t.byte_sym[b0], t.byte_sym[b1]);
    }

    // ---- independent segments (spl_segment.h): which character boundaries can no key cross --------------------
    {
        t.seg_irr.assign(2048, 0u);
        std::vector<std::pair<uint32_t, uint32_t>> regular;        // (A, B) packed
        for (auto& kv : t.encoder) {
            const uint8_t* k = (const uint8_t*)kv.first.data();
            const uint32_t n = (uint32_t)kv.first.size();
            auto word_at = [&](uint32_t i) { uint32_t w = 0; for (uint32_t q = 0; q < 4 && i + q < n; ++q) w |= (uint32_t)k[i + q] << (8 * q); return w; };
            for (uint32_t q = 0; q + 1 < n; ++q) {
                if ((k[q + 1] & 0xC0u) == 0x80u) continue;         // never the first byte of a character of the text
                uint32_t pa = 0, pb = 0;
                int32_t p = (int32_t)q;
                while (p >= 0 && (k[p] & 0xC0u) == 0x80u) --p;
                const uint32_t la = p >= 0 ? spl_u8_char(word_at((uint32_t)p), q + 1 - (uint32_t)p, pa) : 0u;
                const bool left_ok = la != 0 && (uint32_t)p + la == q + 1;
                const uint32_t lb = spl_u8_char(word_at(q + 1), n - (q + 1), pb);
                if (left_ok && lb) {
                    if (la > 1 || lb > 1) regular.emplace_back(pa, pb);
                } else {
                    const uint32_t ci = ((uint32_t)k[q] << 8) | k[q + 1];
                    t.seg_irr[ci >> 5] |= 1u << (ci & 31);
                }
            }
        }
        std::sort(regular.begin(), regular.end());
        regular.erase(std::unique(regular.begin(), regular.end()), regular.end());
        t.seg_pairs = regular.size();
        uint32_t lg = 16;
        while (lg < 26 && ((size_t)1 << lg) < regular.size() * 64) ++lg;       // fill <= 1/64: that many safe boundaries are missed
        t.seg_h2_log2 = lg;
        t.seg_h2.assign((size_t)1 << (lg - 5), 0u);
        for (auto& ab : regular) {
            const uint32_t hb = spl_seg_hash(ab.first, ab.second, lg);
            t.seg_h2[hb >> 5] |= 1u << (hb & 31);
        }
        // what the merge loop makes of a single 2- or 3-byte character: one id in place, two or three ids in char_ids
        // (SPL_CHAR_* in spl_segment.h); characters with a byte the vocabulary does not know stay with the merge loop
        t.char_tok.assign(65536, SPL_RANK_NONE);
        t.char_ids.clear();
        std::vector<uint32_t> ids;
        for (uint32_t cp = 0x80; cp < 0x10000; ++cp) {
            uint8_t u[3];
            uint32_t L;
            if (cp < 0x800) { u[0] = (uint8_t)(0xC0 | (cp >> 6)); u[1] = (uint8_t)(0x80 | (cp & 0x3F)); L = 2; }
            else { u[0] = (uint8_t)(0xE0 | (cp >> 12)); u[1] = (uint8_t)(0x80 | ((cp >> 6) & 0x3F)); u[2] = (uint8_t)(0x80 | (cp & 0x3F)); L = 3; }
            bool known = true;
            for (uint32_t q = 0; q < L; ++q) known &= t.byte_sym[u[q]] < SPL_UNK_BASE;
            if (!known) continue;                 // bpe.rs:187-191 drops unknown bytes one by one: leave that to the loop
            ids.clear();
            spl_host_merge_loop(t, u, L, ids);
            if (ids.size() == 1) t.char_tok[cp] = ids[0];
            else if (ids.size() == L || ids.size() == 2) {
                t.char_tok[cp] = ((uint32_t)(ids.size() - 1) << SPL_CHAR_COUNT_SHIFT) | (uint32_t)t.char_ids.size();
                t.char_ids.insert(t.char_ids.end(), ids.begin(), ids.end());
            }
        }
        t.char_ids.resize(t.char_ids.size() + 4, 0);
    }

    // ---- special tokens -----------------------------------------------------------------------
    t.sp_bytes.clear(); t.sp_off.clear(); t.sp_id.clear();
    memset(t.sp_first, 0, sizeof(t.sp_first));
    std::vector<std::string> sp;
    for (size_t i = 0; i < n_special; ++i) {
        std::string s(special_strs[i]);
        if (s.empty()) { t.error = "empty special token string"; return false; }
        sp.push_back(s);
        t.sp_off.push_back((uint32_t)t.sp_bytes.size());
        t.sp_bytes.insert(t.sp_bytes.end(), s.begin(), s.end());
        t.sp_id.push_back(special_ids[i]);
        uint8_t f = (uint8_t)s[0];
        t.sp_first[f >> 5] |= 1u << (f & 31);
    }
    t.sp_off.push_back((uint32_t)t.sp_bytes.size());
    t.sp_bytes.resize(t.sp_bytes.size() + 16, 0);
    t.specials_unambiguous = true;
    for (size_t i = 0; i < sp.size() && t.specials_unambiguous; ++i) {
        for (size_t j = 0; j < sp.size(); ++j) {
            const std::string &a = sp[i], &b = sp[j];
            if (i != j && a.find(b) != std::string::npos) { t.specials_unambiguous = false; break; }
            // a proper suffix of a that is a proper prefix of b lets two matches overlap
            size_t m = std::min(a.size(), b.size());
            for (size_t l = 1; l < m; ++l)
                if (a.compare(a.size() - l, l, b, 0, l) == 0) { t.specials_unambiguous = false; break; }
            if (!t.specials_unambiguous) break;
        }
    }
    return true;
}
