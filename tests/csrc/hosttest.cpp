// Host-side test harness (tests only): runs the product's __host__ __device__ pre-tokenizer
// rules on the CPU so they can be fuzzed against the oracle's regex engine without a GPU.
#include <vector>
#include <cstring>
#include "../../splintr_b200/csrc/spl_pretok.h"
#include "../../splintr_b200/csrc/unicode_tables.inc"

struct HostText {
    const uint8_t* p;
    uint8_t byte(uint32_t i) const { return p[i]; }
    uint64_t word8(uint32_t i) const { uint64_t v; memcpy(&v, p + i, 8); return v; }      // spl_ingest.h: i + 16 <= end
};

extern "C" {

// sequential scan of one segment [0,n): starts[i]=1 at every piece start
int ht_scan_seq(int pattern, const uint8_t* text, uint32_t n, uint8_t* starts) {
    HostText t{text};
    SplScanner<HostText> sc(t, spl_ucd_stage1, spl_ucd_stage2, pattern);
    memset(starts, 0, n + 1);
    uint32_t p = 0;
    while (p < n) {
        starts[p] = 1;
        uint32_t e = sc.next_end(p, n);
        if (e <= p) return -1;
        p = e;
    }
    return 0;
}

// emulation of the device work split: one "thread" per `chunk` bytes; a thread starts at
// the first sync point inside its chunk and scans until the first piece start that is
// >= chunk end AND a sync point.  `hard` marks segment starts (hard[n] must be 1).
int ht_scan_chunked(int pattern, const uint8_t* text, uint32_t n, uint32_t chunk,
                    const uint8_t* hard, uint8_t* starts, uint32_t* n_sync) {
    HostText t{text};
    SplScanner<HostText> sc(t, spl_ucd_stage1, spl_ucd_stage2, pattern);
    memset(starts, 0, n + 1);
    uint32_t ns = 0;
    for (uint32_t c0 = 0; c0 < n; c0 += chunk) {
        uint32_t c1 = c0 + chunk < n ? c0 + chunk : n;
        // segment containing c0
        uint32_t S = c0; while (!hard[S]) --S;
        uint32_t E = c0 + 1; while (!hard[E]) ++E;
        uint32_t a = c0;
        bool found = false;
        for (; a < c1; ++a) {
            if (hard[a]) { S = a; E = a + 1; while (!hard[E]) ++E; found = true; break; }
            if ((text[a] & 0xC0) == 0x80) continue;
            if (sc.is_sync(a, S, E)) { found = true; break; }
        }
        if (!found) continue;
        ++ns;
        uint32_t p = a;
        for (;;) {
            if (starts[p]) return -2;          // two threads own the same piece
            starts[p] = 1;
            uint32_t e = sc.next_end(p, E);
            if (e <= p || e > E) return -1;
            p = e;
            if (p >= n) break;
            if (p == E) { S = E; E = S + 1; while (!hard[E]) ++E; }
            if (p >= c1 && sc.is_sync(p, S, E)) break;
        }
    }
    if (n_sync) *n_sync = ns;
    return 0;
}

}

// ---------------------------------------------------------------------------------------
// Host emulation of the device piece encoder: same tables, same probes, same merge rule
// (pair table keyed by symbol ids, leftmost minimum, in-place part bitmap).
#include "../../splintr_b200/csrc/spl_host.h"

#include "../../splintr_b200/csrc/spl_segment.h"

struct HostPieceReader {
    const uint8_t* p; uint32_t n;
    uint32_t load4(uint32_t i) const { uint32_t v = 0; for (uint32_t q = 0; q < 4 && i + q < n; ++q) v |= (uint32_t)p[i + q] << (8 * q); return v; }
};

static uint64_t ht_seg_stats[4];     // segments, single-character segments, segments beyond SPL_SEG_MAX, safe boundaries

// The device path for one piece: whole-piece probe (k_probe); on a miss the piece falls apart at its safe boundaries
// (spl_segment.h, k_probe's refining pass) -- a segment that is a single 2- or 3-byte character goes through char_tok,
// any other segment through the merge loop (k_bpe up to SPL_SEG_MAX bytes, k_bpe_long beyond).
// tile_limit: boundaries at or beyond this byte of the piece are not looked for (the device only refines inside the
// tile that owns the piece).
static uint32_t ht_tile_limit = 0xFFFFFFFFu;
static void ht_bpe_piece(const SplHostTables& T, const uint8_t* p, uint32_t n, std::vector<uint32_t>& out, bool segments = true) {
    uint32_t whole = spl_host_lookup_piece(T, p, n);
    if (whole != SPL_RANK_NONE) { out.push_back(whole); return; }
    if (n == 1) { if (T.byte_sym[p[0]] < SPL_UNK_BASE) out.push_back(T.byte_sym[p[0]]); return; }
    if (!segments) { spl_host_merge_loop(T, p, n, out); return; }
    HostPieceReader rd{p, n};
    std::vector<uint32_t> cuts;
    spl_safe_boundaries(rd, n, ht_tile_limit, T.seg_irr.data(), T.seg_h2.data(), T.seg_h2_log2, [&](uint32_t pos) { cuts.push_back(pos); });
    cuts.push_back(n);
    uint32_t pos = 0;
    for (uint32_t end : cuts) {
        const uint32_t sl = end - pos;
        ++ht_seg_stats[0];
        if (end < n) ++ht_seg_stats[3];
        if (sl > SPL_SEG_MAX) ++ht_seg_stats[2];
        uint32_t packed = 0;
        const uint32_t la = spl_u8_char(rd.load4(pos), sl, packed);
        if (sl == 1) {
            const uint32_t sy = T.byte_sym[p[pos]];
            if (sy < SPL_UNK_BASE) out.push_back(sy);
        } else if (la == sl && sl <= 3 && T.char_tok[spl_u8_cp23(packed, sl)] != SPL_RANK_NONE) {
            const uint32_t v = T.char_tok[spl_u8_cp23(packed, sl)], cnt = (v >> SPL_CHAR_COUNT_SHIFT) + 1;
            if (cnt == 1) out.push_back(v & SPL_CHAR_VALUE_MASK);
            else for (uint32_t q = 0; q < cnt; ++q) out.push_back(T.char_ids[(v & SPL_CHAR_VALUE_MASK) + q]);
            ++ht_seg_stats[1];
        } else {
            spl_host_merge_loop(T, p + pos, sl, out);
        }
        pos = end;
    }
}

extern "C" {

void* ht_create(const uint8_t* vocab, size_t vocab_len, int pattern, uint32_t flags,
                const char* const* sp_strs, const uint32_t* sp_ids, size_t n_sp, char* err, size_t errcap) {
    SplHostTables* t = new SplHostTables();
    if (!spl_build_tables(*t, vocab, vocab_len, pattern, flags, sp_strs, sp_ids, n_sp)) {
        strncpy(err, t->error.c_str(), errcap - 1); err[errcap - 1] = 0;
        delete t; return nullptr;
    }
    return t;
}
void ht_destroy(void* h) { delete (SplHostTables*)h; }
void ht_stats(void* h, uint64_t* out) {
    SplHostTables* t = (SplHostTables*)h;
    out[0] = t->encoder.size(); out[1] = t->n_pairs; out[2] = t->t8_log2; out[3] = t->t16_log2;
    out[4] = t->tl_log2; out[5] = t->pair_log2; out[6] = t->max_key_len; out[7] = t->specials_unambiguous;
    out[8] = t->t8_displaced; out[9] = t->pair_displaced;
    out[10] = t->seg_pairs; out[11] = t->seg_h2_log2;
    uint64_t irr = 0; for (uint32_t w : t->seg_irr) irr += __builtin_popcount(w);
    out[12] = irr;
    uint64_t ct = 0; for (uint32_t v : t->char_tok) ct += v != SPL_RANK_NONE;
    out[13] = ct;
}
void ht_set_tile_limit(uint32_t l) { ht_tile_limit = l; }
void ht_seg_counters(uint64_t* out, int reset) { for (int i = 0; i < 4; ++i) { out[i] = ht_seg_stats[i]; if (reset) ht_seg_stats[i] = 0; } }

// one piece (no pre-tokenizer): segments != 0 -> the device path, 0 -> whole-piece probe + the plain merge loop
long ht_encode_piece(void* h, const uint8_t* p, uint32_t n, int segments, uint32_t* ids, size_t cap) {
    std::vector<uint32_t> out;
    ht_bpe_piece(*(SplHostTables*)h, p, n, out, segments != 0);
    if (out.size() > cap) return -2;
    memcpy(ids, out.data(), out.size() * 4);
    return (long)out.size();
}

// encode one segment (no special handling); returns token count or -1
long ht_encode(void* h, const uint8_t* text, uint32_t n, uint32_t* ids, size_t cap) {
    SplHostTables* T = (SplHostTables*)h;
    HostText t{text};
    SplScanner<HostText> sc(t, spl_ucd_stage1, spl_ucd_stage2, T->pattern);
    std::vector<uint32_t> out;
    uint32_t p = 0;
    while (p < n) {
        uint32_t e = sc.next_end(p, n);
        if (e <= p || e > n) return -1;
        ht_bpe_piece(*T, text + p, e - p, out);
        p = e;
    }
    if (out.size() > cap) return -2;
    memcpy(ids, out.data(), out.size() * 4);
    return (long)out.size();
}

}

// ---------------------------------------------------------------------------------------
// spl_pretok_chunk (the device kernel's per-thread routine) under a host environment with
// an artificially small tile / window, to exercise provisional segment ends.
struct HostEnv {
    const uint8_t* text; const uint8_t* hardb; const uint8_t* specb; uint8_t* out; uint32_t W; int* dup;
    uint8_t byte(uint32_t i) const { return text[i]; }
    bool hard(uint32_t i) const { return hardb[i]; }
    bool spec(uint32_t i) const { return specb && specb[i]; }
    uint32_t next_hard(uint32_t from, uint32_t lim) const { while (from < lim && !hardb[from]) ++from; return from; }
    uint32_t win_end() const { return W; }
    void mark(uint32_t p) { if (out[p]) *dup = 1; out[p] = 1; }
};

extern "C" int ht_scan_kernel_emul(int pattern, const uint8_t* text, uint32_t n, uint32_t tile, uint32_t halo,
                                   uint32_t chunk, const uint8_t* hard, const uint8_t* spec, uint8_t* starts) {
    memset(starts, 0, n + 1);
    int dup = 0;
    for (uint32_t t0 = 0; t0 < n; t0 += tile) {
        for (uint32_t c0 = t0; c0 < t0 + tile && c0 < n; c0 += chunk) {
            uint32_t c1 = c0 + chunk; if (c1 > t0 + tile) c1 = t0 + tile; if (c1 > n) c1 = n;
            HostEnv env{text, hard, spec, starts, t0 + tile + halo, &dup};
            spl_pretok_chunk(env, c0, c1, n, spl_ucd_stage1, spl_ucd_stage2, pattern, spec != nullptr);
        }
    }
    return dup ? -2 : 0;
}

// ---------------------------------------------------------------------------------------
// Bit-parallel pre-tokenizer (spl_pretok_fast.h) under a host emulation of the kernel's
// tile / window structure: `payload` words per tile, `halo` words of context on each side.
#include "../../splintr_b200/csrc/spl_pretok_fast.h"

struct HostFastText {
    const uint8_t* p; uint32_t n;
    uint8_t byte(uint32_t i) const { return i < n ? p[i] : 0; }
    uint32_t load4(uint32_t i) const {
        uint32_t v = 0;
        for (uint32_t b = 0; b < 4; ++b) if (i + b < n) v |= (uint32_t)p[i + b] << (8 * b);
        return v;
    }
};

struct HostFastMasks {
    std::vector<uint32_t> m[FM_COUNT];
    std::vector<uint32_t> hardw, specw, validw, sum;
    uint32_t get(int q, int k) const { return m[q][k]; }
    uint32_t hard(int k) const { return hardw[k]; }
    uint32_t spec(int k) const { return specw[k]; }
    uint32_t valid(int k) const { return validw[k]; }
    uint32_t summary(int k) const { return sum[k]; }
};

// The fused kernel (k_pretok_probe) also trusts the first `ht_fast_ext` words of the right halo of a tile that did
// not fall back.  With ht_fast_ext > 0 a carry that enters one of those words from outside the window also sends the
// tile to the fallback, and for the other tiles bit 2 (value 4) of starts[i] marks the piece starts those words
// report, bit 3 (value 8) the bytes they cover.
static int ht_fast_ext = 0;
extern "C" void ht_set_fast_ext(int e) { ht_fast_ext = e; }

// starts[i] = 1 at piece starts found by the fast path; tile_flag[t] = 1 when tile t asked for the fallback
extern "C" int ht_scan_fast(int pattern, const uint8_t* text, uint32_t n, uint32_t payload, uint32_t halo,
                            const uint8_t* hard, const uint8_t* spec, uint8_t* starts, uint8_t* tile_flag) {
    memset(starts, 0, n + 1);
    HostFastText t{text, n};
    const int nw = (int)(payload + 2 * halo);
    uint32_t n_tiles = (n + payload * 32 - 1) / (payload * 32);
    for (uint32_t tile = 0; tile < n_tiles; ++tile) {
        long w0 = (long)tile * payload - halo;            // global word index of window word 0
        HostFastMasks M;
        for (auto& v : M.m) v.assign(nw, 0);
        M.hardw.assign(nw, 0); M.specw.assign(nw, 0); M.validw.assign(nw, 0); M.sum.assign(nw, 0);
        for (int k = 0; k < nw; ++k) {
            long gw = w0 + k;
            if (gw < 0) continue;
            uint32_t base = (uint32_t)gw * 32u;
            if (base > n) continue;
            for (uint32_t b = 0; b < 32; ++b) {
                uint32_t i = base + b;
                if (i <= n && hard[i]) M.hardw[k] |= 1u << b;
                if (i < n && spec && spec[i]) M.specw[k] |= 1u << b;
                if (i < n) M.validw[k] |= 1u << b;
            }
            if (base < n) {
                uint32_t xw[8];
                for (int g = 0; g < 8; ++g) xw[g] = t.load4(base + 4u * g);
                SplFastWord w = spl_fast_classify(t, xw, base, n, spl_ucd_stage1, spl_ucd_stage2, pattern);
                for (int q = 0; q <= FM_BAD; ++q) M.m[q][k] = w.m[q];
            }
        }
        std::vector<SplFastLocal> loc(nw);
        for (int k = 0; k < nw; ++k) {
            M.sum[k] = spl_fast_local(M, k, nw, pattern, loc[k]);
        }
        for (int k = 0; k < nw; ++k) { M.m[FM_A2][k] = loc[k].A2; M.m[FM_A3][k] = loc[k].A3; }
        bool fb = false;
        for (int k = 0; k < nw; ++k) if (M.sum[k] & FS_BAD) fb = true;
        std::vector<uint32_t> out(payload + ht_fast_ext, 0);
        for (int k = 0; k < nw; ++k) {
            bool st = false, un = false;
            uint32_t v = spl_fast_final(M, loc[k], k, nw, pattern, spec != nullptr, st, un);
            if (st) fb = true;
            if (k >= (int)halo && k < (int)(halo + payload + ht_fast_ext)) { out[k - halo] = v; if (un) fb = true; }
        }
        tile_flag[tile] = fb ? 1 : 0;
        if (fb) {
            // what k_pretok_fb does: sequential rules over the tile's 16-byte chunks plus the lead-in worker
            int dup = 0;
            uint32_t t0 = tile * payload * 32u, t1 = t0 + payload * 32u;
            std::vector<uint8_t> tmp(n + 1, 0);
            for (uint32_t c0 = t0; c0 < t1 && c0 < n; c0 += 16) {
                uint32_t c1 = c0 + 16; if (c1 > n) c1 = n;
                HostEnv env{text, hard, spec, tmp.data(), t1 + halo * 32u, &dup};
                spl_pretok_chunk(env, c0, c1, n, spl_ucd_stage1, spl_ucd_stage2, pattern, spec != nullptr);
            }
            {
                std::vector<uint8_t> tmp2(n + 1, 0);
                HostEnv env{text, hard, spec, tmp2.data(), t1 + halo * 32u, &dup};
                spl_pretok_leadin(env, t0, n, spl_ucd_stage1, spl_ucd_stage2, pattern, spec != nullptr);
                for (uint32_t i = 0; i < n; ++i) tmp[i] |= tmp2[i];
            }
            for (uint32_t i = 0; i < n; ++i) if (tmp[i]) starts[i] |= 2;      // bit 1: produced by the fallback
            continue;
        }
        for (uint32_t pk = 0; pk < payload; ++pk)
            for (uint32_t b = 0; b < 32; ++b)
                if ((out[pk] >> b) & 1u) {
                    uint32_t i = ((uint32_t)tile * payload + pk) * 32u + b;
                    if (i < n) starts[i] |= 1;
                }
        for (uint32_t pk = payload; pk < payload + (uint32_t)ht_fast_ext; ++pk)
            for (uint32_t b = 0; b < 32; ++b) {
                uint32_t i = ((uint32_t)tile * payload + pk) * 32u + b;
                if (i >= n) continue;
                starts[i] |= 8;
                if ((out[pk] >> b) & 1u) starts[i] |= 4;
            }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// SentencePiece mode: the five kernels of spl_sentencepiece.cu replayed sequentially with the SAME
// __host__ __device__ rules (spl_sentencepiece.h), word by word and tile by tile like the device does.
#include "../../splintr_b200/csrc/spl_sentencepiece.h"

// hardb[i] = 1 at segment starts (hardb[n] must be 1), specb (optional) = bytes of special spans.
// out_text must hold 3 * n bytes, out_ps / out_spec 3 * n + 1 flags.  Returns the length of T'.
extern "C" long ht_sp_transform(const uint8_t* text, uint32_t n, const uint8_t* hardb, const uint8_t* specb,
                                uint8_t* out_text, uint8_t* out_ps, uint8_t* out_spec) {
    HostText t{text};
    std::vector<uint8_t> w0(n + 1, 0), a(n + 1, 0), rs(n + 1, 0);
    // k_sp_classify (one "thread" per 32-byte word; prev_w0 of the word's first byte is recomputed, as on the device)
    for (uint32_t base = 0; base < n; base += 32) {
        bool prev_w0 = false;
        if (base > 0 && !(specb && specb[base - 1])) prev_w0 = spl_sp_ws_byte(t, base - 1, n, spl_ucd_stage1, spl_ucd_stage2);
        for (uint32_t i = base; i < n && i < base + 32; ++i) {
            uint32_t b = text[i];
            bool sp = specb && specb[i];
            bool ws = !sp && spl_sp_ws_byte(t, i, n, spl_ucd_stage1, spl_ucd_stage2);
            if (ws) {
                w0[i] = 1; a[i] = 1;
                bool char_start = (b & 0xC0u) != 0x80u;
                if (char_start && (hardb[i] || i == 0 || !prev_w0) && !spl_sp_ascii_ws(b)) rs[i] = 1;
            }
            prev_w0 = ws;
        }
    }
    // k_sp_rawruns
    for (uint32_t st = 0; st < n; ++st)
        if (rs[st])
            for (uint32_t i = st; i < n && w0[i] && (i == st || !hardb[i]); ++i) a[i] = 0;
    // k_sp_count / k_sp_scan / k_sp_emit
    memset(out_ps, 0, 3 * (size_t)n + 1);
    memset(out_spec, 0, 3 * (size_t)n + 1);
    uint32_t o = 0;
    for (uint32_t i = 0; i < n; ++i) {
        SplSpPos p;
        p.w0 = w0[i]; p.a = a[i]; p.rs = rs[i]; p.s = hardb[i]; p.b = text[i];
        p.w0_prev = i ? w0[i - 1] : false; p.a_prev = i ? a[i - 1] : false; p.b_prev = i ? text[i - 1] : 0;
        if (spl_sp_piece_start(p)) out_ps[o] = 1;
        if (specb && specb[i]) out_spec[o] = 1;
        if (spl_sp_conv(p)) { out_text[o] = 0xE2; out_text[o + 1] = 0x96; out_text[o + 2] = 0x81; o += 3; }
        else out_text[o++] = text[i];
    }
    out_ps[o] = 1;
    return (long)o;
}

// encode one segment in SentencePiece mode: transform, then every piece of T' through the device-equivalent piece
// encoder.  ids needs room for 3 * n entries.
extern "C" long ht_encode_sp(void* h, const uint8_t* text, uint32_t n, uint32_t* ids, size_t cap) {
    SplHostTables* T = (SplHostTables*)h;
    std::vector<uint8_t> hard(n + 1, 0), t2(3 * (size_t)n + 4), ps(3 * (size_t)n + 1), sp(3 * (size_t)n + 1);
    hard[0] = 1; hard[n] = 1;
    long n2 = ht_sp_transform(text, n, hard.data(), nullptr, t2.data(), ps.data(), sp.data());
    std::vector<uint32_t> out;
    long p = 0;
    while (p < n2) {
        long e = p + 1;
        while (!ps[e]) ++e;
        ht_bpe_piece(*T, t2.data() + p, (uint32_t)(e - p), out);
        p = e;
    }
    if (out.size() > cap) return -2;
    memcpy(ids, out.data(), out.size() * 4);
    return (long)out.size();
}

// ---------------------------------------------------------------------------------------
// JSON Lines ingestion (row N4): the per-line parser of spl_ingest.h over a whole buffer, line by line, as the
// device kernels apply it.  out_text needs n bytes, out_off n + 2 entries.  counts = {docs, text bytes, missing, bad}.
#include "../../splintr_b200/csrc/spl_ingest.h"
extern "C" int ht_jsonl(const uint8_t* text, uint32_t n, const char* field, uint8_t* out_text, uint64_t* out_off, uint64_t* counts) {
    HostText t{text};
    uint32_t flen = (uint32_t)strlen(field);
    uint64_t docs = 0, bytes = 0, missing = 0, bad = 0;
    uint32_t s = 0;
    for (;;) {
        uint32_t e = s;
        while (e < n && text[e] != '\n') ++e;
        SplJlSpan sp = spl_jl_parse_line(t, s, e, (const uint8_t*)field, flen);
        if (sp.flags & SPL_JL_DOC) {
            out_off[docs++] = bytes;
            if (sp.flags & SPL_JL_FOUND) {
                uint32_t j = sp.vs, len = 0;
                while (j < sp.ve) {
                    if (j + 16u <= sp.ve && !spl_jl_has(t.word8(j), '\\')) {         // as k_jl_emit: eight plain bytes at a time
                        for (uint32_t q = 0; q < 8; ++q) out_text[bytes + len + q] = text[j + q];
                        j += 8; len += 8;
                        continue;
                    }
                    uint8_t ch[4]; uint32_t k;
                    j = spl_jl_char(t, j, sp.ve, ch, k);
                    for (uint32_t q = 0; q < k; ++q) out_text[bytes + len + q] = ch[q];
                    len += k;
                }
                if (len != sp.out_len) return -1;
                bytes += len;
            } else if (sp.flags & SPL_JL_BAD) ++bad; else ++missing;
        }
        if (e >= n) break;
        s = e + 1;
    }
    out_off[docs] = bytes;
    counts[0] = docs; counts[1] = bytes; counts[2] = missing; counts[3] = bad;
    return 0;
}

// 32 host threads in lock step: the lane group of the warp-wide snappy decoder (shuffles and ballots through a slot
// array and a barrier; a collective that not every lane reaches hangs the test, as it would hang the warp)
#include <pthread.h>
#include <thread>
struct HostWarpShared {
    pthread_barrier_t bar;
    uint32_t slot[2][32];
    HostWarpShared() { pthread_barrier_init(&bar, nullptr, 32); }
    ~HostWarpShared() { pthread_barrier_destroy(&bar); }
};
struct HostWarp {
    static constexpr uint32_t NL = 32;
    uint32_t lane;
    uint8_t* win; uint8_t* inbuf;
    HostWarpShared* sh;
    mutable uint32_t phase = 0;
    void sync() const { pthread_barrier_wait(&sh->bar); }
    uint32_t shfl(uint32_t v, uint32_t src) const {
        uint32_t* b = sh->slot[phase++ & 1u];
        __atomic_store_n(&b[lane], v, __ATOMIC_RELAXED);
        pthread_barrier_wait(&sh->bar);
        return __atomic_load_n(&b[src & 31u], __ATOMIC_RELAXED);
    }
    uint32_t ballot(bool p) const {
        uint32_t* b = sh->slot[phase++ & 1u];
        __atomic_store_n(&b[lane], p ? 1u : 0u, __ATOMIC_RELAXED);
        pthread_barrier_wait(&sh->bar);
        uint32_t r = 0;
        for (uint32_t i = 0; i < 32; ++i) r |= __atomic_load_n(&b[i], __ATOMIC_RELAXED) << i;
        return r;
    }
};
template <class Fn>
static void run_host_warp(uint8_t* win, uint8_t* inbuf, Fn fn) {
    HostWarpShared sh;
    std::vector<std::thread> th;
    for (uint32_t lane = 0; lane < 32; ++lane)
        th.emplace_back([&, lane] { HostWarp g{lane, win, inbuf, &sh}; fn(g); });
    for (auto& t : th) t.join();
}

// ---- Parquet ingestion (row N4): the host plan (spl_parquet_meta.cpp) and the page decoder of spl_parquet.h with a
// lane group of one, batch by batch, as spl_api.cu drives the device.  Returns the row count, or -1 (malformed: err),
// -2 (unsupported: err), -3 (a page raised error bits: info[0]), -4 (capacities).  info[1] = batches, info[2] = pages.
#include "../../splintr_b200/csrc/spl_parquet_meta.h"
extern "C" long ht_parquet(const uint8_t* file, size_t n, const char* column, uint64_t batch_bytes,
                           uint8_t* out_text, size_t text_cap, uint64_t* out_off, size_t off_cap,
                           char* err, size_t err_cap, uint64_t* info, int staged) {
    SplPqPlan plan;
    info[0] = info[1] = info[2] = 0;
    if (!spl_pq_plan(file, n, column, batch_bytes, plan)) {
        snprintf(err, err_cap, "%s", plan.err.c_str());
        return plan.unsupported ? -2 : -1;
    }
    info[1] = plan.batches.size(); info[2] = plan.pages.size();
    if (plan.n_rows + 1 > off_cap) return -4;
    uint64_t rows = 0, bytes = 0;
    for (const SplPqBatch& b : plan.batches) {
        std::vector<uint8_t> stage(b.stage_bytes + 16, 0), scratch(b.scratch_bytes + 16, 0);
        for (size_t k = b.range0; k < b.range1; ++k) memcpy(stage.data() + plan.ranges[k].stage_off, file + plan.ranges[k].file_off, plan.ranges[k].len);
        std::vector<uint64_t> row_off(b.n_rows + 1), dict_off(b.dict_entries + 1);
        std::vector<uint32_t> row_len(b.n_rows + 1), dict_len(b.dict_entries + 1);
        SplPqSpans R{row_off.data(), row_len.data()}, D{dict_off.data(), dict_len.data()};
        SplPqOneLane g;
        std::vector<SplU128> win(SPL_SNAPPY_WIN / 16), inbuf(SPL_SNAPPY_INBUF / 16);
        if (staged) { g.win = (uint8_t*)win.data(); g.inbuf = (uint8_t*)inbuf.data(); }   // the snappy decoder the device runs
        uint32_t e = 0;
        for (int pass = 0; pass < 2; ++pass)
            for (size_t k = b.page0; k < b.page1; ++k) {
                if (spl_pq_first_pass(plan.pages[k]) != (pass == 0)) continue;
                if (staged == 2)                                    // 32 lanes, as the device runs it
                    run_host_warp(g.win, g.inbuf, [&](const HostWarp& hw) {
                        const uint32_t pe = spl_pq_decode_page(hw, plan.pages[k], stage.data(), scratch.data(), R, D);
                        if (pe) __atomic_fetch_or(&e, pe, __ATOMIC_RELAXED);
                    });
                else e |= spl_pq_decode_page(g, plan.pages[k], stage.data(), scratch.data(), R, D);
            }
        if (e) { info[0] = e; return -3; }
        for (uint64_t r = 0; r < b.n_rows; ++r) {
            out_off[rows + r] = bytes;
            if (bytes + row_len[r] > text_cap) return -4;
            const uint8_t* src = (row_off[r] & SPL_PQ_IN_SCRATCH) ? scratch.data() + (row_off[r] & ~SPL_PQ_IN_SCRATCH) : stage.data() + row_off[r];
            memcpy(out_text + bytes, src, row_len[r]);
            bytes += row_len[r];
        }
        rows += b.n_rows;
    }
    out_off[rows] = bytes;
    return (long)rows;
}

// the two snappy decoders of spl_parquet.h on a raw stream (src and dst 16-byte aligned copies, padded); 1 = decoded
extern "C" int ht_snappy(const uint8_t* src, uint32_t n, uint8_t* out, uint32_t cap, int staged) {
    std::vector<SplU128> in((n + 31) / 16 + 1), dst((cap + 31) / 16 + 1), win(SPL_SNAPPY_WIN / 16), inbuf(SPL_SNAPPY_INBUF / 16);
    uint8_t* s = (uint8_t*)in.data() + 5;                       // an odd start, as a page body has
    memcpy(s, src, n);
    SplPqOneLane g;
    if (staged) { g.win = (uint8_t*)win.data(); g.inbuf = (uint8_t*)inbuf.data(); }
    bool ok;
    if (staged == 2) {
        uint32_t oks = 0;
        run_host_warp((uint8_t*)win.data(), (uint8_t*)inbuf.data(), [&](const HostWarp& hw) {
            if (spl_snappy_decode_warp(hw, s, n, (uint8_t*)dst.data(), cap)) __atomic_fetch_or(&oks, 1u << hw.lane, __ATOMIC_RELAXED);
        });
        if (oks != 0 && oks != 0xFFFFFFFFu) return -1;          // the lanes disagree
        ok = oks != 0;
    } else ok = staged ? spl_snappy_decode_staged(g, s, n, (uint8_t*)dst.data(), cap) : spl_snappy_decode(g, s, n, (uint8_t*)dst.data(), cap);
    if (ok) memcpy(out, dst.data(), cap);
    return ok ? 1 : 0;
}

// ---- windowed merge rounds (spl_bpe_bits.h): the m masks of one group of G lanes, B parts per lane ----------------
// K[0 .. n_pairs): rank of every pair (0x1FFFFF = none).  Builds the lanes' LT / RT masks the way bpe_group does, then
// runs the kernel's boundary-bit iteration (the shuffles become array reads) and the peak step.  m_out[g] = lane g's mask.
#include "../../splintr_b200/csrc/spl_bpe_bits.h"
extern "C" int ht_bpe_window_m(const uint32_t* K, uint32_t n_pairs, uint32_t G, uint32_t B, uint32_t* m_out) {
    const uint32_t NONE = 0x1FFFFFu, L = n_pairs + 1;
    if (B < 2 || B > 32 || (uint64_t)G * B < L || G > 32) return -1;
    uint32_t fV[32], fDL[32], fDR[32], fPK[32], m[32], cin[32] = {0}, cin2[32] = {0};
    auto rank = [&](uint32_t e) { return e + 1 < L ? K[e] : NONE; };
    for (uint32_t g = 0; g < G; ++g) {
        const uint32_t e0 = g * B, nv = e0 < L ? (B < L - e0 ? B : L - e0) : 0u;
        uint32_t LV = 0, LT = 0, RT = 0;
        uint32_t prev = (e0 && e0 < L) ? rank(e0 - 1) : NONE, cur = e0 < L ? rank(e0) : NONE;
        for (uint32_t j = 0; j < nv; ++j) {
            const uint32_t nxt = rank(e0 + j + 1);
            if (cur != NONE) {
                LV |= 1u << j;
                if (prev <= cur) LT |= 1u << j;
                if (nxt < cur) RT |= 1u << j;
            }
            prev = cur; cur = nxt;
        }
        fV[g] = LV & ~LT & ~RT; fDL[g] = LV & LT & ~RT; fDR[g] = LV & RT & ~LT; fPK[g] = LV & LT & RT;
    }
    int passes = 0;
    for (;;) {
        ++passes;
        for (uint32_t g = 0; g < G; ++g) m[g] = spl_window_slopes(fV[g], fDL[g], fDR[g], cin[g], cin2[g], B);
        bool ch = false;
        for (uint32_t g = 0; g < G; ++g) {
            const uint32_t ncin = g ? (m[g - 1] >> (B - 1u)) & 1u : 0u, ncin2 = g + 1u < G ? m[g + 1] & 1u : 0u;
            ch |= (ncin != cin[g] && (fDL[g] & 1u)) || (ncin2 != cin2[g] && ((fDR[g] >> (B - 1u)) & 1u));
            cin[g] = ncin; cin2[g] = ncin2;
        }
        if (!ch) break;
        if (passes > 64) return -2;
    }
    for (uint32_t g = 0; g < G; ++g) m_out[g] = spl_window_peaks(m[g], fPK[g], cin[g], cin2[g], B);
    return passes;
}

// ---------------------------------------------------------------------------------------
// Special-token matches for arbitrary (overlapping) sets: spl_special.h, the walker k_resolve_specials runs per document.
#include "../../splintr_b200/csrc/spl_special.h"

struct HostByteText { const uint8_t* p; uint8_t byte(uint32_t i) const { return p[i]; } };
struct HostCand {
    const std::vector<uint8_t>* c;
    uint32_t next(uint32_t from, uint32_t lim) const { while (from < lim && !(*c)[from]) ++from; return from < lim ? from : lim; }
};

// strs: n strings concatenated with offsets off[n + 1]; out: triples (start, end, k); returns the number of matches
extern "C" long ht_special_walk(const uint8_t* text, uint32_t n_text, const uint8_t* strs, const uint32_t* off, uint32_t n,
                                uint32_t* out, size_t cap) {
    SplSpecialSet S{strs, off, n};
    HostByteText t{text};
    std::vector<uint8_t> cand(n_text + 1, 0);
    for (uint32_t i = 0; i < n_text; ++i)                       // the parallel pass: where does any special string occur?
        for (uint32_t k = 0; k < n; ++k)
            if (spl_special_match(t, S, k, i, n_text)) { cand[i] = 1; break; }
    HostCand c{&cand};
    size_t m = 0;
    spl_special_walk(t, c, S, 0, n_text, [&](uint32_t s, uint32_t e, uint32_t k) {
        if (m + 3 <= cap) { out[m] = s; out[m + 1] = e; out[m + 2] = k; }
        m += 3;
    });
    return (long)(m / 3);
}
