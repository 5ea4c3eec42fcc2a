/* CPU oracle (C) for the splintr `encode_batch` path -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's algorithm, used (a) as a second checker next
 * to oracle/py_oracle.py at sizes the Python loops cannot reach and (b) as the CPU
 * baseline bench.py times beside the GPU path (`cpu_baseline`, `--impl reference`).
 * Only tests/, __graft_entry__.smoke() and bench.py may load it.  The product
 * (splintr_b200/) never does.
 *
 * Restated from (paths under /root/reference):
 *   src/core/vocab.rs:57-89        load_tiktoken_bpe            -> parse_vocab()
 *   src/core/byte_level.rs:46-74   BYTE_TO_CHAR                 -> build_byte_level()
 *   src/core/byte_level.rs:105-107 byte_level_encode            -> byte_level_expand()
 *   src/core/bpe.rs:42-54          Node                         -> struct node
 *   src/core/bpe.rs:67-197         byte_pair_encode             -> byte_pair_encode()
 *   src/core/tokenizer.rs:693-724  encode_chunk_with_position   -> encode_chunk()  (LRU omitted:
 *                                  result-transparent, tokenizer.rs:667-690)
 *   src/core/tokenizer.rs:729-808  encode                       -> encode_text() (regex branch :796-807),
 *                                                                  encode_text_sp() (SentencePiece branch :737-795)
 *   src/core/vocab.rs:101-143      load_tiktoken_bpe_with_decoder -> parse_vocab(first_wins = 1)
 *   src/core/tokenizer.rs:842-874  encode_with_special          -> encode_text_special()
 *   src/core/tokenizer.rs:932-942  encode_batch (Rayon par_iter)-> orc_encode_batch() (OpenMP,
 *                                  schedule(dynamic) over documents)
 *
 * Third-party arithmetic not under /root/reference:
 *   regexr 0.1.0-beta.5 (Cargo.toml:40) -- source absent.  The reference's own alternative
 *   backend is PCRE2 with UTF+UCP (tokenizer.rs:470-488, Cargo.toml:24 `pcre2 = "0.2"`), and
 *   its tests assert regexr == PCRE2 ids (python/tests/test_cl100k.py:436-454).  This file
 *   uses that backend: the system libpcre2-8.so.0 (10.42, Unicode 14), loaded with dlopen
 *   and hand-declared prototypes (no headers in the image), JIT-compiled, find_iter by
 *   repeated matching from the previous end (tokenizer.rs:244-257).
 *   aho-corasick 1.1 (Cargo.toml:36), MatchKind::Standard, non-overlapping find_iter:
 *   the reported match is the one that ends first (ties: the longest), scanning restarts
 *   at its end -> special_next().
 *
 * Parity pinning: checked against every golden id vector of the reference's tests and
 * against oracle/py_oracle.py on fuzz inputs in tests/test_oracle_c.py.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- PCRE2 (8-bit) hand-declared ------------------------------------------------------ */
typedef struct pcre2_code_s pcre2_code;
typedef struct pcre2_md_s pcre2_match_data;
typedef struct pcre2_mc_s pcre2_match_context;
typedef struct pcre2_js_s pcre2_jit_stack;
#define PCRE2_UTF 0x00080000u
#define PCRE2_UCP 0x00020000u
#define PCRE2_NO_UTF_CHECK 0x40000000u
#define PCRE2_NO_JIT 0x00002000u
#define PCRE2_JIT_COMPLETE 0x00000001u

static struct {
    void* so;
    pcre2_code* (*compile)(const uint8_t*, size_t, uint32_t, int*, size_t*, void*);
    int (*jit_compile)(pcre2_code*, uint32_t);
    pcre2_match_data* (*md_create)(const pcre2_code*, void*);
    int (*match)(const pcre2_code*, const uint8_t*, size_t, size_t, uint32_t, pcre2_match_data*, pcre2_match_context*);
    size_t* (*ovector)(pcre2_match_data*);
    void (*md_free)(pcre2_match_data*);
    void (*code_free)(pcre2_code*);
    pcre2_match_context* (*mc_create)(void*);
    void (*mc_free)(pcre2_match_context*);
    pcre2_jit_stack* (*js_create)(size_t, size_t, void*);
    void (*js_assign)(pcre2_match_context*, void*, void*);
    void (*js_free)(pcre2_jit_stack*);
} P;

static int load_pcre2(char* err, size_t cap) {
    if (P.so) return 0;
    const char* names[] = {"libpcre2-8.so.0", "libpcre2-8.so", NULL};
    for (int i = 0; names[i] && !P.so; ++i) P.so = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!P.so) { snprintf(err, cap, "cannot dlopen libpcre2-8.so.0"); return -1; }
#define SYM(field, name) do { *(void**)(&P.field) = dlsym(P.so, name); if (!P.field) { snprintf(err, cap, "missing %s", name); return -1; } } while (0)
    SYM(compile, "pcre2_compile_8"); SYM(jit_compile, "pcre2_jit_compile_8");
    SYM(md_create, "pcre2_match_data_create_from_pattern_8"); SYM(match, "pcre2_match_8");
    SYM(ovector, "pcre2_get_ovector_pointer_8"); SYM(md_free, "pcre2_match_data_free_8");
    SYM(code_free, "pcre2_code_free_8"); SYM(mc_create, "pcre2_match_context_create_8");
    SYM(mc_free, "pcre2_match_context_free_8"); SYM(js_create, "pcre2_jit_stack_create_8");
    SYM(js_assign, "pcre2_jit_stack_assign_8"); SYM(js_free, "pcre2_jit_stack_free_8");
#undef SYM
    return 0;
}

/* ---- byte-keyed encoder map (stands in for FxHashMap<Vec<u8>, u32>) --------------------- */
typedef struct { uint32_t off, len, rank, used; } slot;
typedef struct {
    slot* slots; uint32_t mask;
    uint8_t* pool; size_t pool_len, pool_cap;
} bmap;

static uint64_t hash_bytes(const uint8_t* p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull ^ (n * 0x9E3779B97F4A7C15ull);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, p + i, 8); h = (h ^ w) * 0x100000001b3ull; h ^= h >> 29; }
    for (; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
    h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
    return h;
}

#define RANK_NONE 0xFFFFFFFFu          /* u32::MAX of bpe.rs */

static uint32_t bmap_get(const bmap* m, const uint8_t* p, size_t n) {
    uint32_t h = (uint32_t)hash_bytes(p, n) & m->mask;
    for (;;) {
        const slot* s = &m->slots[h];
        if (!s->used) return RANK_NONE;
        if (s->len == n && memcmp(m->pool + s->off, p, n) == 0) return s->rank;
        h = (h + 1) & m->mask;
    }
}

/* insert overwrites (vocab.rs:85) unless first_wins (SentencePiece vocabularies, vocab.rs:139) */
static int bmap_put(bmap* m, const uint8_t* p, size_t n, uint32_t rank, int first_wins) {
    uint32_t h = (uint32_t)hash_bytes(p, n) & m->mask;
    for (;;) {
        slot* s = &m->slots[h];
        if (!s->used) {
            if (m->pool_len + n > m->pool_cap) {
                size_t nc = (m->pool_cap + n) * 2 + 64;
                uint8_t* np = (uint8_t*)realloc(m->pool, nc);
                if (!np) return -1;
                m->pool = np; m->pool_cap = nc;
            }
            memcpy(m->pool + m->pool_len, p, n);
            s->off = (uint32_t)m->pool_len; s->len = (uint32_t)n; s->rank = rank; s->used = 1;
            m->pool_len += n;
            return 0;
        }
        if (s->len == n && memcmp(m->pool + s->off, p, n) == 0) { if (!first_wins) s->rank = rank; return 0; }
        h = (h + 1) & m->mask;
    }
}

/* ---- tokenizer ----------------------------------------------------------------------------- */
typedef struct {
    bmap enc;
    pcre2_code* re;
    int byte_level, sentencepiece;
    uint8_t bl_utf8[256][2]; uint8_t bl_len[256];
    uint8_t** sp_str; uint32_t* sp_len; uint32_t* sp_id; uint32_t n_sp;
    uint8_t sp_last[256];
    char err[256];
} orc;

static int b64v(uint8_t c) {
    if (c >= 'A' && c <= 'Z') return c - 'A';
    if (c >= 'a' && c <= 'z') return c - 'a' + 26;
    if (c >= '0' && c <= '9') return c - '0' + 52;
    if (c == '+') return 62;
    if (c == '/') return 63;
    return -1;
}

static long b64decode(const uint8_t* s, size_t n, uint8_t* out) {
    if (n % 4) return -1;
    size_t o = 0;
    for (size_t i = 0; i < n; i += 4) {
        int v[4], pad = 0;
        for (int k = 0; k < 4; ++k) {
            if (s[i + k] == '=') { if (i + 4 != n || k < 2) return -1; v[k] = 0; ++pad; }
            else { if (pad) return -1; v[k] = b64v(s[i + k]); if (v[k] < 0) return -1; }
        }
        uint32_t w = (uint32_t)(v[0] << 18 | v[1] << 12 | v[2] << 6 | v[3]);
        out[o++] = (uint8_t)(w >> 16);
        if (pad < 2) out[o++] = (uint8_t)(w >> 8);
        if (pad < 1) out[o++] = (uint8_t)w;
    }
    return (long)o;
}

/* vocab.rs:57-89: `base64 SP rank LF`, split on the LAST space, rank trimmed */
static int parse_vocab(orc* t, const uint8_t* data, size_t len) {
    size_t lines = 0;
    for (size_t i = 0; i < len; ++i) lines += data[i] == '\n';
    uint32_t cap = 16;
    while (cap < 2 * (lines + 2)) cap <<= 1;
    t->enc.slots = (slot*)calloc(cap, sizeof(slot));
    t->enc.mask = cap - 1;
    if (!t->enc.slots) return -1;
    uint8_t* tmp = (uint8_t*)malloc(len + 4);
    size_t pos = 0;
    while (pos < len) {
        size_t eol = pos;
        while (eol < len && data[eol] != '\n') ++eol;
        size_t n = eol - pos;
        if (n) {
            const uint8_t* line = data + pos;
            size_t sp = n;
            while (sp > 0 && line[sp - 1] != ' ') --sp;
            if (sp == 0) { snprintf(t->err, sizeof t->err, "Invalid line format: Missing space separator"); free(tmp); return -1; }
            long tl = b64decode(line, sp - 1, tmp);
            if (tl < 0) { snprintf(t->err, sizeof t->err, "Invalid base64 encoding"); free(tmp); return -1; }
            size_t a = sp, b = n;
            while (a < b && (line[a] == ' ' || (line[a] >= 9 && line[a] <= 13))) ++a;
            while (b > a && (line[b - 1] == ' ' || (line[b - 1] >= 9 && line[b - 1] <= 13))) --b;
            uint64_t r = 0;
            if (a == b) { snprintf(t->err, sizeof t->err, "Invalid line format: Invalid rank"); free(tmp); return -1; }
            for (size_t i = a; i < b; ++i) {
                if (line[i] < '0' || line[i] > '9') { snprintf(t->err, sizeof t->err, "Invalid line format: Invalid rank"); free(tmp); return -1; }
                r = r * 10 + (line[i] - '0');
            }
            if (bmap_put(&t->enc, tmp, (size_t)tl, (uint32_t)r, t->sentencepiece)) { free(tmp); return -1; }
        }
        pos = eol + 1;
    }
    free(tmp);
    return 0;
}

/* byte_level.rs:46-74 */
static void build_byte_level(orc* t) {
    uint32_t next = 256;
    for (int b = 0; b < 256; ++b) {
        int direct = (b >= 33 && b <= 126) || (b >= 161 && b <= 172) || (b >= 174 && b <= 255);
        uint32_t cp = direct ? (uint32_t)b : next++;
        if (cp < 0x80) { t->bl_utf8[b][0] = (uint8_t)cp; t->bl_len[b] = 1; }
        else { t->bl_utf8[b][0] = (uint8_t)(0xC0 | (cp >> 6)); t->bl_utf8[b][1] = (uint8_t)(0x80 | (cp & 63)); t->bl_len[b] = 2; }
    }
}

typedef struct { uint32_t* v; size_t n, cap; } u32vec;
static int push(u32vec* o, uint32_t x) {
    if (o->n == o->cap) {
        size_t nc = o->cap ? o->cap * 2 : 256;
        uint32_t* nv = (uint32_t*)realloc(o->v, nc * 4);
        if (!nv) return -1;
        o->v = nv; o->cap = nc;
    }
    o->v[o->n++] = x;
    return 0;
}

/* bpe.rs:42-54 */
typedef struct { size_t prev, next; uint32_t rank; size_t start, len; } node;
#define NIL ((size_t)-1)

typedef struct {            /* per-thread scratch */
    node* nodes; size_t ncap;
    uint8_t* bl; size_t blcap;
    pcre2_match_data* md; pcre2_match_context* mc; pcre2_jit_stack* js;
} scratch;

/* bpe.rs:99-111 */
static inline uint32_t get_rank(const orc* t, const uint8_t* piece, const node* nodes, size_t li, size_t ri) {
    if (li == NIL || ri == NIL) return RANK_NONE;
    return bmap_get(&t->enc, piece + nodes[li].start, nodes[li].len + nodes[ri].len);
}

/* bpe.rs:67-197 */
static int byte_pair_encode(const orc* t, scratch* sc, const uint8_t* piece, size_t n, u32vec* out) {
    if (n == 0) return 0;                                              /* :68-70 */
    if (n == 1) {                                                      /* :73-75 */
        uint32_t r = bmap_get(&t->enc, piece, 1);
        return r == RANK_NONE ? 0 : push(out, r);
    }
    {                                                                  /* :78-80 */
        uint32_t r = bmap_get(&t->enc, piece, n);
        if (r != RANK_NONE) return push(out, r);
    }
    if (sc->ncap < n) {
        free(sc->nodes);
        sc->ncap = n * 2;
        sc->nodes = (node*)malloc(sc->ncap * sizeof(node));
        if (!sc->nodes) return -1;
    }
    node* nd = sc->nodes;
    for (size_t i = 0; i < n; ++i) {                                   /* :83-97 */
        nd[i].prev = i ? i - 1 : NIL; nd[i].next = i + 1 < n ? i + 1 : NIL;
        nd[i].rank = RANK_NONE; nd[i].start = i; nd[i].len = 1;
    }
    for (size_t i = 0; i + 1 < n; ++i) nd[i].rank = get_rank(t, piece, nd, i, nd[i].next);   /* :114-116 */
    for (;;) {                                                         /* :119-167 */
        uint32_t min_rank = RANK_NONE; size_t min_idx = NIL;
        for (size_t cur = 0; cur != NIL; cur = nd[cur].next)           /* node 0 is always the head */
            if (nd[cur].rank < min_rank) { min_rank = nd[cur].rank; min_idx = cur; }   /* strict <, :133 */
        if (min_rank == RANK_NONE) break;
        size_t nx = nd[min_idx].next;
        nd[min_idx].len += nd[nx].len;
        size_t nn = nd[nx].next;
        nd[min_idx].next = nn;
        if (nn != NIL) nd[nn].prev = min_idx;
        if (nd[min_idx].prev != NIL) { size_t p = nd[min_idx].prev; nd[p].rank = get_rank(t, piece, nd, p, min_idx); }
        nd[min_idx].rank = get_rank(t, piece, nd, min_idx, nd[min_idx].next);
    }
    for (size_t cur = 0; cur != NIL; cur = nd[cur].next) {             /* :170-194 */
        uint32_t r = bmap_get(&t->enc, piece + nd[cur].start, nd[cur].len);
        if (r != RANK_NONE) { if (push(out, r)) return -1; }
        else for (size_t j = 0; j < nd[cur].len; ++j) {
            uint32_t rb = bmap_get(&t->enc, piece + nd[cur].start + j, 1);
            if (rb != RANK_NONE && push(out, rb)) return -1;           /* unknown bytes dropped */
        }
    }
    return 0;
}

/* tokenizer.rs:693-724 (LRU skipped) */
static int encode_chunk(const orc* t, scratch* sc, const uint8_t* piece, size_t n, u32vec* out) {
    if (t->byte_level) {                                               /* :695-700 */
        if (sc->blcap < 2 * n + 2) { free(sc->bl); sc->blcap = 4 * n + 64; sc->bl = (uint8_t*)malloc(sc->blcap); if (!sc->bl) return -1; }
        size_t o = 0;
        for (size_t i = 0; i < n; ++i) { sc->bl[o++] = t->bl_utf8[piece[i]][0]; if (t->bl_len[piece[i]] == 2) sc->bl[o++] = t->bl_utf8[piece[i]][1]; }
        piece = sc->bl; n = o;
    }
    uint32_t r = n ? bmap_get(&t->enc, piece, n) : RANK_NONE;          /* :703-705 */
    if (r != RANK_NONE) return push(out, r);
    return byte_pair_encode(t, sc, piece, n, out);
}

/* tokenizer.rs:666-690 (LRU skipped) */
static int encode_bytes(const orc* t, scratch* sc, const uint8_t* b, size_t n, u32vec* out) {
    uint32_t r = n ? bmap_get(&t->enc, b, n) : RANK_NONE;
    if (r != RANK_NONE) return push(out, r);
    return byte_pair_encode(t, sc, b, n, out);
}

/* k copies of U+2581 (E2 96 81) followed by tail[0, m), through encode_bytes */
static int encode_underscores(const orc* t, scratch* sc, size_t k, const uint8_t* tail, size_t m, u32vec* out) {
    size_t need = 3 * k + m;
    if (sc->blcap < need + 2) { free(sc->bl); sc->blcap = 2 * need + 64; sc->bl = (uint8_t*)malloc(sc->blcap); if (!sc->bl) return -1; }
    for (size_t i = 0; i < k; ++i) { sc->bl[3 * i] = 0xE2; sc->bl[3 * i + 1] = 0x96; sc->bl[3 * i + 2] = 0x81; }
    if (m) memcpy(sc->bl + 3 * k, tail, m);
    return encode_bytes(t, sc, sc->bl, need, out);
}

static int encode_text_sp(const orc* t, scratch* sc, const uint8_t* text, size_t n, u32vec* out);

/* tokenizer.rs:729-808, regex find_iter :244-257 */
static int encode_text(const orc* t, scratch* sc, const uint8_t* text, size_t n, u32vec* out) {
    if (t->sentencepiece) return encode_text_sp(t, sc, text, n, out);
    size_t pos = 0;
    while (pos < n) {
        int rc = P.match(t->re, text, n, pos, PCRE2_NO_UTF_CHECK, sc->md, sc->mc);
        if (rc == -46 /* JIT stack limit */ || rc == -47 /* match limit */)
            rc = P.match(t->re, text, n, pos, PCRE2_NO_UTF_CHECK | PCRE2_NO_JIT, sc->md, sc->mc);
        if (rc == -1) break;                                           /* no further match */
        if (rc < 0) return -2;
        size_t* ov = P.ovector(sc->md);
        size_t s = ov[0], e = ov[1];
        if (e <= s) { pos = s + 1; continue; }
        if (encode_chunk(t, sc, text + s, e - s, out)) return -1;
        pos = e;
    }
    return 0;
}

/* tokenizer.rs:737-795: the SentencePiece branch of encode */
static int encode_text_sp(const orc* t, scratch* sc, const uint8_t* text, size_t n, u32vec* out) {
    size_t pos = 0, pending = 0;                                       /* U+2581 to prepend to the next word */
    while (pos < n) {
        int rc = P.match(t->re, text, n, pos, PCRE2_NO_UTF_CHECK, sc->md, sc->mc);
        if (rc == -46 || rc == -47) rc = P.match(t->re, text, n, pos, PCRE2_NO_UTF_CHECK | PCRE2_NO_JIT, sc->md, sc->mc);
        if (rc == -1) break;
        if (rc < 0) return -2;
        size_t* ov = P.ovector(sc->md);
        size_t s = ov[0], e = ov[1];
        if (e <= s) { pos = s + 1; continue; }
        uint8_t b0 = text[s];
        if (b0 == ' ' || b0 == '\t' || b0 == '\n' || b0 == 0x0C || b0 == '\r') {      /* u8::is_ascii_whitespace, :750 */
            for (size_t i = s; i < e; ++i) {                                           /* byte by byte, :752 */
                if (text[i] == ' ') { ++pending; continue; }
                if (pending) { if (encode_underscores(t, sc, pending, NULL, 0, out)) return -1; pending = 0; }
                if (encode_bytes(t, sc, text + i, 1, out)) return -1;
            }
        } else if (pending) {
            if (encode_underscores(t, sc, pending, text + s, e - s, out)) return -1;
            pending = 0;
        } else if (encode_bytes(t, sc, text + s, e - s, out)) return -1;
        pos = e;
    }
    if (pending && encode_underscores(t, sc, pending, NULL, 0, out)) return -1;        /* :787-790 */
    return 0;
}

/* aho-corasick Standard semantics: earliest end, then longest, at/after `from` */
static int special_next(const orc* t, const uint8_t* text, size_t n, size_t from, size_t* ms, size_t* me, uint32_t* id) {
    for (size_t e = from + 1; e <= n; ++e) {
        if (!t->sp_last[text[e - 1]]) continue;
        int best = -1;
        for (uint32_t k = 0; k < t->n_sp; ++k) {
            uint32_t l = t->sp_len[k];
            if (l == 0 || l > e - from) continue;
            if (memcmp(text + e - l, t->sp_str[k], l) == 0 && (best < 0 || l > t->sp_len[best])) best = (int)k;
        }
        if (best >= 0) { *ms = e - t->sp_len[best]; *me = e; *id = t->sp_id[best]; return 1; }
    }
    return 0;
}

/* tokenizer.rs:842-874 */
static int encode_text_special(const orc* t, scratch* sc, const uint8_t* text, size_t n, u32vec* out) {
    if (t->n_sp == 0) return encode_text(t, sc, text, n, out);
    size_t last = 0, ms, me; uint32_t id;
    while (special_next(t, text, n, last, &ms, &me, &id)) {
        if (ms > last) { int rc = encode_text(t, sc, text + last, ms - last, out); if (rc) return rc; }
        if (push(out, id)) return -1;
        last = me;
    }
    if (last < n) return encode_text(t, sc, text + last, n - last, out);
    return 0;
}

/* ---- exported API ------------------------------------------------------------------------------ */
void orc_destroy(void* h) {
    orc* t = (orc*)h;
    if (!t) return;
    if (t->re) P.code_free(t->re);
    free(t->enc.slots); free(t->enc.pool);
    for (uint32_t k = 0; k < t->n_sp; ++k) free(t->sp_str[k]);
    free(t->sp_str); free(t->sp_len); free(t->sp_id);
    free(t);
}

void* orc_create(const uint8_t* vocab, size_t vocab_len, const char* pattern, int byte_level,
                 const char* const* sp_strs, const uint32_t* sp_ids, size_t n_sp, char* err, size_t errcap) {
    if (load_pcre2(err, errcap)) return NULL;
    orc* t = (orc*)calloc(1, sizeof(orc));
    if (!t) return NULL;
    t->sentencepiece = (byte_level & 2) != 0;      /* `byte_level` carries the mode flags: 1 byte-level, 2 SentencePiece */
    t->byte_level = (byte_level & 1) != 0;
    build_byte_level(t);
    if (parse_vocab(t, vocab, vocab_len)) { snprintf(err, errcap, "%s", t->err[0] ? t->err : "out of memory"); orc_destroy(t); return NULL; }
    int ec = 0; size_t eo = 0;
    t->re = P.compile((const uint8_t*)pattern, strlen(pattern), PCRE2_UTF | PCRE2_UCP, &ec, &eo, NULL);   /* tokenizer.rs:474-481 */
    if (!t->re) { snprintf(err, errcap, "Regex compilation error: pcre2 code %d at %zu", ec, eo); orc_destroy(t); return NULL; }
    P.jit_compile(t->re, PCRE2_JIT_COMPLETE);
    t->n_sp = (uint32_t)n_sp;
    t->sp_str = (uint8_t**)calloc(n_sp + 1, sizeof(uint8_t*));
    t->sp_len = (uint32_t*)calloc(n_sp + 1, 4);
    t->sp_id = (uint32_t*)calloc(n_sp + 1, 4);
    for (size_t k = 0; k < n_sp; ++k) {
        size_t l = strlen(sp_strs[k]);
        t->sp_str[k] = (uint8_t*)malloc(l + 1);
        memcpy(t->sp_str[k], sp_strs[k], l + 1);
        t->sp_len[k] = (uint32_t)l; t->sp_id[k] = sp_ids[k];
        if (l) t->sp_last[(uint8_t)sp_strs[k][l - 1]] = 1;
    }
    return t;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* encode_batch: documents i = bytes[offsets[i] .. offsets[i+1]).  ids_out has room for ids_cap
 * entries; out_offsets n_docs+1.  Returns the id count, or <0 (-3 = ids_cap too small). */
long long orc_encode_batch(void* h, const uint8_t* bytes, const uint64_t* offsets, size_t n_docs,
                           int with_special, int n_threads, uint32_t* ids_out, size_t ids_cap, uint64_t* out_offsets) {
    const orc* t = (const orc*)h;
    u32vec* res = (u32vec*)calloc(n_docs ? n_docs : 1, sizeof(u32vec));
    if (!res) return -1;
    int fail = 0;
    if (n_threads <= 0) n_threads = orc_max_threads();
#pragma omp parallel num_threads(n_threads)
    {
        scratch sc; memset(&sc, 0, sizeof sc);
        sc.md = P.md_create(t->re, NULL);
        sc.mc = P.mc_create(NULL);
        sc.js = P.js_create(64 * 1024, 64 * 1024 * 1024, NULL);
        if (sc.js) P.js_assign(sc.mc, NULL, sc.js);
#pragma omp for schedule(dynamic, 16)
        for (long long d = 0; d < (long long)n_docs; ++d) {
            const uint8_t* p = bytes + offsets[d];
            size_t n = (size_t)(offsets[d + 1] - offsets[d]);
            int rc = with_special ? encode_text_special(t, &sc, p, n, &res[d]) : encode_text(t, &sc, p, n, &res[d]);
            if (rc) {
#pragma omp atomic write
                fail = rc;
            }
        }
        P.md_free(sc.md); P.mc_free(sc.mc); if (sc.js) P.js_free(sc.js);
        free(sc.nodes); free(sc.bl);
    }
    long long total = 0;
    out_offsets[0] = 0;
    for (size_t d = 0; d < n_docs; ++d) { total += (long long)res[d].n; out_offsets[d + 1] = (uint64_t)total; }
    long long ret = total;
    if (fail) ret = fail;
    else if ((size_t)total > ids_cap) ret = -3;
    else {
#pragma omp parallel for num_threads(n_threads) schedule(static)
        for (long long d = 0; d < (long long)n_docs; ++d)
            if (res[d].n) memcpy(ids_out + out_offsets[d], res[d].v, res[d].n * 4);
    }
    for (size_t d = 0; d < n_docs; ++d) free(res[d].v);
    free(res);
    return ret;
}

/* piece boundaries of one text (for pre-tokenizer cross-checks): starts[i] = 1 at piece starts */
int orc_split(void* h, const uint8_t* text, size_t n, uint8_t* starts) {
    const orc* t = (const orc*)h;
    memset(starts, 0, n + 1);
    pcre2_match_data* md = P.md_create(t->re, NULL);
    size_t pos = 0; int ret = 0;
    while (pos < n) {
        int rc = P.match(t->re, text, n, pos, PCRE2_NO_UTF_CHECK | PCRE2_NO_JIT, md, NULL);
        if (rc == -1) break;
        if (rc < 0) { ret = rc; break; }
        size_t* ov = P.ovector(md);
        if (ov[1] <= ov[0]) { pos = ov[0] + 1; continue; }
        starts[ov[0]] = 1;
        pos = ov[1];
    }
    P.md_free(md);
    return ret;
}
