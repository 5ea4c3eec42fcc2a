// Device encode path: packed UTF-8 bytes + document offsets  ->  token ids + offsets.
//
//   k_mark_docs      document starts -> `hard` bitmap, per-tile first-document index
//   k_mark_specials  (encode_with_special) special-token spans -> `hard` edges + `spec` bytes
//                    replaces the Aho-Corasick scan of tokenizer.rs:842-874
//   k_pretok         piece-start bitmap: the split regex as class rules (spl_pretok.h)
//                    replaces regex find_iter, tokenizer.rs:244-257 / :731
//   k_encode         per piece: whole-piece probe (tokenizer.rs:703-705, bpe.rs:73-80), else
//                    leftmost-min-rank BPE merge (bpe.rs:83-194) with one warp per piece;
//                    ordered compaction of the ids through a decoupled look-back scan
//                    replaces encode_chunk_with_position + byte_pair_encode + the Rayon
//                    collect of encode_batch (tokenizer.rs:932-934)
//
// Integer / byte work, bounded by HBM traffic and L2 probe latency; no tensor cores.
#include "spl_kernels.cuh"
#include "spl_pretok.h"
#include "spl_pretok_fast.h"

#define FULL 0xFFFFFFFFu

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sm_next_bit(const uint32_t* w, uint32_t from, uint32_t lim) {
    if (from >= lim) return lim;
    uint32_t wi = from >> 5;
    uint32_t v = w[wi] & (FULL << (from & 31));
    for (;;) {
        if (v) { uint32_t p = (wi << 5) + __ffs(v) - 1; return p < lim ? p : lim; }
        ++wi;
        if ((wi << 5) >= lim) return lim;
        v = w[wi];
    }
}

// last set bit in [lo, before), or SPL_RANK_NONE
__device__ __forceinline__ uint32_t sm_prev_bit(const uint32_t* w, uint32_t before, uint32_t lo) {
    if (before <= lo) return SPL_RANK_NONE;
    uint32_t i = before - 1, wi = i >> 5;
    uint32_t v = w[wi] & (FULL >> (31 - (i & 31)));
    for (;;) {
        if (v) { uint32_t p = (wi << 5) + 31 - __clz(v); return p >= lo ? p : SPL_RANK_NONE; }
        if ((wi << 5) <= lo) return SPL_RANK_NONE;
        --wi;
        v = w[wi];
    }
}

__device__ __forceinline__ uint32_t g_next_bit(const uint32_t* __restrict__ w, uint32_t from, uint32_t lim) {
    if (from >= lim) return lim;
    uint32_t wi = from >> 5;
    uint32_t v = __ldg(w + wi) & (FULL << (from & 31));
    for (;;) {
        if (v) { uint32_t p = (wi << 5) + __ffs(v) - 1; return p < lim ? p : lim; }
        ++wi;
        if ((wi << 5) >= lim) return lim;
        v = __ldg(w + wi);
    }
}

__device__ __forceinline__ uint32_t pair_lookup(const uint64_t* __restrict__ tab, uint32_t log2, uint32_t l, uint32_t r) {
    uint64_t key = spl_pair_key(l, r);
    uint32_t mask = (1u << log2) - 1, h = spl_pair_hash(key, log2);
    for (;;) {
        uint64_t e = __ldg(tab + h);
        if ((e >> SPL_SYM_BITS) == key) return (uint32_t)e & ((1u << SPL_SYM_BITS) - 1);
        if (e == SPL_PAIR_EMPTY) return SPL_RANK_NONE;
        h = (h + 1) & mask;
    }
}

__device__ __forceinline__ uint32_t lookup8(const SplKey8* __restrict__ t, uint32_t log2, uint64_t k0, uint32_t len) {
    uint32_t mask = (1u << log2) - 1, h = spl_hash8(k0, len, log2);
    for (;;) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(t + h));
        if (v.w == 0) return SPL_RANK_NONE;
        if (v.w == len && v.x == (uint32_t)k0 && v.y == (uint32_t)(k0 >> 32)) return v.z;
        h = (h + 1) & mask;
    }
}

__device__ __forceinline__ uint32_t lookup16(const SplKey16* __restrict__ t, uint32_t log2, uint64_t k0, uint64_t k1, uint32_t len) {
    uint32_t mask = (1u << log2) - 1, h = spl_hash16(k0, k1, len, log2);
    for (;;) {
        const uint4* p = reinterpret_cast<const uint4*>(t + h);
        uint4 b = __ldg(p + 1);                       // {id, len, pad, pad}
        if (b.y == 0) return SPL_RANK_NONE;
        if (b.y == len) {
            uint4 a = __ldg(p);                       // {k0, k1}
            if (a.x == (uint32_t)k0 && a.y == (uint32_t)(k0 >> 32) && a.z == (uint32_t)k1 && a.w == (uint32_t)(k1 >> 32)) return b.x;
        }
        h = (h + 1) & mask;
    }
}

// ------------------------------------------------------------------------------------------
// k_mark_docs
// ------------------------------------------------------------------------------------------
__global__ void k_mark_docs(SplWork w) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d > w.n_docs) return;
    uint64_t s = w.doc_off[d] - w.off_base;
    uint64_t prev = d ? w.doc_off[d - 1] - w.off_base : 0;
    bool bad = w.doc_off[d] < w.off_base || s > w.N || s < prev || (d == 0 && s != 0) || (d == w.n_docs && s != w.N);
    if (bad) { atomicOr(&w.counters[1], SPL_DEVERR_OFFSETS); return; }
    uint32_t p = (uint32_t)s;
    atomicOr(&w.hard[p >> 5], 1u << (p & 31));
    uint32_t t_lo = d ? (uint32_t)(prev / SPL_TILE) + 1 : 0, t_hi = p / SPL_TILE;
    for (uint32_t t = t_lo; t <= t_hi; ++t) w.tile_first_doc[t] = d;
    if (d == w.n_docs) {
        w.tile_first_doc[w.n_tiles] = w.n_docs + 1;
        atomicOr(&w.pstart[p >> 5], 1u << (p & 31));            // sentinel piece start at N
    }
}

// ------------------------------------------------------------------------------------------
// k_mark_specials: every occurrence of a special string that lies inside one document
// ------------------------------------------------------------------------------------------
__global__ void k_mark_specials(SplWork w) {
    const SplTables* T = w.T;
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < w.N; i += stride) {
        uint32_t b = w.text[i];
        if (!((T->sp_first[b >> 5] >> (b & 31)) & 1u)) continue;
        for (uint32_t s = 0; s < T->n_special; ++s) {
            uint32_t o = T->sp_off[s], len = T->sp_off[s + 1] - o;
            if (len > w.N - i) continue;
            bool eq = true;
            for (uint32_t j = 0; j < len; ++j)
                if (w.text[i + j] != T->sp_bytes[o + j]) { eq = false; break; }
            if (!eq) continue;
            if (g_next_bit(w.hard, i + 1, i + len) < i + len) continue;     // would span two documents
            atomicOr(&w.hard[i >> 5], 1u << (i & 31));
            atomicOr(&w.hard[(i + len) >> 5], 1u << ((i + len) & 31));
            for (uint32_t j = i; j < i + len; ++j) atomicOr(&w.spec[j >> 5], 1u << (j & 31));
            break;
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_pretok: one tile per block, 16 bytes per thread (spl_pretok_chunk does the work)
// ------------------------------------------------------------------------------------------
#define PT_LEFT 16u
#define PT_HW   ((SPL_WIN / 32u) + 2u)       // hard/spec words staged: one before the tile, one after the halo

struct DevEnv {
    const uint8_t* sm_text; uint32_t w0, wlen; const uint8_t* g_text;
    const uint32_t* sm_hard; const uint32_t* sm_spec; uint32_t hw0;
    const uint32_t* g_hard; const uint32_t* g_spec;
    uint32_t* sm_ps; uint32_t tile0; uint32_t* g_ps;
    uint32_t W;

    __device__ __forceinline__ uint8_t byte(uint32_t i) const {
        uint32_t r = i - w0;
        return r < wlen ? sm_text[r] : __ldg(g_text + i);
    }
    __device__ __forceinline__ uint32_t hword(uint32_t wi) const {
        uint32_t r = wi - hw0;
        return r < PT_HW ? sm_hard[r] : __ldg(g_hard + wi);
    }
    __device__ __forceinline__ bool hard(uint32_t i) const { return (hword(i >> 5) >> (i & 31)) & 1u; }
    __device__ __forceinline__ bool spec(uint32_t i) const {
        uint32_t wi = i >> 5, r = wi - hw0;
        uint32_t v = r < PT_HW ? sm_spec[r] : __ldg(g_spec + wi);
        return (v >> (i & 31)) & 1u;
    }
    __device__ __forceinline__ uint32_t next_hard(uint32_t from, uint32_t lim) const {
        if (from >= lim) return lim;
        uint32_t wi = from >> 5;
        uint32_t v = hword(wi) & (FULL << (from & 31));
        for (;;) {
            if (v) { uint32_t p = (wi << 5) + __ffs(v) - 1; return p < lim ? p : lim; }
            ++wi;
            if ((wi << 5) >= lim) return lim;
            v = hword(wi);
        }
    }
    __device__ __forceinline__ uint32_t win_end() const { return W; }
    __device__ __forceinline__ void mark(uint32_t p) {
        uint32_t r = p - tile0;
        if (r < SPL_TILE) atomicOr(&sm_ps[r >> 5], 1u << (r & 31));
        else atomicOr(&g_ps[p >> 5], 1u << (p & 31));
    }
};

struct PretokSmem {
    __align__(16) uint8_t text[PT_LEFT + SPL_WIN];
    uint32_t hard[PT_HW];
    uint32_t spec[PT_HW];
    uint32_t ps[SPL_TILE / 32];
};

// sequential rules over one SPL_TILE-byte tile (whole block); leadin: also cover the pieces that start between
// tile0 and the tile's first sync point (needed when the tile to the left is not processed by this routine)
__device__ void pretok_tile(const SplWork& w, PretokSmem& sm, uint32_t tile0, bool leadin) {
    const uint32_t tid = threadIdx.x;
    const uint32_t N = w.N;
    const uint32_t w0 = tile0 >= PT_LEFT ? tile0 - PT_LEFT : 0;       // window start (16-aligned)
    const uint32_t lead = tile0 - w0;                                  // 0 or 16
    const uint32_t Nup = (N + 15u) & ~15u;

    // stage text [w0, tile0 + WIN) with 16-byte loads
    for (uint32_t v = tid; v < (lead + SPL_WIN) / 16; v += SPL_THREADS) {
        uint32_t g = w0 + v * 16;
        uint4 x = make_uint4(0, 0, 0, 0);
        if (g < Nup) x = __ldg(reinterpret_cast<const uint4*>(w.text + g));
        *reinterpret_cast<uint4*>(sm.text + v * 16) = x;
    }
    const uint32_t hw0 = (tile0 >> 5) - (tile0 ? 1u : 0u);
    for (uint32_t v = tid; v < PT_HW; v += SPL_THREADS) {
        sm.hard[v] = __ldg(w.hard + hw0 + v);
        sm.spec[v] = w.with_special ? __ldg(w.spec + hw0 + v) : 0u;
    }
    if (tid < SPL_TILE / 32) sm.ps[tid] = 0;
    __syncthreads();

    DevEnv env;
    env.sm_text = sm.text; env.w0 = w0; env.wlen = lead + SPL_WIN; env.g_text = w.text;
    env.sm_hard = sm.hard; env.sm_spec = sm.spec; env.hw0 = hw0; env.g_hard = w.hard; env.g_spec = w.spec;
    env.sm_ps = sm.ps; env.tile0 = tile0; env.g_ps = w.pstart;
    env.W = tile0 + SPL_WIN;

    const SplTables* T = w.T;
    uint32_t c0 = tile0 + tid * 16;
    if (c0 < N) {
        uint32_t c1 = c0 + 16 < N ? c0 + 16 : N;
        spl_pretok_chunk(env, c0, c1, N, T->ucd_stage1, T->ucd_stage2, w.pattern, w.with_special);
    }
    if (leadin && tid == SPL_THREADS - 1)
        spl_pretok_leadin(env, tile0, N, T->ucd_stage1, T->ucd_stage2, w.pattern, w.with_special);
    __syncthreads();
    if (tid < SPL_TILE / 32) {
        uint32_t v = sm.ps[tid];
        if (v) atomicOr(&w.pstart[(tile0 >> 5) + tid], v);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SPL_THREADS) k_pretok(SplWork w) {
    __shared__ PretokSmem sm;
    pretok_tile(w, sm, blockIdx.x * SPL_TILE, false);
}

// ------------------------------------------------------------------------------------------
// k_pretok_fast: the bit-parallel formulation (spl_pretok_fast.h).  One thread per 32-byte word,
// FAST_HALO words of context on each side of FAST_PAYLOAD payload words.  A tile that cannot be
// decided here is appended to the fallback list and left untouched.
// ------------------------------------------------------------------------------------------
struct FastGText {
    const uint8_t* p;
    __device__ __forceinline__ uint8_t byte(uint32_t i) const { return __ldg(p + i); }
};

struct FastSmem {
    uint32_t m[FM_COUNT][SPL_FAST_THREADS];
    uint32_t hardw[SPL_FAST_THREADS];
    uint32_t specw[SPL_FAST_THREADS];
    uint32_t sum[SPL_FAST_THREADS];
};

struct FastMasks {
    const FastSmem* s; int gw0; uint32_t N;
    __device__ __forceinline__ uint32_t get(int q, int k) const { return s->m[q][k]; }
    __device__ __forceinline__ uint32_t hard(int k) const { return s->hardw[k]; }
    __device__ __forceinline__ uint32_t spec(int k) const { return s->specw[k]; }
    __device__ __forceinline__ uint32_t summary(int k) const { return s->sum[k]; }
    __device__ __forceinline__ uint32_t valid(int k) const {
        int gw = gw0 + k;
        if (gw < 0) return 0u;
        uint32_t base = (uint32_t)gw * 32u;
        if (base >= N) return 0u;
        return (N - base >= 32u) ? 0xFFFFFFFFu : ((1u << (N - base)) - 1u);
    }
};

__global__ void __launch_bounds__(SPL_FAST_THREADS) k_pretok_fast(SplWork w) {
    __shared__ FastSmem sm;
    const int k = threadIdx.x;
    const int gw0 = (int)(blockIdx.x * SPL_FAST_PAYLOAD) - (int)SPL_FAST_HALO;
    const int gw = gw0 + k;
    const uint32_t N = w.N;
    const SplTables* T = w.T;
    const uint32_t last_word = N >> 5;                        // the word that holds the sentinel bit N

    // ---- phase A: classify my word ------------------------------------------------------------------
    SplFastWord fw;
#pragma unroll
    for (int q = 0; q <= FM_BAD; ++q) fw.m[q] = 0;
    uint32_t hw = 0, sw = 0;
    if (gw >= 0 && (uint32_t)gw <= last_word) {
        const uint32_t base = (uint32_t)gw * 32u;
        hw = __ldg(w.hard + gw);
        if (w.with_special) sw = __ldg(w.spec + gw);
        if (base < N) {
            uint32_t xw[8];
            const uint4* p4 = reinterpret_cast<const uint4*>(w.text + base);
            uint4 a = __ldg(p4);
            uint4 b = (base + 16u < ((N + 15u) & ~15u)) ? __ldg(p4 + 1) : make_uint4(0, 0, 0, 0);
            xw[0] = a.x; xw[1] = a.y; xw[2] = a.z; xw[3] = a.w; xw[4] = b.x; xw[5] = b.y; xw[6] = b.z; xw[7] = b.w;
            FastGText t{w.text};
            fw = spl_fast_classify(t, xw, base, N, T->ucd_stage1, T->ucd_stage2, w.pattern);
        }
    }
#pragma unroll
    for (int q = 0; q <= FM_BAD; ++q) sm.m[q][k] = fw.m[q];
    sm.hardw[k] = hw; sm.specw[k] = sw;
    __syncthreads();

    // ---- phase B: local masks, fills, summary ------------------------------------------------------------
    FastMasks M{&sm, gw0, N};
    SplFastLocal loc;
    uint32_t s = spl_fast_local(M, k, SPL_FAST_THREADS, w.pattern, loc);
    sm.sum[k] = s;
    sm.m[FM_A2][k] = loc.A2; sm.m[FM_A3][k] = loc.A3;
    __syncthreads();

    // ---- phase C: carries, piece starts --------------------------------------------------------------------
    bool structural = false, unknown = false;
    uint32_t start = spl_fast_final(M, loc, k, SPL_FAST_THREADS, w.pattern, w.with_special, structural, unknown);
    const bool payload = k >= (int)SPL_FAST_HALO && k < (int)(SPL_FAST_HALO + SPL_FAST_PAYLOAD);
    int flag = ((s & FS_BAD) != 0) || structural || (payload && unknown);
    flag = __syncthreads_or(flag);
    if (flag) {
        if (k == 0) {
            uint32_t idx = atomicAdd(&w.counters[3], 1u);
            w.fb_list[idx] = blockIdx.x;
        }
        return;
    }
    if (payload && (uint32_t)gw <= last_word) {
        if ((uint32_t)gw == last_word) start |= 1u << (N & 31u);
        w.pstart[gw] = start;
    }
}

// tiles the fast path declined: the sequential rules, two SPL_TILE tiles per fast tile
__global__ void __launch_bounds__(SPL_THREADS) k_pretok_fb(SplWork w) {
    __shared__ PretokSmem sm;
    const uint32_t n = w.counters[3];
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint32_t t0 = w.fb_list[i] * (SPL_FAST_PAYLOAD * 32u);
#pragma unroll 1
        for (uint32_t sub = 0; sub < SPL_FAST_PAYLOAD * 32u / SPL_TILE; ++sub) {
            uint32_t tile0 = t0 + sub * SPL_TILE;
            if (tile0 < w.N) pretok_tile(w, sm, tile0, sub == 0);
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_encode
// ------------------------------------------------------------------------------------------
#define EN_WORDS ((SPL_WIN / 32u) + 1u)       // bitmap words covering window positions 0 .. SPL_WIN
#define EN_WARPS (SPL_THREADS / 32)

struct EncSmem {
    uint32_t text[SPL_WIN / 4 + 4];   // staged bytes (+ slack for unaligned 8-byte key loads)
    uint32_t pb[EN_WORDS + 1];        // piece-start bits
    uint32_t tb[EN_WORDS + 1];        // token-start bits
    uint32_t tok[SPL_WIN];            // token id (or symbol during merging) at its first byte
    uint32_t rnk[SPL_WIN];            // rank of the pair (part at i, next part)
    uint32_t wpre[EN_WORDS + 1];      // exclusive token count before each bitmap word
    uint16_t plist[SPL_TILE + 2];     // window positions of the tile's piece starts, in order (+ end of the last piece)
    uint16_t mlist[SPL_TILE];         // pieces that missed the whole-piece probe: short ones from the bottom, long from the top
    uint32_t wtot[EN_WARPS];
    uint32_t n_short, n_long;
    uint32_t tile, huge_start, huge_end, huge_cnt, huge_off;
    uint64_t prefix;
    uint64_t red[SPL_THREADS];        // block reductions of the out-of-window path
};

__device__ __forceinline__ uint8_t sm_byte(const uint32_t* words, uint32_t i) {
    return reinterpret_cast<const uint8_t*>(words)[i];
}

__device__ __forceinline__ uint64_t sm_load8(const uint32_t* words, uint32_t s) {
    uint32_t wi = s >> 2, sh = (s & 3u) * 8u;
    uint32_t a = words[wi], b = words[wi + 1], c = words[wi + 2];
    uint32_t lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);
    return (uint64_t)lo | ((uint64_t)hi << 32);
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        uint32_t lo = __shfl_xor_sync(FULL, (uint32_t)v, o), hi = __shfl_xor_sync(FULL, (uint32_t)(v >> 32), o);
        v += (uint64_t)lo | ((uint64_t)hi << 32);
    }
    return v;
}

// special-token id of the span text[s, s+len) (linear search; special spans are rare)
template <class ByteAt>
__device__ uint32_t special_id(const SplTables* T, ByteAt at, uint32_t len) {
    for (uint32_t k = 0; k < T->n_special; ++k) {
        uint32_t o = T->sp_off[k];
        if (T->sp_off[k + 1] - o != len) continue;
        bool eq = true;
        for (uint32_t j = 0; j < len; ++j)
            if (at(j) != T->sp_bytes[o + j]) { eq = false; break; }
        if (eq) return T->sp_id[k];
    }
    return SPL_RANK_NONE;
}

// One warp encodes the piece occupying window bytes [s, e): long-key whole-piece probe, then
// the merge loop.  Parts are delimited by the bits of sm.tb; sm.tok holds each part's symbol
// at its first byte, sm.rnk the rank of (part, next part).
__device__ void bpe_piece_warp(EncSmem& sm, const SplTables* T, uint32_t s, uint32_t e) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t len = e - s;
    if (len > 16 && len <= T->max_key_len) {
        uint64_t sum = 0;
        for (uint32_t i = lane; i * 8 < len; i += 32) {
            uint64_t wv = sm_load8(sm.text, s + i * 8);
            uint32_t rem = len - i * 8;
            if (rem < 8) wv &= (1ull << (8 * rem)) - 1;
            sum += spl_hashL_word(wv, i);
        }
        sum = warp_sum_u64(sum);
        uint64_t hv = spl_hashL_final(sum, len);
        uint32_t mask = (1u << T->tl_log2) - 1, h = (uint32_t)(hv >> (64 - T->tl_log2));
        for (;;) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(T->tl + h));     // {hash lo, hash hi, id, len}
            if (v.w == 0) break;
            if (v.w == len && v.x == (uint32_t)hv && v.y == (uint32_t)(hv >> 32)) {
                const uint8_t* kb = T->tok_bytes + __ldg(T->tok_off + v.z);
                bool ok = true;
                for (uint32_t j = lane; j < len; j += 32) ok &= (sm_byte(sm.text, s + j) == __ldg(kb + j));
                if (__all_sync(FULL, ok)) {
                    if (lane == 0) sm.tok[s] = v.z;
                    return;
                }
            }
            h = (h + 1) & mask;
        }
    }
    // every byte becomes a part
    for (uint32_t j = s + lane; j < e; j += 32) {
        sm.tok[j] = T->byte_sym[sm_byte(sm.text, j)];
        atomicOr(&sm.tb[j >> 5], 1u << (j & 31));
    }
    __syncwarp();
    for (uint32_t j = s + lane; j < e; j += 32)
        sm.rnk[j] = (j + 1 < e) ? pair_lookup(T->pair, T->pair_log2, sm.tok[j], sm.tok[j + 1]) : SPL_RANK_NONE;
    __syncwarp();
    for (;;) {
        uint32_t best = SPL_RANK_NONE, bpos = SPL_RANK_NONE;
        for (uint32_t j = s + lane; j < e; j += 32) {
            uint32_t r = sm.rnk[j];
            if (r < best) { best = r; bpos = j; }
        }
        uint32_t m = __reduce_min_sync(FULL, best);
        if (m == SPL_RANK_NONE) break;
        uint32_t pos = __reduce_min_sync(FULL, best == m ? bpos : SPL_RANK_NONE);   // leftmost minimum
        uint32_t j = sm_next_bit(sm.tb, pos + 1, e);          // the part being absorbed
        uint32_t k = sm_next_bit(sm.tb, j + 1, e);            // its right neighbour (e if none)
        uint32_t h = sm_prev_bit(sm.tb, pos, s);              // left neighbour (NONE if none)
        uint32_t symk = k < e ? sm.tok[k] : 0u;
        uint32_t symh = h != SPL_RANK_NONE ? sm.tok[h] : 0u;
        __syncwarp();
        if (lane == 0) {
            sm.tok[pos] = m;                                   // merged id == its rank
            sm.rnk[j] = SPL_RANK_NONE;
            atomicAnd(&sm.tb[j >> 5], ~(1u << (j & 31)));
            sm.rnk[pos] = k < e ? pair_lookup(T->pair, T->pair_log2, m, symk) : SPL_RANK_NONE;
        } else if (lane == 1 && h != SPL_RANK_NONE) {
            sm.rnk[h] = pair_lookup(T->pair, T->pair_log2, symh, m);
        }
        __syncwarp();
    }
    // bytes that are not in the vocabulary produce no id (bpe.rs:187-191)
    for (uint32_t j = s + lane; j < e; j += 32)
        if (((sm.tb[j >> 5] >> (j & 31)) & 1u) && sm.tok[j] >= SPL_UNK_BASE)
            atomicAnd(&sm.tb[j >> 5], ~(1u << (j & 31)));
    __syncwarp();
}

// The whole block encodes one piece that does not fit the staging window: text bytes
// [gs, ge) read from global memory; sym / rnk / next / prev arrays live in the scratch pool.
// Returns (to every thread) the number of ids; they stay in sym[] (dead parts = NONE).
__device__ uint32_t bpe_piece_block(EncSmem& sm, const SplWork& w, uint32_t gs, uint32_t ge, uint32_t* scratch) {
    const SplTables* T = w.T;
    const uint32_t tid = threadIdx.x, len = ge - gs;
    uint32_t* sym = scratch;
    uint32_t* rnk = scratch + len;
    uint32_t* nxt = scratch + 2 * (size_t)len;
    uint32_t* prv = scratch + 3 * (size_t)len;
    const uint8_t* tx = w.text + gs;

    if (w.with_special && ((__ldg(w.spec + (gs >> 5)) >> (gs & 31)) & 1u)) {
        if (tid == 0) {
            uint32_t id = special_id(T, [&](uint32_t j) { return __ldg(tx + j); }, len);
            sym[0] = id;
        }
        for (uint32_t j = tid + 1; j < len; j += SPL_THREADS) sym[j] = SPL_RANK_NONE;
        __syncthreads();
        return sym[0] == SPL_RANK_NONE ? 0u : 1u;
    }
    if (len <= T->max_key_len) {                     // only for vocabularies with keys longer than the halo
        if (tid == 0) {
            uint64_t sum = 0;
            for (uint32_t i = 0; i * 8 < len; ++i) {
                uint64_t wv = 0;
                for (uint32_t b = 0; b < 8 && i * 8 + b < len; ++b) wv |= (uint64_t)__ldg(tx + i * 8 + b) << (8 * b);
                sum += spl_hashL_word(wv, i);
            }
            uint64_t hv = spl_hashL_final(sum, len);
            uint32_t mask = (1u << T->tl_log2) - 1, h = (uint32_t)(hv >> (64 - T->tl_log2));
            uint32_t found = SPL_RANK_NONE;
            for (;;) {
                SplKeyL k = T->tl[h];
                if (k.len == 0) break;
                if (k.len == len && k.hash == hv) {
                    const uint8_t* kb = T->tok_bytes + T->tok_off[k.id];
                    bool ok = true;
                    for (uint32_t j = 0; j < len; ++j) if (kb[j] != __ldg(tx + j)) { ok = false; break; }
                    if (ok) { found = k.id; break; }
                }
                h = (h + 1) & mask;
            }
            sm.huge_cnt = found;
        }
        __syncthreads();
        uint32_t found = sm.huge_cnt;
        __syncthreads();
        if (found != SPL_RANK_NONE) {
            for (uint32_t j = tid; j < len; j += SPL_THREADS) sym[j] = j ? SPL_RANK_NONE : found;
            __syncthreads();
            return 1u;
        }
    }
    for (uint32_t j = tid; j < len; j += SPL_THREADS) {
        sym[j] = T->byte_sym[__ldg(tx + j)];
        nxt[j] = j + 1;                              // len == "no next"
        prv[j] = j ? j - 1 : SPL_RANK_NONE;
    }
    __syncthreads();
    for (uint32_t j = tid; j < len; j += SPL_THREADS)
        rnk[j] = (j + 1 < len) ? pair_lookup(T->pair, T->pair_log2, sym[j], sym[j + 1]) : SPL_RANK_NONE;
    __syncthreads();
    for (;;) {
        uint64_t best = ~0ull;                       // (rank << 32) | position : min = leftmost minimum
        for (uint32_t j = tid; j < len; j += SPL_THREADS) {
            uint64_t v = ((uint64_t)rnk[j] << 32) | j;
            if (v < best) best = v;
        }
        sm.red[tid] = best;
        __syncthreads();
        for (uint32_t o = SPL_THREADS / 2; o; o >>= 1) {
            if (tid < o && sm.red[tid + o] < sm.red[tid]) sm.red[tid] = sm.red[tid + o];
            __syncthreads();
        }
        uint64_t mn = sm.red[0];
        __syncthreads();
        uint32_t m = (uint32_t)(mn >> 32), pos = (uint32_t)mn;
        if (m == SPL_RANK_NONE) break;
        if (tid == 0) {
            uint32_t j = nxt[pos], k = nxt[j], h = prv[pos];
            sym[pos] = m; sym[j] = SPL_RANK_NONE; rnk[j] = SPL_RANK_NONE;
            nxt[pos] = k;
            if (k < len) prv[k] = pos;
            rnk[pos] = k < len ? pair_lookup(T->pair, T->pair_log2, m, sym[k]) : SPL_RANK_NONE;
            if (h != SPL_RANK_NONE) rnk[h] = pair_lookup(T->pair, T->pair_log2, sym[h], m);
        }
        __syncthreads();
    }
    // count surviving known symbols; unknown single bytes are dropped
    uint32_t cnt = 0;
    for (uint32_t j = tid; j < len; j += SPL_THREADS) {
        uint32_t sv = sym[j];
        if (sv != SPL_RANK_NONE && sv >= SPL_UNK_BASE) { sym[j] = SPL_RANK_NONE; sv = SPL_RANK_NONE; }
        cnt += (sv != SPL_RANK_NONE);
    }
    sm.red[tid] = cnt;
    __syncthreads();
    for (uint32_t o = SPL_THREADS / 2; o; o >>= 1) {
        if (tid < o) sm.red[tid] += sm.red[tid + o];
        __syncthreads();
    }
    uint32_t total = (uint32_t)sm.red[0];
    __syncthreads();
    return total;
}

// two independent pair probes issued back to back (the two re-ranks after a merge)
__device__ __forceinline__ void pair_lookup2(const uint64_t* __restrict__ tab, uint32_t log2,
                                             bool va, uint32_t la, uint32_t ra, bool vb, uint32_t lb, uint32_t rb,
                                             uint32_t& outa, uint32_t& outb) {
    const uint32_t mask = (1u << log2) - 1, symmask = (1u << SPL_SYM_BITS) - 1;
    uint64_t ka = spl_pair_key(la, ra), kb = spl_pair_key(lb, rb);
    uint32_t ha = spl_pair_hash(ka, log2), hb = spl_pair_hash(kb, log2);
    uint64_t ea = va ? __ldg(tab + ha) : SPL_PAIR_EMPTY;
    uint64_t eb = vb ? __ldg(tab + hb) : SPL_PAIR_EMPTY;
    outa = SPL_RANK_NONE; outb = SPL_RANK_NONE;
    while (ea != SPL_PAIR_EMPTY) {
        if ((ea >> SPL_SYM_BITS) == ka) { outa = (uint32_t)ea & symmask; break; }
        ha = (ha + 1) & mask; ea = __ldg(tab + ha);
    }
    while (eb != SPL_PAIR_EMPTY) {
        if ((eb >> SPL_SYM_BITS) == kb) { outb = (uint32_t)eb & symmask; break; }
        hb = (hb + 1) & mask; eb = __ldg(tab + hb);
    }
}

// whole-piece probe of a 17..32-byte piece by one thread (long-key table, verified against the token bytes)
__device__ uint32_t lookupL_thread(const SplTables* T, const uint32_t* text, uint32_t s, uint32_t len) {
    if (len > T->max_key_len) return SPL_RANK_NONE;
    uint64_t sum = 0;
    for (uint32_t i = 0; i * 8 < len; ++i) {
        uint64_t wv = sm_load8(text, s + i * 8);
        uint32_t rem = len - i * 8;
        if (rem < 8) wv &= (1ull << (8 * rem)) - 1;
        sum += spl_hashL_word(wv, i);
    }
    uint64_t hv = spl_hashL_final(sum, len);
    uint32_t mask = (1u << T->tl_log2) - 1, h = (uint32_t)(hv >> (64 - T->tl_log2));
    for (;;) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(T->tl + h));     // {hash lo, hash hi, id, len}
        if (v.w == 0) return SPL_RANK_NONE;
        if (v.w == len && v.x == (uint32_t)hv && v.y == (uint32_t)(hv >> 32)) {
            const uint8_t* kb = T->tok_bytes + __ldg(T->tok_off + v.z);
            bool ok = true;
            for (uint32_t j = 0; j < len; ++j) ok &= (sm_byte(text, s + j) == __ldg(kb + j));
            if (ok) return v.z;
        }
        h = (h + 1) & mask;
    }
}

// One THREAD merges the piece at window bytes [s, s+n), 1 <= n <= 32 (bpe.rs:83-194): parts are the set bits of
// `live` (bit i = a part starts at byte s+i), their symbols sit in sm.tok, the rank of (part, next part) in sm.rnk.
// 32 pieces merge side by side in a warp, so the probe latency of the re-ranks overlaps across pieces.
__device__ void bpe_piece_thread(EncSmem& sm, const SplTables* T, uint32_t s, uint32_t n) {
    const uint64_t* __restrict__ ptab = T->pair;
    const uint32_t plog = T->pair_log2;
    for (uint32_t i = 0; i < n; ++i) sm.tok[s + i] = T->byte_sym[sm_byte(sm.text, s + i)];
    for (uint32_t i = 0; i + 1 < n; i += 2) {
        uint32_t ra, rb;
        bool vb = i + 2 < n;
        pair_lookup2(ptab, plog, true, sm.tok[s + i], sm.tok[s + i + 1], vb, vb ? sm.tok[s + i + 1] : 0u, vb ? sm.tok[s + i + 2] : 0u, ra, rb);
        sm.rnk[s + i] = ra;
        if (vb) sm.rnk[s + i + 1] = rb;
    }
    sm.rnk[s + n - 1] = SPL_RANK_NONE;
    uint32_t live = n >= 32u ? 0xFFFFFFFFu : ((1u << n) - 1u);
    for (;;) {
        uint32_t best = SPL_RANK_NONE, bpos = 0;
        for (uint32_t m = live; m; m &= m - 1) {
            uint32_t i = __ffs(m) - 1;
            uint32_t r = sm.rnk[s + i];
            if (r < best) { best = r; bpos = i; }                 // strict <: leftmost minimum (bpe.rs:133)
        }
        if (best == SPL_RANK_NONE) break;
        uint32_t above = bpos >= 31u ? 0u : (live & ~((2u << bpos) - 1u));
        uint32_t nx = __ffs(above) - 1;                            // the absorbed part (exists: its pair has a rank)
        uint32_t above2 = above & (above - 1);
        uint32_t below = live & ((1u << bpos) - 1u);
        bool has_nn = above2 != 0, has_pv = below != 0;
        uint32_t nn = has_nn ? __ffs(above2) - 1 : 0u, pv = has_pv ? 31u - __clz(below) : 0u;
        live &= ~(1u << nx);
        sm.tok[s + bpos] = best;                                   // merged id == its rank
        uint32_t r1, r0;
        pair_lookup2(ptab, plog, has_nn, best, has_nn ? sm.tok[s + nn] : 0u, has_pv, has_pv ? sm.tok[s + pv] : 0u, best, r1, r0);
        sm.rnk[s + bpos] = r1;
        if (has_pv) sm.rnk[s + pv] = r0;
    }
    // bytes that are not in the vocabulary produce no id (bpe.rs:187-191)
    for (uint32_t m = live; m; m &= m - 1) {
        uint32_t i = __ffs(m) - 1;
        if (sm.tok[s + i] >= SPL_UNK_BASE) live &= ~(1u << i);
    }
    // publish the parts: the piece-start bit is already set in tb
    uint32_t w0 = s >> 5, sh = s & 31u;
    uint32_t lo = live << sh, hi = sh ? (live >> (32u - sh)) : 0u;
    if (lo & ~(1u << sh)) atomicOr(&sm.tb[w0], lo);
    if (hi) atomicOr(&sm.tb[w0 + 1], hi);
    if (!(live & 1u)) atomicAnd(&sm.tb[w0], ~(1u << sh));
}

// ------------------------------------------------------------------------------------------
// k_encode: one tile per block iteration (ticketed).  No block waits for another one: the ids of a tile go to a
// bump-allocated chunk of the staging buffer; k_tile_scan + k_gather put them in document order afterwards.
// ------------------------------------------------------------------------------------------
extern __shared__ __align__(16) uint8_t spl_dyn_smem[];

__global__ void __launch_bounds__(SPL_THREADS) k_encode(SplWork w) {
    EncSmem& sm = *reinterpret_cast<EncSmem*>(spl_dyn_smem);
    const SplTables* T = w.T;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t N = w.N, Nup = (N + 15u) & ~15u;

    if (tid == 0) sm.tile = atomicAdd(&w.counters[0], 1u);
    __syncthreads();
    for (;;) {
        const uint32_t tile = sm.tile;
        if (tile >= w.n_tiles) return;
        const uint32_t tile0 = tile * SPL_TILE;
        __syncthreads();                                          // everybody has read sm.tile
        if (tid == 0) sm.tile = atomicAdd(&w.counters[0], 1u);    // next ticket: its latency hides behind this tile

        // ---- stage the window -----------------------------------------------------------
        for (uint32_t v = tid; v < SPL_WIN / 16 + 1; v += SPL_THREADS) {
            uint32_t g = tile0 + v * 16;
            uint4 x = make_uint4(0, 0, 0, 0);
            if (g < Nup) x = __ldg(reinterpret_cast<const uint4*>(w.text + g));
            reinterpret_cast<uint4*>(sm.text)[v] = x;
        }
        for (uint32_t v = tid; v <= EN_WORDS; v += SPL_THREADS) {
            uint32_t pbv = __ldg(w.pstart + (tile0 >> 5) + v);
            sm.pb[v] = pbv;
            sm.tb[v] = v < SPL_TILE / 32 ? pbv : 0u;     // beyond the tile only this tile's last piece adds bits
        }
        if (tid == 0) { sm.huge_start = SPL_RANK_NONE; sm.huge_cnt = 0; sm.n_short = 0; sm.n_long = 0; }
        __syncthreads();

        // ---- piece list: positions of the piece starts of this tile, in order ------------------
        const uint32_t avail = N - tile0;                          // text bytes from tile0 on
        uint32_t my = (sm.pb[tid >> 1] >> ((tid & 1u) * 16u)) & 0xFFFFu;
        if (tid * 16u + 16u > avail) my &= (tid * 16u >= avail) ? 0u : ((1u << (avail - tid * 16u)) - 1u);   // sentinel bit at N
        uint32_t cnt = __popc(my), incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) sm.wtot[warp] = incl;
        __syncthreads();
        uint32_t base = incl - cnt;
#pragma unroll
        for (uint32_t q = 0; q < EN_WARPS; ++q) base += (q < warp) ? sm.wtot[q] : 0u;
        uint32_t P = 0;
#pragma unroll
        for (uint32_t q = 0; q < EN_WARPS; ++q) P += sm.wtot[q];
        while (my) {
            uint32_t b = __ffs(my) - 1;
            my &= my - 1;
            sm.plist[base++] = (uint16_t)(tid * 16u + b);
        }
        if (tid == 0) {
            uint32_t e = sm_next_bit(sm.pb, SPL_TILE < avail ? SPL_TILE : avail, SPL_WIN + 1);
            sm.plist[P] = (uint16_t)(e > SPL_WIN ? SPL_WIN + 1 : e);      // end of the last piece (WIN+1: it leaves the window)
        }
        __syncthreads();

        // ---- one thread per piece: whole-piece probe (tokenizer.rs:703-705, bpe.rs:73-80) --------------
        for (uint32_t j = tid; j < P; j += SPL_THREADS) {
            uint32_t s = sm.plist[j], e = sm.plist[j + 1];
            if (e > SPL_WIN) { sm.huge_start = s; continue; }
            uint32_t len = e - s;
            uint32_t id = SPL_RANK_NONE;
            if (w.with_special && ((__ldg(w.spec + ((tile0 + s) >> 5)) >> ((tile0 + s) & 31)) & 1u)) {
                id = special_id(T, [&](uint32_t q) { return sm_byte(sm.text, s + q); }, len);
                if (id == SPL_RANK_NONE) { atomicAnd(&sm.tb[s >> 5], ~(1u << (s & 31))); continue; }
            } else if (len <= 8) {
                uint64_t k0 = sm_load8(sm.text, s);
                if (len < 8) k0 &= (1ull << (8 * len)) - 1;
                id = lookup8(T->t8, T->t8_log2, k0, len);
            } else if (len <= 16) {
                uint64_t k0 = sm_load8(sm.text, s), k1 = sm_load8(sm.text, s + 8);
                if (len < 16) k1 &= (1ull << (8 * (len - 8))) - 1;
                id = lookup16(T->t16, T->t16_log2, k0, k1, len);
            } else if (len <= 32) {
                id = lookupL_thread(T, sm.text, s, len);
            }
            if (id != SPL_RANK_NONE) sm.tok[s] = id;
            else if (len <= 32) sm.mlist[atomicAdd(&sm.n_short, 1u)] = (uint16_t)j;
            else sm.mlist[SPL_TILE - 1 - atomicAdd(&sm.n_long, 1u)] = (uint16_t)j;      // long pieces fill from the top
        }
        __syncthreads();

        // ---- misses: leftmost-min-rank merge; one thread per short piece, one warp per long piece ---------
        {
            const uint32_t n_short = sm.n_short, n_long = sm.n_long;
            for (uint32_t i = tid; i < n_short; i += SPL_THREADS) {
                uint32_t j = sm.mlist[i];
                uint32_t s = sm.plist[j];
                bpe_piece_thread(sm, T, s, sm.plist[j + 1] - s);
            }
            for (uint32_t i = warp; i < n_long; i += EN_WARPS) {
                uint32_t j = sm.mlist[SPL_TILE - 1 - i];
                bpe_piece_warp(sm, T, sm.plist[j], sm.plist[j + 1]);
            }
        }
        __syncthreads();

        // ---- a piece that outgrew the window: whole block, global scratch ----------------------
        uint32_t huge_cnt = 0;
        uint32_t* huge_scratch = nullptr;
        uint32_t huge_len = 0;
        if (sm.huge_start != SPL_RANK_NONE) {
            uint32_t gs = tile0 + sm.huge_start;
            if (tid == 0) {
                uint32_t ge = g_next_bit(w.pstart, tile0 + SPL_WIN + 1, N + 1);
                sm.huge_end = ge;
                uint32_t need = 4u * (ge - gs);
                uint32_t off = atomicAdd(&w.counters[2], need);
                if ((uint64_t)off + need > w.huge_pool_words) { atomicOr(&w.counters[1], SPL_DEVERR_HUGE_POOL); off = SPL_RANK_NONE; }
                sm.huge_off = off;
                atomicAnd(&sm.tb[sm.huge_start >> 5], ~(1u << (sm.huge_start & 31)));   // its ids are appended separately
            }
            __syncthreads();
            if (sm.huge_off != SPL_RANK_NONE) {
                huge_scratch = w.huge_pool + sm.huge_off;
                huge_len = sm.huge_end - gs;
                huge_cnt = bpe_piece_block(sm, w, gs, sm.huge_end, huge_scratch);
            }
            __syncthreads();
        }

        // ---- count ids, word prefixes, staging chunk ---------------------------------------------------
        if (warp == 0) {
            // exclusive scan of the per-word popcounts (EN_WORDS <= 160: five words per lane)
            uint32_t local[5], run = 0;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                uint32_t v = lane * 5 + q;
                local[q] = v < EN_WORDS ? __popc(sm.tb[v]) : 0u;
                run += local[q];
            }
            uint32_t inc2 = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(FULL, inc2, o);
                if (lane >= (uint32_t)o) inc2 += t;
            }
            uint32_t b2 = inc2 - run;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                uint32_t v = lane * 5 + q;
                if (v <= EN_WORDS) sm.wpre[v] = b2;
                b2 += local[q];
            }
            uint32_t win_cnt = __shfl_sync(FULL, inc2, 31);
            if (lane == 0) {
                uint32_t total = win_cnt + huge_cnt;
                uint64_t off = atomicAdd(reinterpret_cast<unsigned long long*>(w.stage_bump), (unsigned long long)total);
                sm.prefix = off;
                w.tile_cnt[tile] = total;
                w.tile_soff[tile] = off;
            }
        }
        __syncthreads();
        const uint64_t soff = sm.prefix;

        // ---- ids of the tile, in order, into its staging chunk --------------------------------------------
        for (uint32_t hw = tid; hw < EN_WORDS * 2; hw += SPL_THREADS) {
            uint32_t word = sm.tb[hw >> 1], sh = (hw & 1u) * 16u;
            uint32_t mine = (word >> sh) & 0xFFFFu;
            uint64_t o = soff + sm.wpre[hw >> 1] + __popc(word & ((1u << sh) - 1u));
            while (mine) {
                uint32_t b = __ffs(mine) - 1;
                mine &= mine - 1;
                w.stage[o++] = sm.tok[hw * 16 + b];
            }
        }
        uint32_t win_total = sm.wpre[EN_WORDS];
        if (huge_cnt) {
            // block-ordered compaction of the survivors in huge_scratch[0 .. huge_len)
            uint64_t hb = soff + win_total;
            uint32_t per = (huge_len + SPL_THREADS - 1) / SPL_THREADS;
            uint32_t lo = tid * per, hi = lo + per < huge_len ? lo + per : huge_len;
            uint32_t c = 0;
            for (uint32_t j = lo; j < hi; ++j) c += (huge_scratch[j] != SPL_RANK_NONE);
            sm.red[tid] = c;
            __syncthreads();
            if (tid == 0) { uint64_t run = 0; for (int q = 0; q < SPL_THREADS; ++q) { uint64_t t = sm.red[q]; sm.red[q] = run; run += t; } }
            __syncthreads();
            uint64_t o = hb + sm.red[tid];
            for (uint32_t j = lo; j < hi; ++j) { uint32_t v = huge_scratch[j]; if (v != SPL_RANK_NONE) w.stage[o++] = v; }
        }

        // ---- per-document output offsets, relative to the tile (k_gather adds the tile's prefix) -----------
        {
            uint32_t d0 = __ldg(w.tile_first_doc + tile), d1 = __ldg(w.tile_first_doc + tile + 1);
            for (uint32_t d = d0 + tid; d < d1; d += SPL_THREADS) {
                uint32_t x = (uint32_t)(w.doc_off[d] - w.off_base - tile0);
                w.out_off[d] = sm.wpre[x >> 5] + __popc(sm.tb[x >> 5] & ((1u << (x & 31)) - 1u));
            }
        }
        __syncthreads();
    }
}

// exclusive prefix of the per-tile id counts (one block; tiles are few: N / 4096)
__global__ void __launch_bounds__(1024) k_tile_scan(SplWork w) {
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < w.n_tiles; b0 += 1024) {
        uint32_t i = b0 + tid;
        uint64_t v = i < w.n_tiles ? (uint64_t)w.tile_cnt[i] : 0ull, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t lo = __shfl_up_sync(FULL, (uint32_t)incl, o), hi = __shfl_up_sync(FULL, (uint32_t)(incl >> 32), o);
            if (lane >= (uint32_t)o) incl += (uint64_t)lo | ((uint64_t)hi << 32);
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t x = s_w[lane], xi = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t lo = __shfl_up_sync(FULL, (uint32_t)xi, o), hi = __shfl_up_sync(FULL, (uint32_t)(xi >> 32), o);
                if (lane >= (uint32_t)o) xi += (uint64_t)lo | ((uint64_t)hi << 32);
            }
            s_w[lane] = xi - x;
        }
        __syncthreads();
        uint64_t carry = s_carry;
        if (i < w.n_tiles) w.tile_state[i] = carry + s_w[warp] + incl - v;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_w[warp] + incl;
        __syncthreads();
    }
}

// staging chunk of every tile -> its place in document order; tile-relative document offsets -> global
__global__ void __launch_bounds__(256) k_gather(SplWork w) {
    for (uint32_t tile = blockIdx.x; tile < w.n_tiles; tile += gridDim.x) {
        const uint64_t prefix = w.tile_state[tile], soff = w.tile_soff[tile];
        const uint32_t cnt = w.tile_cnt[tile];
        for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) w.ids[prefix + i] = w.stage[soff + i];
        uint32_t d0 = __ldg(w.tile_first_doc + tile), d1 = __ldg(w.tile_first_doc + tile + 1);
        for (uint32_t d = d0 + threadIdx.x; d < d1; d += blockDim.x) w.out_off[d] += prefix;
    }
}

// ------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------
void spl_kernels_init() {
    cudaFuncSetAttribute(k_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmem));
    cudaFuncSetAttribute(k_encode, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
}

int spl_launch_encode(const SplWork& w, int num_sms, cudaStream_t stream, SplKernelProfile* prof) {
    int launches = 0;
    if (prof) { prof->n = 0; cudaEventRecord(prof->ev[0], stream); }
    auto mark = [&](const char* name) {
        ++launches;
        if (prof && prof->n < SPL_PROF_MAX) { prof->name[prof->n] = name; ++prof->n; cudaEventRecord(prof->ev[prof->n], stream); }
    };
    {
        uint32_t n = w.n_docs + 1;
        k_mark_docs<<<(n + 255) / 256, 256, 0, stream>>>(w);
        mark("k_mark_docs");
    }
    if (w.with_special && w.N) {
        uint32_t blocks = (w.N + 255) / 256;
        uint32_t cap = (uint32_t)num_sms * 16;
        k_mark_specials<<<blocks < cap ? blocks : cap, 256, 0, stream>>>(w);
        mark("k_mark_specials");
    }
    if (w.N && w.pattern == SPL_PAT_MISTRAL_V3) {
        k_pretok<<<(w.N + SPL_TILE - 1) / SPL_TILE, SPL_THREADS, 0, stream>>>(w);
        mark("k_pretok");
    } else if (w.N) {
        k_pretok_fast<<<w.n_fast_tiles, SPL_FAST_THREADS, 0, stream>>>(w);
        mark("k_pretok_fast");
        uint32_t cap = (uint32_t)num_sms * 4;
        k_pretok_fb<<<w.n_fast_tiles < cap ? w.n_fast_tiles : cap, SPL_THREADS, 0, stream>>>(w);
        mark("k_pretok_fb");
    }
    {
        uint32_t cap = (uint32_t)num_sms * 4;
        uint32_t blocks = w.n_tiles < cap ? w.n_tiles : cap;
        k_encode<<<blocks, SPL_THREADS, sizeof(EncSmem), stream>>>(w);
        mark("k_encode");
        k_tile_scan<<<1, 1024, 0, stream>>>(w);
        mark("k_tile_scan");
        uint32_t gcap = (uint32_t)num_sms * 16;
        k_gather<<<w.n_tiles < gcap ? w.n_tiles : gcap, 256, 0, stream>>>(w);
        mark("k_gather");
    }
    return launches;
}
