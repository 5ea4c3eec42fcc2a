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
    uint32_t mb[SPL_TILE / 32];       // pieces that missed the whole-piece probe
    uint32_t tok[SPL_WIN];            // token id (or symbol during merging) at its first byte
    uint32_t rnk[SPL_WIN];            // rank of the pair (part at i, next part)
    uint32_t wpre[EN_WORDS + 1];      // exclusive token count before each bitmap word
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

#define ST_AGG  (1ull << 62)
#define ST_INCL (2ull << 62)
#define ST_MASK ((1ull << 62) - 1)

__global__ void __launch_bounds__(SPL_THREADS) k_encode(SplWork w) {
    __shared__ __align__(16) EncSmem sm;
    const SplTables* T = w.T;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t N = w.N, Nup = (N + 15u) & ~15u;

    for (;;) {
        if (tid == 0) sm.tile = atomicAdd(&w.counters[0], 1u);
        __syncthreads();
        const uint32_t tile = sm.tile;
        if (tile >= w.n_tiles) return;
        const uint32_t tile0 = tile * SPL_TILE;

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
        if (tid < SPL_TILE / 32) sm.mb[tid] = 0;
        if (tid == 0) { sm.huge_start = SPL_RANK_NONE; sm.huge_cnt = 0; }
        __syncthreads();

        // ---- fast path: one thread per piece start in its 16 bytes ------------------------
        {
            uint32_t my = (sm.pb[tid >> 1] >> ((tid & 1u) * 16u)) & 0xFFFFu;
            const uint32_t avail = N - tile0;            // text bytes from tile0 on (>= 1 piece start only below this)
            while (my) {
                uint32_t b = __ffs(my) - 1;
                my &= my - 1;
                uint32_t s = tid * 16 + b;
                if (s >= avail) break;                   // sentinel bit at N
                uint32_t e = sm_next_bit(sm.pb, s + 1, SPL_WIN + 1);
                if (e > SPL_WIN) { sm.huge_start = s; break; }    // the tile's last piece leaves the window
                uint32_t len = e - s;
                uint32_t id = SPL_RANK_NONE;
                if (w.with_special && ((__ldg(w.spec + ((tile0 + s) >> 5)) >> ((tile0 + s) & 31)) & 1u)) {
                    id = special_id(T, [&](uint32_t j) { return sm_byte(sm.text, s + j); }, len);
                    if (id == SPL_RANK_NONE) { atomicAnd(&sm.tb[s >> 5], ~(1u << (s & 31))); continue; }
                } else if (len <= 8) {
                    uint64_t k0 = sm_load8(sm.text, s);
                    if (len < 8) k0 &= (1ull << (8 * len)) - 1;
                    id = lookup8(T->t8, T->t8_log2, k0, len);
                } else if (len <= 16) {
                    uint64_t k0 = sm_load8(sm.text, s), k1 = sm_load8(sm.text, s + 8);
                    if (len < 16) k1 &= (1ull << (8 * (len - 8))) - 1;
                    id = lookup16(T->t16, T->t16_log2, k0, k1, len);
                }
                if (id != SPL_RANK_NONE) sm.tok[s] = id;
                else atomicOr(&sm.mb[s >> 5], 1u << (s & 31));
            }
        }
        __syncthreads();

        // ---- slow path: one warp per missed piece --------------------------------------------
        for (uint32_t wi = warp; wi < SPL_TILE / 32; wi += EN_WARPS) {
            uint32_t bits = sm.mb[wi];
            while (bits) {
                uint32_t b = __ffs(bits) - 1;
                bits &= bits - 1;
                uint32_t s = wi * 32 + b;
                uint32_t e = sm_next_bit(sm.pb, s + 1, SPL_WIN + 1);
                bpe_piece_warp(sm, T, s, e);
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

        // ---- count ids, word prefixes, tile prefix ------------------------------------------------
        if (warp == 0) {
            // exclusive scan of the per-word popcounts (EN_WORDS <= 160: five words per lane)
            uint32_t local[5], run = 0;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                uint32_t v = lane * 5 + q;
                local[q] = v < EN_WORDS ? __popc(sm.tb[v]) : 0u;
                run += local[q];
            }
            uint32_t incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(FULL, incl, o);
                if (lane >= (uint32_t)o) incl += t;
            }
            uint32_t base = incl - run;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                uint32_t v = lane * 5 + q;
                if (v <= EN_WORDS) sm.wpre[v] = base;
                base += local[q];
            }
            // ---- decoupled look-back over tiles ------------------------------------------------
            uint32_t win_cnt = __shfl_sync(FULL, incl, 31);
            uint64_t total = (uint64_t)win_cnt + huge_cnt;
            volatile uint64_t* st = w.tile_state;
            if (lane == 0) {
                __threadfence();
                st[tile] = (tile == 0 ? ST_INCL : ST_AGG) | total;
            }
            uint64_t prefix = 0;
            if (tile > 0) {
                int64_t look = (int64_t)tile - 1;
                for (;;) {
                    int64_t idx = look - lane;
                    uint64_t v = ST_INCL;                         // lanes before tile 0 contribute an inclusive 0
                    if (idx >= 0) { do { v = st[idx]; } while ((v >> 62) == 0); }
                    uint32_t incl_mask = __ballot_sync(FULL, (v >> 62) == 2);
                    uint32_t first = incl_mask ? (uint32_t)__ffs(incl_mask) - 1 : 32u;
                    uint64_t contrib = lane <= first ? (v & ST_MASK) : 0ull;
                    prefix += warp_sum_u64(contrib);
                    if (incl_mask) break;
                    look -= 32;
                }
                if (lane == 0) { __threadfence(); st[tile] = ST_INCL | (prefix + total); }
            }
            if (lane == 0) sm.prefix = prefix;
        }
        __syncthreads();
        const uint64_t prefix = sm.prefix;

        // ---- ordered id output -----------------------------------------------------------------
        for (uint32_t hw = tid; hw < EN_WORDS * 2; hw += SPL_THREADS) {
            uint32_t word = sm.tb[hw >> 1], sh = (hw & 1u) * 16u;
            uint32_t my = (word >> sh) & 0xFFFFu;
            uint64_t o = prefix + sm.wpre[hw >> 1] + __popc(word & ((1u << sh) - 1u));
            while (my) {
                uint32_t b = __ffs(my) - 1;
                my &= my - 1;
                w.ids[o++] = sm.tok[hw * 16 + b];
            }
        }
        uint32_t win_total = sm.wpre[EN_WORDS];
        if (huge_cnt) {
            // block-ordered compaction of the survivors in huge_scratch[0 .. huge_len)
            uint64_t base = prefix + win_total;
            uint32_t per = (huge_len + SPL_THREADS - 1) / SPL_THREADS;
            uint32_t lo = tid * per, hi = lo + per < huge_len ? lo + per : huge_len;
            uint32_t c = 0;
            for (uint32_t j = lo; j < hi; ++j) c += (huge_scratch[j] != SPL_RANK_NONE);
            sm.red[tid] = c;
            __syncthreads();
            if (tid == 0) { uint64_t run = 0; for (int q = 0; q < SPL_THREADS; ++q) { uint64_t t = sm.red[q]; sm.red[q] = run; run += t; } }
            __syncthreads();
            uint64_t o = base + sm.red[tid];
            for (uint32_t j = lo; j < hi; ++j) { uint32_t v = huge_scratch[j]; if (v != SPL_RANK_NONE) w.ids[o++] = v; }
        }

        // ---- per-document output offsets -----------------------------------------------------------
        {
            uint32_t d0 = __ldg(w.tile_first_doc + tile), d1 = __ldg(w.tile_first_doc + tile + 1);
            for (uint32_t d = d0 + tid; d < d1; d += SPL_THREADS) {
                uint32_t x = (uint32_t)(w.doc_off[d] - w.off_base - tile0);
                w.out_off[d] = prefix + sm.wpre[x >> 5] + __popc(sm.tb[x >> 5] & ((1u << (x & 31)) - 1u));
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------
void spl_kernels_init() {
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
        uint32_t cap = (uint32_t)num_sms * 5;
        uint32_t blocks = w.n_tiles < cap ? w.n_tiles : cap;
        k_encode<<<blocks, SPL_THREADS, 0, stream>>>(w);
        mark("k_encode");
    }
    return launches;
}
