// Device encode path, first half: packed UTF-8 bytes + document offsets  ->  piece-start bitmap.
//
//   k_mark_docs      document starts -> `hard` bitmap, per-tile first-document index
//   k_mark_specials  (encode_with_special) special-token spans -> `hard` edges + `spec` bytes
//                    replaces the Aho-Corasick scan of tokenizer.rs:842-874
//   k_pretok_fast    piece-start bitmap, bit-parallel (spl_pretok_fast.h); k_pretok / k_pretok_fb: the split
//                    regex as sequential class rules (spl_pretok.h)
//                    replaces regex find_iter, tokenizer.rs:244-257 / :731
//
// The second half (whole-piece probe, merge loop, ordered id output) is spl_encode.cu.
// Integer / byte work, bounded by instruction issue and HBM traffic; no tensor cores.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "spl_device.cuh"
#include "spl_pretok.h"
#include "spl_pretok_fast.h"
#include "spl_fast_dev.cuh"
#include "spl_special.h"

// ------------------------------------------------------------------------------------------
// k_mark_docs
// ------------------------------------------------------------------------------------------
__global__ void k_mark_docs(SplWork w) {
    SPL_PDL_ENTER();
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d > w.n_docs) return;
    uint64_t s = w.doc_off[d] - w.off_base;
    uint64_t prev = d ? w.doc_off[d - 1] - w.off_base : 0;
    bool bad = w.doc_off[d] < w.off_base || s > w.N || s < prev || (d == 0 && s != 0) || (d == w.n_docs && s != w.N);
    if (bad) { atomicOr(&w.counters[SPL_CTR_ERR], SPL_DEVERR_OFFSETS); return; }
    uint32_t p = (uint32_t)s;
    atomicOr(&w.hard[p >> 5], 1u << (p & 31));
    uint32_t t_lo = d ? (uint32_t)(prev / SPL_TILE) + 1 : 0, t_hi = p / SPL_TILE;
    for (uint32_t t = t_lo; t <= t_hi; ++t) w.tinfo[t].first_doc = d;
    if (d == w.n_docs) {
        w.tinfo[w.n_tiles].first_doc = w.n_docs + 1;
        atomicOr(&w.pstart[p >> 5], 1u << (p & 31));            // sentinel piece start at N
    }
}

// ------------------------------------------------------------------------------------------
// k_mark_specials: every occurrence of a special string that lies inside one document.
// Sets where no string contains or overlaps another (all bundled ones): every occurrence is a match, marked right here.
// Other sets (w.cand != nullptr): occurrences compete (spl_special.h); this pass only records where one STARTS, and
// k_resolve_specials walks every document from candidate to candidate.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mark_special_span(const SplWork& w, uint32_t i, uint32_t e) {
    atomicOr(&w.hard[i >> 5], 1u << (i & 31));
    atomicOr(&w.hard[e >> 5], 1u << (e & 31));
    for (uint32_t j = i; j < e; ++j) atomicOr(&w.spec[j >> 5], 1u << (j & 31));
}

__global__ void k_mark_specials(SplWork w) {
    SPL_RETURN_IF_BAD_OFFSETS(w);
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
            if (w.cand) { atomicOr(&w.cand[i >> 5], 1u << (i & 31)); break; }
            mark_special_span(w, i, i + len);
            break;
        }
    }
}

struct GlobalByteText { const uint8_t* __restrict__ p; __device__ __forceinline__ uint8_t byte(uint32_t i) const { return __ldg(p + i); } };
struct GlobalCand {
    const uint32_t* __restrict__ bits;
    __device__ __forceinline__ uint32_t next(uint32_t from, uint32_t lim) const { return g_next_bit(bits, from, lim); }
};

// one thread per document: the matches aho-corasick's Standard non-overlapping find_iter reports (tokenizer.rs:851)
__global__ void k_resolve_specials(SplWork w) {
    SPL_RETURN_IF_BAD_OFFSETS(w);
    const SplTables* T = w.T;
    const SplSpecialSet S{T->sp_bytes, T->sp_off, T->n_special};
    const GlobalByteText t{w.text};
    const GlobalCand c{w.cand};
    for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < w.n_docs; d += gridDim.x * blockDim.x) {
        const uint32_t d0 = (uint32_t)(w.doc_off[d] - w.off_base), d1 = (uint32_t)(w.doc_off[d + 1] - w.off_base);
        if (g_next_bit(w.cand, d0, d1) >= d1) continue;
        spl_special_walk(t, c, S, d0, d1, [&](uint32_t s, uint32_t e, uint32_t) { mark_special_span(w, s, e); });
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
    SPL_RETURN_IF_BAD_OFFSETS(w);
    pretok_tile(w, sm, blockIdx.x * SPL_TILE, false);
}

// ------------------------------------------------------------------------------------------
// k_pretok_fast: the bit-parallel formulation (spl_pretok_fast.h).  One thread per 32-byte word,
// FAST_HALO words of context on each side of FAST_PAYLOAD payload words.  A tile that cannot be
// decided here is appended to the fallback list and left untouched.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPL_FAST_THREADS) k_pretok_fast(SplWork w) {
    __shared__ FastSmem sm;
    SPL_RETURN_IF_BAD_OFFSETS(w);
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
            uint32_t idx = atomicAdd(&w.counters[SPL_CTR_FB], 1u);
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
    SPL_RETURN_IF_BAD_OFFSETS(w);
    const uint32_t n = w.counters[SPL_CTR_FB];
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
// host launcher
// ------------------------------------------------------------------------------------------
void spl_kernels_init() {
    spl_encode_init();
}

namespace {
void launch_desc(const SplLaunchDesc& d, const SplWork& w, cudaStream_t stream) {
    SplWork wc = w;
    uint32_t u = d.u32;
    void* args[2] = {&wc, &u};
    cudaLaunchKernel(d.func, dim3(d.grid), dim3(d.block), args, d.smem, stream);
}
void sync_each(cudaStream_t stream, const char* name) {
    // SPL_SYNC_EACH=1 (debugging): wait for every kernel and name the one that faults
    static const bool on = [] { const char* e = getenv("SPL_SYNC_EACH"); return e && e[0] == '1'; }();
    if (!on) return;
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) fprintf(stderr, "[spl] kernel %s: %s\n", name, cudaGetErrorString(e));
}
}  // namespace

static int describe_mark(const SplWork& w, int num_sms, SplLaunchDesc* out) {
    int n = 0;
    out[n++] = SplLaunchDesc{(const void*)k_mark_docs, "k_mark_docs", (w.n_docs + 1 + 255) / 256, 256, 0, false, 0};
    if (w.with_special && w.N && !w.pretok_done) {
        const uint32_t blocks = (w.N + 255) / 256, cap = (uint32_t)num_sms * 16;
        out[n++] = SplLaunchDesc{(const void*)k_mark_specials, "k_mark_specials", blocks < cap ? blocks : cap, 256, 0, false, 0};
        if (w.cand)
            out[n++] = SplLaunchDesc{(const void*)k_resolve_specials, "k_resolve_specials", std::min<uint32_t>((w.n_docs + 127u) / 128u, cap), 128, 0, false, 0};
    }
    return n;
}

int spl_launch_mark(const SplWork& w, int num_sms, cudaStream_t stream) {
    SplLaunchDesc d[4];
    SplWork v = w;
    v.pretok_done = false;
    const int n = describe_mark(v, num_sms, d);
    for (int i = 0; i < n; ++i) { launch_desc(d[i], w, stream); sync_each(stream, d[i].name); }
    return n;
}

int spl_describe_encode(const SplWork& w, int num_sms, SplLaunchDesc* out) {
    int n = describe_mark(w, num_sms, out);
    if (w.pretok_done) {
        // SentencePiece mode: k_sp_emit has written the piece starts of the transformed text
    } else if (w.N && w.pattern == SPL_PAT_MISTRAL_V3) {
        out[n++] = SplLaunchDesc{(const void*)k_pretok, "k_pretok", (w.N + SPL_TILE - 1) / SPL_TILE, SPL_THREADS, 0, false, 0};
    } else if (w.N) {
        const uint32_t cap = (uint32_t)num_sms * 4;
        out[n++] = SplLaunchDesc{(const void*)k_pretok_fast, "k_pretok_fast", w.n_fast_tiles, SPL_FAST_THREADS, 0, false, 0};
        out[n++] = SplLaunchDesc{(const void*)k_pretok_fb, "k_pretok_fb", w.n_fast_tiles < cap ? w.n_fast_tiles : cap, SPL_THREADS, 0, false, 0};
    }
    n += spl_describe_encode_stage(w, num_sms, out + n);
    return n;
}

int spl_launch_encode(const SplWork& w, int num_sms, cudaStream_t stream, SplKernelProfile* prof) {
    SplLaunchDesc d[SPL_MAX_LAUNCHES];
    const int n = spl_describe_encode(w, num_sms, d);
    if (prof) { prof->n = 0; cudaEventRecord(prof->ev[0], stream); }
    for (int i = 0; i < n; ++i) {
        launch_desc(d[i], w, stream);
        if (prof && prof->n < SPL_PROF_MAX) { prof->name[prof->n] = d[i].name; ++prof->n; cudaEventRecord(prof->ev[prof->n], stream); }
        sync_each(stream, d[i].name);
    }
    return n;
}
