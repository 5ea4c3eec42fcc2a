// SentencePiece mode on the device (mistral / mistral_v2; SURVEY 8f row N3, second half): the sequential
// `pending_underscores` walk of tokenizer.rs:737-795 as five data-parallel kernels that turn the packed text T into
// the transformed text T' (every converted space = E2 96 81), its piece-start bitmap, its special-span bitmap and
// the document offsets inside T'.  The ordinary encode stage (k_probe / k_bpe / k_emit) then runs over T' unchanged.
// Rules and their derivation: spl_sentencepiece.h.
//
//   k_sp_classify   one thread per 32-byte word: W0 (byte of a \s character), A = copy of W0, RS (raw chunk starts)
//   k_sp_rawruns    one thread per RS word: clears the A bits of every raw chunk (rare: chunks led by 0B / U+00A0 ...)
//   k_sp_count      converted spaces per 4 KiB tile
//   k_sp_scan       exclusive prefix over the tiles (one block), total -> counters
//   k_sp_emit       one block per tile: bytes of T', piece-start / special bits at their new positions, document
//                   offsets in T'
// Byte / bit work bounded by HBM traffic and the scattered bit updates; no tensor cores.
#include "spl_device.cuh"
#include "spl_sentencepiece.h"

namespace {

struct GText {
    const uint8_t* p;
    __device__ __forceinline__ uint8_t byte(uint32_t i) const { return __ldg(p + i); }
};

__device__ __forceinline__ bool bit_at(const uint32_t* __restrict__ w, uint32_t i) { return (__ldg(w + (i >> 5)) >> (i & 31)) & 1u; }

__global__ void __launch_bounds__(256) k_sp_classify(SplSpWork s) {
    const uint32_t gw = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t base = gw * 32u;
    if (base >= s.N) return;
    const SplTables* T = s.T;
    const GText t{s.text};
    const uint32_t hw = __ldg(s.hard + gw), sw = s.spec ? __ldg(s.spec + gw) : 0u;
    const uint32_t n = s.N - base < 32u ? s.N - base : 32u;
    uint32_t x[8];
    {
        const uint32_t Nup = (s.N + 15u) & ~15u;
        const uint4* p4 = reinterpret_cast<const uint4*>(s.text + base);
        const uint4 a = __ldg(p4), b = (base + 16u < Nup) ? __ldg(p4 + 1) : make_uint4(0, 0, 0, 0);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    }
    const uint32_t pb = base ? t.byte(base - 1) : 0u;                          // the byte in front of the word
    const uint32_t any_hi = (x[0] | x[1] | x[2] | x[3] | x[4] | x[5] | x[6] | x[7] | pb) & 0x80808080u;
    uint32_t w0 = 0, rs = 0;
    if (!any_hi) {
        // ASCII only (the usual word): whitespace = 09..0D, 20; of those only 0B is not u8::is_ascii_whitespace
        bool prev_w0 = base && !(s.spec && bit_at(s.spec, base - 1)) && (pb == 0x20u || pb - 9u <= 4u);
        const uint8_t* xb = reinterpret_cast<const uint8_t*>(x);
#pragma unroll 8
        for (uint32_t k = 0; k < 32u; ++k) {
            const uint32_t b = xb[k];
            const bool ws = k < n && !((sw >> k) & 1u) && (b == 0x20u || b - 9u <= 4u);
            if (ws) {
                w0 |= 1u << k;
                if (b == 0x0Bu && (((hw >> k) & 1u) || base + k == 0 || !prev_w0)) rs |= 1u << k;
            }
            prev_w0 = ws;
        }
    } else {
        bool prev_w0 = false;
        if (base > 0 && !(s.spec && bit_at(s.spec, base - 1))) prev_w0 = spl_sp_ws_byte(t, base - 1, s.N, T->ucd_stage1, T->ucd_stage2);
        for (uint32_t k = 0; k < n; ++k) {
            const uint32_t i = base + k, b = t.byte(i);
            const bool sp = (sw >> k) & 1u;
            const bool ws = !sp && spl_sp_ws_byte(t, i, s.N, T->ucd_stage1, T->ucd_stage2);
            if (ws) {
                w0 |= 1u << k;
                const bool char_start = (b & 0xC0u) != 0x80u;
                if (char_start && (((hw >> k) & 1u) || i == 0 || !prev_w0) && !spl_sp_ascii_ws(b)) rs |= 1u << k;
            }
            prev_w0 = ws;
        }
    }
    s.w0[gw] = w0; s.a[gw] = w0; s.rs[gw] = rs;
}

__global__ void __launch_bounds__(256) k_sp_rawruns(SplSpWork s) {
    const uint32_t gw = blockIdx.x * blockDim.x + threadIdx.x;
    if (gw * 32u >= s.N) return;
    uint32_t r = s.rs[gw];
    while (r) {
        const uint32_t st = gw * 32u + __ffs(r) - 1;
        r &= r - 1;
        // the chunk: W0 bytes from st up to the next segment start (W0 itself is never modified)
        for (uint32_t i = st; i < s.N && bit_at(s.w0, i) && (i == st || !bit_at(s.hard, i)); ++i)
            atomicAnd(&s.a[i >> 5], ~(1u << (i & 31)));
    }
}

// bit k set iff byte k of the 32-byte word is 0x20
__device__ __forceinline__ uint32_t eq20_mask(const uint32_t (&x)[8]) {
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (((x[q] >> (8 * b)) & 0xFFu) == 0x20u) m |= 1u << (q * 4 + b);
    return m;
}

__device__ __forceinline__ void load_word32(const SplSpWork& s, uint32_t base, uint32_t (&x)[8]) {
    const uint32_t Nup = (s.N + 15u) & ~15u;
    const uint4* p4 = reinterpret_cast<const uint4*>(s.text + base);
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (base < Nup) a = __ldg(p4);
    if (base + 16u < Nup) b = __ldg(p4 + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}

#define SP_TW (SPL_TILE / 32u)              // words (= threads) per tile

__global__ void __launch_bounds__(SP_TW) k_sp_count(SplSpWork s) {
    __shared__ uint32_t wsum[SP_TW / 32];
    const uint32_t tid = threadIdx.x, gw = blockIdx.x * SP_TW + tid, base = gw * 32u;
    uint32_t c = 0;
    if (base < s.N) {
        uint32_t x[8];
        load_word32(s, base, x);
        uint32_t valid = s.N - base >= 32u ? FULL : ((1u << (s.N - base)) - 1u);
        c = __popc(s.a[gw] & eq20_mask(x) & valid);
    }
    c = __reduce_add_sync(FULL, c);
    if ((tid & 31u) == 0) wsum[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) { uint32_t tot = 0; for (uint32_t q = 0; q < SP_TW / 32; ++q) tot += wsum[q]; s.tile_cnt[blockIdx.x] = tot; }
}

__global__ void __launch_bounds__(1024) k_sp_scan(SplSpWork s) {
    __shared__ uint32_t sw[32];
    __shared__ uint32_t carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < s.n_tiles; t0 += 1024) {
        const uint32_t t = t0 + tid;
        const uint32_t v = t < s.n_tiles ? s.tile_cnt[t] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += u; }
        if (lane == 31) sw[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t x = sw[lane], xi = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(FULL, xi, o); if (lane >= (uint32_t)o) xi += u; }
            sw[lane] = xi - x;
        }
        __syncthreads();
        const uint32_t excl = carry + sw[warp] + incl - v;
        if (t < s.n_tiles) s.tile_pref[t] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) { s.tile_pref[s.n_tiles] = carry; s.counters[SPL_SPCTR_CONV] = carry; }
}

// bytes and bits of a tile's image in T' are staged in shared memory and leave as coalesced stores: the image of tile t
// is the contiguous range [tile0 + 2 * before, ... + SPL_TILE + 2 * conv(t)) and no other tile writes into it
#define SP_OUT_MAX (3u * SPL_TILE)
#define SP_BM_WORDS (SP_OUT_MAX / 32u + 2u)
struct SpEmitSmem {
    uint8_t out[SP_OUT_MAX];
    uint32_t ps[SP_BM_WORDS], sp[SP_BM_WORDS];       // bit r <-> output position (obase_tile & ~31) + r
    uint32_t excl[SP_TW], conv[SP_TW], wsum[SP_TW / 32];
};

__global__ void __launch_bounds__(SP_TW) k_sp_emit(SplSpWork s) {
    __shared__ SpEmitSmem sm;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile = blockIdx.x, tile0 = tile * SPL_TILE, gw = tile * SP_TW + tid, base = gw * 32u;
    uint32_t x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t w0 = 0, a = 0, rs = 0, hw = 0, sw = 0, conv = 0, valid = 0;
    if (base < s.N) {
        load_word32(s, base, x);
        valid = s.N - base >= 32u ? FULL : ((1u << (s.N - base)) - 1u);
        w0 = s.w0[gw]; a = s.a[gw]; rs = s.rs[gw]; hw = __ldg(s.hard + gw);
        if (s.spec) sw = __ldg(s.spec + gw);
        conv = a & eq20_mask(x) & valid;
    }
    for (uint32_t v = tid; v < SP_BM_WORDS; v += SP_TW) { sm.ps[v] = 0; sm.sp[v] = 0; }
    // exclusive count of converted spaces in front of my word, inside the tile
    const uint32_t c = __popc(conv);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += u; }
    if (lane == 31) sm.wsum[warp] = incl;
    __syncthreads();
    uint32_t excl = incl - c, tile_conv = 0;
    for (uint32_t q = 0; q < SP_TW / 32; ++q) { const uint32_t t = sm.wsum[q]; excl += q < warp ? t : 0u; tile_conv += t; }
    sm.excl[tid] = excl; sm.conv[tid] = conv;
    const uint32_t before = s.tile_pref[tile];
    const uint32_t obase_tile = tile0 + 2u * before;                    // first output position of the tile
    const uint32_t bit0 = obase_tile & ~31u;
    const uint32_t tile_in = s.N > tile0 ? (s.N - tile0 < SPL_TILE ? s.N - tile0 : SPL_TILE) : 0u;
    const uint32_t tile_out = tile_in + 2u * tile_conv;
    if (base < s.N) {
        bool w0_prev = false, a_prev = false;
        uint32_t b_prev = 0;
        if (base > 0) {
            w0_prev = (s.w0[gw - 1] >> 31) & 1u; a_prev = (s.a[gw - 1] >> 31) & 1u;
            b_prev = __ldg(s.text + base - 1);
        }
        uint32_t lo = tid * 32u + 2u * excl;                            // tile-relative output position of my first byte
        const uint8_t* xb = reinterpret_cast<const uint8_t*>(x);
        const uint32_t n = s.N - base < 32u ? s.N - base : 32u;
        for (uint32_t k = 0; k < n; ++k) {
            SplSpPos p;
            p.w0 = (w0 >> k) & 1u; p.a = (a >> k) & 1u; p.rs = (rs >> k) & 1u; p.s = (hw >> k) & 1u; p.b = xb[k];
            p.w0_prev = w0_prev; p.a_prev = a_prev; p.b_prev = b_prev;
            const uint32_t r = obase_tile + lo - bit0;                  // bit index in the staged bitmaps
            if (spl_sp_piece_start(p)) atomicOr(&sm.ps[r >> 5], 1u << (r & 31));
            if ((sw >> k) & 1u) atomicOr(&sm.sp[r >> 5], 1u << (r & 31));
            if ((conv >> k) & 1u) { sm.out[lo] = 0xE2u; sm.out[lo + 1] = 0x96u; sm.out[lo + 2] = 0x81u; lo += 3; }
            else sm.out[lo++] = (uint8_t)p.b;
            w0_prev = p.w0; a_prev = p.a; b_prev = p.b;
        }
    }
    __syncthreads();
    // bytes: consecutive threads store consecutive bytes
    uint8_t* __restrict__ dst = s.text2 + obase_tile;
    for (uint32_t i = tid; i < tile_out; i += SP_TW) dst[i] = sm.out[i];
    // bits: the first and the last word of the range can be shared with the neighbouring tiles
    const uint32_t nbw = tile_out ? ((obase_tile + tile_out - 1u) >> 5) - (bit0 >> 5) + 1u : 0u;
    for (uint32_t v = tid; v < nbw; v += SP_TW) {
        const uint32_t pv = sm.ps[v], sv = sm.sp[v];
        if (pv) atomicOr(&s.pstart2[(bit0 >> 5) + v], pv);
        if (sv) atomicOr(&s.spec2[(bit0 >> 5) + v], sv);
    }
    // document starts of this tile -> positions in T'
    const uint32_t d0 = s.tinfo[tile].first_doc, d1 = s.tinfo[tile + 1].first_doc;
    for (uint32_t d = d0 + tid; d < d1 && d <= s.n_docs; d += SP_TW) {
        const uint32_t xo = (uint32_t)(s.doc_off[d] - s.off_base) - tile0;
        const uint32_t wq = xo >> 5;
        const uint32_t cnt = before + sm.excl[wq] + __popc(sm.conv[wq] & ((1u << (xo & 31)) - 1u));
        s.doc_off2[d] = (uint64_t)tile0 + xo + 2ull * cnt;
    }
}

}  // namespace

int spl_launch_sp_scan(const SplSpWork& s, cudaStream_t stream) {
    const uint32_t words = (s.N + 31u) / 32u;
    if (words) {
        k_sp_classify<<<(words + 255) / 256, 256, 0, stream>>>(s);
        k_sp_rawruns<<<(words + 255) / 256, 256, 0, stream>>>(s);
    }
    k_sp_count<<<s.n_tiles, SP_TW, 0, stream>>>(s);
    k_sp_scan<<<1, 1024, 0, stream>>>(s);
    return words ? 4 : 2;
}

int spl_launch_sp_emit(const SplSpWork& s, cudaStream_t stream) {
    k_sp_emit<<<s.n_tiles, SP_TW, 0, stream>>>(s);
    return 1;
}
