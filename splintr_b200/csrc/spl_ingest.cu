// JSON Lines ingestion on the device: file bytes -> packed UTF-8 text + document offsets,
// the input format of spl_encode_batch_device.  Contract and per-line parser: spl_ingest.h.
//
//   k_jl_count   newlines per 4 KiB tile                       k_jl_scan    exclusive prefix over tiles (one block)
//   k_jl_lines   start offset of every line (rank of its newline = tile prefix + rank inside the tile)
//   k_jl_parse   one thread per line: locate the member, length of its unescaped value (spl_jl_parse_line)
//   k_jl_scan2   exclusive prefixes over lines: document index, text offset; totals and error counts
//   k_jl_emit    one thread per line: unescape the value to its place, write the document's offset
// A first, parity-oriented version: one thread walks one line byte by byte (lines are the unit of parallelism), so
// loads and stores are not coalesced across a warp; it is bounded by L1/L2 sector traffic, not by HBM.
#include "spl_device.cuh"
#include "spl_ingest.h"

namespace {

// Text accessor of the per-line walks: a 16-byte window in registers.  A thread reads its line front to back with
// short look-aheads, so one 16-byte load serves ~16 byte() calls; without it every byte is its own sector request and
// the walk is a chain of ~1 000 dependent L2 round trips per line.
struct GText {
    const uint8_t* p;
    mutable uint32_t blk;
    mutable uint64_t lo, hi;
    __device__ __forceinline__ explicit GText(const uint8_t* p_) : p(p_), blk(0xFFFFFFFFu), lo(0), hi(0) {}
    __device__ __forceinline__ uint8_t byte(uint32_t i) const {
        const uint32_t b = i >> 4;
        if (b != blk) {
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p) + b);     // p is 16-byte aligned and padded
            lo = v.x; hi = v.y; blk = b;
        }
        const uint64_t w = (i & 8u) ? hi : lo;
        return (uint8_t)(w >> ((i & 7u) * 8u));
    }
    // the eight bytes from i on (two aligned loads; the walk is sequential, so the second one is the next call's first)
    __device__ __forceinline__ uint64_t word8(uint32_t i) const {
        const uint64_t* q = reinterpret_cast<const uint64_t*>(p + (i & ~7u));
        const uint64_t a = __ldg(q), b = __ldg(q + 1);
        const uint32_t sh = (i & 7u) * 8u;
        return sh ? (a >> sh) | (b << (64u - sh)) : a;
    }
};

#define JL_THREADS 256

__global__ void __launch_bounds__(JL_THREADS) k_jl_count(SplJlWork w) {
    __shared__ uint32_t wsum[JL_THREADS / 32];
    const uint32_t tid = threadIdx.x, base = blockIdx.x * SPL_TILE + tid * 16u;
    uint32_t c = 0;
    if (base < w.N) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(w.text + base));       // readable up to N rounded up to 16
        const uint32_t v[4] = {x.x, x.y, x.z, x.w};
        const uint32_t n = w.N - base < 16u ? w.N - base : 16u;
        for (uint32_t k = 0; k < n; ++k) c += ((v[k >> 2] >> ((k & 3u) * 8u)) & 0xFFu) == '\n';
    }
    c = __reduce_add_sync(FULL, c);
    if ((tid & 31u) == 0) wsum[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) { uint32_t t = 0; for (uint32_t q = 0; q < JL_THREADS / 32; ++q) t += wsum[q]; w.tile_cnt[blockIdx.x] = t; }
}

// exclusive prefix of in[0, n) -> out[0, n], out[n] = total; one block of 1024 threads
__device__ void block_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n) {
    __shared__ uint32_t sw[32];
    __shared__ uint32_t carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < n; t0 += 1024) {
        const uint32_t t = t0 + tid;
        const uint32_t v = t < n ? in[t] : 0u;
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
        if (t < n) out[t] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry;
}

__global__ void __launch_bounds__(1024) k_jl_scan(SplJlWork w) {
    block_scan_u32(w.tile_cnt, w.tile_pref, w.n_tiles);
    if (threadIdx.x == 0) w.counters[SPL_JLCTR_NEWLINES] = w.tile_pref[w.n_tiles];
}

__global__ void __launch_bounds__(JL_THREADS) k_jl_lines(SplJlWork w) {
    __shared__ uint32_t wsum[JL_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, base = blockIdx.x * SPL_TILE + tid * 16u;
    uint32_t mask = 0;
    if (base < w.N) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(w.text + base));
        const uint32_t v[4] = {x.x, x.y, x.z, x.w};
        const uint32_t n = w.N - base < 16u ? w.N - base : 16u;
        for (uint32_t k = 0; k < n; ++k) if (((v[k >> 2] >> ((k & 3u) * 8u)) & 0xFFu) == '\n') mask |= 1u << k;
    }
    const uint32_t c = __popc(mask);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += u; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t rank = w.tile_pref[blockIdx.x] + incl - c;
    for (uint32_t q = 0; q < warp; ++q) rank += wsum[q];
    while (mask) {
        const uint32_t k = __ffs(mask) - 1;
        mask &= mask - 1;
        w.line_start[++rank] = base + k + 1;                     // line `rank` starts right behind newline number rank - 1
    }
    if (blockIdx.x == 0 && tid == 0) { w.line_start[0] = 0; w.line_start[w.n_lines] = w.N + 1; }
}

__global__ void __launch_bounds__(JL_THREADS) k_jl_parse(SplJlWork w) {
    const uint32_t k = blockIdx.x * JL_THREADS + threadIdx.x;
    if (k >= w.n_lines) return;
    const GText t(w.text);
    const SplJlSpan sp = spl_jl_parse_line(t, w.line_start[k], w.line_start[k + 1] - 1, w.field, w.flen);
    w.span[k] = sp;
    w.is_doc[k] = sp.flags & SPL_JL_DOC;
    w.out_len[k] = sp.out_len;
    if ((sp.flags & SPL_JL_DOC) && !(sp.flags & SPL_JL_FOUND))
        atomicAdd(&w.counters[(sp.flags & SPL_JL_BAD) ? SPL_JLCTR_BAD : SPL_JLCTR_MISSING], 1u);
}

__global__ void __launch_bounds__(1024) k_jl_scan2(SplJlWork w) {
    block_scan_u32(w.is_doc, w.doc_idx, w.n_lines);
    __syncthreads();
    block_scan_u32(w.out_len, w.text_off, w.n_lines);
    if (threadIdx.x == 0) {
        w.counters[SPL_JLCTR_DOCS] = w.doc_idx[w.n_lines];
        w.counters[SPL_JLCTR_TEXT] = w.text_off[w.n_lines];
    }
}

__global__ void __launch_bounds__(JL_THREADS) k_jl_emit(SplJlWork w) {
    const uint32_t k = blockIdx.x * JL_THREADS + threadIdx.x;
    if (k > w.n_lines) return;
    if (k == w.n_lines) {                                        // closing entry of the offsets
        if (w.doc_idx[k] < w.off_capacity) w.out_off[w.doc_idx[k]] = w.text_off[k];
        return;
    }
    const SplJlSpan sp = w.span[k];
    if (!(sp.flags & SPL_JL_DOC)) return;
    const uint32_t d = w.doc_idx[k];
    if (d >= w.off_capacity) return;                             // the host reports the needed capacity
    uint32_t o = w.text_off[k];
    w.out_off[d] = o;
    if (!(sp.flags & SPL_JL_FOUND) || (uint64_t)w.text_off[w.n_lines] > w.text_capacity) return;
    const GText t(w.text);
    uint32_t j = sp.vs;
    while (j < sp.ve) {
        if (j + 16u <= sp.ve) {
            const uint64_t x = t.word8(j);
            if (!spl_jl_has(x, '\\')) {                          // eight bytes without an escape: copied as they are,
                const uint32_t mis = o & 7u;                     // with one 8-byte store once the destination is aligned
                if (mis == 0) {
                    *reinterpret_cast<uint64_t*>(w.out_text + o) = x;
                    o += 8; j += 8;
                } else {
                    const uint32_t k = 8u - mis;
                    for (uint32_t q = 0; q < k; ++q) w.out_text[o + q] = (uint8_t)(x >> (8u * q));
                    o += k; j += k;
                }
                continue;
            }
        }
        uint8_t ch[4];
        uint32_t n;
        j = spl_jl_char(t, j, sp.ve, ch, n);
        for (uint32_t q = 0; q < n; ++q) w.out_text[o++] = ch[q];
    }
}

}  // namespace

int spl_launch_jsonl_count(const SplJlWork& w, cudaStream_t stream) {
    k_jl_count<<<w.n_tiles, JL_THREADS, 0, stream>>>(w);
    k_jl_scan<<<1, 1024, 0, stream>>>(w);
    return 2;
}

int spl_launch_jsonl_extract(const SplJlWork& w, cudaStream_t stream) {
    k_jl_lines<<<w.n_tiles, JL_THREADS, 0, stream>>>(w);
    k_jl_parse<<<(w.n_lines + JL_THREADS - 1) / JL_THREADS, JL_THREADS, 0, stream>>>(w);
    k_jl_scan2<<<1, 1024, 0, stream>>>(w);
    k_jl_emit<<<(w.n_lines + JL_THREADS) / JL_THREADS, JL_THREADS, 0, stream>>>(w);
    return 4;
}
