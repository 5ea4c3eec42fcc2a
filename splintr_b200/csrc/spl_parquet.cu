// Parquet ingestion on the device: the pages of one string column -> packed UTF-8 text +
// document offsets, the input format of spl_encode_batch_device.  Formats and per-page decoder: spl_parquet.h; footer
// and page headers are read on the host (spl_parquet_meta.cpp).
//
//   k_pq_pages    one warp per page: snappy (all lanes parse the element stream alike, all lanes copy; input slots and
//                 a 64 KiB ring of output in shared memory, so an element touches no global memory), definition
//                 levels, PLAIN lengths or dictionary indices -> one (source offset, length) span per row;
//                 launched twice: dictionary pages and PLAIN data pages, then the dictionary-encoded data pages (which
//                 read the dictionary's spans)
//   k_pq_bsum     row lengths summed per block of 2 048 rows        k_pq_bscan   exclusive prefix over the blocks
//   k_pq_offsets  u64 document offsets (row r: bytes of the rows in front of it), total -> counters
//   k_pq_copy     one block per 16 KiB of output text: finds its first row by bisection, a warp per row copies the part of
//                 the row that lies in the block's range with aligned 4-byte stores (funnel-shifted aligned loads)
// Pages are the unit of parallelism of the first kernel (a 1 MiB page is one warp's work: ~0.1 ms uncompressed PLAIN,
// a few ms of snappy); the copy is bounded by HBM traffic.
#include <cstdlib>

#include "spl_device.cuh"
#include "spl_parquet.h"

namespace {

struct WarpLanes {
    static constexpr uint32_t NL = 32;
    uint32_t lane;
    uint8_t* win; uint8_t* inbuf;          // snappy: 64 KiB ring of output + two 4 KiB input slots in shared memory (or null)
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, uint32_t src) const { return __shfl_sync(FULL, v, src); }
    __device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(FULL, p); }
};
#define PQ_RING_SMEM (SPL_SNAPPY_WIN + SPL_SNAPPY_INBUF)

#define PQ_THREADS 256
#define PQ_ROWS_PER_BLOCK 2048u
#define PQ_COPY_BYTES 16384u

// RING: one warp per block with 72 KiB of shared memory (three pages in flight per SM); else four warps per block
template <bool RING>
__global__ void __launch_bounds__(RING ? 32 : 128) k_pq_pages(SplPqWork w, const uint32_t first, const uint32_t count, const bool dict_pass) {
    extern __shared__ __align__(16) uint8_t pq_smem[];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= count) return;
    const SplPqPage pg = w.pages[first + warp];
    if (spl_pq_first_pass(pg) != dict_pass) return;
    WarpLanes g{threadIdx.x & 31u, RING ? pq_smem : nullptr, RING ? pq_smem + SPL_SNAPPY_WIN : nullptr};
    const uint32_t err = spl_pq_decode_page(g, pg, w.file, w.scratch, SplPqSpans{w.row_off, w.row_len}, SplPqSpans{w.dict_off, w.dict_len});
    if (err && g.lane == 0) atomicOr(&w.counters[SPL_PQCTR_ERR], err);
}

__global__ void __launch_bounds__(PQ_THREADS) k_pq_bsum(SplPqWork w) {
    __shared__ unsigned long long wsum[PQ_THREADS / 32];
    const uint32_t tid = threadIdx.x;
    const uint64_t r0 = (uint64_t)blockIdx.x * PQ_ROWS_PER_BLOCK;
    unsigned long long s = 0;
    for (uint32_t k = 0; k < PQ_ROWS_PER_BLOCK / PQ_THREADS; ++k) {
        const uint64_t r = r0 + k * PQ_THREADS + tid;
        if (r < w.n_rows) s += w.row_len[r];
    }
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(FULL, s, o);
    if ((tid & 31u) == 0) wsum[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) { unsigned long long t = 0; for (uint32_t q = 0; q < PQ_THREADS / 32; ++q) t += wsum[q]; w.bsum[blockIdx.x] = t; }
}

// exclusive prefix of bsum[0, n_blocks) in place, bsum[n_blocks] = total; one block
__global__ void __launch_bounds__(1024) k_pq_bscan(SplPqWork w) {
    __shared__ unsigned long long wtot[32];
    __shared__ unsigned long long carry_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < w.n_blocks; base += 1024) {
        const uint32_t i = base + tid;
        const unsigned long long v = i < w.n_blocks ? w.bsum[i] : 0ull;
        unsigned long long incl = v;
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += t; }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        unsigned long long before = carry_s;
        for (uint32_t q = 0; q < warp; ++q) before += wtot[q];
        if (i < w.n_blocks) w.bsum[i] = before + incl - v;
        __syncthreads();
        if (tid == 1023) carry_s = before + incl;
        __syncthreads();
    }
    if (tid == 0) {
        w.bsum[w.n_blocks] = carry_s;
        *reinterpret_cast<unsigned long long*>(&w.counters[SPL_PQCTR_TEXT]) = carry_s;
    }
}

__global__ void __launch_bounds__(PQ_THREADS) k_pq_offsets(SplPqWork w) {
    __shared__ unsigned long long wtot[PQ_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr uint32_t PER = PQ_ROWS_PER_BLOCK / PQ_THREADS;
    const uint64_t r0 = (uint64_t)blockIdx.x * PQ_ROWS_PER_BLOCK + (uint64_t)tid * PER;   // a thread owns PER consecutive rows
    uint32_t len[PER];
    unsigned long long s = 0;
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) { len[k] = r0 + k < w.n_rows ? w.row_len[r0 + k] : 0u; s += len[k]; }
    unsigned long long incl = s;
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += t; }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    unsigned long long run = w.bsum[blockIdx.x] + incl - s;
    for (uint32_t q = 0; q < warp; ++q) run += wtot[q];
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) {
        if (r0 + k < w.n_rows) w.out_off[r0 + k] = run;
        run += len[k];
    }
    if (blockIdx.x == 0 && tid == 0) w.out_off[w.n_rows] = w.bsum[w.n_blocks];
}

__device__ __forceinline__ const uint8_t* span_ptr(const SplPqWork& w, uint64_t off) {
    return (off & SPL_PQ_IN_SCRATCH) ? w.scratch + (off & ~SPL_PQ_IN_SCRATCH) : w.file + off;
}

// n bytes src -> dst by one warp; src is readable up to 3 bytes beyond its end rounded to a word (buffers are padded)
__device__ __forceinline__ void warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane) {
    uint32_t h = (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u);
    if (h > n) h = n;
    if (lane < h) dst[lane] = src[lane];
    const uint32_t nw = (n - h) >> 2;
    const uint8_t* s = src + h;
    const uint32_t sh = ((uint32_t)(uintptr_t)s & 3u) * 8u;
    const uint32_t* sa = reinterpret_cast<const uint32_t*>((uintptr_t)s & ~(uintptr_t)3);
    uint32_t* da = reinterpret_cast<uint32_t*>(dst + h);
    if (sh == 0) {
        for (uint32_t i = lane; i < nw; i += 32) da[i] = sa[i];
    } else {
        for (uint32_t i = lane; i < nw; i += 32) da[i] = __funnelshift_r(sa[i], sa[i + 1], sh);
    }
    const uint32_t t = (n - h) & 3u, o = h + nw * 4u;
    if (lane < t) dst[o + lane] = src[o + lane];
}

__global__ void __launch_bounds__(PQ_THREADS) k_pq_copy(SplPqWork w) {
    __shared__ uint64_t s_r0;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t total = w.out_off[w.n_rows];
    const uint64_t B0 = (uint64_t)blockIdx.x * PQ_COPY_BYTES;
    if (B0 >= total) return;
    const uint64_t B1 = B0 + PQ_COPY_BYTES < total ? B0 + PQ_COPY_BYTES : total;
    if (tid == 0) {                                              // the first row that ends behind B0
        uint64_t lo = 0, hi = w.n_rows;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (w.out_off[mid + 1] > B0) hi = mid; else lo = mid + 1;
        }
        s_r0 = lo;
    }
    __syncthreads();
    for (uint64_t r = s_r0 + warp; r < w.n_rows; r += PQ_THREADS / 32) {
        const uint64_t o0 = w.out_off[r];
        if (o0 >= B1) break;
        const uint64_t o1 = w.out_off[r + 1];
        const uint64_t c0 = o0 > B0 ? o0 : B0, c1 = o1 < B1 ? o1 : B1;
        if (c1 > c0) warp_copy(w.out_text + c0, span_ptr(w, w.row_off[r]) + (c0 - o0), (uint32_t)(c1 - c0), lane);
    }
}

}  // namespace

int spl_launch_pq_spans(const SplPqWork& w, uint32_t first_page, uint32_t n_pages, bool has_dict, bool has_snappy, cudaStream_t stream) {
    if (n_pages == 0) return 0;
    // snappy out of shared memory (SPL_PQ_SNAPPY_RING=0: the plain decoder, every element through global memory -- A/B)
    static const bool ring_on = [] { const char* e = getenv("SPL_PQ_SNAPPY_RING"); return !e || e[0] != '0'; }();
    int n = 0;
    for (int pass = 0; pass < (has_dict ? 2 : 1); ++pass, ++n) {            // without a dictionary page there is no second pass
        if (has_snappy && ring_on) {
            cudaFuncSetAttribute(k_pq_pages<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PQ_RING_SMEM);
            k_pq_pages<true><<<n_pages, 32, PQ_RING_SMEM, stream>>>(w, first_page, n_pages, pass == 0);
        } else {
            k_pq_pages<false><<<(n_pages + 3) / 4, 128, 0, stream>>>(w, first_page, n_pages, pass == 0);
        }
    }
    k_pq_bsum<<<w.n_blocks, PQ_THREADS, 0, stream>>>(w);
    k_pq_bscan<<<1, 1024, 0, stream>>>(w);
    k_pq_offsets<<<w.n_blocks, PQ_THREADS, 0, stream>>>(w);
    return n + 3;
}

int spl_launch_pq_copy(const SplPqWork& w, uint64_t text_bytes, cudaStream_t stream) {
    if (text_bytes == 0) return 0;
    k_pq_copy<<<(unsigned)((text_bytes + PQ_COPY_BYTES - 1) / PQ_COPY_BYTES), PQ_THREADS, 0, stream>>>(w);
    return 1;
}
