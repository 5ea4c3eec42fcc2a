// Device decode path: token ids + per-document token offsets -> bytes + byte offsets.
//
// Replaces, for batches, Tokenizer::decode_bytes / decode_batch (/root/reference/src/core/tokenizer.rs:877-897,
// :945-958): per id a table lookup (vocabulary bytes, byte-level keys already translated back to raw bytes; special
// token strings where the vocabulary has no entry; unknown ids contribute nothing) and concatenation.
//
//   k_dec_len    byte length of every id, summed per tile of SPL_DEC_TILE ids
//   k_dec_scan   exclusive prefix of the tile sums (one block) + the batch total
//   k_dec_emit   per tile: exclusive scan of the lengths, bytes staged in shared memory and written with coalesced
//                16-byte-wide rows; byte offset of every document that starts in the tile
//
// A pure gather: bounded by HBM traffic (4 B per id in, ~3.6 B per id out on English text) and by the L2-resident
// id -> bytes table.
#include "spl_device.cuh"

#define DEC_THREADS 256
#define DEC_PER     8u                                  // ids per thread
#define DEC_STAGE   16384u                              // bytes of a tile staged in shared memory (else: direct stores)

struct SplDecWork {
    const uint32_t* ids;         // [n_tok]
    uint64_t        n_tok;
    const uint64_t* tok_off;     // [n_docs+1] token offsets of the documents
    uint64_t        n_docs;
    uint32_t        n_tiles;     // n_tok / SPL_DEC_TILE + 1
    uint32_t*       tile_sum;    // [n_tiles]
    uint64_t*       tile_pref;   // [n_tiles + 1]; [n_tiles] = total bytes
    uint8_t*        out;         // [capacity]
    uint64_t        capacity;
    uint64_t*       out_off;     // [n_docs+1] byte offsets
    const SplTables* T;
};

__device__ __forceinline__ uint32_t dec_len(const SplTables* T, uint32_t id, uint32_t& src) {
    if (id >= T->n_dec) { src = 0; return 0u; }
    const uint32_t a = __ldg(T->dec_off + id), b = __ldg(T->dec_off + id + 1);
    src = a;
    return b - a;
}

__global__ void __launch_bounds__(DEC_THREADS) k_dec_len(SplDecWork w) {
    __shared__ uint32_t s_w[DEC_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t t0 = (uint64_t)blockIdx.x * SPL_DEC_TILE;
    uint32_t sum = 0;
#pragma unroll
    for (uint32_t q = 0; q < DEC_PER; ++q) {
        const uint64_t i = t0 + q * DEC_THREADS + tid;            // coalesced
        uint32_t src;
        if (i < w.n_tok) sum += dec_len(w.T, __ldg(w.ids + i), src);
    }
    sum = __reduce_add_sync(FULL, sum);
    if (lane == 0) s_w[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int q = 0; q < DEC_THREADS / 32; ++q) t += s_w[q];
        w.tile_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) k_dec_scan(SplDecWork w) {
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < w.n_tiles; c0 += 1024) {
        const uint32_t i = c0 + tid;
        const uint64_t v = i < w.n_tiles ? (uint64_t)w.tile_sum[i] : 0ull;
        uint64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t a = __shfl_up_sync(FULL, (uint32_t)incl, o), b = __shfl_up_sync(FULL, (uint32_t)(incl >> 32), o);
            if (lane >= (uint32_t)o) incl += (uint64_t)a | ((uint64_t)b << 32);
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t x = s_w[lane], xi = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t a = __shfl_up_sync(FULL, (uint32_t)xi, o), b = __shfl_up_sync(FULL, (uint32_t)(xi >> 32), o);
                if (lane >= (uint32_t)o) xi += (uint64_t)a | ((uint64_t)b << 32);
            }
            s_w[lane] = xi - x;
        }
        __syncthreads();
        const uint64_t carry = s_carry;
        if (i < w.n_tiles) w.tile_pref[i] = carry + s_w[warp] + incl - v;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_w[warp] + incl;
        __syncthreads();
    }
    if (tid == 0) w.tile_pref[w.n_tiles] = s_carry;
}

struct DecSmem {
    __align__(16) uint8_t stage[DEC_STAGE + 16];
    uint32_t pos[SPL_DEC_TILE + 1];                    // byte offset of every id inside the tile; [TILE] = tile bytes
    __align__(16) uint32_t wtot[DEC_THREADS / 32];     // aligned: the compiler reads it with LDS.128, which must not cover pos[TILE]
    uint64_t d0, d1;
};

__global__ void __launch_bounds__(DEC_THREADS) k_dec_emit(SplDecWork w) {
    __shared__ DecSmem sm;
    const SplTables* T = w.T;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t t0 = (uint64_t)blockIdx.x * SPL_DEC_TILE;
    const uint64_t pref = w.tile_pref[blockIdx.x];
    if (w.tile_pref[w.n_tiles] > w.capacity) return;              // the caller learns the needed size and retries

    // every thread owns DEC_PER consecutive ids (so that the scan is a per-thread running sum)
    uint32_t len[DEC_PER], src[DEC_PER], sum = 0;
#pragma unroll
    for (uint32_t q = 0; q < DEC_PER; ++q) {
        const uint64_t i = t0 + tid * DEC_PER + q;
        len[q] = 0; src[q] = 0;
        if (i < w.n_tok) len[q] = dec_len(T, __ldg(w.ids + i), src[q]);
        sum += len[q];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) sm.wtot[warp] = incl;
    if (tid == 0) {
        // documents whose first id lies in this tile: [d0, d1) by binary search over the token offsets
        uint64_t lo = 0, hi = w.n_docs + 1;
        while (lo < hi) { uint64_t m = (lo + hi) >> 1; if (w.tok_off[m] < t0) lo = m + 1; else hi = m; }
        sm.d0 = lo;
        hi = w.n_docs + 1;
        while (lo < hi) { uint64_t m = (lo + hi) >> 1; if (w.tok_off[m] < t0 + SPL_DEC_TILE) lo = m + 1; else hi = m; }
        sm.d1 = lo;
    }
    __syncthreads();
    uint32_t base = incl - sum, total = 0;
#pragma unroll
    for (uint32_t q = 0; q < DEC_THREADS / 32; ++q) {
        const uint32_t t = sm.wtot[q];
        base += q < warp ? t : 0u;
        total += t;
    }
    {
        uint32_t run = base;
#pragma unroll
        for (uint32_t q = 0; q < DEC_PER; ++q) { sm.pos[tid * DEC_PER + q] = run; run += len[q]; }
        if (tid == DEC_THREADS - 1) sm.pos[SPL_DEC_TILE] = run;
    }
    uint8_t* __restrict__ out = w.out + pref;
    if (total <= DEC_STAGE) {
        // bytes of the tile into shared memory, then coalesced to global memory
        uint32_t run = base;
#pragma unroll
        for (uint32_t q = 0; q < DEC_PER; ++q) {
            const uint8_t* __restrict__ s = T->dec_bytes + src[q];
            for (uint32_t b = 0; b < len[q]; ++b) sm.stage[run + b] = __ldg(s + b);
            run += len[q];
        }
        __syncthreads();
        // head up to the first 16-byte boundary of the destination, 16-byte rows, tail
        const uint32_t mis = (uint32_t)((16u - ((uintptr_t)out & 15u)) & 15u);
        const uint32_t head = mis < total ? mis : total;
        if (tid < head) out[tid] = sm.stage[tid];
        const uint32_t rows = (total - head) / 16u;
        for (uint32_t r = tid; r < rows; r += DEC_THREADS) {
            const uint8_t* s = sm.stage + head + r * 16u;          // not 16-byte aligned in shared memory: assemble from words
            uint4 v;
            uint32_t x[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                x[k] = (uint32_t)s[4 * k] | ((uint32_t)s[4 * k + 1] << 8) | ((uint32_t)s[4 * k + 2] << 16) | ((uint32_t)s[4 * k + 3] << 24);
            v.x = x[0]; v.y = x[1]; v.z = x[2]; v.w = x[3];
            *reinterpret_cast<uint4*>(out + head + r * 16u) = v;
        }
        const uint32_t done = head + rows * 16u;
        if (tid < total - done) out[done + tid] = sm.stage[done + tid];
    } else {
        uint32_t run = base;
#pragma unroll
        for (uint32_t q = 0; q < DEC_PER; ++q) {
            const uint8_t* __restrict__ s = T->dec_bytes + src[q];
            for (uint32_t b = 0; b < len[q]; ++b) out[run + b] = __ldg(s + b);
            run += len[q];
        }
        __syncthreads();
    }
    // byte offsets of the documents that start here (document n_docs "starts" at n_tok)
    for (uint64_t d = sm.d0 + tid; d < sm.d1; d += DEC_THREADS)
        w.out_off[d] = pref + sm.pos[(uint32_t)(w.tok_off[d] - t0)];
}

// host side ----------------------------------------------------------------------------------------------
void spl_launch_decode_count(const SplDecLaunch& L, cudaStream_t stream) {
    SplDecWork w;
    w.ids = L.ids; w.n_tok = L.n_tok; w.tok_off = L.tok_off; w.n_docs = L.n_docs; w.n_tiles = L.n_tiles;
    w.tile_sum = L.tile_sum; w.tile_pref = L.tile_pref; w.out = L.out; w.capacity = L.capacity; w.out_off = L.out_off; w.T = L.T;
    k_dec_len<<<L.n_tiles, DEC_THREADS, 0, stream>>>(w);
    k_dec_scan<<<1, 1024, 0, stream>>>(w);
}

void spl_launch_decode_emit(const SplDecLaunch& L, cudaStream_t stream) {
    SplDecWork w;
    w.ids = L.ids; w.n_tok = L.n_tok; w.tok_off = L.tok_off; w.n_docs = L.n_docs; w.n_tiles = L.n_tiles;
    w.tile_sum = L.tile_sum; w.tile_pref = L.tile_pref; w.out = L.out; w.capacity = L.capacity; w.out_off = L.out_off; w.T = L.T;
    k_dec_emit<<<L.n_tiles, DEC_THREADS, 0, stream>>>(w);
}
