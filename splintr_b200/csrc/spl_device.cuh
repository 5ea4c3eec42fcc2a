// Device helpers shared by spl_kernels.cu (pre-tokenizer) and spl_encode.cu (encode stage):
// bitmap searches and the table probes that mirror spl_host.cpp bit for bit.
#pragma once
#include "spl_kernels.cuh"

#define FULL 0xFFFFFFFFu

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
