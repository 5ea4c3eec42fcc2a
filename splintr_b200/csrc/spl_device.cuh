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

// one bucket of the pair table: four tags + four values = one 32-byte sector, read with ONE 256-bit load
// (LDG.E.256 -- sm_100 and later)
struct PairBucket { uint4 tag, val; };
__device__ __forceinline__ PairBucket pair_bucket_load(const uint32_t* __restrict__ tab, uint32_t b) {
    const uint32_t* p = tab + (size_t)b * SPL_PAIR_WORDS;
    unsigned long long a, c, d, e;
    asm("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(c), "=l"(d), "=l"(e) : "l"(p));
    PairBucket r;
    r.tag = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)c, (uint32_t)(c >> 32));
    r.val = make_uint4((uint32_t)d, (uint32_t)(d >> 32), (uint32_t)e, (uint32_t)(e >> 32));
    return r;
}
// the key as the bucket holds it
struct PairKey { uint32_t tag, hi; };
__device__ __forceinline__ PairKey pair_key(uint32_t l, uint32_t r) { return PairKey{spl_pair_tag(l, r), spl_pair_hi(l)}; }
// 0: key absent (bucket not full), 1: found (out set), 2: bucket full, go on with the next one
__device__ __forceinline__ int pair_bucket_match(const PairBucket& k, const PairKey key, uint32_t& out) {
    const bool m0 = k.tag.x == key.tag && (k.val.x & ~SPL_SYM_MASK) == key.hi;
    const bool m1 = k.tag.y == key.tag && (k.val.y & ~SPL_SYM_MASK) == key.hi;
    const bool m2 = k.tag.z == key.tag && (k.val.z & ~SPL_SYM_MASK) == key.hi;
    const bool m3 = k.tag.w == key.tag && (k.val.w & ~SPL_SYM_MASK) == key.hi;
    if (m0 | m1 | m2 | m3) { out = (m0 ? k.val.x : m1 ? k.val.y : m2 ? k.val.z : k.val.w) & SPL_SYM_MASK; return 1; }
    return k.val.w == SPL_PAIR_EMPTY ? 0 : 2;
}

__device__ __forceinline__ uint32_t pair_lookup(const uint32_t* __restrict__ tab, uint32_t log2, uint32_t l, uint32_t r) {
    const PairKey key = pair_key(l, r);
    const uint32_t mask = (1u << log2) - 1;
    uint32_t b = spl_pair_hash(l, r, log2), out = SPL_RANK_NONE;
    for (;;) {
        PairBucket k = pair_bucket_load(tab, b);
        int m = pair_bucket_match(k, key, out);
        if (m != 2) return out;
        b = (b + 1) & mask;
    }
}

// whole-piece probe, 1..8 bytes: buckets of SPL_T8_WAYS entries (one sector)
__device__ __forceinline__ uint32_t lookup8(const SplKey8* __restrict__ t, uint32_t log2, uint64_t k0, uint32_t len) {
    const uint32_t lo = (uint32_t)k0, hi = (uint32_t)(k0 >> 32);
    const uint32_t mask = (1u << log2) - 1;
    uint32_t b = spl_hash8(lo, hi, len, log2);
    for (;;) {
        const uint4* p = reinterpret_cast<const uint4*>(t + (size_t)b * SPL_T8_WAYS);
        uint4 v0 = __ldg(p), v1 = __ldg(p + 1);               // {k0 lo, k0 hi, id, len}
        if (v0.w == len && v0.x == lo && v0.y == hi) return v0.z;
        if (v1.w == len && v1.x == lo && v1.y == hi) return v1.z;
        if (v1.w == 0) return SPL_RANK_NONE;
        b = (b + 1) & mask;
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

// three independent pair probes issued back to back (the ranks a merge window needs per candidate, k_bpe_long)
__device__ __forceinline__ void pair_lookup3(const uint32_t* __restrict__ tab, uint32_t log2,
                                             bool va, uint32_t la, uint32_t ra, bool vb, uint32_t lb, uint32_t rb,
                                             bool vc, uint32_t lc, uint32_t rc,
                                             uint32_t& outa, uint32_t& outb, uint32_t& outc) {
    const uint32_t mask = (1u << log2) - 1;
    const PairKey ka = pair_key(la, ra), kb = pair_key(lb, rb), kc = pair_key(lc, rc);
    uint32_t ba = spl_pair_hash(la, ra, log2), bb = spl_pair_hash(lb, rb, log2), bc = spl_pair_hash(lc, rc, log2);
    PairBucket xa, xb, xc;
    xa.tag = xa.val = make_uint4(SPL_PAIR_EMPTY, SPL_PAIR_EMPTY, SPL_PAIR_EMPTY, SPL_PAIR_EMPTY);
    xb = xa; xc = xa;
    if (va) xa = pair_bucket_load(tab, ba);
    if (vb) xb = pair_bucket_load(tab, bb);
    if (vc) xc = pair_bucket_load(tab, bc);
    outa = SPL_RANK_NONE; outb = SPL_RANK_NONE; outc = SPL_RANK_NONE;
    if (va)
        while (pair_bucket_match(xa, ka, outa) == 2) { ba = (ba + 1) & mask; xa = pair_bucket_load(tab, ba); }
    if (vb)
        while (pair_bucket_match(xb, kb, outb) == 2) { bb = (bb + 1) & mask; xb = pair_bucket_load(tab, bb); }
    if (vc)
        while (pair_bucket_match(xc, kc, outc) == 2) { bc = (bc + 1) & mask; xc = pair_bucket_load(tab, bc); }
}

// two independent pair probes issued back to back (the two re-ranks after a merge)
__device__ __forceinline__ void pair_lookup2(const uint32_t* __restrict__ tab, uint32_t log2,
                                             bool va, uint32_t la, uint32_t ra, bool vb, uint32_t lb, uint32_t rb,
                                             uint32_t& outa, uint32_t& outb) {
    const uint32_t mask = (1u << log2) - 1;
    const PairKey ka = pair_key(la, ra), kb = pair_key(lb, rb);
    uint32_t ba = spl_pair_hash(la, ra, log2), bb = spl_pair_hash(lb, rb, log2);
    PairBucket xa, xb;
    xa.tag = xa.val = make_uint4(SPL_PAIR_EMPTY, SPL_PAIR_EMPTY, SPL_PAIR_EMPTY, SPL_PAIR_EMPTY);
    xb = xa;
    if (va) xa = pair_bucket_load(tab, ba);
    if (vb) xb = pair_bucket_load(tab, bb);
    outa = SPL_RANK_NONE; outb = SPL_RANK_NONE;
    if (va)
        while (pair_bucket_match(xa, ka, outa) == 2) { ba = (ba + 1) & mask; xa = pair_bucket_load(tab, ba); }
    if (vb)
        while (pair_bucket_match(xb, kb, outb) == 2) { bb = (bb + 1) & mask; xb = pair_bucket_load(tab, bb); }
}
