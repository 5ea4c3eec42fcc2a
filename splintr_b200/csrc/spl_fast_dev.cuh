// Device-side plumbing of the bit-parallel pre-tokenizer (spl_pretok_fast.h) shared by k_pretok_fast (spl_kernels.cu)
// and the fused k_pretok_probe (spl_encode.cu): the shared-memory mask arrays and their accessor.
#pragma once
#include "spl_device.cuh"
#include "spl_pretok_fast.h"

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

