// Bit tricks of a windowed merge round (k_bpe_long, spl_encode.cu bpe_group): host + device, so that tests/ can
// hold them against a part-by-part evaluation on the CPU (tests/test_bpe_rounds_host.py).
//
// A lane's masks cover its B <= 32 consecutive parts, bit j = pair (e0 + j, e0 + j + 1).  Order pairs by
// key = (rank, position).  If no pair created in the window ranks below the window's end, the sequential loop of
// bpe.rs:119-167 merges pair i unless a neighbour merged first:
//     m(i) = !(m(i-1) && key(i-1) < key(i)) && !(m(i+1) && key(i+1) < key(i))
// V: neither neighbour is lower (valley, m = 1); DL: only the left one is (slope that rises to the right: m(i) = !m(i-1));
// DR: only the right one is (m(i) = !m(i+1)); PK: both are (m = 1 iff neither neighbour has m = 1).
#pragma once
#include "spl_common.h"

SPL_HD uint32_t spl_brev32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

// the runs of D that start right above the bits of st (the addition ripples through exactly those)
SPL_HD uint32_t spl_runs_above(uint32_t D, uint32_t st) { return D & ~(D + (st << 1)); }

// m along upward slopes: D = slope bits, V = valleys (m = 1); m alternates with the distance from the valley.  A run
// that starts at bit `first` continues the slope of the neighbouring lane, whose boundary part has m = cin.
SPL_HD uint32_t spl_slope_m(uint32_t D, uint32_t V, uint32_t cin, uint32_t first) {
    const uint32_t E = 0x55555555u;
    const uint32_t run0 = D & ~(D + first);                                        // run from `first` upwards (empty if D lacks that bit)
    const uint32_t P = (first & E) ? ~E : E;                                       // positions at an even distance from the part below `first`
    return (spl_runs_above(D, V & E) & E) | (spl_runs_above(D, V & ~E) & ~E) | (run0 & (cin ? P : ~P));
}

// m of the valleys and both kinds of slopes of one lane, given the boundary bits of its neighbours:
// cin = m of the last part of the lane before, cin2 = m of the first part of the lane after.  (__brev turns the
// slopes that rise to the left into the same problem.)  Peaks come last: spl_window_peaks.
SPL_HD uint32_t spl_window_slopes(uint32_t fV, uint32_t fDL, uint32_t fDR, uint32_t cin, uint32_t cin2, uint32_t B) {
    return fV | spl_slope_m(fDL, fV, cin, 1u) | spl_brev32(spl_slope_m(spl_brev32(fDR), spl_brev32(fV), cin2, 1u << (32u - B)));
}

// a peak merges iff neither neighbour does
SPL_HD uint32_t spl_window_peaks(uint32_t m, uint32_t fPK, uint32_t cin, uint32_t cin2, uint32_t B) {
    return m | (fPK & ~((m << 1) | cin) & ~((m >> 1) | (cin2 << (B - 1u))));
}
