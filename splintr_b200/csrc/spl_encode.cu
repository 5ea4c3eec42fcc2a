// Device encode path, second half: piece-start bitmap -> token ids in document order.
//
//   k_probe      one thread per piece: whole-piece vocabulary probe (tokenizer.rs:703-705, bpe.rs:73-80); the id of a
//                hit goes to pv[] in piece order, a miss is appended to the miss list of its length class
//   k_bpe        the leftmost-min-rank merge loop of bpe.rs:83-194 for every listed piece -- one thread per piece up
//                to 32 bytes, one warp up to 3968 bytes, one block beyond -- ids to pool[] at the piece's byte position
//   k_chunk_scan exclusive prefix of the id counts of 32-tile chunks (k_emit adds the tiles inside its chunk)
//   k_emit       pv[] + pool[] -> ids in document order (the collect of tokenizer.rs:806 and the Rayon collect of
//                encode_batch, tokenizer.rs:932-934) and the per-document output offsets
//
// No block waits for another block inside a kernel; the miss lists decouple the rare, latency-bound merge loop from
// the streaming probe so that both run at full occupancy.  Integer / byte work; no tensor cores.
#include "spl_device.cuh"

// ------------------------------------------------------------------------------------------
// probes that only this stage uses
// ------------------------------------------------------------------------------------------

// special-token id of the span tx[0, len) (linear search; special spans are rare)
__device__ uint32_t special_id_g(const SplTables* T, const uint8_t* __restrict__ tx, uint32_t len) {
    for (uint32_t k = 0; k < T->n_special; ++k) {
        uint32_t o = T->sp_off[k];
        if (T->sp_off[k + 1] - o != len) continue;
        bool eq = true;
        for (uint32_t j = 0; j < len; ++j)
            if (__ldg(tx + j) != T->sp_bytes[o + j]) { eq = false; break; }
        if (eq) return T->sp_id[k];
    }
    return SPL_RANK_NONE;
}

// whole-piece probe of a 17..128-byte piece by one thread (long-key table, verified against the token bytes)
__device__ uint32_t lookupL_thread(const SplTables* T, const uint32_t* text, uint32_t s, uint32_t len) {
    uint64_t sum = 0;
    for (uint32_t i = 0; i * 8 < len; ++i) {
        uint64_t wv = sm_load8(text, s + i * 8);
        uint32_t rem = len - i * 8;
        if (rem < 8) wv &= (1ull << (8 * rem)) - 1;
        sum += spl_hashL_word(wv, i);
    }
    uint64_t hv = spl_hashL_final(sum, len);
    uint32_t mask = (1u << T->tl_log2) - 1, h = (uint32_t)(hv >> (64 - T->tl_log2));
    for (;;) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(T->tl + h));     // {hash lo, hash hi, id, len}
        if (v.w == 0) return SPL_RANK_NONE;
        if (v.w == len && v.x == (uint32_t)hv && v.y == (uint32_t)(hv >> 32)) {
            const uint8_t* kb = T->tok_bytes + __ldg(T->tok_off + v.z);
            bool ok = true;
            for (uint32_t j = 0; j < len; ++j) ok &= (sm_byte(text, s + j) == __ldg(kb + j));
            if (ok) return v.z;
        }
        h = (h + 1) & mask;
    }
}

// the same probe for a piece in global memory, by one warp (pieces longer than the probe halo; only vocabularies with
// keys beyond 128 bytes get here).  Every lane returns the id or SPL_RANK_NONE.
__device__ uint32_t lookupL_warp_g(const SplTables* T, const uint8_t* __restrict__ tx, uint32_t len) {
    const uint32_t lane = threadIdx.x & 31u;
    uint64_t sum = 0;
    for (uint32_t i = lane; i * 8 < len; i += 32) {
        uint64_t wv = 0;
        for (uint32_t b = 0; b < 8 && i * 8 + b < len; ++b) wv |= (uint64_t)__ldg(tx + i * 8 + b) << (8 * b);
        sum += spl_hashL_word(wv, i);
    }
    sum = warp_sum_u64(sum);
    uint64_t hv = spl_hashL_final(sum, len);
    uint32_t mask = (1u << T->tl_log2) - 1, h = (uint32_t)(hv >> (64 - T->tl_log2));
    for (;;) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(T->tl + h));
        if (v.w == 0) return SPL_RANK_NONE;
        if (v.w == len && v.x == (uint32_t)hv && v.y == (uint32_t)(hv >> 32)) {
            const uint8_t* kb = T->tok_bytes + __ldg(T->tok_off + v.z);
            bool ok = true;
            for (uint32_t j = lane; j < len; j += 32) ok &= (__ldg(tx + j) == __ldg(kb + j));
            if (__all_sync(FULL, ok)) return v.z;
        }
        h = (h + 1) & mask;
    }
}

__device__ __forceinline__ uint64_t ml_entry(uint32_t gpos, uint32_t len, uint32_t j) {
    uint32_t l = len < SPL_ML_LEN_SAT ? len : SPL_ML_LEN_SAT;
    return (uint64_t)gpos | ((uint64_t)l << 32) | ((uint64_t)j << 52);
}

// ------------------------------------------------------------------------------------------
// k_probe: one tile per block
// ------------------------------------------------------------------------------------------
#define PB_WORDS (SPL_PROBE_WIN / 32u + 1u)          // piece-start words staged: bits 0 .. SPL_PROBE_WIN + 31
#define PB_BITS  (PB_WORDS * 32u)

struct ProbeSmem {
    uint32_t text[SPL_PROBE_WIN / 4 + 4];     // staged bytes (+ slack for the unaligned 8-byte key loads)
    uint32_t pb[PB_WORDS];                    // piece-start bits
    uint32_t spw[SPL_TILE / 32];              // special-span bits of the tile (with_special)
    uint16_t plist[SPL_TILE + 2];             // window positions of the tile's piece starts, in order (+ end of the last piece)
    uint16_t slow[SPL_TILE];                  // pieces the one-sector probe did not settle (each warp: its own range)
    uint16_t mloc[SPL_TILE];                  // missed pieces: thread class from the bottom of the warp's range, warp class from its top
    uint32_t wtot[SPL_THREADS / 32];
    uint32_t last_end;                        // window position of the end of the tile's last piece
};

__global__ void __launch_bounds__(SPL_THREADS) k_probe(SplWork w) {
    __shared__ ProbeSmem sm;
    const SplTables* T = w.T;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t N = w.N, Nup = (N + 15u) & ~15u;
    const uint32_t tile = blockIdx.x, tile0 = tile * SPL_TILE;
    const SplKey8* __restrict__ t8 = T->t8;
    const uint32_t t8_log2 = T->t8_log2;

    // ---- stage the window -----------------------------------------------------------
    for (uint32_t v = tid; v < SPL_PROBE_WIN / 16 + 1; v += SPL_THREADS) {
        uint32_t g = tile0 + v * 16;
        uint4 x = make_uint4(0, 0, 0, 0);
        if (g < Nup) x = __ldg(reinterpret_cast<const uint4*>(w.text + g));
        reinterpret_cast<uint4*>(sm.text)[v] = x;
    }
    for (uint32_t v = tid; v < PB_WORDS; v += SPL_THREADS) sm.pb[v] = __ldg(w.pstart + (tile0 >> 5) + v);
    if (w.with_special && tid < SPL_TILE / 32) sm.spw[tid] = __ldg(w.spec + (tile0 >> 5) + tid);
    __syncthreads();

    // ---- piece list: positions of the piece starts of this tile, in order ------------------
    const uint32_t avail = N - tile0;                          // text bytes from tile0 on
    uint32_t my = (sm.pb[tid >> 1] >> ((tid & 1u) * 16u)) & 0xFFFFu;
    if (tid * 16u + 16u > avail) my &= (tid * 16u >= avail) ? 0u : ((1u << (avail - tid * 16u)) - 1u);   // sentinel bit at N
    uint32_t cnt = __popc(my), incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) sm.wtot[warp] = incl;
    __syncthreads();
    uint32_t base = incl - cnt, P = 0;
#pragma unroll
    for (uint32_t q = 0; q < SPL_THREADS / 32; ++q) {
        uint32_t t = sm.wtot[q];
        base += (q < warp) ? t : 0u;
        P += t;
    }
    while (my) {
        uint32_t b = __ffs(my) - 1;
        my &= my - 1;
        sm.plist[base++] = (uint16_t)(tid * 16u + b);
    }
    if (tid == 0) {
        uint32_t e = sm_next_bit(sm.pb, SPL_TILE < avail ? SPL_TILE : avail, PB_BITS);
        if (P && e >= PB_BITS) e = g_next_bit(w.pstart, tile0 + PB_BITS, N + 1) - tile0;    // the last piece leaves the staged bits
        sm.last_end = e;
        sm.plist[P] = (uint16_t)(e >= PB_BITS ? 0xFFFFu : e);      // 0xFFFF: see last_end
        w.tinfo[tile].np = P;
        if (P) atomicAdd(&w.chunk_cnt[tile / SPL_CHUNK_TILES], (int32_t)P);
    }
    __syncthreads();

    // From here on every warp works alone on its own range of pieces [jlo, jhi): no block barrier, a warp with
    // slow pieces does not hold the others back.
    const uint32_t per = (((P + SPL_THREADS / 32 - 1) / (SPL_THREADS / 32)) + 31u) & ~31u;
    const uint32_t jlo = warp * per < P ? warp * per : P, jhi = jlo + per < P ? jlo + per : P;
    const uint32_t pvbase = tile * SPL_TILE;

    // ---- hot loop, one thread per piece: pieces of up to 8 bytes, one sector of the bucketed table -------------
    uint32_t n_slow = 0;
    for (uint32_t j0 = jlo; j0 < jhi; j0 += 32) {
        const uint32_t j = j0 + lane;
        const bool valid = j < jhi;
        uint32_t s = 0, len = 0;
        if (valid) { s = sm.plist[j]; len = sm.plist[j + 1] - s; }      // end 0xFFFF: len is large, not probed here
        bool fast = valid && len <= 8;
        if (w.with_special && ((sm.spw[s >> 5] >> (s & 31)) & 1u)) fast = false;
        bool found = false;
        if (fast) {
            const uint32_t wi = s >> 2, sh = (s & 3u) * 8u;
            const uint32_t a = sm.text[wi], b = sm.text[wi + 1], c = sm.text[wi + 2];
            uint32_t lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);
            const uint32_t nb = len * 8u;
            lo &= nb >= 32u ? FULL : ((1u << nb) - 1u);
            hi &= nb <= 32u ? 0u : (nb >= 64u ? FULL : ((1u << (nb - 32u)) - 1u));
            const uint4* p = reinterpret_cast<const uint4*>(t8 + (size_t)spl_hash8(lo, hi, len, t8_log2) * SPL_T8_WAYS);
            const uint4 v0 = __ldg(p), v1 = __ldg(p + 1);       // {k0 lo, k0 hi, id, len}
            const bool h0 = v0.w == len && v0.x == lo && v0.y == hi;
            const bool h1 = v1.w == len && v1.x == lo && v1.y == hi;
            found = h0 || h1;
            if (found) w.pv[pvbase + j] = h0 ? v0.z : v1.z;
        }
        const bool sl = valid && !found;
        const uint32_t bal = __ballot_sync(FULL, sl);
        if (sl) sm.slow[jlo + n_slow + __popc(bal & lt_mask)] = (uint16_t)j;
        n_slow += __popc(bal);
    }
    __syncwarp();

    // ---- the rest: specials, longer pieces, second buckets, misses ------------------------------------------
    uint32_t n_short = 0, n_warp = 0;
    for (uint32_t i0 = 0; i0 < n_slow; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint32_t cls = 0, j = 0;                               // 0 settled, 1 thread class, 2 warp class, 3 big, 4 huge
        if (i < n_slow) {
            j = sm.slow[jlo + i];
            const uint32_t s = sm.plist[j];
            uint32_t e = sm.plist[j + 1];
            if (e == 0xFFFFu) e = sm.last_end;
            const uint32_t len = e - s, gpos = tile0 + s;
            uint32_t val = SPL_PV_NONE;
            if (w.with_special && ((sm.spw[s >> 5] >> (s & 31)) & 1u)) {
                val = special_id_g(T, w.text + gpos, len);
            } else if (len == 1) {
                uint32_t sy = T->byte_sym[sm_byte(sm.text, s)];
                val = sy < SPL_UNK_BASE ? sy : SPL_PV_NONE;   // unknown byte: no id (bpe.rs:73-75)
            } else {
                uint32_t id = SPL_RANK_NONE;
                if (len <= 8) {
                    uint64_t k0 = sm_load8(sm.text, s);
                    if (len < 8) k0 &= (1ull << (8 * len)) - 1;
                    id = lookup8(t8, t8_log2, k0, len);
                } else if (len <= 16) {
                    uint64_t k0 = sm_load8(sm.text, s), k1 = sm_load8(sm.text, s + 8);
                    if (len < 16) k1 &= (1ull << (8 * (len - 8))) - 1;
                    id = lookup16(T->t16, T->t16_log2, k0, k1, len);
                } else if (len <= SPL_PROBE_HALO && len <= T->max_key_len) {
                    id = lookupL_thread(T, sm.text, s, len);
                }
                if (id != SPL_RANK_NONE) val = id;
                else cls = len <= SPL_SHORT_MAX ? 1u : len <= SPL_WARP_MAX ? 2u : len <= SPL_BIG_MAX ? 3u : 4u;
            }
            if (cls == 0) {
                w.pv[pvbase + j] = val;
                if (val == SPL_PV_NONE) { atomicAdd(&w.tinfo[tile].extra, -1); atomicAdd(&w.chunk_cnt[tile / SPL_CHUNK_TILES], -1); }
            } else if (cls >= 3) {                             // rare: straight to the global list of its class
                uint32_t midx = cls == 3 ? w.ml_r0 + atomicAdd(&w.counters[SPL_CTR_BIG], 1u)
                                         : w.ml_r1 + atomicAdd(&w.counters[SPL_CTR_HUGE], 1u);
                w.mlist[midx] = ml_entry(gpos, len, j);
                w.pv[pvbase + j] = SPL_PV_MISS | midx;
            }
        }
        uint32_t bal = __ballot_sync(FULL, cls == 1);
        if (cls == 1) sm.mloc[jlo + n_short + __popc(bal & lt_mask)] = (uint16_t)j;
        n_short += __popc(bal);
        bal = __ballot_sync(FULL, cls == 2);
        if (cls == 2) sm.mloc[jhi - 1 - (n_warp + __popc(bal & lt_mask))] = (uint16_t)j;
        n_warp += __popc(bal);
    }
    __syncwarp();

    // ---- publish the warp's misses ------------------------------------------------------------
    if (n_short | n_warp) {
        uint32_t g_short = 0, g_warp = 0;
        if (lane == 0) {
            if (n_short) g_short = atomicAdd(&w.counters[SPL_CTR_SHORT], n_short);
            if (n_warp) g_warp = atomicAdd(&w.counters[SPL_CTR_WARP], n_warp);
        }
        g_short = __shfl_sync(FULL, g_short, 0);
        g_warp = __shfl_sync(FULL, g_warp, 0);
        for (uint32_t i = lane; i < n_short; i += 32) {
            uint32_t j = sm.mloc[jlo + i], s = sm.plist[j], e = sm.plist[j + 1], midx = g_short + i;
            if (e == 0xFFFFu) e = sm.last_end;
            w.mlist[midx] = ml_entry(tile0 + s, e - s, j);
            w.pv[pvbase + j] = SPL_PV_MISS | midx;
        }
        for (uint32_t i = lane; i < n_warp; i += 32) {
            uint32_t j = sm.mloc[jhi - 1 - i], s = sm.plist[j], e = sm.plist[j + 1], midx = w.ml_r0 - 1 - (g_warp + i);
            if (e == 0xFFFFu) e = sm.last_end;
            w.mlist[midx] = ml_entry(tile0 + s, e - s, j);
            w.pv[pvbase + j] = SPL_PV_MISS | midx;
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_bpe: the merge loop (bpe.rs:83-194) for the listed pieces
// ------------------------------------------------------------------------------------------
struct BpeWarpSmem {
    uint32_t sym[SPL_WARP_MAX];
    uint32_t rnk[SPL_WARP_MAX];
    uint32_t tb[SPL_WARP_MAX / 32 + 1];
};
struct BpeBigSmem {
    uint32_t sym[SPL_BIG_MAX];
    uint32_t rnk[SPL_BIG_MAX];
    uint32_t tb[SPL_BIG_MAX / 32 + 1];
};
union BpeSmem {
    struct { uint32_t sym[SPL_SHORT_MAX][SPL_BPE_THREADS]; uint32_t rnk[SPL_SHORT_MAX][SPL_BPE_THREADS]; } th;
    BpeWarpSmem wp[SPL_BPE_THREADS / 32];
    BpeBigSmem big;
    uint64_t red[SPL_BPE_THREADS];
};

__device__ __forceinline__ void bpe_finish(const SplWork& w, uint64_t* slot, uint32_t gpos, uint32_t cnt) {
    *slot = (uint64_t)gpos | ((uint64_t)cnt << 32);
    if (cnt != 1u) {
        atomicAdd(&w.tinfo[gpos / SPL_TILE].extra, (int32_t)cnt - 1);
        atomicAdd(&w.chunk_cnt[gpos / (SPL_TILE * SPL_CHUNK_TILES)], (int32_t)cnt - 1);
    }
}

// One THREAD merges the piece tx[0, n), 2 <= n <= 32: parts are the set bits of `live` (bit i = a part starts at byte
// i), their symbols in S(i), the rank of (part, next part) in R(i).  32 pieces merge side by side in a warp, so the
// probe latency of the re-ranks overlaps across pieces.  Returns the id count; ids go to out[0 ..].
__device__ uint32_t bpe_piece_thread(BpeSmem& sm, const SplTables* T, const uint8_t* __restrict__ tx, uint32_t n, uint32_t* __restrict__ out) {
    const uint64_t* __restrict__ ptab = T->pair;
    const uint32_t plog = T->pair_log2;
    const uint32_t t = threadIdx.x;
#define S(i) sm.th.sym[(i)][t]
#define R(i) sm.th.rnk[(i)][t]
#pragma unroll 4
    for (uint32_t i = 0; i < n; ++i) S(i) = __ldg(tx + i);
#pragma unroll 4
    for (uint32_t i = 0; i < n; ++i) S(i) = T->byte_sym[S(i)];
    {
        const uint32_t pmask = (1u << plog) - 1;
        for (uint32_t i = 0; i + 1 < n; i += 4) {                  // four independent probes in flight
            PairBucket bk[4];
            uint64_t key[4];
            uint32_t bb[4];
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q)
                if (i + q + 1 < n) {
                    key[q] = spl_pair_key(S(i + q), S(i + q + 1));
                    bb[q] = spl_pair_hash(key[q], plog);
                    bk[q] = pair_bucket_load(ptab, bb[q]);
                }
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q)
                if (i + q + 1 < n) {
                    uint32_t r = SPL_RANK_NONE;
                    while (pair_bucket_match(bk[q], key[q], r) == 2) { bb[q] = (bb[q] + 1) & pmask; bk[q] = pair_bucket_load(ptab, bb[q]); }
                    R(i + q) = r;
                }
        }
    }
    R(n - 1) = SPL_RANK_NONE;
    uint32_t live = n >= 32u ? 0xFFFFFFFFu : ((1u << n) - 1u);
    for (;;) {
        uint32_t best = SPL_RANK_NONE, bpos = 0;
        for (uint32_t m = live; m; m &= m - 1) {
            uint32_t i = __ffs(m) - 1;
            uint32_t r = R(i);
            if (r < best) { best = r; bpos = i; }                 // strict <: leftmost minimum (bpe.rs:133)
        }
        if (best == SPL_RANK_NONE) break;
        uint32_t above = bpos >= 31u ? 0u : (live & ~((2u << bpos) - 1u));
        uint32_t nx = __ffs(above) - 1;                            // the absorbed part (exists: its pair has a rank)
        uint32_t above2 = above & (above - 1);
        uint32_t below = live & ((1u << bpos) - 1u);
        bool has_nn = above2 != 0, has_pv = below != 0;
        uint32_t nn = has_nn ? __ffs(above2) - 1 : 0u, pv = has_pv ? 31u - __clz(below) : 0u;
        live &= ~(1u << nx);
        S(bpos) = best;                                            // merged id == its rank
        uint32_t r1, r0;
        pair_lookup2(ptab, plog, has_nn, best, has_nn ? S(nn) : 0u, has_pv, has_pv ? S(pv) : 0u, best, r1, r0);
        R(bpos) = r1;
        if (has_pv) R(pv) = r0;
    }
    uint32_t c = 0;
    for (uint32_t m = live; m; m &= m - 1) {
        uint32_t sy = S(__ffs(m) - 1);
        if (sy < SPL_UNK_BASE) out[c++] = sy;                      // bytes that are not in the vocabulary produce no id (bpe.rs:187-191)
    }
#undef S
#undef R
    return c;
}

// One WARP merges the piece tx[0, len): parts are delimited by the bits of tb, sym holds each part's symbol at its
// first byte, rnk the rank of (part, next part).  Returns (to every lane) the id count; ids go to out[0 ..].
__device__ uint32_t bpe_piece_warp(uint32_t* sym, uint32_t* rnk, uint32_t* tb, const SplTables* T,
                                   const uint8_t* __restrict__ tx, uint32_t len, uint32_t* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t nwords = (len + 31u) >> 5;
    if (len > SPL_PROBE_HALO && len <= T->max_key_len) {           // k_probe has already tried the shorter ones
        uint32_t id = lookupL_warp_g(T, tx, len);
        if (id != SPL_RANK_NONE) {
            if (lane == 0) out[0] = id;
            return 1u;
        }
    }
    for (uint32_t j = lane; j < len; j += 32) sym[j] = T->byte_sym[__ldg(tx + j)];
    for (uint32_t v = lane; v <= nwords; v += 32) {
        uint32_t lo = v * 32u;
        tb[v] = lo + 32u <= len ? FULL : (lo < len ? ((1u << (len - lo)) - 1u) : 0u);
    }
    __syncwarp();
    for (uint32_t j = lane; j < len; j += 32)
        rnk[j] = (j + 1 < len) ? pair_lookup(T->pair, T->pair_log2, sym[j], sym[j + 1]) : SPL_RANK_NONE;
    __syncwarp();
    for (;;) {
        uint32_t best = SPL_RANK_NONE, bpos = SPL_RANK_NONE;
        for (uint32_t j = lane; j < len; j += 32) {
            uint32_t r = rnk[j];
            if (r < best) { best = r; bpos = j; }
        }
        uint32_t m = __reduce_min_sync(FULL, best);
        if (m == SPL_RANK_NONE) break;
        uint32_t pos = __reduce_min_sync(FULL, best == m ? bpos : SPL_RANK_NONE);   // leftmost minimum (bpe.rs:133)
        uint32_t j = sm_next_bit(tb, pos + 1, len);            // the part being absorbed
        uint32_t k = sm_next_bit(tb, j + 1, len);              // its right neighbour (len if none)
        uint32_t h = sm_prev_bit(tb, pos, 0);                  // left neighbour (NONE if none)
        uint32_t symk = k < len ? sym[k] : 0u;
        uint32_t symh = h != SPL_RANK_NONE ? sym[h] : 0u;
        __syncwarp();
        if (lane == 0) {
            sym[pos] = m;                                      // merged id == its rank
            rnk[j] = SPL_RANK_NONE;
            tb[j >> 5] &= ~(1u << (j & 31));
            rnk[pos] = k < len ? pair_lookup(T->pair, T->pair_log2, m, symk) : SPL_RANK_NONE;
        } else if (lane == 1 && h != SPL_RANK_NONE) {
            rnk[h] = pair_lookup(T->pair, T->pair_log2, symh, m);
        }
        __syncwarp();
    }
    // surviving known parts, in order
    uint32_t run = 0;
    for (uint32_t w0 = 0; w0 < nwords; w0 += 32) {
        uint32_t wi = w0 + lane;
        uint32_t bits = wi < nwords ? tb[wi] : 0u, keep = 0;
        for (uint32_t mm = bits; mm; mm &= mm - 1) {
            uint32_t b = __ffs(mm) - 1;
            if (sym[wi * 32u + b] < SPL_UNK_BASE) keep |= 1u << b;     // unknown bytes produce no id (bpe.rs:187-191)
        }
        uint32_t c = __popc(keep), incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        uint32_t o = run + incl - c;
        for (uint32_t mm = keep; mm; mm &= mm - 1) out[o++] = sym[wi * 32u + __ffs(mm) - 1];
        run += __shfl_sync(FULL, incl, 31);
    }
    return run;
}

// The whole block merges one piece that does not fit shared memory: text bytes tx[0, len) in global memory; the
// sym / rnk / next / prev arrays live in the scratch pool.  Returns (to every thread) the number of ids, written in
// order to out[0 ..].
__device__ uint32_t bpe_piece_block(BpeSmem& sm, uint32_t* s_bcast, const SplTables* T, const uint8_t* __restrict__ tx, uint32_t len,
                                    uint32_t* scratch, uint32_t* __restrict__ out) {
    const uint32_t tid = threadIdx.x;
    uint32_t* sym = scratch;
    uint32_t* rnk = scratch + len;
    uint32_t* nxt = scratch + 2 * (size_t)len;
    uint32_t* prv = scratch + 3 * (size_t)len;

    if (len <= T->max_key_len) {                     // only for vocabularies with very long keys
        if (tid < 32) {
            uint32_t id = lookupL_warp_g(T, tx, len);
            if (tid == 0) *s_bcast = id;
        }
        __syncthreads();
        uint32_t found = *s_bcast;
        __syncthreads();
        if (found != SPL_RANK_NONE) {
            if (tid == 0) out[0] = found;
            return 1u;
        }
    }
    for (uint32_t j = tid; j < len; j += SPL_BPE_THREADS) {
        sym[j] = T->byte_sym[__ldg(tx + j)];
        nxt[j] = j + 1;                              // len == "no next"
        prv[j] = j ? j - 1 : SPL_RANK_NONE;
    }
    __syncthreads();
    for (uint32_t j = tid; j < len; j += SPL_BPE_THREADS)
        rnk[j] = (j + 1 < len) ? pair_lookup(T->pair, T->pair_log2, sym[j], sym[j + 1]) : SPL_RANK_NONE;
    __syncthreads();
    for (;;) {
        uint64_t best = ~0ull;                       // (rank << 32) | position : min = leftmost minimum
        for (uint32_t j = tid; j < len; j += SPL_BPE_THREADS) {
            uint64_t v = ((uint64_t)rnk[j] << 32) | j;
            if (v < best) best = v;
        }
        sm.red[tid] = best;
        __syncthreads();
        for (uint32_t o = SPL_BPE_THREADS / 2; o; o >>= 1) {
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
    // ordered compaction of the surviving known symbols (unknown single bytes are dropped)
    const uint32_t per = (len + SPL_BPE_THREADS - 1) / SPL_BPE_THREADS;
    const uint32_t lo = tid * per < len ? tid * per : len, hi = lo + per < len ? lo + per : len;
    uint32_t c = 0;
    for (uint32_t j = lo; j < hi; ++j) { uint32_t sv = sym[j]; c += (sv != SPL_RANK_NONE && sv < SPL_UNK_BASE); }
    sm.red[tid] = c;
    __syncthreads();
    if (tid == 0) { uint64_t run = 0; for (int q = 0; q < SPL_BPE_THREADS; ++q) { uint64_t t = sm.red[q]; sm.red[q] = run; run += t; } *s_bcast = (uint32_t)run; }
    __syncthreads();
    uint32_t o = (uint32_t)sm.red[tid];
    for (uint32_t j = lo; j < hi; ++j) { uint32_t sv = sym[j]; if (sv != SPL_RANK_NONE && sv < SPL_UNK_BASE) out[o++] = sv; }
    uint32_t total = *s_bcast;
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(SPL_BPE_THREADS) k_bpe(SplWork w) {
    __shared__ BpeSmem sm;
    __shared__ uint32_t s_bcast, s_off;
    const SplTables* T = w.T;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    // ---- thread class ------------------------------------------------------------------------
    {
        const uint32_t n = w.counters[SPL_CTR_SHORT];
        for (uint32_t i = blockIdx.x * SPL_BPE_THREADS + tid; i < n; i += gridDim.x * SPL_BPE_THREADS) {
            uint64_t e = w.mlist[i];
            uint32_t gpos = (uint32_t)e, len = (uint32_t)(e >> 32) & SPL_ML_LEN_SAT;
            uint32_t c = bpe_piece_thread(sm, T, w.text + gpos, len, w.pool + gpos);
            bpe_finish(w, &w.mlist[i], gpos, c);
        }
    }
    // ---- warp class --------------------------------------------------------------------------
    {
        const uint32_t n = w.counters[SPL_CTR_WARP];
        if (n) {
            __syncthreads();
            BpeWarpSmem& ws = sm.wp[warp];
            for (uint32_t i = blockIdx.x * (SPL_BPE_THREADS / 32) + warp; i < n; i += gridDim.x * (SPL_BPE_THREADS / 32)) {
                uint64_t* slot = &w.mlist[w.ml_r0 - 1 - i];
                uint64_t e = *slot;
                uint32_t gpos = (uint32_t)e, len = (uint32_t)(e >> 32) & SPL_ML_LEN_SAT;
                uint32_t c = bpe_piece_warp(ws.sym, ws.rnk, ws.tb, T, w.text + gpos, len, w.pool + gpos);
                if (lane == 0) bpe_finish(w, slot, gpos, c);
                __syncwarp();
            }
        }
    }
    // ---- big class: warp 0 with the whole block's shared memory ---------------------------------------
    {
        const uint32_t n = w.counters[SPL_CTR_BIG];
        if (n) {
            __syncthreads();
            if (warp == 0)
                for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
                    uint64_t* slot = &w.mlist[w.ml_r0 + i];
                    uint64_t e = *slot;
                    uint32_t gpos = (uint32_t)e, len = (uint32_t)(e >> 32) & SPL_ML_LEN_SAT;
                    uint32_t c = bpe_piece_warp(sm.big.sym, sm.big.rnk, sm.big.tb, T, w.text + gpos, len, w.pool + gpos);
                    if (lane == 0) bpe_finish(w, slot, gpos, c);
                    __syncwarp();
                }
        }
    }
    // ---- huge class: whole block, global scratch ----------------------------------------------------
    {
        const uint32_t n = w.counters[SPL_CTR_HUGE];
        if (n) {
            __syncthreads();
            for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
                uint64_t* slot = &w.mlist[w.ml_r1 + i];
                uint64_t e = *slot;
                uint32_t gpos = (uint32_t)e;
                if (tid == 0) {
                    uint32_t ge = g_next_bit(w.pstart, gpos + 1, w.N + 1);
                    uint32_t need = 4u * (ge - gpos);
                    uint32_t off = atomicAdd(&w.counters[SPL_CTR_HUGE_POOL], need);
                    if ((uint64_t)off + need > w.huge_pool_words) { atomicOr(&w.counters[SPL_CTR_ERR], SPL_DEVERR_HUGE_POOL); off = SPL_RANK_NONE; }
                    s_off = off; s_bcast = ge - gpos;
                }
                __syncthreads();
                const uint32_t off = s_off, len = s_bcast;
                __syncthreads();
                uint32_t c = 0;
                if (off != SPL_RANK_NONE) c = bpe_piece_block(sm, &s_bcast, T, w.text + gpos, len, w.huge_pool + off, w.pool + gpos);
                if (tid == 0) bpe_finish(w, slot, gpos, c);
                __syncthreads();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_chunk_scan: exclusive prefix of the id counts of the chunks (SPL_CHUNK_TILES tiles each; k_probe and k_bpe keep
// the chunk totals up to date).  One block: chunks are few (N / 128 KiB); k_emit adds the tiles inside a chunk itself.
// ------------------------------------------------------------------------------------------
#define TS_PER 4u                                    // chunks per thread and round
__global__ void __launch_bounds__(1024) k_chunk_scan(SplWork w) {
    __shared__ uint32_t s_cnt[1024 * TS_PER];
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n = (w.n_tiles + SPL_CHUNK_TILES - 1) / SPL_CHUNK_TILES;
    if (tid == 0) s_carry = 0;
    for (uint32_t c0 = 0; c0 < n; c0 += 1024 * TS_PER) {
        // coalesced, independent loads; then every thread owns TS_PER consecutive chunks
#pragma unroll
        for (uint32_t q = 0; q < TS_PER; ++q) {
            uint32_t i = c0 + q * 1024 + tid;
            s_cnt[q * 1024 + tid] = i < n ? (uint32_t)w.chunk_cnt[i] : 0u;
        }
        __syncthreads();
        uint32_t loc[TS_PER];
        uint64_t v = 0;
#pragma unroll
        for (uint32_t q = 0; q < TS_PER; ++q) { loc[q] = s_cnt[tid * TS_PER + q]; v += loc[q]; }
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
        uint64_t run = s_carry + s_w[warp] + incl - v;
#pragma unroll
        for (uint32_t q = 0; q < TS_PER; ++q) {
            uint32_t i = c0 + tid * TS_PER + q;
            if (i < n) w.chunk_state[i] = run;
            run += loc[q];
        }
        __syncthreads();
        if (tid == 1023) s_carry = run;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// k_emit: one tile per block: pv[] (+ pool[] for merged pieces) -> ids in document order, output offsets.
// Four pieces per thread and round (one 16-byte load of pv), so a warp owns 128 consecutive pieces.
// ------------------------------------------------------------------------------------------
#define EM_ROUNDS (SPL_TILE / (SPL_THREADS * 4))     // 4 rounds cover the 4096 pieces a tile can have
#define EM_WARPS (SPL_THREADS / 32)
#define EM_INLINE 8u                                 // ids of a merged piece copied by its own thread up to this many
#define EM_BIGCAP (SPL_TILE / (EM_INLINE + 1u) + 1u) // pieces of a tile that can have more ids than that

struct EmitSmem {
    __align__(16) uint32_t spos[SPL_TILE + 4];       // ids of the warp's round before piece j (see wtot)
    uint32_t pbw[SPL_TILE / 32];
    uint32_t wpre[SPL_TILE / 32];
    uint32_t wtot[EM_ROUNDS * EM_WARPS];             // ids of the tile before each (round, warp)
    uint32_t bigpos[EM_BIGCAP], biggp[EM_BIGCAP], bigcnt[EM_BIGCAP];
    uint32_t n_big, total;
    uint64_t prefix;
};

__device__ __forceinline__ uint32_t emit_count(const SplWork& w, uint32_t v, bool valid) {
    if (!valid) return 0u;
    if (v < SPL_PV_MISS) return 1u;
    if (v == SPL_PV_NONE) return 0u;
    return (uint32_t)(w.mlist[v & ~SPL_PV_MISS] >> 32);
}

__global__ void __launch_bounds__(SPL_THREADS, 8) k_emit(SplWork w) {
    __shared__ EmitSmem sm;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile = blockIdx.x, tile0 = tile * SPL_TILE;
    const uint32_t* __restrict__ pv = w.pv + tile0;

    // everything the block needs to know comes from one round of independent loads
    const uint4 ti = __ldg(reinterpret_cast<const uint4*>(w.tinfo + tile));          // {np, extra, first_doc, -}
    const uint32_t d1 = __ldg(&w.tinfo[tile + 1].first_doc);
    uint4 v_first = make_uint4(SPL_PV_NONE, SPL_PV_NONE, SPL_PV_NONE, SPL_PV_NONE);
    if (tid < 128) v_first = __ldg(reinterpret_cast<const uint4*>(pv + tid * 4u));   // the first 512 pieces: every ordinary tile has them
    const uint32_t pbw_mine = tid < SPL_TILE / 32 ? __ldg(w.pstart + (tile0 >> 5) + tid) : 0u;
    if (warp == EM_WARPS - 1) {
        // ids before this tile: the chunk's prefix plus the tiles of the chunk in front of this one
        const uint32_t t = (tile & ~(SPL_CHUNK_TILES - 1u)) + lane;
        uint32_t c = 0;
        if (t < tile) { const uint4 x = __ldg(reinterpret_cast<const uint4*>(w.tinfo + t)); c = x.x + x.y; }
        c = __reduce_add_sync(FULL, c);
        if (lane == 0) sm.prefix = w.chunk_state[tile / SPL_CHUNK_TILES] + c;
    }
    const uint32_t P = ti.x, d0 = ti.z;
    const uint32_t rounds = (P + SPL_THREADS * 4 - 1) / (SPL_THREADS * 4);
    if (tid == 0) sm.n_big = 0;
    if (tid < SPL_TILE / 32) sm.pbw[tid] = pbw_mine;

    // ---- pass 1: id count of every piece, warp-level prefixes ---------------------------------------
    for (uint32_t k = 0; k < rounds; ++k) {
        const uint32_t j4 = (k * SPL_THREADS + tid) * 4u;
        uint4 v = v_first;
        if ((k > 0 || tid >= 128) && j4 < P) v = __ldg(reinterpret_cast<const uint4*>(pv + j4));
        const uint32_t c0 = emit_count(w, v.x, j4 < P), c1 = emit_count(w, v.y, j4 + 1 < P),
                       c2 = emit_count(w, v.z, j4 + 2 < P), c3 = emit_count(w, v.w, j4 + 3 < P);
        const uint32_t c = c0 + c1 + c2 + c3;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        const uint32_t ex = incl - c;
        *reinterpret_cast<uint4*>(&sm.spos[j4]) = make_uint4(ex, ex + c0, ex + c0 + c1, ex + c0 + c1 + c2);
        if (lane == 31) sm.wtot[k * EM_WARPS + warp] = incl;
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive scan of the rounds * EM_WARPS (<= 32) warp totals; word prefixes of the piece bits
        uint32_t x = lane < rounds * EM_WARPS ? sm.wtot[lane] : 0u, incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        sm.wtot[lane] = incl - x;
        if (lane == 31) sm.total = incl;                       // ids of the whole tile
        if (d1 > d0) {
            uint32_t loc[4], run = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) { loc[q] = __popc(sm.pbw[lane * 4 + q]); run += loc[q]; }
            incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(FULL, incl, o);
                if (lane >= (uint32_t)o) incl += t;
            }
            uint32_t b = incl - run;
#pragma unroll
            for (int q = 0; q < 4; ++q) { sm.wpre[lane * 4 + q] = b; b += loc[q]; }
        }
    }
    __syncthreads();

    // ---- pass 2: ids to their place ----------------------------------------------------------------------
    const uint64_t prefix = sm.prefix;
    uint32_t* __restrict__ out = w.ids + prefix;
    for (uint32_t k = 0; k < rounds; ++k) {
        const uint32_t j4 = (k * SPL_THREADS + tid) * 4u;
        if (j4 < P) {
            uint4 v = v_first;
            if (k > 0 || tid >= 128) v = __ldg(reinterpret_cast<const uint4*>(pv + j4));
            uint32_t pos = sm.spos[j4] + sm.wtot[k * EM_WARPS + warp];
            const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) {
                if (j4 + q < P) {
                    const uint32_t x = vv[q];
                    if (x < SPL_PV_MISS) out[pos++] = x;
                    else if (x != SPL_PV_NONE) {
                        const uint64_t e = w.mlist[x & ~SPL_PV_MISS];
                        const uint32_t gp = (uint32_t)e, c = (uint32_t)(e >> 32);
                        if (c <= EM_INLINE) {
                            for (uint32_t r = 0; r < c; ++r) out[pos + r] = w.pool[gp + r];
                        } else {
                            uint32_t b = atomicAdd(&sm.n_big, 1u);
                            sm.bigpos[b] = pos; sm.biggp[b] = gp; sm.bigcnt[b] = c;
                        }
                        pos += c;
                    }
                }
            }
        }
    }
    if (d1 > d0 || P > 0) __syncthreads();
    {
        const uint32_t nb = sm.n_big;
        for (uint32_t b = 0; b < nb; ++b) {
            const uint32_t pos = sm.bigpos[b], gp = sm.biggp[b], c = sm.bigcnt[b];
            for (uint32_t q = tid; q < c; q += SPL_THREADS) out[pos + q] = w.pool[gp + q];
        }
    }

    // ---- output offset of every document that starts in this tile ----------------------------------------
    for (uint32_t d = d0 + tid; d < d1; d += SPL_THREADS) {
        const uint32_t x = (uint32_t)(w.doc_off[d] - w.off_base - tile0);
        const uint32_t pi = sm.wpre[x >> 5] + __popc(sm.pbw[x >> 5] & ((1u << (x & 31)) - 1u));
        const uint32_t rel = pi >= P ? sm.total : sm.spos[pi] + sm.wtot[(pi / (SPL_THREADS * 4u)) * EM_WARPS + ((pi / 128u) & (EM_WARPS - 1u))];
        w.out_off[d] = prefix + rel;
        if (d == w.n_docs && w.host_meta) {                    // the call's summary, straight to the host (no copy on this stream)
            w.host_meta[0] = prefix + rel;
            w.host_meta[1] = (uint64_t)w.counters[SPL_CTR_ERR] | ((uint64_t)w.counters[SPL_CTR_HUGE_POOL] << 32);
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
void spl_encode_init() {
    cudaFuncSetAttribute(k_probe, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
    cudaFuncSetAttribute(k_emit, cudaFuncAttributePreferredSharedMemoryCarveout, 75);
    cudaGetLastError();
}

void spl_launch_encode_stage(const SplWork& w, int num_sms, cudaStream_t stream, SplMarkFn mark, void* ctx) {
    k_probe<<<w.n_tiles, SPL_THREADS, 0, stream>>>(w);
    mark(ctx, "k_probe");
    k_bpe<<<(uint32_t)num_sms * 6u, SPL_BPE_THREADS, 0, stream>>>(w);
    mark(ctx, "k_bpe");
    k_chunk_scan<<<1, 1024, 0, stream>>>(w);
    mark(ctx, "k_chunk_scan");
    k_emit<<<w.n_tiles, SPL_THREADS, 0, stream>>>(w);
    mark(ctx, "k_emit");
}
