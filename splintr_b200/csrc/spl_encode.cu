// Device encode path, second half: piece-start bitmap -> token ids in document order.
//
//   k_probe      one thread per piece: whole-piece vocabulary probe (tokenizer.rs:703-705, bpe.rs:73-80); the id of a
//                hit goes to pv[] in piece order, a miss is appended to the miss list of its length class
//   k_bpe        the leftmost-min-rank merge loop of bpe.rs:83-194 for every listed piece -- a group of 1..32 lanes per
//                piece by length class, 32 / group pieces side by side in a warp -- ids to pool[] at the piece's byte position
//   (chunk scan) exclusive prefix of the id counts of 32-tile chunks, by the last block of k_bpe_long (k_emit adds the
//                tiles inside its chunk)
//   k_emit       pv[] + pool[] -> ids in document order (the collect of tokenizer.rs:806 and the Rayon collect of
//                encode_batch, tokenizer.rs:932-934) and the per-document output offsets
//
// No block waits for another block inside a kernel; the miss lists decouple the rare, latency-bound merge loop from
// the streaming probe so that both run at full occupancy.  Integer / byte work; no tensor cores.
#include <algorithm>
#include "spl_device.cuh"
#include "spl_segment.h"
#include "spl_bpe_bits.h"

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

__device__ __forceinline__ uint64_t ml_entry(uint32_t gpos, uint32_t len) {
    return (uint64_t)gpos | ((uint64_t)(len & SPL_ML_LEN_MASK) << 32);
}

// ------------------------------------------------------------------------------------------
// probe_tile: the whole-piece probe of one 4 KiB tile whose text and piece-start bits are staged in shared memory.
// Called by all threads of the block (barriers inside); threads beyond SPL_THREADS only take part in the barriers.
// k_probe stages the tile from global memory (text + the piece-start bitmap the pre-tokenizer wrote).
// ------------------------------------------------------------------------------------------
#define PB_WORDS (SPL_PROBE_WIN / 32u + 1u)          // piece-start words staged: bits 0 .. SPL_PROBE_WIN + 31
#define PB_BITS  (PB_WORDS * 32u)
#define PROBE_WARPS (SPL_THREADS / 32)

// Piece list of a tile: sm.plist[0 .. P) = window positions of the piece starts in pb (bits below SPL_TILE), in order,
// sm.plist[P] = end of the last piece (0xFFFF: beyond the staged bits, see sm.last_end).  All threads of the block call
// this (two barriers inside).  publish: record P for k_emit and add it to the chunk total.
__device__ __forceinline__ uint32_t build_piece_list(const SplWork& w, SplProbeScratch& sm, const uint32_t* pb, const uint32_t tile,
                                                     const bool publish, const uint32_t my_hi = 0u, uint32_t* n_hi = nullptr) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t N = w.N, tile0 = tile * SPL_TILE;
    const uint32_t avail = N - tile0;                          // text bytes from tile0 on
    uint32_t my = (pb[tid >> 1] >> ((tid & 1u) * 16u)) & 0xFFFFu;
    if (tid * 16u + 16u > avail) my &= (tid * 16u >= avail) ? 0u : ((1u << (avail - tid * 16u)) - 1u);   // sentinel bit at N
    uint32_t cnt = __popc(my), incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) sm.wtot[warp] = incl;
    if (n_hi) {                                                // (block-uniform) bytes >= 0x80 of the tile, summed alongside
        const uint32_t h = __reduce_add_sync(FULL, my_hi);
        if (lane == 0) sm.whi[warp] = h;
    }
    __syncthreads();
    uint32_t base = incl - cnt, P = 0, H = 0;
#pragma unroll
    for (uint32_t q = 0; q < PROBE_WARPS; ++q) {
        uint32_t t = sm.wtot[q];
        base += (q < warp) ? t : 0u;
        P += t;
        if (n_hi) H += sm.whi[q];
    }
    if (n_hi) *n_hi = H;
    while (my) {
        uint32_t b = __ffs(my) - 1;
        my &= my - 1;
        sm.plist[base++] = (uint16_t)(tid * 16u + b);
    }
    if (tid == 0) {
        uint32_t e = sm_next_bit(pb, SPL_TILE < avail ? SPL_TILE : avail, PB_BITS);
        if (P && e >= PB_BITS) {                                   // the last piece leaves the staged bits
            e = g_next_bit(w.pstart, tile0 + PB_BITS, N + 1) - tile0;
        }
        sm.last_end = e;
        sm.plist[P] = (uint16_t)(e >= PB_BITS ? 0xFFFFu : e);      // 0xFFFF: see last_end
        if (publish) {
            w.tinfo[tile].np = P;
            if (P) atomicAdd(&w.chunk_cnt[tile / SPL_CHUNK_TILES], (int32_t)P);
        }
    }
    __syncthreads();
    return P;
}

__device__ __forceinline__ void probe_tile(const SplWork& w, SplProbeScratch& sm, const uint32_t* text, const uint32_t* pb,
                                           const uint32_t tile, const uint32_t P) {
    const SplTables* T = w.T;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t N = w.N;
    const uint32_t tile0 = tile * SPL_TILE;
    const SplKey8* __restrict__ t8 = T->t8;
    const uint32_t t8_log2 = T->t8_log2;

    // (sm.spw, the special-span bits of the tile, is filled by the caller together with text and pb; so is the piece list)
    if (tid == 0) {
        w.tinfo[tile].np = P;
        if (P) atomicAdd(&w.chunk_cnt[tile / SPL_CHUNK_TILES], (int32_t)P);
    }
    // From here on every warp works alone on its own range of pieces [jlo, jhi): no block barrier, a warp with
    // slow pieces does not hold the others back.
    const uint32_t per = (((P + PROBE_WARPS - 1) / PROBE_WARPS) + 31u) & ~31u;
    const uint32_t jlo = warp * per < P ? warp * per : P, jhi = jlo + per < P ? jlo + per : P;
    const uint32_t pvbase = tile * SPL_TILE;

    // ---- hot loop, one thread per piece: pieces of up to 8 bytes against the FIRST slot of their home bucket ------
    // A scattered 16-byte load costs one L1 wavefront per lane, and that is what bounds this loop, so only one load
    // is spent here: keys were inserted in rank order, the frequent tokens sit in the first slot.  Everything else
    // (second slot, next buckets, longer pieces, specials) goes to the slow list.
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
            const uint32_t a = text[wi], b = text[wi + 1], c = text[wi + 2];
            uint32_t lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);
            const uint32_t nb = len * 8u;
            lo &= nb >= 32u ? FULL : ((1u << nb) - 1u);
            hi &= nb <= 32u ? 0u : (nb >= 64u ? FULL : ((1u << (nb - 32u)) - 1u));
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(t8 + (size_t)spl_hash8(lo, hi, len, t8_log2) * SPL_T8_WAYS));
            found = v0.w == len && v0.x == lo && v0.y == hi;            // {k0 lo, k0 hi, id, len}
            if (found) w.pv[pvbase + j] = v0.z;
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
        uint32_t cls = 0, j = 0, mlen = 0, mpos = 0;           // cls: 0 settled, else 1 + length class of the miss
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
                uint32_t sy = T->byte_sym[sm_byte(text, s)];
                val = sy < SPL_UNK_BASE ? sy : SPL_PV_NONE;   // unknown byte: no id (bpe.rs:73-75)
            } else {
                uint32_t id = SPL_RANK_NONE;
                if (len <= 8) {
                    uint64_t k0 = sm_load8(text, s);
                    if (len < 8) k0 &= (1ull << (8 * len)) - 1;
                    id = lookup8(t8, t8_log2, k0, len);
                } else if (len <= 16) {
                    uint64_t k0 = sm_load8(text, s), k1 = sm_load8(text, s + 8);
                    if (len < 16) k1 &= (1ull << (8 * (len - 8))) - 1;
                    id = lookup16(T->t16, T->t16_log2, k0, k1, len);
                } else if (len <= SPL_PROBE_HALO && len <= T->max_key_len) {
                    id = lookupL_thread(T, text, s, len);
                }
                if (id != SPL_RANK_NONE) val = id;
                else cls = 1u + spl_len_class(len);
            }
            if (cls == 0) {
                w.pv[pvbase + j] = val;
                if (val == SPL_PV_NONE) { atomicAdd(&w.tinfo[tile].extra, -1); atomicAdd(&w.chunk_cnt[tile / SPL_CHUNK_TILES], -1); }
            }
            mlen = len; mpos = gpos;
        }
        // classes 0 and 1 (up to 64 bytes) are collected per warp and published below with one atomic each
        uint32_t bal = __ballot_sync(FULL, cls == 1);
        if (cls == 1) sm.mloc[jlo + n_short + __popc(bal & lt_mask)] = (uint16_t)j;
        n_short += __popc(bal);
        bal = __ballot_sync(FULL, cls == 2);
        if (cls == 2) sm.mloc[jhi - 1 - (n_warp + __popc(bal & lt_mask))] = (uint16_t)j;
        n_warp += __popc(bal);
        // longer pieces: straight to the list of their class, one atomic per warp and class
        if (__any_sync(FULL, cls > 2)) {
            for (uint32_t c = 2; c < SPL_NCLS; ++c) {
                bal = __ballot_sync(FULL, cls == c + 1);
                if (!bal) continue;
                uint32_t leader = __ffs(bal) - 1, b0 = 0;
                if (lane == leader) b0 = atomicAdd(&w.counters[SPL_CTR_CLS + c], (uint32_t)__popc(bal));
                b0 = __shfl_sync(FULL, b0, leader);
                if (cls == c + 1) {
                    const uint32_t midx = w.ml_base[c] + b0 + __popc(bal & lt_mask);
                    w.mlist[midx] = ml_entry(mpos, mlen);
                    w.pv[pvbase + j] = SPL_PV_MISS | midx;
                }
            }
        }
    }
    __syncwarp();

    // ---- publish the warp's misses ------------------------------------------------------------
    if (n_short | n_warp) {
        uint32_t g_short = 0, g_warp = 0;
        if (lane == 0) {
            if (n_short) g_short = atomicAdd(&w.counters[SPL_CTR_CLS + 0], n_short);
            if (n_warp) g_warp = atomicAdd(&w.counters[SPL_CTR_CLS + 1], n_warp);
        }
        g_short = __shfl_sync(FULL, g_short, 0);
        g_warp = __shfl_sync(FULL, g_warp, 0);
        for (uint32_t i = lane; i < n_short; i += 32) {
            uint32_t j = sm.mloc[jlo + i], s = sm.plist[j], e = sm.plist[j + 1], midx = w.ml_base[0] + g_short + i;
            if (e == 0xFFFFu) e = sm.last_end;
            w.mlist[midx] = ml_entry(tile0 + s, e - s);
            w.pv[pvbase + j] = SPL_PV_MISS | midx;
        }
        for (uint32_t i = lane; i < n_warp; i += 32) {
            uint32_t j = sm.mloc[jhi - 1 - i], s = sm.plist[j], e = sm.plist[j + 1], midx = w.ml_base[1] + g_warp + i;
            if (e == 0xFFFFu) e = sm.last_end;
            w.mlist[midx] = ml_entry(tile0 + s, e - s);
            w.pv[pvbase + j] = SPL_PV_MISS | midx;
        }
    }
}

// ------------------------------------------------------------------------------------------
// probe_tile_refine: the probe of a tile with many multi-byte characters (CJK text).  On such text most pieces miss the
// whole-piece probe and would go through the merge loop as they are -- 24 bytes on average, 20 merges each.  Instead
// every missed piece is cut at its safe boundaries (spl_segment.h: no vocabulary key can lie across one, so the merge
// loop runs per segment and the id lists concatenate); the segments become pieces of their own:
//   pass 1   whole-piece probe of every piece (results by start position in val_at, misses marked in mstw)
//   pass R   one thread per missed piece walks its characters and marks the safe boundaries inside the tile in segw
//   pass 2   piece list over the old and the new starts; a segment that is one 2- or 3-byte character is ONE table
//            load (char_tok: what the merge loop makes of that character), any other segment is filed in the miss
//            list of its length class with SPL_ML_SEG; the piece-start bitmap in global memory gets the new bits
//            (k_emit counts pieces with it)
// Only boundaries inside the tile are looked for (the tile that owns a piece owns its segments; what lies beyond
// stays one segment), and a piece that is a candidate for the long-key probe of the merge kernels (longer than the
// probe halo, not longer than the longest key) is left whole.
// ------------------------------------------------------------------------------------------
#define PV_MISSMARK 0xFFFFFFFEu

struct SmemPieceReader {
    const uint32_t* text; uint32_t s;
    __device__ __forceinline__ uint32_t load4(uint32_t i) const {
        const uint32_t q = s + i, wi = q >> 2;
        return __funnelshift_r(text[wi], text[wi + 1], (q & 3u) * 8u);
    }
};

// whole-piece probe of the piece [s, s + len) of the staged window; returns its pv value or PV_MISSMARK
__device__ __forceinline__ uint32_t probe_whole(const SplWork& w, const SplTables* T, const SplProbeScratch& sm,
                                                const uint32_t* text, uint32_t s, uint32_t len, uint32_t gpos) {
    if (w.with_special && ((sm.spw[s >> 5] >> (s & 31)) & 1u)) return special_id_g(T, w.text + gpos, len);
    if (len == 1) {
        const uint32_t sy = T->byte_sym[sm_byte(text, s)];
        return sy < SPL_UNK_BASE ? sy : SPL_PV_NONE;                // unknown byte: no id (bpe.rs:73-75)
    }
    uint32_t id = SPL_RANK_NONE;
    if (len <= 8) {
        uint64_t k0 = sm_load8(text, s);
        if (len < 8) k0 &= (1ull << (8 * len)) - 1;
        id = lookup8(T->t8, T->t8_log2, k0, len);
    } else if (len <= 16) {
        uint64_t k0 = sm_load8(text, s), k1 = sm_load8(text, s + 8);
        if (len < 16) k1 &= (1ull << (8 * (len - 8))) - 1;
        id = lookup16(T->t16, T->t16_log2, k0, k1, len);
    } else if (len <= SPL_PROBE_HALO && len <= T->max_key_len) {
        id = lookupL_thread(T, text, s, len);
    }
    return id != SPL_RANK_NONE ? id : PV_MISSMARK;
}

__device__ __forceinline__ void probe_tile_refine(const SplWork& w, SplProbeScratch& sm, const uint32_t* text, uint32_t* pb,
                                                  const uint32_t tile, const uint32_t P1) {
    const SplTables* T = w.T;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t tile0 = tile * SPL_TILE, pvbase = tile * SPL_TILE;
    uint32_t* val_at = reinterpret_cast<uint32_t*>(sm.slow);           // [SPL_TILE] (slow + mloc: unused on this path)
    static_assert(sizeof(sm.slow) + sizeof(sm.mloc) >= SPL_TILE * 4u && offsetof(SplProbeScratch, mloc) == offsetof(SplProbeScratch, slow) + sizeof(sm.slow), "val_at");
    if (tid < SPL_TILE / 32) { sm.segw[tid] = 0u; sm.mstw[tid] = 0u; }
    if (tid == 0) atomicAdd(&w.counters[SPL_CTR_REFINED], 1u);
    __syncthreads();

    // ---- pass 1 + pass R: one thread per piece ------------------------------------------------------------
    for (uint32_t j = tid; j < P1; j += SPL_THREADS) {
        const uint32_t s = sm.plist[j];
        uint32_t e = sm.plist[j + 1];
        if (e == 0xFFFFu) e = sm.last_end;
        const uint32_t len = e - s;
        const uint32_t v = probe_whole(w, T, sm, text, s, len, tile0 + s);
        val_at[s] = v;
        if (v != PV_MISSMARK) continue;
        if (len > SPL_PROBE_HALO && len <= T->max_key_len) continue;      // the merge kernels still owe it the long-key probe: stays whole
        atomicOr(&sm.mstw[s >> 5], 1u << (s & 31u));
    }
    __syncthreads();

    // ---- pass R: safe boundaries inside the missed pieces, one thread per 16 BYTES of the tile (a thread per piece
    // walked its characters one dependent table load after the other while the block waited for the longest piece)
    // R1: which bytes belong to a missed piece?  Word k: a piece start switches membership on (missed) or off.
    if (tid < SPL_TILE / 32) {
        const uint32_t S = pb[tid], M = sm.mstw[tid];
        const bool has = S != 0u;
        const uint32_t last = has ? (M >> (31u - (uint32_t)__clz(S))) & 1u : 0u;           // membership behind the word's last start
        const uint32_t hmask = __ballot_sync(FULL, has), lmask = __ballot_sync(FULL, last != 0u);
        if (lane == 0) sm.wtot[tid >> 5] = hmask ? (2u | ((lmask >> (31u - (uint32_t)__clz(hmask))) & 1u)) : 0u;
        const uint32_t below = hmask & lt_mask;
        uint32_t carry = below ? (lmask >> (31u - (uint32_t)__clz(below))) & 1u : 2u;         // 2: ask the warps in front
        // (the four warps of this branch meet at a named barrier: the other four are not here)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (carry == 2u) {
            carry = 0u;                                            // a piece that started in an earlier tile is not ours
            for (int q = (int)(tid >> 5) - 1; q >= 0; --q)
                if (sm.wtot[q]) { carry = sm.wtot[q] & 1u; break; }
        }
        uint32_t mask = 0, pos = 0, cur = carry;
        for (uint32_t m = S; m; m &= m - 1u) {
            const uint32_t b = (uint32_t)__ffs(m) - 1u;
            if (cur) mask |= ((1u << b) - 1u) & ~((1u << pos) - 1u);
            cur = (M >> b) & 1u; pos = b;
        }
        if (cur) mask |= ~((1u << pos) - 1u);
        sm.inmw[tid] = mask;
    }
    __syncthreads();
    // R2: my 16 bytes
    {
        const uint32_t sh = (tid & 1u) * 16u;
        const uint32_t in16 = (sm.inmw[tid >> 1] >> sh) & 0xFFFFu, st16 = (pb[tid >> 1] >> sh) & 0xFFFFu;
        uint32_t cand = in16 & ~st16;
        if (cand) {
            // only bytes that start a character: not 10xxxxxx (four bytes per word -> four mask bits)
            const uint4 x = reinterpret_cast<const uint4*>(text)[tid];
            const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
            uint32_t lead = 0;
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) {
                const uint32_t t = (xs[k] & 0xC0C0C0C0u) ^ 0x80808080u;          // a byte of t is zero iff the text byte is a continuation byte
                const uint32_t y = ((t | (t << 1)) >> 7) & 0x01010101u;
                lead |= ((y & 1u) | ((y >> 7) & 2u) | ((y >> 14) & 4u) | ((y >> 21) & 8u)) << (4u * k);
            }
            cand &= lead;
        }
        const SmemPieceReader rd{text, 0u};
        const uint32_t avail_tile = w.N - tile0;                    // text bytes from the tile start on
        uint32_t found = 0;
        for (uint32_t m = cand; m; m &= m - 1u) {
            const uint32_t bq = (uint32_t)__ffs(m) - 1u, q = tid * 16u + bq;
            if (spl_boundary_safe_at(rd, q, 0u, min(4u, avail_tile - q), T->seg_irr, T->seg_h2, T->seg_h2_log2)) found |= 1u << bq;
        }
        if (found) atomicOr(&sm.segw[tid >> 1], found << sh);
    }
    __syncthreads();

    // ---- pass 2: the new piece list ----------------------------------------------------------------------------
    if (tid < SPL_TILE / 32) {
        const uint32_t nw = sm.segw[tid];
        if (nw) { pb[tid] |= nw; w.pstart[(tile0 >> 5) + tid] = pb[tid]; }     // only this block writes the words of its tile
    }
    __syncthreads();
    const uint32_t P = build_piece_list(w, sm, pb, tile, true);
    // 2a: settle what can be settled; count the misses of the block per length class (shared counters: the global
    // class counters are ONE address each for the whole grid -- one atomic per warp was 48 per tile and the kernel's
    // whole cost on CJK text).  "Class" SPL_NCLS: characters that are two or three ids (char_tok): settled here too,
    // but more than one id needs a miss-list entry and room in the pool.
    uint32_t* s_cnt = sm.cls_cnt;                                     // [SPL_NCLS + 1] block counts, then global bases
    if (tid <= SPL_NCLS) s_cnt[tid] = 0u;
    __syncthreads();
    int32_t extra = 0;                                                // ids minus pieces of what this thread settles
    for (uint32_t j0 = 0; j0 < P; j0 += SPL_THREADS) {
        const uint32_t j = j0 + tid;
        uint32_t cls = 0, s = 0;                                      // cls: 0 settled, else 1 + class of the entry to file
        if (j < P) {
            s = sm.plist[j];
            uint32_t e = sm.plist[j + 1];
            if (e == 0xFFFFu) e = sm.last_end;
            const uint32_t len = e - s;
            const bool seg = ((sm.segw[s >> 5] | sm.mstw[s >> 5]) >> (s & 31u)) & 1u;
            uint32_t val;
            if (!seg) {
                val = val_at[s];
            } else if (len == 1) {
                const uint32_t sy = T->byte_sym[sm_byte(text, s)];
                val = sy < SPL_UNK_BASE ? sy : SPL_PV_NONE;
            } else {
                val = PV_MISSMARK;
                if (len <= 3) {                                       // one 2- or 3-byte character?
                    const SmemPieceReader rd{text, s};
                    uint32_t packed = 0;
                    if (spl_u8_char(rd.load4(0), len, packed) == len) {
                        const uint32_t ct = __ldg(T->char_tok + spl_u8_cp23(packed, len));
                        if (ct != SPL_RANK_NONE) {
                            const uint32_t more = ct >> SPL_CHAR_COUNT_SHIFT;             // ids - 1
                            if (more == 0u) val = ct;                                     // one id
                            else if (w.charref) {                                         // two or three: k_emit reads them from the table
                                val = SPL_PV_CHARREF | (more << 28) | (ct & SPL_CHAR_VALUE_MASK);
                                extra += (int32_t)more;
                            } else cls = 1u + SPL_NCLS;                                   // ... or 2b writes them to the pool
                        }
                    }
                }
            }
            if (cls == 0u) {
                if (val == PV_MISSMARK) cls = 1u + spl_len_class(len);
                else {
                    w.pv[pvbase + j] = val;
                    if (val == SPL_PV_NONE) --extra;
                    val_at[s] = 0xFFFFFFFFu;                          // settled (2b files what is not)
                }
            }
        }
        const uint32_t anyb = __ballot_sync(FULL, cls != 0u);
        if (anyb) {
            const uint32_t cmax = __ballot_sync(FULL, cls > 1u) ? SPL_NCLS : 0u;          // usually only the shortest class is there
            for (uint32_t c = 0; c <= cmax; ++c) {
                const uint32_t bal = __ballot_sync(FULL, cls == c + 1u);
                if (!bal) continue;
                const uint32_t leader = __ffs(bal) - 1;
                uint32_t b0 = 0;
                if (lane == leader) b0 = atomicAdd(&s_cnt[c], (uint32_t)__popc(bal));
                b0 = __shfl_sync(FULL, b0, leader);
                if (cls == c + 1u) val_at[s] = (c << 16) | (b0 + __popc(bal & lt_mask));       // class | index within the block
            }
        }
    }
    __syncthreads();
    if (tid <= SPL_NCLS) {
        const uint32_t nb = s_cnt[tid];
        s_cnt[tid] = nb ? atomicAdd(&w.counters[tid < SPL_NCLS ? SPL_CTR_CLS + tid : SPL_CTR_FINAL], nb) : 0u;
    }
    __syncthreads();
    // 2b: file the entries
    for (uint32_t j = tid; j < P; j += SPL_THREADS) {
        const uint32_t s = sm.plist[j];
        const uint32_t code = val_at[s];
        if (code == 0xFFFFFFFFu) continue;
        uint32_t e = sm.plist[j + 1];
        if (e == 0xFFFFu) e = sm.last_end;
        const uint32_t c = code >> 16, gpos = tile0 + s;
        if (c < SPL_NCLS) {
            const bool seg = ((sm.segw[s >> 5] | sm.mstw[s >> 5]) >> (s & 31u)) & 1u;
            const uint32_t midx = w.ml_base[c] + s_cnt[c] + (code & 0xFFFFu);
            w.mlist[midx] = ml_entry(gpos, e - s) | (seg ? SPL_ML_SEG : 0ull);
            w.pv[pvbase + j] = SPL_PV_MISS | midx;
        } else {
            // a character of two or three ids: ids to the pool, a settled entry from the top of class 0's region down
            // (pending and settled entries of the region are disjoint pieces of two bytes or more: they fit together)
            const SmemPieceReader rd{text, s};
            uint32_t packed = 0;
            const uint32_t len = e - s;
            spl_u8_char(rd.load4(0), len, packed);
            const uint32_t ct = __ldg(T->char_tok + spl_u8_cp23(packed, len));
            const uint32_t n_ids = (ct >> SPL_CHAR_COUNT_SHIFT) + 1u;
            const uint32_t* __restrict__ src = T->char_ids + (ct & SPL_CHAR_VALUE_MASK);
            for (uint32_t q = 0; q < n_ids; ++q) w.pool[gpos + q] = __ldg(src + q);
            const uint32_t midx = w.ml_base[1] - 1u - (s_cnt[SPL_NCLS] + (code & 0xFFFFu));
            w.mlist[midx] = (uint64_t)gpos | ((uint64_t)n_ids << 32) | SPL_ML_DONE | SPL_ML_SEG;
            w.pv[pvbase + j] = SPL_PV_MISS | midx;
            extra += (int32_t)n_ids - 1;
        }
    }
    extra = __reduce_add_sync(FULL, extra);
    if (lane == 0 && extra) { atomicAdd(&w.tinfo[tile].extra, extra); atomicAdd(&w.chunk_cnt[tile / SPL_CHUNK_TILES], extra); }
}

struct __align__(16) ProbeSmem {
    uint32_t text[SPL_PROBE_WIN / 4 + 4];     // staged bytes (+ slack for the unaligned 8-byte key loads)
    uint32_t pb[PB_WORDS + 3];                // piece-start bits (a multiple of 16 bytes: bulk copies move 16-byte units)
    SplProbeScratch ps;
    unsigned long long mbar;                  // arrival barrier of the bulk copies
};
#define PROBE_TEXT_BYTES (SPL_PROBE_WIN + 16u)
#define PROBE_PB_BYTES   ((PB_WORDS + 3u) / 4u * 16u)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Stage one tile (text window + piece-start bits) from global memory.
// BULK: two 1-D bulk copies (cp.async.bulk -> SASS UBLKCP, the TMA engine of sm_90+/sm_100) issued by one thread and
// awaited on an mbarrier -- no registers, no LDG/STS pairs in the 256 threads; else 16-byte loads through registers.
// Returns the number of bytes >= 0x80 among the thread's 16 bytes of the tile.
template <bool BULK>
__device__ __forceinline__ uint32_t probe_stage(const SplWork& w, ProbeSmem& sm, const uint32_t tile) {
    const uint32_t tid = threadIdx.x, tile0 = tile * SPL_TILE, Nup = (w.N + 15u) & ~15u;
    uint32_t hi = 0;
    if (BULK) {
        const uint32_t tb = min(PROBE_TEXT_BYTES, Nup - tile0);                  // bytes of the window that exist (16-byte units)
        const uint32_t bar = smem_addr(&sm.mbar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(tb + PROBE_PB_BYTES) : "memory");
            if (tb) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(smem_addr(sm.text)), "l"(w.text + tile0), "r"(tb), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_addr(sm.pb)), "l"(w.pstart + (tile0 >> 5)), "r"(PROBE_PB_BYTES), "r"(bar) : "memory");
        }
        for (uint32_t v = tb / 16u + tid; v < PROBE_TEXT_BYTES / 16u; v += SPL_THREADS)       // beyond the end of the text: zeros
            reinterpret_cast<uint4*>(sm.text)[v] = make_uint4(0, 0, 0, 0);
        if (w.with_special && tid < SPL_TILE / 32) sm.ps.spw[tid] = __ldg(w.spec + (tile0 >> 5) + tid);
        __syncthreads();                                                          // the barrier is initialised for everybody
        uint32_t ok, spins = 0;
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar) : "memory");
            if (!ok && ++spins > (1u << 24)) __trap();                            // a copy that never lands must not hang the device
        } while (!ok);
        const uint4 x = reinterpret_cast<const uint4*>(sm.text)[tid];            // the tile proper: 256 x 16 bytes
        hi = __popc(x.x & 0x80808080u) + __popc(x.y & 0x80808080u) + __popc(x.z & 0x80808080u) + __popc(x.w & 0x80808080u);
    } else {
        for (uint32_t v = tid; v < PROBE_TEXT_BYTES / 16u; v += SPL_THREADS) {
            uint32_t g = tile0 + v * 16;
            uint4 x = make_uint4(0, 0, 0, 0);
            if (g < Nup) x = __ldg(reinterpret_cast<const uint4*>(w.text + g));
            reinterpret_cast<uint4*>(sm.text)[v] = x;
            if (v < SPL_TILE / 16u) hi += __popc(x.x & 0x80808080u) + __popc(x.y & 0x80808080u) + __popc(x.z & 0x80808080u) + __popc(x.w & 0x80808080u);
        }
        for (uint32_t v = tid; v < PB_WORDS; v += SPL_THREADS) sm.pb[v] = __ldg(w.pstart + (tile0 >> 5) + v);
        if (w.with_special && tid < SPL_TILE / 32) sm.ps.spw[tid] = __ldg(w.spec + (tile0 >> 5) + tid);
        __syncthreads();
    }
    return hi;
}

// a tile goes through the refining pass when more than 1/16 of its bytes belong to multi-byte characters
template <bool BULK>
__global__ void __launch_bounds__(SPL_THREADS) k_probe(SplWork w) {
    __shared__ ProbeSmem sm;
    SPL_RETURN_IF_BAD_OFFSETS(w);
    const uint32_t my_hi = probe_stage<BULK>(w, sm, blockIdx.x);
    uint32_t n_hi = 0;
    const uint32_t P = build_piece_list(w, sm.ps, sm.pb, blockIdx.x, false, my_hi, &n_hi);
    if (n_hi * 16u > SPL_TILE) probe_tile_refine(w, sm.ps, sm.text, sm.pb, blockIdx.x, P);
    else probe_tile(w, sm.ps, sm.text, sm.pb, blockIdx.x, P);
}

// ------------------------------------------------------------------------------------------
// k_bpe: the merge loop (bpe.rs:83-194) for the listed pieces
//
// A piece of length class c is merged by a GROUP of G = 2^log2group(c) lanes, up to 32 parts per lane, so a warp
// merges 32 / G pieces side by side.  The live parts of a piece are an ARRAY (compacted after every round) in the
// warp's 12 KiB of shared memory:
//     S[i] = symbol of part i     K[i] = rank of the pair (part i, part i+1), BG_RANK_NONE if none     X[i] = scratch
// Part i sits at word (i / G) * 32 + p * G + i % G (p = the group's index in the warp), so row r of a group is G
// consecutive parts in G consecutive lanes and a ballot over a row is a bitmap of G consecutive parts.
//
// The reference merges one pair per step: the lowest rank, leftmost on ties (bpe.rs:121-138).  Here one ROUND
// performs every merge of a whole window of that sequence at once; the result is the same list of parts:
//   order pairs by key = (rank, position).  If no pair created in the window had a rank below the window's end,
//   the sequential loop would visit the existing pairs in key order and merge pair i unless a neighbour merged
//   first (its part is gone then):  m(i) = !(m(i-1) && key(i-1) < key(i)) && !(m(i+1) && key(i+1) < key(i)).
//   A pair below both neighbours merges; along a slope m alternates with the distance from the valley (the parity
//   comes from the row bitmaps); a pair above both neighbours merges iff neither does.
//   Every pair with m = 1 looks up the ranks its merge can create: (left part, T), (T, right part) and, when the
//   pair two to the right has m = 1 too, (T, T').  theta = the minimum of all of them.  No created pair ranks below
//   theta, so up to theta the sequential loop does exactly the m = 1 merges: the round commits those with
//   rank < theta (and always the global minimum, which bpe.rs merges first whatever follows).
//   The three looked-up ranks are also the K values of the new neighbours, so a round needs no second probe pass.
// tests/bpe_batch_sim.py is the same round in Python, checked against the oracle; random letter strings of
// 32..512 bytes take ~3 rounds instead of ~80 merges steps, runs of one character ~6 instead of ~120.
// ------------------------------------------------------------------------------------------
#define BG_RANK_NONE 0x1FFFFFu
#define SPL_BPE_WIN_CLS 3u                             // pieces beyond 64 bytes: windowed rounds (measured: cfg4 2.3 ms at 3, 2.6 at 4, 3.9 at 5)
#define BG_ARR       1056u                              // words per array: 1024 parts + one pad word per 32
#define BG_WORDS     (3u * BG_ARR + 256u + 64u)         // words per warp: S + K + X + worklist (512 x u16) + 2 x 32

// whole-piece probe of a piece in global memory by ONE thread (vocabularies with keys beyond 128 bytes only)
__device__ uint32_t lookupL_serial_g(const SplTables* T, const uint8_t* __restrict__ tx, uint32_t len) {
    uint64_t sum = 0;
    for (uint32_t i = 0; i * 8 < len; ++i) {
        uint64_t wv = 0;
        for (uint32_t b = 0; b < 8 && i * 8 + b < len; ++b) wv |= (uint64_t)__ldg(tx + i * 8 + b) << (8 * b);
        sum += spl_hashL_word(wv, i);
    }
    uint64_t hv = spl_hashL_final(sum, len);
    uint32_t mask = (1u << T->tl_log2) - 1, h = (uint32_t)(hv >> (64 - T->tl_log2));
    for (;;) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(T->tl + h));
        if (v.w == 0) return SPL_RANK_NONE;
        if (v.w == len && v.x == (uint32_t)hv && v.y == (uint32_t)(hv >> 32)) {
            const uint8_t* kb = T->tok_bytes + __ldg(T->tok_off + v.z);
            bool ok = true;
            for (uint32_t j = 0; j < len && ok; ++j) ok = (__ldg(tx + j) == __ldg(kb + j));
            if (ok) return v.z;
        }
        h = (h + 1) & mask;
    }
}

// The same for the pieces of a whole warp (ALL 32 lanes call this; has: the lane has a finished piece).  Pieces of one
// warp task are neighbours in the text, so their id-count corrections go to the same tile and the same chunk: lanes
// with the same tile add up first and ONE atomic goes out -- with millions of short segments (CJK text) per-piece
// atomics on a few hundred chunk counters were the whole cost of the kernel.
__device__ __forceinline__ void bpe_finish_warp(const SplWork& w, const bool has, uint64_t* slot, uint32_t gpos, uint32_t cnt) {
    const uint32_t lane = threadIdx.x & 31u;
    if (has) *slot = (uint64_t)gpos | ((uint64_t)cnt << 32) | SPL_ML_DONE;
    const int32_t delta = has ? (int32_t)cnt - 1 : 0;
    const uint32_t tile = has ? gpos / SPL_TILE : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(FULL, tile);
    const int32_t tsum = __reduce_add_sync(peers, delta);
    if (has && lane == (uint32_t)__ffs(peers) - 1u && tsum) atomicAdd(&w.tinfo[tile].extra, tsum);
    const uint32_t chunk = has ? tile / SPL_CHUNK_TILES : 0xFFFFFFFFu;
    const uint32_t cpeers = __match_any_sync(FULL, chunk);
    const int32_t csum = __reduce_add_sync(cpeers, delta);
    if (has && lane == (uint32_t)__ffs(cpeers) - 1u && csum) atomicAdd(&w.chunk_cnt[chunk], csum);
}

__device__ __forceinline__ void bpe_finish(const SplWork& w, uint64_t* slot, uint32_t gpos, uint32_t cnt) {
    *slot = (uint64_t)gpos | ((uint64_t)cnt << 32) | SPL_ML_DONE;
    if (cnt != 1u) {
        atomicAdd(&w.tinfo[gpos / SPL_TILE].extra, (int32_t)cnt - 1);
        atomicAdd(&w.chunk_cnt[gpos / (SPL_TILE * SPL_CHUNK_TILES)], (int32_t)cnt - 1);
    }
}

// Short pieces (classes below SplBpeCfg's window class): the loop of bpe.rs:119-167 as it stands, one merge per step,
// by a group of G = 2^LG lanes (one lane up to 32 bytes) -- a piece of a few dozen parts takes fewer instructions that
// way than the rounds of the windowed form below cost.
// The parts form a doubly linked list (bpe.rs:42-54) in the warp's shared memory:
//     A[e] = symbol:21 | next:11     R[e] = rank of (part e, next part):21 | e:11     P[e] = prev (u16)
// R doubles as the scan key: the minimum over R is the lowest rank at the leftmost position (bpe.rs:133).
// Part e of the piece of group p sits at word (e / G) * 32 + p * G + e % G, so the strided scan of a group is conflict free.
// One merge = strided min-scan of the ranks, a shuffle reduction inside the group, the splice and the two re-ranks
// (bpe.rs:146-166), done by lanes 0 and 1 of the group.
// All 32 lanes call this (valid: the group has a piece).  Returns the id count in lane g == 0; ids go to out[0 ..] in order.
#define BG_LINK_NONE 0x7FFu
template <uint32_t LG>
__device__ uint32_t bpe_seq(uint32_t* reg, const bool valid, const bool allow_whole, const SplTables* T,
                            const uint8_t* __restrict__ tx, const uint32_t n, uint32_t* __restrict__ out) {
    constexpr uint32_t G = 1u << LG;
    const uint32_t lane = threadIdx.x & 31u, p = lane >> LG, g = lane & (G - 1u);
    uint32_t* A = reg;                                             // symbol:21 | next:11
    uint32_t* R = reg + 1024;                                      // rank:21 | own position:11  (the scan key)
    uint16_t* P = reinterpret_cast<uint16_t*>(reg + 2048);         // prev
    const uint32_t* __restrict__ ptab = T->pair;
    const uint32_t plog = T->pair_log2;
    const uint32_t* __restrict__ bpair = T->bpair;
#define IDX(e) ((((e) >> LG) << 5) + (p << LG) + ((e) & (G - 1u)))
    bool act = valid;
    {
        // whole-piece probe of pieces beyond the probe halo (k_probe has already tried the shorter ones)
        const bool tryw = act && allow_whole && n > SPL_PROBE_HALO && n <= T->max_key_len;
        if (__any_sync(FULL, tryw)) {
            uint32_t id = SPL_RANK_NONE;
            if (tryw && g == 0) id = lookupL_serial_g(T, tx, n);
            id = __shfl_sync(FULL, id, p << LG);
            if (tryw && id != SPL_RANK_NONE) { if (g == 0) out[0] = id; act = false; }
        }
    }
    const bool whole = valid && !act;
    const uint32_t rows = act ? (n + G - 1u - g) >> LG : 0u;      // parts of this lane: e = g + it * G, at word it * 32 + lane
    const uint32_t rows4 = (rows + 3u) & ~3u;                      // the scan runs four rows at a time
    // ---- every byte becomes a part; its first rank comes from the dense byte x byte table (no hashing) ----------
#pragma unroll 4
    for (uint32_t it = 0; it < rows; ++it) A[it * 32u + lane] = __ldg(tx + g + (it << LG));
    for (uint32_t it = rows; it < rows4; ++it) R[it * 32u + lane] = BG_RANK_NONE << 11;
    __syncwarp();
#pragma unroll 4
    for (uint32_t it = 0; it < rows; ++it) {
        const uint32_t e = g + (it << LG);
        const uint32_t r = e + 1 < n ? __ldg(bpair + ((A[it * 32u + lane] << 8) | A[IDX(e + 1)])) : SPL_RANK_NONE;
        R[it * 32u + lane] = ((r & BG_RANK_NONE) << 11) | e;
        P[it * 32u + lane] = (uint16_t)(e ? e - 1 : BG_LINK_NONE);
    }
    __syncwarp();
#pragma unroll 4
    for (uint32_t it = 0; it < rows; ++it) {
        const uint32_t e = g + (it << LG);
        A[it * 32u + lane] = (T->byte_sym[A[it * 32u + lane]] << 11) | (e + 1 < n ? e + 1 : BG_LINK_NONE);
    }
    __syncwarp();
    // ---- merge loop --------------------------------------------------------------------------------------
    for (;;) {
        // leftmost minimum (bpe.rs:133): the key orders by rank, then by position
        uint32_t key = 0xFFFFFFFFu;
        for (uint32_t it = 0; it < rows4; it += 4) {
            const uint32_t* r4 = R + it * 32u + lane;
            key = min(min(key, r4[0]), min(r4[32], min(r4[64], r4[96])));
        }
#pragma unroll
        for (uint32_t o = G >> 1; o; o >>= 1) key = min(key, __shfl_xor_sync(FULL, key, o));
        const uint32_t best = key >> 11, bpos = key & BG_LINK_NONE;
        const bool go = act && best != BG_RANK_NONE;
        if (!__any_sync(FULL, go)) break;
        uint32_t j = 0, k = 0, h = 0, symk = 0, symh = 0;
        bool has_k = false, has_h = false;
        if (go) {
            j = A[IDX(bpos)] & BG_LINK_NONE;                       // the part being absorbed (exists: its pair has a rank)
            k = A[IDX(j)] & BG_LINK_NONE;                          // its right neighbour
            h = P[IDX(bpos)];                                      // left neighbour
            has_k = k != BG_LINK_NONE; has_h = h != BG_LINK_NONE;
            if (has_k) symk = A[IDX(k)] >> 11;
            if (has_h) symh = A[IDX(h)] >> 11;
        }
        __syncwarp();
        if (go) {
            const bool do_a = g == 0, do_b = g == (G > 1u ? 1u : 0u);
            uint32_t ra, rb;
            pair_lookup2(ptab, plog, do_a && has_k, best, symk, do_b && has_h, symh, best, ra, rb);
            if (do_a) {
                A[IDX(bpos)] = (best << 11) | k;                   // merged id == its rank
                R[IDX(bpos)] = ((ra & BG_RANK_NONE) << 11) | bpos;
                R[IDX(j)] = (BG_RANK_NONE << 11) | j;              // unlinked
                if (has_k) P[IDX(k)] = (uint16_t)bpos;
            }
            if (do_b && has_h) R[IDX(h)] = ((rb & BG_RANK_NONE) << 11) | h;
        }
        __syncwarp();
    }
    // ---- surviving known parts, in order (part 0 is never absorbed: it heads the list) --------------------
    uint32_t c = whole ? 1u : 0u;
    if (valid && !whole && g == 0)
        for (uint32_t e = 0; e != BG_LINK_NONE;) {
            const uint32_t a = A[IDX(e)], sy = a >> 11;
            if (sy < SPL_UNK_BASE) out[c++] = sy;                  // unknown bytes produce no id (bpe.rs:187-191)
            e = a & BG_LINK_NONE;
        }
#undef IDX
    return c;
}

// All 32 lanes call this; lane = p * G + g works on piece p of the warp's task (valid: the piece exists), G = 1 << LG.
// Returns the id count in every lane of the group; ids go to out[0 ..] in order.
template <uint32_t LG>
__device__ uint32_t bpe_group(uint32_t* reg, const bool valid, const bool allow_whole, const SplTables* T,
                              const uint8_t* __restrict__ tx, const uint32_t n, uint32_t* __restrict__ out) {
    constexpr uint32_t G = 1u << LG;
    constexpr uint32_t NONE = BG_RANK_NONE;
    const uint32_t lane = threadIdx.x & 31u, g = lane & (G - 1u), gsh = lane & ~(G - 1u);
    uint32_t* S = reg;                                             // symbols of the live parts
    uint32_t* K = reg + BG_ARR;                                    // rank of (part e, part e + 1)
    uint32_t* X = reg + 2u * BG_ARR;                               // ranks looked up in the current round
    uint16_t* WL = reinterpret_cast<uint16_t*>(reg + 3u * BG_ARR); // worklist of the round: every m-pair of the warp (<= 512)
    uint32_t* LW = reg + 3u * BG_ARR + 256u;                       // per lane: part count of its piece | B << 16
    uint32_t* TH = LW + 32u;                                       // per group (at its first lane): theta
    const uint32_t* __restrict__ ptab = T->pair;
    const uint32_t plog = T->pair_log2;
    const uint32_t* __restrict__ bpair = T->bpair;
    // part e of the group that starts at lane gs: one pad word per 32 parts, so lanes that walk their blocks in step hit 32 banks
#define ADR(gs, e) ((gs) * 33u + (e) + ((e) >> 5))
#define AD(e) ADR(gsh, e)
    // inclusive sum over the lanes of the group
#define GROUP_SCAN(v) _Pragma("unroll") for (uint32_t o_ = 1; o_ < G; o_ <<= 1) { const uint32_t t_ = __shfl_up_sync(FULL, (v), o_); if (g >= o_) (v) += t_; }
    bool act = valid;
    {
        // whole-piece probe of pieces beyond the probe halo (k_probe has already tried the shorter ones)
        const bool tryw = act && allow_whole && n > SPL_PROBE_HALO && n <= T->max_key_len;
        if (__any_sync(FULL, tryw)) {
            uint32_t id = SPL_RANK_NONE;
            if (tryw && g == 0) id = lookupL_serial_g(T, tx, n);
            id = __shfl_sync(FULL, id, gsh);
            if (tryw && id != SPL_RANK_NONE) { if (g == 0) out[0] = id; act = false; }
        }
    }
    const bool whole = valid && !act;
    uint32_t L = act ? n : 0u;                                     // live parts (uniform in the group)
    uint32_t B = max(2u, (L + G - 1u) >> LG);                      // parts per lane: lane g owns [g * B, g * B + B)
    {
        const uint32_t e0 = g * B, nv = e0 < L ? min(B, L - e0) : 0u;
        // ---- every byte becomes a part (the lanes of the group read consecutive bytes); its first rank comes from
        // the dense byte x byte table (no hashing)
#pragma unroll 4
        for (uint32_t i = g; i < L; i += G) S[AD(i)] = __ldg(tx + i);
        __syncwarp();
#pragma unroll 4
        for (uint32_t j = 0; j < nv; ++j) {
            const uint32_t e = e0 + j;
            const uint32_t r = e + 1 < L ? __ldg(bpair + ((S[AD(e)] << 8) | S[AD(e + 1)])) : SPL_RANK_NONE;
            K[AD(e)] = r & NONE;
        }
        __syncwarp();
#pragma unroll 4
        for (uint32_t i = g; i < L; i += G) S[AD(i)] = T->byte_sym[S[AD(i)]];
        __syncwarp();
    }
    // ---- merge rounds ------------------------------------------------------------------------------------
    for (;;) {
        const uint32_t e0 = g * B, nv = e0 < L ? min(B, L - e0) : 0u;
        // (1) every pair of the lane's block against its neighbours
        uint32_t LV = 0, LT = 0, RT = 0, gmin = 0xFFFFFFFFu;
        {
            // (K[L - 1] is BG_RANK_NONE at all times, so the walk needs no bound checks: what lies beyond is never used)
            uint32_t prev = (e0 && e0 < L) ? K[AD(e0 - 1u)] : NONE, cur = e0 < L ? K[AD(e0)] : NONE;
            for (uint32_t j = 0; j < nv; ++j) {
                const uint32_t e = e0 + j, nxt = K[AD(e + 1u)];
                if (cur != NONE) {
                    const uint32_t bit = 1u << j;
                    LV |= bit;
                    if (prev <= cur) LT |= bit;                                      // the left neighbour merges first (it wins ties)
                    if (nxt < cur) RT |= bit;
                    gmin = min(gmin, (cur << 11) | e);
                }
                prev = cur; cur = nxt;
            }
        }
#pragma unroll
        for (uint32_t o = G >> 1; o; o >>= 1) gmin = min(gmin, __shfl_xor_sync(FULL, gmin, o));
        if (!__any_sync(FULL, gmin != 0xFFFFFFFFu)) break;                          // no pair with a rank anywhere in the warp
        const uint32_t fV = LV & ~LT & ~RT, fDL = LV & LT & ~RT, fDR = LV & RT & ~LT, fPK = LV & LT & RT;
        // (2) m: valleys merge; along a slope m alternates with the distance from its valley; slopes that cross into the
        // next lane's block take that lane's boundary bit, until nothing changes (one extra pass per lane a slope spans)
        uint32_t m, cin = 0, cin2 = 0;
        {
            for (;;) {
                m = spl_window_slopes(fV, fDL, fDR, cin, cin2, B);
                const uint32_t up = __shfl_up_sync(FULL, m, 1), dn = __shfl_down_sync(FULL, m, 1);
                const uint32_t ncin = g ? (up >> (B - 1u)) & 1u : 0u, ncin2 = g + 1u < G ? dn & 1u : 0u;
                const bool ch = (ncin != cin && (fDL & 1u)) || (ncin2 != cin2 && ((fDR >> (B - 1u)) & 1u));
                cin = ncin; cin2 = ncin2;
                if (!__any_sync(FULL, ch)) break;
            }
            m = spl_window_peaks(m, fPK, cin, cin2, B);
        }
        // (3) the ranks the merges can create; theta = their minimum.  The m-pairs of the whole warp go through a
        // worklist so that all 32 lanes probe.
        {
            const uint32_t dn = __shfl_down_sync(FULL, m, 1);
            const uint32_t m2 = (uint32_t)(((uint64_t)m | ((uint64_t)(g + 1u < G ? dn : 0u) << B)) >> 2);   // m of the pair two to the right
            const uint32_t cntm = __popc(m);
            uint32_t off = cntm;
#pragma unroll
            for (uint32_t o = 1; o < 32u; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, off, o); if (lane >= o) off += t; }
            const uint32_t total = __shfl_sync(FULL, off, 31);
            off -= cntm;
            for (uint32_t mm = m; mm; mm &= mm - 1u) {
                const uint32_t j = (uint32_t)__ffs(mm) - 1u;
                WL[off++] = (uint16_t)(j | (lane << 5) | (((m2 >> j) & 1u) << 10));
            }
            LW[lane] = L | (B << 16);
            TH[lane] = NONE;
            __syncwarp();
            for (uint32_t q = lane; q < total; q += 32u) {
                const uint32_t wv = WL[q], ln = (wv >> 5) & 31u, g2 = ln & (G - 1u), gs2 = ln & ~(G - 1u), info = LW[ln];
                const uint32_t L2 = info & 0xFFFFu, e = g2 * (info >> 16) + (wv & 31u);
                const uint32_t tm = K[ADR(gs2, e)];                                 // merged id == its rank
                const bool hasL = e > 0u, hasR = e + 2u < L2, hasC = (wv >> 10) & 1u;
                const uint32_t sl = hasL ? S[ADR(gs2, e - 1u)] : 0u, sr = hasR ? S[ADR(gs2, e + 2u)] : 0u, tc = hasC ? K[ADR(gs2, e + 2u)] : 0u;
                uint32_t ra, rb, rc;
                pair_lookup3(ptab, plog, hasL, sl, tm, hasR, tm, sr, hasC, tm, tc, ra, rb, rc);
                ra &= NONE; rb &= NONE; rc &= NONE;
                atomicMin(&TH[gs2], min(ra, min(rb, rc)));
                X[ADR(gs2, e)] = ra | (rc << 21);                                   // rc: low 11 bits here, high 10 bits in the next word
                X[ADR(gs2, e + 1u)] = rb | ((rc >> 11) << 21);
            }
            __syncwarp();
        }
        // (4) commit: m-pairs below theta, and the global minimum in any case
        const uint32_t theta = TH[gsh];
        uint32_t cm = 0;
        for (uint32_t mm = m; mm; mm &= mm - 1u) {
            const uint32_t j = (uint32_t)__ffs(mm) - 1u, kc = K[AD(e0 + j)];
            if (kc < theta || ((kc << 11) | (e0 + j)) == gmin) cm |= 1u << j;
        }
        // (5) new symbol and rank of every surviving part, in place; then the parts move to their new index through
        // the free array (the three arrays swap roles)
        uint32_t surv;
        {
            const uint32_t up = __shfl_up_sync(FULL, cm, 1), dn = __shfl_down_sync(FULL, cm, 1);
            const uint64_t cw = (uint64_t)cm | ((uint64_t)(g + 1u < G ? dn : 0u) << B);
            const uint32_t cm1 = (uint32_t)(cw >> 1), cm2 = (uint32_t)(cw >> 2);
            const uint32_t ex = nv >= 32u ? 0xFFFFFFFFu : (1u << nv) - 1u;
            surv = ex & ~((cm << 1) | (g ? (up >> (B - 1u)) & 1u : 0u));            // the right part of a committed pair is absorbed
            for (uint32_t mm = surv & (cm | cm1); mm; mm &= mm - 1u) {
                const uint32_t j = (uint32_t)__ffs(mm) - 1u, e = e0 + j, x1 = X[AD(e + 1u)];
                if ((cm >> j) & 1u) {
                    const uint32_t x0 = X[AD(e)];
                    S[AD(e)] = K[AD(e)];
                    K[AD(e)] = ((cm2 >> j) & 1u) ? ((x0 >> 21) | ((x1 >> 21) << 11)) : (x1 & NONE);
                } else {
                    K[AD(e)] = x1 & NONE;
                }
            }
        }
        __syncwarp();
        uint32_t lb = __popc(surv);
        GROUP_SCAN(lb);
        const uint32_t newL = __shfl_sync(FULL, lb, gsh + G - 1u);
        lb -= __popc(surv);
        {
            uint32_t d = lb;
            for (uint32_t mm = surv; mm; mm &= mm - 1u, ++d) X[AD(d)] = S[AD(e0 + (uint32_t)__ffs(mm) - 1u)];
            __syncwarp();
            d = lb;
            for (uint32_t mm = surv; mm; mm &= mm - 1u, ++d) S[AD(d)] = K[AD(e0 + (uint32_t)__ffs(mm) - 1u)];
            __syncwarp();
            uint32_t* t = K; K = S; S = X; X = t;
        }
        L = newL;
        B = max(2u, (L + G - 1u) >> LG);
    }
    // ---- surviving known parts, in order -------------------------------------------------------------------
    uint32_t c = whole ? 1u : 0u;
    {
        const uint32_t e0 = g * B, nv = e0 < L ? min(B, L - e0) : 0u;
        uint32_t keep = 0;
        for (uint32_t j = 0; j < nv; ++j) keep |= (S[AD(e0 + j)] < SPL_UNK_BASE ? 1u : 0u) << j;   // unknown bytes produce no id (bpe.rs:187-191)
        uint32_t lb = __popc(keep);
        GROUP_SCAN(lb);
        c += __shfl_sync(FULL, lb, gsh + G - 1u);
        uint32_t d = lb - __popc(keep) + (whole ? 1u : 0u);
        for (uint32_t mm = keep; mm; mm &= mm - 1u, ++d) out[d] = S[AD(e0 + (uint32_t)__ffs(mm) - 1u)];
    }
#undef ADR
#undef AD
#undef GROUP_SCAN
    return c;
}

// ------------------------------------------------------------------------------------------
// k_bpe: the merge loop for pieces and segments of up to SPL_SEG_MAX (32) bytes.  ONE LANE PER PIECE.
//
// The whole state of a piece lives in the lane's own two rows of shared memory and one register:
//     S[i] = symbol of the part that starts at byte i          K[i] = rank of (part at i, next part) << 5 | i
//     live = bit i set while a part starts at byte i
// so the neighbours of a part are two bit scans and the reference's linked list (bpe.rs:42-54) needs no memory.
// The minimum over K is the lowest rank at the leftmost position -- exactly the pair bpe.rs:121-138 selects; it is
// found with 16-byte loads (rows are 36 words apart: 16-byte aligned, and the eight lanes of a quarter warp cover all 32
// banks).  The first rank of every part comes from the dense byte x byte table (one 4-byte load, no hashing); the two
// re-ranks after a merge (bpe.rs:146-166) are two 256-bit bucket loads in flight together.
// The whole-piece probe is k_probe's business; CJK text arrives here already cut into segments (probe_tile_refine).
// ------------------------------------------------------------------------------------------
#define WK_STRIDE 36u                                  // words between the rows of two lanes
#define WK_WORDS  (2u * 32u * WK_STRIDE)               // per warp: 32 K rows, 32 S rows
#define WK_NONE   (BG_RANK_NONE << 5)
#define WK_SMEM_BYTES ((SPL_BPE_THREADS / 32) * WK_WORDS * 4)

// four text bytes from byte index gi on (little endian); never reads beyond the 16-byte padded end of the text
__device__ __forceinline__ uint32_t text_load4(const uint8_t* __restrict__ text, uint32_t gi, uint32_t n_up) {
    const uint32_t a0 = gi & ~3u;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(text + a0);
    const uint32_t a = __ldg(wp), b = (a0 + 4u < n_up) ? __ldg(wp + 1) : 0u;
    return __funnelshift_r(a, b, (gi & 3u) * 8u);
}

// merge loop over the n <= 32 bytes from text[gpos] on; ids to out[0 ..] in order; returns their number
__device__ __forceinline__ uint32_t bpe_lane(const SplTables* __restrict__ T, const uint8_t* __restrict__ text, const uint32_t n_up,
                                             uint32_t* __restrict__ Krow, uint32_t* __restrict__ Srow,
                                             const uint32_t gpos, const uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t* __restrict__ ptab = T->pair;
    const uint32_t plog = T->pair_log2;
    const uint32_t* __restrict__ bpair = T->bpair;
    {
        uint32_t prevb = 0;
        for (uint32_t i = 0; i < n; i += 4u) {
            uint32_t w4 = text_load4(text, gpos + i, n_up);
#pragma unroll
            for (uint32_t q = 0; q < 4u; ++q) {
                const uint32_t b = w4 & 0xFFu;
                w4 >>= 8;
                if (i + q < n) {
                    Srow[i + q] = T->byte_sym[b];
                    if (i + q) Krow[i + q - 1u] = ((__ldg(bpair + ((prevb << 8) | b)) & BG_RANK_NONE) << 5) | (i + q - 1u);
                    prevb = b;
                }
            }
        }
        if (n) Krow[n - 1u] = WK_NONE | (n - 1u);
        for (uint32_t i = n; i < ((n + 3u) & ~3u); ++i) Krow[i] = 0xFFFFFFFFu;
    }
    uint32_t live = n >= 32u ? 0xFFFFFFFFu : (1u << n) - 1u;
    const uint32_t n4 = (n + 3u) >> 2;
    for (;;) {
        uint32_t key = 0xFFFFFFFFu;
        const uint4* k4 = reinterpret_cast<const uint4*>(Krow);
        for (uint32_t it = 0; it < n4; ++it) {
            const uint4 v = k4[it];
            key = min(min(key, v.x), min(v.y, min(v.z, v.w)));
        }
        const uint32_t r = key >> 5;
        if (r >= BG_RANK_NONE) break;                                   // (also n == 0: key is all ones)
        const uint32_t p = key & 31u;                                   // the pair (part at p, next part): p <= 30
        const uint32_t upper = live & ~((2u << p) - 1u);
        const uint32_t j = (uint32_t)__ffs(upper) - 1u;                 // the part being absorbed
        const uint32_t upper2 = upper & (upper - 1u);
        const uint32_t lower = live & ((1u << p) - 1u);
        const bool has_k = upper2 != 0u, has_h = lower != 0u;
        const uint32_t k = has_k ? (uint32_t)__ffs(upper2) - 1u : 0u, h = has_h ? 31u - (uint32_t)__clz(lower) : 0u;
        const uint32_t symk = Srow[k], symh = Srow[h];
        live &= ~(1u << j);
        Srow[p] = r;                                                    // merged id == its rank
        Krow[j] = WK_NONE | j;
        uint32_t ra, rb;
        pair_lookup2(ptab, plog, has_k, r, symk, has_h, symh, r, ra, rb);
        Krow[p] = ((ra & BG_RANK_NONE) << 5) | p;
        if (has_h) Krow[h] = ((rb & BG_RANK_NONE) << 5) | h;
    }
    uint32_t cnt = 0;
    for (uint32_t m = live; m; m &= m - 1u) {
        const uint32_t sy = Srow[(uint32_t)__ffs(m) - 1u];
        if (sy < SPL_UNK_BASE) out[cnt++] = sy;                         // unknown bytes produce no id (bpe.rs:187-191)
    }
    return cnt;
}

__global__ void __launch_bounds__(SPL_BPE_THREADS, 6) k_bpe(SplWork w) {
    extern __shared__ __align__(16) uint32_t bpe_smem[];
    SPL_RETURN_IF_BAD_OFFSETS(w);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t gwarp = blockIdx.x * (SPL_BPE_THREADS / 32) + warp, nwarps = gridDim.x * (SPL_BPE_THREADS / 32);
    uint32_t* Krow = bpe_smem + warp * WK_WORDS + lane * WK_STRIDE;
    uint32_t* Srow = Krow + 32u * WK_STRIDE;
    const SplTables* T = w.T;
    const uint32_t n_up = (w.N + 15u) & ~15u;
    // the longer pieces first, so that the tail of the kernel is short work
    for (int c = 1; c >= 0; --c) {
        const uint32_t n = w.counters[SPL_CTR_CLS + c];
        // few pieces: spread them over all warps of the grid (a warp's time is its slowest lane's)
        uint32_t per = 32u;
        while (per > 1u && (size_t)n * 2u <= (size_t)nwarps * per) per >>= 1;
        const uint32_t tasks = (n + per - 1u) / per;
        for (uint32_t t = gwarp; t < tasks; t += nwarps) {
            const uint32_t pi = t * per + lane;
            const bool has = lane < per && pi < n;
            uint64_t* slot = &w.mlist[w.ml_base[c] + (has ? pi : 0u)];
            uint32_t gpos = 0, cnt = 0;
            if (has) {
                const uint64_t e = *slot;
                gpos = (uint32_t)e;
                cnt = bpe_lane(T, w.text, n_up, Krow, Srow, gpos, (uint32_t)(e >> 32) & SPL_ML_LEN_MASK, w.pool + gpos);
            }
            bpe_finish_warp(w, has, slot, gpos, cnt);
        }
    }
}

// 64-bit hash of the len bytes at text[gpos] by the G = 2^LG lanes of a group (every lane gets it)
template <uint32_t LG>
__device__ __forceinline__ uint64_t group_hash(const uint8_t* __restrict__ text, uint32_t n_up, uint32_t gpos, uint32_t len) {
    constexpr uint32_t G = 1u << LG;
    const uint32_t g = threadIdx.x & (G - 1u);
    uint64_t sum = 0;
    for (uint32_t i = g * 4u; i < len; i += G * 4u) {
        uint32_t w4 = text_load4(text, gpos + i, n_up);
        if (len - i < 4u) w4 &= (1u << (8u * (len - i))) - 1u;
        sum += spl_mix64((uint64_t)w4 + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1u));
    }
#pragma unroll
    for (uint32_t o = G >> 1; o; o >>= 1) {
        const uint32_t lo = __shfl_xor_sync(FULL, (uint32_t)sum, o), hi = __shfl_xor_sync(FULL, (uint32_t)(sum >> 32), o);
        sum += (uint64_t)lo | ((uint64_t)hi << 32);
    }
    return spl_hashL_final(sum, len);
}

// Duplicate detection for the pieces of a warp task (all 32 lanes call this; valid: the lane's group has a piece).
// Returns true in every lane of a group whose piece has the same bytes as an EARLIER entry of the miss list: that
// entry's ids will do (k_bpe_fin copies the result), the merge loop is skipped.  Exact: equal hash tags only nominate
// a candidate, the bytes (and the piece length) are compared.
template <uint32_t LG>
__device__ __forceinline__ bool dedup_check(const SplWork& w, const bool valid, const uint32_t midx, uint64_t* slot,
                                            const uint32_t gpos, const uint32_t len) {
    constexpr uint32_t G = 1u << LG;
    const uint32_t lane = threadIdx.x & 31u, g = lane & (G - 1u), gsh = lane & ~(G - 1u);
    const uint32_t n_up = (w.N + 15u) & ~15u;
    const uint64_t hv = group_hash<LG>(w.text, n_up, valid ? gpos : 0u, valid ? len : 0u);
    uint32_t cand = SPL_RANK_NONE;                                    // miss-list index of the entry that may hold the same bytes
    if (valid && g == 0) {
        const uint32_t tag = (uint32_t)(hv >> 32) | 1u;
        uint32_t idx = (uint32_t)hv & w.dd_mask;
        for (uint32_t probe = 0; probe < 4u; ++probe, idx = (idx + 1u) & w.dd_mask) {
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&w.dd_tab[idx]);
            if (cur == 0ull) cur = atomicCAS(&w.dd_tab[idx], 0ull, ((unsigned long long)tag << 32) | (midx + 1u));
            if (cur == 0ull) break;                                   // inserted: this entry stands for its bytes
            if ((uint32_t)(cur >> 32) == tag) { cand = (uint32_t)cur - 1u; break; }
        }
    }
    cand = __shfl_sync(FULL, cand, gsh);
    const bool check = valid && cand != SPL_RANK_NONE && cand != midx;
    if (!__any_sync(FULL, check)) return false;
    bool same = check;
    uint32_t rpos = 0;
    if (check) {
        rpos = (uint32_t)w.mlist[cand];                               // (the low word of an entry is its position, before and after)
        for (uint32_t i = g * 4u; i < len && same; i += G * 4u) {
            uint32_t a = text_load4(w.text, gpos + i, n_up), b = text_load4(w.text, rpos + i, n_up);
            if (len - i < 4u) { const uint32_t m = (1u << (8u * (len - i))) - 1u; a &= m; b &= m; }
            same = a == b;
        }
        if (g == 0 && same) same = g_next_bit(w.pstart, rpos + 1u, w.N + 1u) == rpos + len;     // the same length, too
    }
    // every lane of the group must agree
    uint32_t bad = __ballot_sync(FULL, check && !same);
    const uint32_t gmask = (G >= 32u ? FULL : ((1u << G) - 1u) << gsh);
    const bool dup = check && !(bad & gmask);
    if (dup && g == 0) {
        w.dup_of[midx - w.ml_base[2]] = cand;
        *slot = (uint64_t)gpos | ((uint64_t)len << 32) | SPL_ML_DUP;
        atomicAdd(&w.counters[SPL_CTR_DUP], 1u);
    }
    return dup;
}

template <uint32_t LG, bool WINDOWED>
__device__ void bpe_class(const SplWork& w, uint32_t c, uint32_t* reg, uint32_t gwarp, uint32_t nwarps) {
    const uint32_t n = w.counters[SPL_CTR_CLS + c];
    if (!n) return;
    const uint32_t lane = threadIdx.x & 31u;
    // pieces per warp task: all 32 / G groups when there is plenty of work; fewer when the list is short, so that the
    // pieces spread over all warps of the grid and a warp's merge rounds are the maximum over fewer pieces
    uint32_t ppw = 32u >> LG;
    while (ppw > 1u && (size_t)n * 2u <= (size_t)nwarps * ppw) ppw >>= 1;
    const uint32_t tasks = (n + ppw - 1) / ppw;
    for (uint32_t t = gwarp; t < tasks; t += nwarps) {
        const uint32_t pi = t * ppw + (lane >> LG);
        bool valid = (lane >> LG) < ppw && pi < n;
        const uint32_t midx = w.ml_base[c] + (valid ? pi : 0u);
        uint64_t* slot = &w.mlist[midx];
        const uint64_t e = *slot;
        const uint32_t gpos = (uint32_t)e, len = (uint32_t)(e >> 32) & SPL_ML_LEN_MASK;
        if (w.dd_tab) {                                               // (all 32 lanes: the check shuffles and votes)
            const bool dup = dedup_check<LG>(w, valid, midx, slot, gpos, len);
            valid = valid && !dup;
        }
        if (!__any_sync(FULL, valid)) continue;
        uint32_t cnt;
        const bool allow_whole = !(e & SPL_ML_SEG);                   // a segment of a piece is not a piece: no whole-piece probe
        if constexpr (WINDOWED) cnt = bpe_group<LG>(reg, valid, allow_whole, w.T, w.text + gpos, len, w.pool + gpos);
        else cnt = bpe_seq<LG>(reg, valid, allow_whole, w.T, w.text + gpos, len, w.pool + gpos);
        bpe_finish_warp(w, valid && (lane & ((1u << LG) - 1u)) == 0, slot, gpos, cnt);
        __syncwarp();
    }
}

// The whole block merges one piece that does not fit shared memory: text bytes tx[0, len) in global memory; the
// sym / rnk / next / prev arrays live in the scratch pool.  Returns (to every thread) the number of ids, written in
// order to out[0 ..].
__device__ uint32_t bpe_piece_block(uint64_t* red, uint32_t* s_bcast, const SplTables* T, const bool allow_whole,
                                    const uint8_t* __restrict__ tx, uint32_t len, uint32_t* scratch, uint32_t* __restrict__ out) {
    const uint32_t tid = threadIdx.x;
    uint32_t* sym = scratch;
    uint32_t* rnk = scratch + len;
    uint32_t* nxt = scratch + 2 * (size_t)len;
    uint32_t* prv = scratch + 3 * (size_t)len;

    if (allow_whole && len <= T->max_key_len) {      // only for vocabularies with very long keys
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
        red[tid] = best;
        __syncthreads();
        for (uint32_t o = SPL_BPE_THREADS / 2; o; o >>= 1) {
            if (tid < o && red[tid + o] < red[tid]) red[tid] = red[tid + o];
            __syncthreads();
        }
        uint64_t mn = red[0];
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
    red[tid] = c;
    __syncthreads();
    if (tid == 0) { uint64_t run = 0; for (int q = 0; q < SPL_BPE_THREADS; ++q) { uint64_t t = red[q]; red[q] = run; run += t; } *s_bcast = (uint32_t)run; }
    __syncthreads();
    uint32_t o = (uint32_t)red[tid];
    for (uint32_t j = lo; j < hi; ++j) { uint32_t sv = sym[j]; if (sv != SPL_RANK_NONE && sv < SPL_UNK_BASE) out[o++] = sv; }
    uint32_t total = *s_bcast;
    __syncthreads();
    return total;
}

#define BPE_SMEM_BYTES ((SPL_BPE_THREADS / 32) * BG_WORDS * 4)

// Exclusive prefix of the id counts of the chunks (SPL_CHUNK_TILES tiles each; k_probe and the merge kernels keep the
// chunk totals up to date with atomics) -> chunk_state; k_emit adds the tiles inside a chunk itself.  Run by ONE block
// of SPL_BPE_THREADS threads: the block of k_bpe_long that finishes last (chunks are few, N / 128 KiB).
// smem: CS_TILE + CS_TILE / 32 + 2 * SPL_BPE_THREADS + 2 words.
#define CS_PER  32u                                  // chunks per thread and round
#define CS_TILE (SPL_BPE_THREADS * CS_PER)
__device__ void chunk_scan_block(const SplWork& w, uint32_t* smem) {
    uint32_t* cnt = smem;                            // one pad word per 32: a thread's CS_PER chunks start in its own bank
    uint64_t* part = reinterpret_cast<uint64_t*>(smem + CS_TILE + CS_TILE / 32u);
    uint64_t* s_carry = part + SPL_BPE_THREADS;
    const uint32_t tid = threadIdx.x;
    const uint32_t n = (w.n_tiles + SPL_CHUNK_TILES - 1) / SPL_CHUNK_TILES;
    if (tid == 0) *s_carry = 0;
    for (uint32_t c0 = 0; c0 < n; c0 += CS_TILE) {
        // coalesced, independent loads (past the L1: other blocks' atomics wrote these words)
#pragma unroll 8
        for (uint32_t i = tid; i < CS_TILE; i += SPL_BPE_THREADS)
            cnt[i + (i >> 5)] = c0 + i < n ? (uint32_t)__ldcg(w.chunk_cnt + c0 + i) : 0u;
        __syncthreads();
        const uint32_t lo = tid * CS_PER;
        uint64_t sum = 0;
#pragma unroll 8
        for (uint32_t q = 0; q < CS_PER; ++q) sum += cnt[lo + q + tid];               // (lo + q) + ((lo + q) >> 5)
        part[tid] = sum;
        __syncthreads();
        if (tid < 32) {                              // one warp turns the per-thread sums into exclusive offsets
            uint64_t carry = *s_carry;
            for (uint32_t q0 = 0; q0 < SPL_BPE_THREADS; q0 += 32) {
                const uint64_t x = part[q0 + tid];
                uint64_t incl = x;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t a = __shfl_up_sync(FULL, (uint32_t)incl, o), b = __shfl_up_sync(FULL, (uint32_t)(incl >> 32), o);
                    if (tid >= (uint32_t)o) incl += (uint64_t)a | ((uint64_t)b << 32);
                }
                part[q0 + tid] = carry + incl - x;
                const uint32_t ta = __shfl_sync(FULL, (uint32_t)incl, 31), tb = __shfl_sync(FULL, (uint32_t)(incl >> 32), 31);
                carry += (uint64_t)ta | ((uint64_t)tb << 32);
            }
            __syncwarp();                            // every lane has read *s_carry (the shuffles order that in practice; this says so)
            if (tid == 0) *s_carry = carry;
        }
        __syncthreads();
        uint64_t run = part[tid];
        for (uint32_t q = 0; q < CS_PER && c0 + lo + q < n; ++q) { w.chunk_state[c0 + lo + q] = run; run += cnt[lo + q + tid]; }
        __syncthreads();
    }
}

// k_bpe_long: what the segment walker left (pieces with a segment beyond SPL_SEG_MAX bytes).  Class 2 (33..64 bytes):
// one merge per step by two lanes per piece; the classes from win_cls up: windowed rounds (13.6 KiB of shared memory
// per warp); then the huge class.
#define BPE_SEQ_WORDS 2560u                            // A[1024] + R[1024] + P[1024 x u16]
__global__ void __launch_bounds__(SPL_BPE_THREADS, 4) k_bpe_long(SplWork w, const uint32_t win_cls) {
    extern __shared__ __align__(16) uint32_t bpe_smem[];
    __shared__ uint32_t s_bcast;
    __shared__ uint64_t s_off;
    SPL_RETURN_IF_BAD_OFFSETS(w);
    const SplTables* T = w.T;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t gwarp = blockIdx.x * (SPL_BPE_THREADS / 32) + warp, nwarps = gridDim.x * (SPL_BPE_THREADS / 32);
    uint32_t* reg = bpe_smem + warp * BG_WORDS;
    // the longest pieces first, so that the tail of the kernel is short work
    bpe_class<5, true>(w, 6, reg, gwarp, nwarps);
    if (win_cls <= 5u) bpe_class<4, true>(w, 5, reg, gwarp, nwarps); else bpe_class<4, false>(w, 5, reg, gwarp, nwarps);
    if (win_cls <= 4u) bpe_class<3, true>(w, 4, reg, gwarp, nwarps); else bpe_class<3, false>(w, 4, reg, gwarp, nwarps);
    if (win_cls <= 3u) bpe_class<2, true>(w, 3, reg, gwarp, nwarps); else bpe_class<2, false>(w, 3, reg, gwarp, nwarps);
    if (win_cls <= 2u) bpe_class<1, true>(w, 2, reg, gwarp, nwarps); else bpe_class<1, false>(w, 2, reg, gwarp, nwarps);
    // ---- huge class: whole block, global scratch ----------------------------------------------------
    {
        const uint32_t n = w.counters[SPL_CTR_CLS + SPL_NCLS - 1];
        if (n) {
            __syncthreads();
            uint64_t* red = reinterpret_cast<uint64_t*>(bpe_smem);
            for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
                uint64_t* slot = &w.mlist[w.ml_base[SPL_NCLS - 1] + i];
                uint64_t e = *slot;
                uint32_t gpos = (uint32_t)e;
                if (tid == 0) {
                    uint32_t ge = g_next_bit(w.pstart, gpos + 1, w.N + 1);
                    const unsigned long long need = 4ull * (ge - gpos);
                    unsigned long long off = atomicAdd(reinterpret_cast<unsigned long long*>(&w.counters[SPL_CTR_HUGE_POOL]), need);
                    if (off + need > w.huge_pool_words) { atomicOr(&w.counters[SPL_CTR_ERR], SPL_DEVERR_HUGE_POOL); off = ~0ull; }
                    s_off = off; s_bcast = ge - gpos;
                }
                __syncthreads();
                const uint64_t off = s_off;
                const uint32_t len = s_bcast;
                __syncthreads();
                uint32_t c = 0;
                if (off != ~0ull) c = bpe_piece_block(red, &s_bcast, T, !(e & SPL_ML_SEG), w.text + gpos, len, w.huge_pool + off, w.pool + gpos);
                if (tid == 0) bpe_finish(w, slot, gpos, c);
                __syncthreads();
            }
        }
    }
}

// k_bpe_fin: (1) an entry that was found to repeat an earlier one (dedup_check) takes over that entry's result: its
// ids are read from the earlier piece's place in the pool, only the id count is added to the entry's own tile;
// (2) the block that finishes last scans the chunk totals for k_emit.
#define FIN_SMEM_WORDS (CS_TILE + CS_TILE / 32u + 4u * SPL_BPE_THREADS + 4u)
__global__ void __launch_bounds__(SPL_BPE_THREADS) k_bpe_fin(SplWork w) {
    __shared__ __align__(16) uint32_t fin_smem[FIN_SMEM_WORDS];
    __shared__ uint32_t s_last;
    SPL_RETURN_IF_BAD_OFFSETS(w);
    const uint32_t tid = threadIdx.x;
    if (w.dd_tab && w.counters[SPL_CTR_DUP]) {
        const uint32_t lo = w.ml_base[2], hi = w.ml_base[SPL_NCLS - 1];           // the classes dedup_check looks at
        for (uint32_t c = 2; c < SPL_NCLS - 1; ++c) {
            const uint32_t n = w.counters[SPL_CTR_CLS + c];
            for (uint32_t i = blockIdx.x * SPL_BPE_THREADS + tid; i < n; i += gridDim.x * SPL_BPE_THREADS) {
                const uint32_t midx = w.ml_base[c] + i;
                const uint64_t e = w.mlist[midx];
                if (!(e & SPL_ML_DUP)) continue;
                const uint32_t rep = w.dup_of[midx - lo];
                if (rep < lo || rep >= hi) continue;                            // (cannot happen)
                const uint64_t re = w.mlist[rep];
                const uint32_t cnt = (uint32_t)(re >> 32) & SPL_ML_LEN_MASK, gpos = (uint32_t)e;
                // the ids stay where the earlier piece put them; only the count is added to this piece's own tile
                // (repeats are scattered over the text: nothing to aggregate, unlike bpe_finish_warp)
                w.mlist[midx] = (uint64_t)(uint32_t)re | ((uint64_t)cnt << 32) | SPL_ML_DONE;
                if (cnt != 1u) {
                    atomicAdd(&w.tinfo[gpos / SPL_TILE].extra, (int32_t)cnt - 1);
                    atomicAdd(&w.chunk_cnt[gpos / (SPL_TILE * SPL_CHUNK_TILES)], (int32_t)cnt - 1);
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(&w.counters[SPL_CTR_TICKET], 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        chunk_scan_block(w, fin_smem);
    }
}

// ------------------------------------------------------------------------------------------
// k_emit: one tile per block: pv[] (+ pool[] for merged pieces, char_ids[] for characters of several ids) -> ids in
// document order, output offsets.  Four pieces per thread and round (one 16-byte load of pv), so a warp owns 128
// consecutive pieces.
//
// The ids of a tile are assembled in shared memory and leave in whole 128-byte lines.  (A thread that stores its own
// four-odd ids straight to global memory makes every warp store touch ~20 sectors for 32 words, one partial-sector
// write per id at the L2: 29 M sector operations per 100 MB of cfg2.  Measured, that was NOT what bounds the kernel:
// cfg2 / cfg3 / cfg4 are unchanged (80 / 102 / 107 us against 80 / 103 / 104 us at 64 MB), cfg5 -- two ids per
// piece -- gains 6 % (272 against 288 us).  The kernel waits on its barriers and on the pv -> mlist -> pool chain and
// issues 415 instructions per warp and tile.  k_emit_direct below is the other version, kept for the A/B measurement:
// SPL_EMIT_STAGE=0.)
// ------------------------------------------------------------------------------------------
#define EM_ROUNDS (SPL_TILE / (SPL_THREADS * 4))     // 4 rounds cover the 4096 pieces a tile can have
#define EM_WARPS (SPL_THREADS / 32)
#define EM_INLINE 8u                                 // ids of a merged piece copied by its own thread up to this many
#define EM_BIGCAP (SPL_TILE / (EM_INLINE + 1u) + 1u) // pieces of a tile that can have more ids than that
#define EM_STAGE 3072u                               // ids assembled at a time (a tile with more takes several phases)

struct EmitSmem {
    // ids of the warp's round before piece j.  16 bits: every piece but the tile's last one ends inside the tile, so
    // the ids in front of any piece of the tile are at most SPL_TILE
    __align__(16) uint16_t spos[SPL_TILE + 8];
    uint32_t stage[EM_STAGE];
    uint32_t pbw[SPL_TILE / 32];
    uint32_t wpre[SPL_TILE / 32];
    uint32_t wtot[EM_ROUNDS * EM_WARPS];             // ids of each (round, warp)
    uint32_t wexc[EM_ROUNDS * EM_WARPS];             // ids of the tile before each (round, warp)
    uint16_t bigj[EM_BIGCAP];                        // pieces with more than EM_INLINE ids
    uint32_t n_big;
    uint64_t prefix;
};

__device__ __forceinline__ uint32_t emit_count(const SplWork& w, uint32_t v, bool valid) {
    if (!valid) return 0u;
    if (v < SPL_PV_MISS) return 1u;
    if (v == SPL_PV_NONE) return 0u;
    if (w.charref && SPL_PV_IS_CHARREF(v)) return ((v >> 28) & 3u) + 1u;
    return (uint32_t)(w.mlist[v & ~SPL_PV_MISS] >> 32) & SPL_ML_LEN_MASK;
}

// ids [lo, lo + EM_STAGE) of the tile into sm.stage.  CHECK = false: the tile has at most EM_STAGE ids (lo == 0).
template <bool CHECK>
__device__ __forceinline__ void emit_stage(const SplWork& w, EmitSmem& sm, const uint32_t* __restrict__ pv, const uint4 v_first,
                                           const uint32_t P, const uint32_t rounds, const uint32_t my_excl, const uint32_t lo) {
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    auto put = [&](uint32_t pos, uint32_t x) {
        const uint32_t r = pos - lo;
        if (!CHECK || r < EM_STAGE) sm.stage[r] = x;
    };
    for (uint32_t k = 0; k < rounds; ++k) {
        const uint32_t j4 = (k * SPL_THREADS + tid) * 4u;
        const uint32_t base_k = __shfl_sync(FULL, my_excl, k * EM_WARPS + warp);
        if (j4 >= P) continue;
        uint32_t pos = sm.spos[j4] + base_k;
        if (CHECK && lo != 0u && pos >= lo + EM_STAGE) continue;  // all beyond the phase (phase 0 visits every piece: bigj)
        uint4 v = v_first;
        if (k > 0 || tid >= 128) v = __ldg(reinterpret_cast<const uint4*>(pv + j4));
        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            if (j4 + q >= P) break;
            const uint32_t x = vv[q];
            if (x < SPL_PV_MISS) put(pos++, x);
            else if (w.charref && SPL_PV_IS_CHARREF(x)) {
                const uint32_t* __restrict__ src = w.T->char_ids + (x & 0x0FFFFFFFu);
                const uint32_t c = ((x >> 28) & 3u) + 1u;
                uint32_t t[4];                                   // all loads first: they are in flight together
#pragma unroll
                for (uint32_t r = 0; r < 4u; ++r) t[r] = r < c ? __ldg(src + r) : 0u;
#pragma unroll
                for (uint32_t r = 0; r < 4u; ++r) if (r < c) put(pos + r, t[r]);
                pos += c;
            } else if (x != SPL_PV_NONE) {
                const uint64_t e = w.mlist[x & ~SPL_PV_MISS];
                const uint32_t gp = (uint32_t)e, c = (uint32_t)(e >> 32) & SPL_ML_LEN_MASK;
                if (c <= EM_INLINE) {
                    for (uint32_t r0 = 0; r0 < c; r0 += 4u) {       // four loads in flight, then their stores
                        uint32_t t[4];
#pragma unroll
                        for (uint32_t r = 0; r < 4u; ++r) t[r] = r0 + r < c ? w.pool[gp + r0 + r] : 0u;
#pragma unroll
                        for (uint32_t r = 0; r < 4u; ++r) if (r0 + r < c) put(pos + r0 + r, t[r]);
                    }
                } else if (lo == 0) {
                    sm.bigj[atomicAdd(&sm.n_big, 1u)] = (uint16_t)(j4 + q);      // by a warp, after the barrier
                }
                pos += c;
            }
        }
    }
}

__global__ void __launch_bounds__(SPL_THREADS, 8) k_emit(SplWork w) {
    __shared__ EmitSmem sm;
    SPL_PDL_ENTER();
    if (w.counters[SPL_CTR_ERR] & SPL_DEVERR_OFFSETS) {            // rejected offsets: nothing was computed; deliver the flag
        if (blockIdx.x == 0 && threadIdx.x == 0 && w.host_meta) { w.host_meta[0] = 0; w.host_meta[2] = 0; w.host_meta[1] = w.counters[SPL_CTR_ERR]; }
        return;
    }
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
        if ((k * SPL_THREADS + warp * 32u) * 4u >= P) {            // the whole warp is past the last piece (warp-uniform)
            if (lane == 31) sm.wtot[k * EM_WARPS + warp] = 0u;
            continue;
        }
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
        const uint32_t ex = incl - c, e1 = ex + c0, e2 = e1 + c1, e3 = e2 + c2;
        // (an entry behind the tile's last piece -- the one piece whose count is not bounded by the tile -- may be
        // truncated: it is never read)
        *reinterpret_cast<uint2*>(&sm.spos[j4]) = make_uint2((ex & 0xFFFFu) | (e1 << 16), (e2 & 0xFFFFu) | (e3 << 16));
        if (lane == 31) sm.wtot[k * EM_WARPS + warp] = incl;
    }
    __syncthreads();
    // every warp turns the (<= 32) warp totals into exclusive offsets for itself: no serial scan, no second barrier
    uint32_t my_tot = lane < rounds * EM_WARPS ? sm.wtot[lane] : 0u, my_incl = my_tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, my_incl, o);
        if (lane >= (uint32_t)o) my_incl += t;
    }
    const uint32_t my_excl = my_incl - my_tot;                   // lane l: ids before (round, warp) pair l
    const uint32_t tile_total = __shfl_sync(FULL, my_incl, 31);
    if (d1 > d0) {                                               // for the document offsets at the end (after the next barrier)
        if (warp == 0) sm.wexc[lane] = my_excl;
        if (warp == 1) {
            // word prefixes of the piece bits (document start -> piece index)
            uint32_t loc[4], run = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) { loc[q] = __popc(sm.pbw[lane * 4 + q]); run += loc[q]; }
            uint32_t incl = run;
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

    // ---- pass 2: ids to their place in the staging buffer, then out in whole lines ------------------------
    const uint64_t prefix = sm.prefix;
    uint32_t* __restrict__ out = w.ids + prefix;
    for (uint32_t lo = 0; lo < tile_total; lo += EM_STAGE) {     // (block-uniform; one phase unless the tile is dense in ids)
        if (tile_total <= EM_STAGE) emit_stage<false>(w, sm, pv, v_first, P, rounds, my_excl, 0u);
        else emit_stage<true>(w, sm, pv, v_first, P, rounds, my_excl, lo);
        __syncthreads();
        const uint32_t nb = sm.n_big;
        if (nb) {                                                // one warp per long piece: pool -> staging buffer
            for (uint32_t b = warp; b < nb; b += EM_WARPS) {
                const uint32_t j = sm.bigj[b];
                const uint32_t pos = sm.spos[j] + __shfl_sync(FULL, my_excl, (j / (SPL_THREADS * 4u)) * EM_WARPS + ((j / 128u) & (EM_WARPS - 1u)));
                const uint64_t e = w.mlist[__ldg(pv + j) & ~SPL_PV_MISS];
                const uint32_t gp = (uint32_t)e, c = (uint32_t)(e >> 32) & SPL_ML_LEN_MASK;
                // the part of [pos, pos + c) inside this phase
                const uint32_t q0 = pos < lo ? lo - pos : 0u;
                const uint32_t q1 = pos + c > lo + EM_STAGE ? (lo + EM_STAGE > pos ? lo + EM_STAGE - pos : 0u) : c;
                for (uint32_t q = q0 + lane; q < q1; q += 32) sm.stage[pos + q - lo] = w.pool[gp + q];
            }
            __syncthreads();
        }
        const uint32_t n = tile_total - lo < EM_STAGE ? tile_total - lo : EM_STAGE;
        for (uint32_t i = tid; i < n; i += SPL_THREADS) out[lo + i] = sm.stage[i];
        if (lo + EM_STAGE < tile_total) __syncthreads();         // the buffer is refilled
    }
    if (tile_total == 0u) __syncthreads();                       // wexc / wpre for the document offsets

    // ---- output offset of every document that starts in this tile ----------------------------------------
    for (uint32_t d = d0 + tid; d < d1; d += SPL_THREADS) {
        const uint32_t x = (uint32_t)(w.doc_off[d] - w.off_base - tile0);
        const uint32_t pi = sm.wpre[x >> 5] + __popc(sm.pbw[x >> 5] & ((1u << (x & 31)) - 1u));
        const uint32_t rel = pi >= P ? tile_total : sm.spos[pi] + sm.wexc[(pi / (SPL_THREADS * 4u)) * EM_WARPS + ((pi / 128u) & (EM_WARPS - 1u))];
        // pipelined host call: the ids of the shard's earlier chunks (kept on the device, so that the host gets
        // shard-relative offsets in one copy at the end instead of one small copy and a rebase per chunk)
        const uint64_t base = w.tok_base_in ? *w.tok_base_in : 0ull;
        w.out_off[d] = base + prefix + rel;
        if (d == w.n_docs) {
            if (w.tok_total_out) *w.tok_total_out = base + prefix + rel;
            if (w.host_meta) {                                 // the call's summary, straight to the host (no copy on this stream)
                w.host_meta[0] = prefix + rel;
                w.host_meta[2] = *reinterpret_cast<const uint64_t*>(&w.counters[SPL_CTR_HUGE_POOL]);
                w.host_meta[1] = w.counters[SPL_CTR_ERR];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_emit_direct: the same with every thread storing its ids straight to global memory (SPL_EMIT_STAGE=0).
// ------------------------------------------------------------------------------------------

struct EmitDirectSmem {
    __align__(16) uint32_t spos[SPL_TILE + 4];       // ids of the warp's round before piece j (see wtot)
    uint32_t pbw[SPL_TILE / 32];
    uint32_t wpre[SPL_TILE / 32];
    uint32_t wtot[EM_ROUNDS * EM_WARPS];             // ids of each (round, warp)
    uint32_t wexc[EM_ROUNDS * EM_WARPS];             // ids of the tile before each (round, warp)
    uint32_t bigpos[EM_BIGCAP], biggp[EM_BIGCAP], bigcnt[EM_BIGCAP];
    uint32_t n_big;
    uint64_t prefix;
};

__global__ void __launch_bounds__(SPL_THREADS, 8) k_emit_direct(SplWork w) {
    __shared__ EmitDirectSmem sm;
    SPL_PDL_ENTER();
    if (w.counters[SPL_CTR_ERR] & SPL_DEVERR_OFFSETS) {            // rejected offsets: nothing was computed; deliver the flag
        if (blockIdx.x == 0 && threadIdx.x == 0 && w.host_meta) { w.host_meta[0] = 0; w.host_meta[2] = 0; w.host_meta[1] = w.counters[SPL_CTR_ERR]; }
        return;
    }
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
        if ((k * SPL_THREADS + warp * 32u) * 4u >= P) {            // the whole warp is past the last piece (warp-uniform)
            if (lane == 31) sm.wtot[k * EM_WARPS + warp] = 0u;
            continue;
        }
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
    // every warp turns the (<= 32) warp totals into exclusive offsets for itself: no serial scan, no second barrier
    uint32_t my_tot = lane < rounds * EM_WARPS ? sm.wtot[lane] : 0u, my_incl = my_tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, my_incl, o);
        if (lane >= (uint32_t)o) my_incl += t;
    }
    const uint32_t my_excl = my_incl - my_tot;                   // lane l: ids before (round, warp) pair l
    const uint32_t tile_total = __shfl_sync(FULL, my_incl, 31);
    if (d1 > d0) {                                               // for the document offsets at the end (after the next barrier)
        if (warp == 0) sm.wexc[lane] = my_excl;
        if (warp == 1) {
            // word prefixes of the piece bits (document start -> piece index)
            uint32_t loc[4], run = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) { loc[q] = __popc(sm.pbw[lane * 4 + q]); run += loc[q]; }
            uint32_t incl = run;
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

    // ---- pass 2: ids to their place ----------------------------------------------------------------------
    const uint64_t prefix = sm.prefix;
    uint32_t* __restrict__ out = w.ids + prefix;
    for (uint32_t k = 0; k < rounds; ++k) {
        const uint32_t j4 = (k * SPL_THREADS + tid) * 4u;
        const uint32_t base_k = __shfl_sync(FULL, my_excl, k * EM_WARPS + warp);
        if (j4 < P) {
            uint4 v = v_first;
            if (k > 0 || tid >= 128) v = __ldg(reinterpret_cast<const uint4*>(pv + j4));
            uint32_t pos = sm.spos[j4] + base_k;
            const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) {
                if (j4 + q < P) {
                    const uint32_t x = vv[q];
                    if (x < SPL_PV_MISS) out[pos++] = x;
                    else if (w.charref && SPL_PV_IS_CHARREF(x)) {
                        const uint32_t* __restrict__ src = w.T->char_ids + (x & 0x0FFFFFFFu);
                        const uint32_t c = ((x >> 28) & 3u) + 1u;
                        for (uint32_t r = 0; r < c; ++r) out[pos + r] = __ldg(src + r);
                        pos += c;
                    } else if (x != SPL_PV_NONE) {
                        const uint64_t e = w.mlist[x & ~SPL_PV_MISS];
                        const uint32_t gp = (uint32_t)e, c = (uint32_t)(e >> 32) & SPL_ML_LEN_MASK;
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
    __syncthreads();
    {
        const uint32_t nb = sm.n_big;                          // one warp per long piece
        for (uint32_t b = warp; b < nb; b += EM_WARPS) {
            const uint32_t pos = sm.bigpos[b], gp = sm.biggp[b], c = sm.bigcnt[b];
            for (uint32_t q = lane; q < c; q += 32) out[pos + q] = w.pool[gp + q];
        }
    }

    // ---- output offset of every document that starts in this tile ----------------------------------------
    for (uint32_t d = d0 + tid; d < d1; d += SPL_THREADS) {
        const uint32_t x = (uint32_t)(w.doc_off[d] - w.off_base - tile0);
        const uint32_t pi = sm.wpre[x >> 5] + __popc(sm.pbw[x >> 5] & ((1u << (x & 31)) - 1u));
        const uint32_t rel = pi >= P ? tile_total : sm.spos[pi] + sm.wexc[(pi / (SPL_THREADS * 4u)) * EM_WARPS + ((pi / 128u) & (EM_WARPS - 1u))];
        // pipelined host call: the ids of the shard's earlier chunks (kept on the device, so that the host gets
        // shard-relative offsets in one copy at the end instead of one small copy and a rebase per chunk)
        const uint64_t base = w.tok_base_in ? *w.tok_base_in : 0ull;
        w.out_off[d] = base + prefix + rel;
        if (d == w.n_docs) {
            if (w.tok_total_out) *w.tok_total_out = base + prefix + rel;
            if (w.host_meta) {                                 // the call's summary, straight to the host (no copy on this stream)
                w.host_meta[0] = prefix + rel;
                w.host_meta[2] = *reinterpret_cast<const uint64_t*>(&w.counters[SPL_CTR_HUGE_POOL]);
                w.host_meta[1] = w.counters[SPL_CTR_ERR];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
void spl_encode_init() {
    cudaFuncSetAttribute(k_bpe_long, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BPE_SMEM_BYTES);
    cudaFuncSetAttribute(k_bpe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WK_SMEM_BYTES);
    cudaFuncSetAttribute(k_probe<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
    cudaFuncSetAttribute(k_probe<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
    cudaFuncSetAttribute(k_emit, cudaFuncAttributePreferredSharedMemoryCarveout, 75);
    cudaFuncSetAttribute(k_emit_direct, cudaFuncAttributePreferredSharedMemoryCarveout, 75);
    cudaGetLastError();
}

int spl_describe_encode_stage(const SplWork& w, int num_sms, SplLaunchDesc* out) {
    // staging by bulk copy (TMA engine) or through registers: SPL_PROBE_BULK=0 selects the latter (A/B measurements)
    static const bool bulk = [] { const char* e = getenv("SPL_PROBE_BULK"); return !e || e[0] != '0'; }();
    // first length class merged by windowed rounds (measured crossover; SPL_BPE_WIN_CLS overrides it for experiments)
    static const uint32_t win_cls = [] { const char* e = getenv("SPL_BPE_WIN_CLS"); return e ? (uint32_t)atoi(e) : SPL_BPE_WIN_CLS; }();
    // ids assembled in shared memory and written in whole lines, or stored by the thread that has them (A/B)
    static const bool emit_stage_on = [] { const char* e = getenv("SPL_EMIT_STAGE"); return !e || e[0] != '0'; }();
    // persistent grids, but no larger than the text can feed: a small batch must not pay for launching (and draining)
    // hundreds of idle blocks
    const uint32_t g_bpe = std::min<uint32_t>((uint32_t)num_sms * 6u, std::max<uint32_t>(1u, w.N / 2048u));
    const uint32_t g_long = std::min<uint32_t>((uint32_t)num_sms * 4u, std::max<uint32_t>(1u, w.N / 4096u));
    const uint32_t g_fin = std::min<uint32_t>((uint32_t)num_sms, std::max<uint32_t>(1u, w.n_tiles / 8u));
    int n = 0;
    out[n++] = SplLaunchDesc{bulk ? (const void*)k_probe<true> : (const void*)k_probe<false>, "k_probe", w.n_tiles, SPL_THREADS, 0, false, 0};
    out[n++] = SplLaunchDesc{(const void*)k_bpe, "k_bpe", g_bpe, SPL_BPE_THREADS, WK_SMEM_BYTES, false, 0};
    out[n++] = SplLaunchDesc{(const void*)k_bpe_long, "k_bpe_long", g_long, SPL_BPE_THREADS, BPE_SMEM_BYTES, true, win_cls};
    out[n++] = SplLaunchDesc{(const void*)k_bpe_fin, "k_bpe_fin", g_fin, SPL_BPE_THREADS, 0, false, 0};      // duplicates + the chunk scan, by its last block
    out[n++] = SplLaunchDesc{emit_stage_on ? (const void*)k_emit : (const void*)k_emit_direct, "k_emit", w.n_tiles, SPL_THREADS, 0, false, 0};
    return n;
}
