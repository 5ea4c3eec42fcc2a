// Kernel launch interface between spl_api.cu (C-ABI, memory management) and
// spl_kernels.cu (the device encode path).
#pragma once
#include <cuda_runtime.h>
#include "spl_common.h"

#define SPL_TILE 4096u          // text bytes per tile (both kernels)
#define SPL_HALO 256u           // bytes staged beyond a tile's end
#define SPL_WIN  (SPL_TILE + SPL_HALO)
#define SPL_THREADS 256
// bit-parallel pre-tokenizer: one thread per 32-byte word, halo words of context on both sides
#define SPL_FAST_HALO    16u
#define SPL_FAST_PAYLOAD 256u         // words per tile = 8192 bytes = 2 * SPL_TILE
#define SPL_FAST_THREADS (SPL_FAST_PAYLOAD + 2u * SPL_FAST_HALO)

// per-call device workspace (all pointers on the current device)
struct SplWork {
    const uint8_t*  text;        // [N], 16-byte aligned, readable up to N rounded up to 16
    uint32_t        N;
    const uint64_t* doc_off;     // [n_docs+1] document starts; doc_off[d] - off_base indexes text
    uint64_t        off_base;
    uint32_t        n_docs;
    uint32_t        n_tiles;     // N / SPL_TILE + 1
    uint32_t*       hard;        // bitmap words: segment boundaries (doc starts, special-span edges, N)
    uint32_t*       spec;        // bitmap words: bytes inside special-token spans (with_special only)
    uint32_t*       pstart;      // bitmap words: piece starts (incl. sentinel bit N)
    size_t          bitmap_words;
    uint32_t*       tile_first_doc;   // [n_tiles+1]
    uint64_t*       tile_state;       // [n_tiles] exclusive prefix of tile_cnt (k_tile_scan)
    uint32_t*       tile_cnt;         // [n_tiles] ids produced by each tile
    uint64_t*       tile_soff;        // [n_tiles] where the tile's ids sit in `stage`
    uint32_t*       stage;            // [>= N] ids in tile-completion order
    uint64_t*       stage_bump;       // bump allocator of `stage` (counters + 8 bytes .. 16-byte aligned)
    uint32_t*       counters;         // [8]: 0 = encode ticket, 1 = error flags, 2 = huge pool bump, 3 = fallback tiles
    uint32_t*       fb_list;          // [n_fast_tiles] fast-path tiles handed to the sequential rules
    uint32_t        n_fast_tiles;
    uint32_t*       huge_pool;        // scratch for pieces that outgrow the staging window
    uint32_t        huge_pool_words;
    uint32_t*       ids;              // [>= N]
    uint64_t*       out_off;          // [n_docs+1]
    const SplTables* T;               // device copy of the tables
    int             pattern;
    bool            with_special;
};

enum : uint32_t { SPL_DEVERR_OFFSETS = 1u, SPL_DEVERR_HUGE_POOL = 2u };

// optional per-kernel device timing: ev[i] is recorded before kernel i, ev[n] after the last
#define SPL_PROF_MAX 8
struct SplKernelProfile {
    cudaEvent_t ev[SPL_PROF_MAX + 1];
    const char* name[SPL_PROF_MAX];
    int n;
};

// Enqueue the whole encode path on `stream`.  Returns the number of kernels launched.
int spl_launch_encode(const SplWork& w, int num_sms, cudaStream_t stream, SplKernelProfile* prof = nullptr);

// Per-device one-time kernel attribute setup (shared-memory carveout).
void spl_kernels_init();
