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

// encode stage (spl_encode.cu)
#define SPL_PROBE_HALO  128u          // bytes staged beyond a tile for the whole-piece probe (bundled keys are <= 128 B)
#define SPL_PROBE_WIN   (SPL_TILE + SPL_PROBE_HALO)
#define SPL_CHUNK_TILES 32u           // tiles per chunk of the two-level id-count prefix
#define SPL_BPE_THREADS 128           // k_bpe block size
// Length classes of the pieces that go through the merge loop.  A piece of class c is merged by a group of
// 2^spl_class_log2group(c) lanes (32 parts per lane), so a warp always works on 32 / group pieces side by side:
//   class  0      1       2       3        4         5          6           7
//   bytes  2..16  17..32  33..64  65..128  129..256  257..512   513..1024   1025.. (whole block, global scratch)
//   group  1      1       2       4        8         16         32          -
#define SPL_NCLS 8
#define SPL_GROUP_MAXLEN 1024u
__host__ __device__ inline uint32_t spl_len_class(uint32_t len) {
    return len <= 16u ? 0u : len <= 32u ? 1u : len <= 64u ? 2u : len <= 128u ? 3u : len <= 256u ? 4u :
           len <= 512u ? 5u : len <= SPL_GROUP_MAXLEN ? 6u : 7u;
}
__host__ __device__ inline uint32_t spl_class_minlen(uint32_t c) {
    return c == 0 ? 2u : c == 7 ? SPL_GROUP_MAXLEN + 1u : (16u << (c - 1)) + 1u;
}
__host__ __device__ inline uint32_t spl_class_log2group(uint32_t c) { return c <= 1u ? 0u : c - 1u; }

// values of SplWork::pv (one per piece, in text order)
#define SPL_PV_MISS  0x80000000u      // | index of the piece's entry in `mlist`
#define SPL_PV_NONE  0xFFFFFFFFu      // the piece produces no id (byte unknown to the vocabulary)
// a single character of two or three ids (SplWork::charref): 0xC0000000 | (ids - 1) << 28 | index of the first id in
// SplTables::char_ids.  Top nibble C..E; needs miss-list indices below 2^30 (the host checks) -- F is NONE / internal.
#define SPL_PV_CHARREF 0xC0000000u
#define SPL_PV_IS_CHARREF(v) ((v) >= 0xC0000000u && (v) < 0xF0000000u)
// miss-list entry before the merge kernels:  gpos:32 | len:31
// miss-list entry after:                      gpos:32 | id count:31 | SPL_ML_DONE    (ids at pool[gpos ..])
// SPL_ML_SEG: the entry is a SEGMENT of a piece (spl_segment.h): the whole-piece probe does not apply to it
// SPL_ML_DUP: same bytes as an earlier entry (dup_of[] names it): k_bpe_fin copies that entry's result
#define SPL_ML_LEN_MASK 0x1FFFFFFFu
#define SPL_ML_DONE (1ull << 63)
#define SPL_ML_SEG  (1ull << 62)
#define SPL_ML_DUP  (1ull << 61)

// one record per tile, so that k_emit learns everything about its tile (and its chunk) from one line
struct SplTileInfo {
    uint32_t np;          // pieces that start in the tile (k_probe)
    int32_t  extra;       // ids minus pieces (k_probe: dropped bytes, k_bpe: merged pieces)
    uint32_t first_doc;   // first document that starts at or after the tile (k_mark_docs); [n_tiles] = n_docs + 1
    uint32_t pad;
};

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
    uint32_t*       cand;        // bitmap words: starts of special-string occurrences; only for sets whose strings can overlap (spl_special.h), else nullptr
    uint32_t*       pstart;      // bitmap words: piece starts (incl. sentinel bit N)
    size_t          bitmap_words;
    SplTileInfo*    tinfo;            // [n_tiles+1] per-tile record (zero-initialised)
    int32_t*        chunk_cnt;        // [n_tiles / SPL_CHUNK_TILES + 1] ids of the chunk's tiles (kept by k_probe and k_bpe)
    uint64_t*       chunk_state;      // [n_tiles / SPL_CHUNK_TILES + 1] exclusive prefix of chunk_cnt (chunk_scan_block, last block of k_bpe_long)
    uint32_t*       pv;               // [n_tiles * SPL_TILE] per-piece value, tile t at pv[t * SPL_TILE ..]
    uint32_t*       pool;             // [N] ids of the pieces that went through the merge loop, at the piece's byte position
    uint64_t*       mlist;            // miss lists, one region per length class
    uint32_t        ml_base[SPL_NCLS + 1];   // class c owns mlist[ml_base[c] .. ml_base[c+1])
    uint32_t*       counters;         // [SPL_CTR_WORDS]: see SPL_CTR_*
    uint32_t*       fb_list;          // [n_fast_tiles] fast-path tiles handed to the sequential rules
    uint32_t        n_fast_tiles;
    // duplicate long pieces (the merge loop runs once per distinct byte string, like the reference's chunk cache,
    // tokenizer.rs:667-724): open-addressing table hash tag:32 | miss-list index + 1, zero-initialised; dd_tab == nullptr: off
    unsigned long long* dd_tab;
    uint32_t        dd_mask;          // slots - 1
    uint32_t*       dup_of;           // [entries of the classes >= 2] miss-list index of the entry with the same bytes
    uint32_t*       huge_pool;        // scratch for pieces that outgrow shared memory
    uint64_t        huge_pool_words;
    uint32_t*       ids;              // [>= N]
    uint64_t*       out_off;          // [n_docs+1]
    uint64_t*       host_meta;        // optional, mapped pinned host memory: [0] id count, [1] error flags, [2] words of huge-piece scratch the pass asked for
    const uint64_t* tok_base_in;      // optional: ids of the shard's earlier chunks, added to out_off (not to the id positions)
    uint64_t*       tok_total_out;    // optional: receives *tok_base_in + this pass's id count
    const SplTables* T;               // device copy of the tables
    int             pattern;
    bool            with_special;
    bool            charref;          // pv may hold SPL_PV_CHARREF values (else such characters get a settled miss-list entry)
    bool            pretok_done;      // pstart / spec are already filled in (SentencePiece mode: written by k_sp_emit)
};

// SplWork::counters
enum : uint32_t { SPL_CTR_ERR = 1, SPL_CTR_HUGE_POOL = 2 /* 64 bit: words 2 and 3 */, SPL_CTR_FB = 4,
                  SPL_CTR_DUP = 5,            // long pieces that were duplicates of an earlier one
                  SPL_CTR_REFINED = 6,        // tiles that went through k_probe's refining pass
                  SPL_CTR_FINAL = 7,          // entries k_probe filed already settled (characters of two or three ids): from the top of class 0's region down
                  SPL_CTR_CLS = 8,            // [8 .. 8 + SPL_NCLS): entries in the miss list of each class
                  SPL_CTR_TICKET = 16,        // blocks of k_bpe_fin that are done (the last one scans the chunk totals)
                  SPL_CTR_WORDS = 32 };

enum : uint32_t { SPL_DEVERR_OFFSETS = 1u, SPL_DEVERR_HUGE_POOL = 2u };
// Every kernel behind k_mark_docs starts with this: document offsets that k_mark_docs rejected must never be used as
// indices (the host reports SPL_ERR_INVALID_ARG from the flag; k_emit still delivers it)
// Programmatic dependent launch: the kernels of a pass are the nodes of a graph whose edges are of type Programmatic
// (spl_api.cu: enqueue_encode_graph).  A kernel first waits for the grid in front of it (all of its memory is visible
// then), and at once lets the grid behind it be scheduled: that grid's blocks come in as this one's last blocks leave
// and sit at their own wait -- the launch latency and the ramp of every kernel are hidden in the tail of the one
// before.  Both instructions do nothing in a plain launch.  Off by default (SPL_PDL=1): inside a graph the gaps between
// the kernels are already too small for this to show, and on cfg4 the early blocks of the kernels behind k_bpe_long take
// resources from its tail (0.80 -> 0.89 ms).
#if defined(__CUDA_ARCH__)
#define SPL_PDL_ENTER() do { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); } while (0)
#else
#define SPL_PDL_ENTER() do { } while (0)
#endif
#define SPL_RETURN_IF_BAD_OFFSETS(w) do { SPL_PDL_ENTER(); if ((w).counters[SPL_CTR_ERR] & SPL_DEVERR_OFFSETS) return; } while (0)

// optional per-kernel device timing: ev[i] is recorded before kernel i, ev[n] after the last
#define SPL_PROF_MAX 16
struct SplKernelProfile {
    cudaEvent_t ev[SPL_PROF_MAX + 1];
    const char* name[SPL_PROF_MAX];
    int n;
};

// One kernel launch of the encode path, described instead of performed: the same list feeds plain launches
// (spl_launch_encode) and the nodes of a CUDA graph (spl_api.cu: one graph launch per pipeline chunk).
// Every kernel takes the SplWork by value; k_bpe_long a second u32 argument.
#define SPL_MAX_LAUNCHES 16
struct SplLaunchDesc {
    const void* func; const char* name;
    unsigned grid, block; size_t smem;
    bool has_u32; uint32_t u32;
};
int spl_describe_encode(const SplWork& w, int num_sms, SplLaunchDesc* out);      // returns the number of launches
int spl_describe_encode_stage(const SplWork& w, int num_sms, SplLaunchDesc* out);

// Enqueue the whole encode path on `stream`.  Returns the number of kernels launched.
int spl_launch_encode(const SplWork& w, int num_sms, cudaStream_t stream, SplKernelProfile* prof = nullptr);

// Per-device one-time kernel attribute setup (shared-memory carveout).
void spl_kernels_init();

// spl_sentencepiece.cu: SentencePiece mode (tokenizer.rs:737-795).  Text T -> transformed text T' + its bitmaps.
// All bitmaps are over T (zero-initialised by the caller); hard / spec / tinfo are filled by k_mark_docs /
// k_mark_specials run on a view of T before spl_launch_sp_scan.
#define SPL_SPCTR_CONV 4u             // counters[]: number of converted spaces (N' = N + 2 * that)
struct SplSpWork {
    const uint8_t*  text; uint32_t N;
    const uint64_t* doc_off; uint64_t off_base; uint32_t n_docs;
    uint32_t        n_tiles;                    // N / SPL_TILE + 1
    const uint32_t* hard; const uint32_t* spec; // spec == nullptr without special tokens
    const SplTileInfo* tinfo;                   // first_doc per tile
    uint32_t *w0, *a, *rs;                      // bitmaps, see spl_sentencepiece.h
    uint32_t *tile_cnt, *tile_pref;             // [n_tiles], [n_tiles + 1]
    uint32_t* counters;
    const SplTables* T;
    // outputs (spl_launch_sp_emit)
    uint8_t*  text2; uint32_t* pstart2; uint32_t* spec2; uint64_t* doc_off2;
};
int spl_launch_sp_scan(const SplSpWork& s, cudaStream_t stream);    // classify, raw runs, count, scan
int spl_launch_sp_emit(const SplSpWork& s, cudaStream_t stream);
// k_mark_docs (+ k_mark_specials) only: the segment bitmaps of a text (SentencePiece mode runs them over T)
int spl_launch_mark(const SplWork& w, int num_sms, cudaStream_t stream);

// spl_ingest.cu: JSON Lines -> packed text + offsets (row N4; contract in spl_ingest.h)
enum : uint32_t { SPL_JLCTR_NEWLINES = 0, SPL_JLCTR_DOCS = 1, SPL_JLCTR_MISSING = 2, SPL_JLCTR_BAD = 3, SPL_JLCTR_TEXT = 4 };
struct SplJlSpan;
struct SplJlWork {
    const uint8_t* text; uint32_t N;              // file bytes, 16-byte aligned, readable up to N rounded up to 16
    uint32_t n_tiles;                             // N / SPL_TILE + 1
    uint32_t* tile_cnt; uint32_t* tile_pref;      // newlines per tile, exclusive prefix [n_tiles + 1]
    uint32_t n_lines;                             // newlines + 1 (known after spl_launch_jsonl_count)
    uint32_t* line_start;                         // [n_lines + 1]
    SplJlSpan* span;                              // [n_lines]
    uint32_t *is_doc, *out_len;                   // [n_lines]
    uint32_t *doc_idx, *text_off;                 // [n_lines + 1] exclusive prefixes
    uint32_t* counters;                           // SPL_JLCTR_* (zero-initialised)
    uint8_t field[64]; uint32_t flen;
    uint8_t* out_text; uint64_t text_capacity;
    uint64_t* out_off; uint64_t off_capacity;     // entries
};
int spl_launch_jsonl_count(const SplJlWork& w, cudaStream_t stream);     // k_jl_count, k_jl_scan
int spl_launch_jsonl_extract(const SplJlWork& w, cudaStream_t stream);   // k_jl_lines, k_jl_parse, k_jl_scan2, k_jl_emit

// spl_parquet.cu: the pages of one Parquet string column -> packed text + offsets (row N4; formats: spl_parquet.h)
#define SPL_PQCTR_ERR 0                // SPL_PQ_ERR_* bits
#define SPL_PQCTR_TEXT 2               // (64-bit) bytes of packed text
struct SplPqPage;
struct SplPqWork {
    const uint8_t* file;               // staged file bytes (the column chunks of the batch), padded by 16 bytes
    uint8_t* scratch;                  // decompressed pages, padded by 16 bytes
    const SplPqPage* pages;
    uint64_t* row_off; uint32_t* row_len;      // [n_rows] span of every row
    uint64_t* dict_off; uint32_t* dict_len;    // [dict entries of the batch]
    uint64_t n_rows;
    uint32_t n_blocks;                 // max(1, ceil(n_rows / 2048))
    unsigned long long* bsum;          // [n_blocks + 1]
    uint32_t* counters;                // SPL_PQCTR_* (zero-initialised, 8 words)
    uint8_t* out_text;                 // packed text (known size: counters[SPL_PQCTR_TEXT] after spl_launch_pq_spans)
    uint64_t* out_off;                 // [n_rows + 1]
};
int spl_launch_pq_spans(const SplPqWork& w, uint32_t first_page, uint32_t n_pages, bool has_dict, bool has_snappy, cudaStream_t stream);
int spl_launch_pq_copy(const SplPqWork& w, uint64_t text_bytes, cudaStream_t stream);

// spl_decode.cu: ids -> bytes (row N2)
#define SPL_DEC_TILE 2048u            // ids per tile
struct SplDecLaunch {
    const uint32_t* ids; uint64_t n_tok;
    const uint64_t* tok_off; uint64_t n_docs;
    uint32_t n_tiles;                  // n_tok / SPL_DEC_TILE + 1
    uint32_t* tile_sum; uint64_t* tile_pref;    // [n_tiles], [n_tiles + 1] (last = total bytes)
    uint8_t* out; uint64_t capacity; uint64_t* out_off;
    const SplTables* T;
};
void spl_launch_decode_count(const SplDecLaunch& L, cudaStream_t stream);   // k_dec_len + k_dec_scan
void spl_launch_decode_emit(const SplDecLaunch& L, cudaStream_t stream);    // k_dec_emit (no-op on the device if capacity is too small)

// spl_encode.cu: the encode stage behind the pre-tokenizer (k_probe, k_bpe, k_tile_scan, k_emit)
void spl_encode_init();

// per-tile scratch of the whole-piece probe (spl_encode.cu: probe_tile)
struct SplProbeScratch {
    uint32_t spw[SPL_TILE / 32];              // special-span bits of the tile (with_special)
    uint16_t plist[SPL_TILE + 2];             // window positions of the tile's piece starts, in order (+ end of the last piece)
    uint16_t slow[SPL_TILE];                  // pieces the one-sector probe did not settle (each warp: its own range)
    uint16_t mloc[SPL_TILE];                  // missed pieces: class 0 from the bottom of the warp's range, class 1 from its top
    uint32_t wtot[SPL_THREADS / 32];
    uint32_t last_end;                        // window position of the end of the tile's last piece
    uint32_t segw[SPL_TILE / 32];             // refining pass: safe boundaries found inside missed pieces (new piece starts)
    uint32_t mstw[SPL_TILE / 32];             // refining pass: starts of the pieces the whole-piece probe missed
    uint32_t inmw[SPL_TILE / 32];             // refining pass: bytes that belong to those pieces
    uint32_t whi[SPL_THREADS / 32];           // bytes >= 0x80 of the tile per warp (decides whether the tile is refined)
    uint32_t cls_cnt[SPL_NCLS + 1];           // refining pass: misses of the block per length class (+ settled multi-id characters), then list bases
};
