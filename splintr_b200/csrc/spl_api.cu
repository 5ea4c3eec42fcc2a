// C-ABI of splintr_b200 (include/splintr_b200.h): handle lifetime, device tables,
// per-call workspace, host<->device staging, document sharding over the handle's devices.
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/splintr_b200.h"
#include "spl_host.h"
#include "spl_kernels.cuh"
#include "spl_parquet_meta.h"
#include "unicode_tables.inc"

namespace {

thread_local std::string g_create_error;

#define CUDA_TRY(expr, errstr)                                                            \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            (errstr) = std::string(#expr) + ": " + cudaGetErrorString(_e);                \
            return _e == cudaErrorMemoryAllocation ? SPL_ERR_OOM : SPL_ERR_CUDA;          \
        }                                                                                 \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, std::string& err) {
        if (bytes <= cap) return SPL_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { err = std::string("cudaMalloc: ") + cudaGetErrorString(e); p = nullptr; return SPL_ERR_OOM; }
        cap = want;
        return SPL_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct DevCtx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;          // kernels (+ the small per-chunk results)
    cudaStream_t s_in = nullptr, s_out = nullptr;   // host->device staging, ids device->host
    std::vector<cudaEvent_t> pipe_ev;       // per-chunk timing events of spl_encode_batch, grown on demand
    std::vector<cudaEvent_t> sync_ev;       // per-chunk ordering events (cudaEventDisableTiming: a timing event on a copy
                                            // stream makes the copy engine drain before the next transfer starts)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    void* table_blob = nullptr;
    SplTables* d_tables = nullptr;
    DevBuf text, doc_off, ids, out_off;          // spl_encode_batch: the shard's buffers
    DevBuf zero, tstate, pv, pool, mlist, fbl, huge, dupof;
    DevBuf dec_ids, dec_off, dec_ws, dec_out, dec_out_off;       // spl_decode_batch   // per-pass workspace (zero: everything that starts cleared)
    DevBuf jl_tiles, jl_lines;                                   // spl_ingest_jsonl_device: tile counts, per-line arrays
    DevBuf jl_text, jl_off, jl_out_off[2];                       // spl_encode_jsonl: ingested text + offsets, output offsets (alternating)
    DevBuf run_tot;                                              // spl_encode_batch: cumulative id count after each pipeline chunk
    DevBuf pq_stage, pq_scratch, pq_pages, pq_rows, pq_dict, pq_small;   // Parquet ingestion: staged column chunks, decompressed pages, page descriptors, spans, block sums + counters
    DevBuf sp_zero, sp_tiles, sp_text, sp_doc;                   // SentencePiece mode: bitmaps over T, tile counts, T', offsets in T'
    size_t huge_words = 0;
    // the kernels of one pipeline chunk as ONE graph launch (spl_encode_batch): [with_special]
    struct GraphCache {
        cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
        cudaGraphNode_t mnode = nullptr; std::vector<cudaGraphNode_t> knodes; std::vector<const void*> funcs;
    } gc[2];
    SplKernelProfile prof;
    bool prof_ready = false;
};

struct PinnedBuf { void* p; size_t cap; };

}  // namespace

struct spl_tokenizer {
    bool profiling = false;
    bool trace = false;                     // SPL_TRACE=1: per-chunk timeline of spl_encode_batch on stderr
    bool use_graph = true;                  // SPL_GRAPH=0: plain launches in spl_encode_batch as well
    bool charref = true;                    // SPL_NO_CHARREF=1: characters of two or three ids get miss-list entries (the path for passes beyond ~1.7 GB)
    bool use_pdl = false;                   // SPL_PDL=1: programmatic edges between the kernel nodes of a pass's graph (measured: nothing on cfg2 / cfg3 / cfg5, -10 % on cfg4)
    int bpe_overlap = 0;                    // SPL_BPE_OVERLAP: 0 the graph of a pass is a plain chain; 1: k_bpe_long beside k_bpe, its node first; 2: k_bpe's node first, on 4 blocks per SM so that one block of k_bpe_long fits beside them (measured: +-3 % either way)
    bool dedup = true;                      // SPL_NO_DEDUP=1: every long piece goes through the merge loop, repeated or not
    int trace_chunk = -1;                   // SPL_TRACE_CHUNK=k: with SPL_TRACE, per-kernel times of the k-th chunk
    uint64_t chunk_bytes = 0;               // pipeline chunk size of spl_encode_batch (0 = automatic)
    uint64_t chunks_per_dev = 8;            // automatic chunk size = shard / this (SPL_CHUNKS_PER_DEV)
    std::vector<uint64_t> ramp_div{4, 2};   // the first pipeline chunks of a device are target / 4, target / 2 (SPL_RAMP="8,4,2" to experiment)
    SplHostTables host;
    std::vector<DevCtx> devs;
    std::string err;
    std::mutex pool_mu;
    std::vector<PinnedBuf> pinned_pool;
};

struct spl_result {
    spl_tokenizer* owner;
    PinnedBuf ids_buf, off_buf;
    size_t n_docs, n_tokens;
    size_t n_bytes = 0;         // spl_decode_batch: ids_buf holds this many bytes
    spl_stats stats;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int upload_tables(spl_tokenizer* tk, DevCtx& dc) {
    const SplHostTables& h = tk->host;
    struct Part { const void* src; size_t bytes; size_t off; };
    std::vector<Part> parts;
    size_t total = 0;
    auto add = [&](const void* src, size_t bytes) { Part p{src, bytes, total}; total = align_up(total + std::max<size_t>(bytes, 16), 256); parts.push_back(p); return parts.size() - 1; };
    size_t i_s1 = add(spl_ucd_stage1, sizeof(spl_ucd_stage1));
    size_t i_s2 = add(spl_ucd_stage2, sizeof(spl_ucd_stage2));
    size_t i_t8 = add(h.t8.data(), h.t8.size() * sizeof(SplKey8));
    size_t i_t16 = add(h.t16.data(), h.t16.size() * sizeof(SplKey16));
    size_t i_tl = add(h.tl.data(), h.tl.size() * sizeof(SplKeyL));
    size_t i_tb = add(h.tok_bytes.data(), h.tok_bytes.size());
    size_t i_to = add(h.tok_off.data(), h.tok_off.size() * 4);
    size_t i_pair = add(h.pair.data(), h.pair.size() * 4);
    size_t i_bpair = add(h.bpair.data(), h.bpair.size() * 4);
    size_t i_irr = add(h.seg_irr.data(), h.seg_irr.size() * 4);
    size_t i_h2 = add(h.seg_h2.data(), h.seg_h2.size() * 4);
    size_t i_ctok = add(h.char_tok.data(), h.char_tok.size() * 4);
    size_t i_cids = add(h.char_ids.data(), h.char_ids.size() * 4);
    size_t i_decb = add(h.dec_bytes.data(), h.dec_bytes.size());
    size_t i_deco = add(h.dec_off.data(), h.dec_off.size() * 4);
    size_t i_spb = add(h.sp_bytes.data(), h.sp_bytes.size());
    size_t i_spo = add(h.sp_off.data(), h.sp_off.size() * 4);
    size_t i_spi = add(h.sp_id.data(), h.sp_id.size() * 4);
    size_t i_struct = add(nullptr, sizeof(SplTables));
    CUDA_TRY(cudaMalloc(&dc.table_blob, total), tk->err);
    uint8_t* base = (uint8_t*)dc.table_blob;
    for (auto& p : parts)
        if (p.src && p.bytes) CUDA_TRY(cudaMemcpy(base + p.off, p.src, p.bytes, cudaMemcpyHostToDevice), tk->err);
    SplTables t;
    memset(&t, 0, sizeof(t));
    t.ucd_stage1 = base + parts[i_s1].off;
    t.ucd_stage2 = base + parts[i_s2].off;
    t.t8 = (const SplKey8*)(base + parts[i_t8].off);     t.t8_log2 = h.t8_log2;
    t.t16 = (const SplKey16*)(base + parts[i_t16].off);  t.t16_log2 = h.t16_log2;
    t.tl = (const SplKeyL*)(base + parts[i_tl].off);     t.tl_log2 = h.tl_log2;
    t.tok_bytes = base + parts[i_tb].off;
    t.tok_off = (const uint32_t*)(base + parts[i_to].off);
    t.n_ids = h.n_ids;
    t.max_key_len = h.max_key_len;
    t.pair = (const uint32_t*)(base + parts[i_pair].off); t.pair_log2 = h.pair_log2;
    t.bpair = (const uint32_t*)(base + parts[i_bpair].off);
    t.seg_irr = (const uint32_t*)(base + parts[i_irr].off);
    t.seg_h2 = (const uint32_t*)(base + parts[i_h2].off); t.seg_h2_log2 = h.seg_h2_log2;
    t.char_tok = (const uint32_t*)(base + parts[i_ctok].off);
    t.char_ids = (const uint32_t*)(base + parts[i_cids].off);
    memcpy(t.byte_sym, h.byte_sym, sizeof(t.byte_sym));
    t.dec_bytes = base + parts[i_decb].off;
    t.dec_off = (const uint32_t*)(base + parts[i_deco].off);
    t.n_dec = h.dec_off.empty() ? 0u : (uint32_t)h.dec_off.size() - 1u;
    t.sp_bytes = base + parts[i_spb].off;
    t.sp_off = (const uint32_t*)(base + parts[i_spo].off);
    t.sp_id = (const uint32_t*)(base + parts[i_spi].off);
    t.n_special = (uint32_t)h.sp_id.size();
    memcpy(t.sp_first, h.sp_first, sizeof(t.sp_first));
    t.pattern = h.pattern;
    dc.d_tables = (SplTables*)(base + parts[i_struct].off);
    CUDA_TRY(cudaMemcpy(dc.d_tables, &t, sizeof(t), cudaMemcpyHostToDevice), tk->err);
    return SPL_OK;
}

void destroy_ctx(DevCtx& dc) {
    cudaSetDevice(dc.device);
    for (DevBuf* b : {&dc.text, &dc.doc_off, &dc.ids, &dc.out_off, &dc.dec_ids, &dc.dec_off, &dc.dec_ws, &dc.dec_out, &dc.dec_out_off, &dc.jl_tiles, &dc.jl_lines, &dc.jl_text, &dc.jl_off, &dc.jl_out_off[0], &dc.jl_out_off[1], &dc.run_tot, &dc.sp_zero, &dc.sp_tiles, &dc.sp_text, &dc.sp_doc, &dc.zero, &dc.tstate, &dc.pv, &dc.pool, &dc.mlist, &dc.fbl, &dc.huge, &dc.dupof,
                       &dc.pq_stage, &dc.pq_scratch, &dc.pq_pages, &dc.pq_rows, &dc.pq_dict, &dc.pq_small})
        b->release();
    for (auto& g : dc.gc) { if (g.exec) cudaGraphExecDestroy(g.exec); if (g.graph) cudaGraphDestroy(g.graph); }
    if (dc.table_blob) cudaFree(dc.table_blob);
    for (auto& e : dc.ev) if (e) cudaEventDestroy(e);
    if (dc.prof_ready) for (auto& e : dc.prof.ev) cudaEventDestroy(e);
    for (auto& e : dc.pipe_ev) cudaEventDestroy(e);
    for (auto& e : dc.sync_ev) cudaEventDestroy(e);
    if (dc.stream) cudaStreamDestroy(dc.stream);
    if (dc.s_in) cudaStreamDestroy(dc.s_in);
    if (dc.s_out) cudaStreamDestroy(dc.s_out);
}

// limits of one device shard: 32-bit positions inside the kernels
const uint64_t kMaxShardBytes = 0xFFFFFFFFull - 4ull * SPL_WIN;

// layout of the zero-initialised region: counters | duplicate table | chunk_cnt | tinfo | hard | pstart | spec
struct ZeroLayout { size_t words, n_tiles, dd_slots, off_dd, off_chunk, off_extra, off_hard, off_pstart, off_spec, off_cand, total; };
ZeroLayout zero_layout(uint64_t N, bool with_special, bool dedup, bool ambiguous = false) {
    ZeroLayout z;
    z.words = (size_t)((N + SPL_WIN) / 32 + 16);
    z.n_tiles = (size_t)(N / SPL_TILE) + 1;
    // duplicate table: one slot per 64 bytes of text, 4 Ki .. 256 Ki slots (at most 2 MiB to clear per pass); when it
    // fills up, further pieces simply go through the merge loop
    z.dd_slots = 0;
    if (dedup) { z.dd_slots = 4096; while (z.dd_slots < ((size_t)1 << 18) && z.dd_slots < N / 64) z.dd_slots <<= 1; }
    z.off_dd = 256;
    z.off_chunk = z.off_dd + z.dd_slots * 8;
    z.off_extra = align_up(z.off_chunk + (z.n_tiles / SPL_CHUNK_TILES + 2) * 4, 256);
    z.off_hard = align_up(z.off_extra + (z.n_tiles + 2) * sizeof(SplTileInfo), 256);
    z.off_pstart = align_up(z.off_hard + z.words * 4, 256);
    z.off_spec = align_up(z.off_pstart + z.words * 4, 256);
    z.off_cand = with_special ? align_up(z.off_spec + z.words * 4, 256) : z.off_spec;
    z.total = (with_special && ambiguous) ? align_up(z.off_cand + z.words * 4, 256) : z.off_cand;
    return z;
}

struct MissLayout { uint32_t base[SPL_NCLS + 1]; };
MissLayout miss_layout(uint64_t N) {
    MissLayout m;
    m.base[0] = 0;
    for (uint32_t c = 0; c < SPL_NCLS; ++c)                       // pieces are disjoint: at most N / minlen of a class
        m.base[c + 1] = m.base[c] + (uint32_t)(N / spl_class_minlen(c) + 16);
    return m;
}

// Size the internal workspace of `dc` for N bytes / n_docs documents (may reallocate: nothing may be in flight).
int reserve_work(spl_tokenizer* tk, DevCtx& dc, uint64_t N, uint64_t n_docs, bool with_special) {
    if (N > kMaxShardBytes || n_docs > 0xFFFFFFF0ull) {
        tk->err = "one device pass is limited to 4 GiB of text";
        return SPL_ERR_UNSUPPORTED;
    }
    const ZeroLayout z = zero_layout(N, with_special, tk->dedup, !tk->host.specials_unambiguous);
    const MissLayout m = miss_layout(N);
    int rc;
    if ((rc = dc.zero.ensure(z.total, tk->err))) return rc;
    if ((rc = dc.tstate.ensure((z.n_tiles / SPL_CHUNK_TILES + 2) * 8, tk->err))) return rc;
    if ((rc = dc.pv.ensure(z.n_tiles * SPL_TILE * 4, tk->err))) return rc;
    if ((rc = dc.pool.ensure((size_t)(N + 64) * 4, tk->err))) return rc;
    if ((rc = dc.mlist.ensure((size_t)m.base[SPL_NCLS] * 8, tk->err))) return rc;
    if (tk->dedup && (rc = dc.dupof.ensure((size_t)(m.base[SPL_NCLS] - m.base[2] + 16) * 4, tk->err))) return rc;
    uint32_t n_fast_tiles = (uint32_t)((N + SPL_FAST_PAYLOAD * 32u - 1) / (SPL_FAST_PAYLOAD * 32u));
    if ((rc = dc.fbl.ensure((size_t)(n_fast_tiles + 1) * 4, tk->err))) return rc;
    if (dc.huge_words == 0) dc.huge_words = (size_t)16 << 20;             // 64 MiB of scratch
    if ((rc = dc.huge.ensure(dc.huge_words * 4, tk->err))) return rc;
    return SPL_OK;
}

// Prepare the internal workspace of `dc` for N bytes / n_docs documents and fill `w`
// (text / doc_off / ids / out_off are set by the caller).  Enqueues the zero-fill.
int prepare_work(spl_tokenizer* tk, DevCtx& dc, uint64_t N, uint64_t n_docs, bool with_special,
                 cudaStream_t st, SplWork& w, size_t* zero_bytes = nullptr) {
    int rc = reserve_work(tk, dc, N, n_docs, with_special);
    if (rc) return rc;
    const ZeroLayout z = zero_layout(N, with_special, tk->dedup, !tk->host.specials_unambiguous);
    const MissLayout m = miss_layout(N);
    uint32_t n_fast_tiles = (uint32_t)((N + SPL_FAST_PAYLOAD * 32u - 1) / (SPL_FAST_PAYLOAD * 32u));
    if (zero_bytes) *zero_bytes = z.total;                       // the caller clears the region itself (graph memset node)
    else CUDA_TRY(cudaMemsetAsync(dc.zero.p, 0, z.total, st), tk->err);
    uint8_t* zb = (uint8_t*)dc.zero.p;
    w.N = (uint32_t)N;
    w.n_docs = (uint32_t)n_docs;
    w.n_tiles = (uint32_t)z.n_tiles;
    w.counters = (uint32_t*)zb;
    w.chunk_cnt = (int32_t*)(zb + z.off_chunk);
    w.tinfo = (SplTileInfo*)(zb + z.off_extra);
    w.hard = (uint32_t*)(zb + z.off_hard);
    w.pstart = (uint32_t*)(zb + z.off_pstart);
    w.spec = with_special ? (uint32_t*)(zb + z.off_spec) : nullptr;
    w.cand = (with_special && !tk->host.specials_unambiguous) ? (uint32_t*)(zb + z.off_cand) : nullptr;
    w.bitmap_words = z.words;
    w.chunk_state = (uint64_t*)dc.tstate.p;
    w.pv = (uint32_t*)dc.pv.p;
    w.pool = (uint32_t*)dc.pool.p;
    w.mlist = (uint64_t*)dc.mlist.p;
    for (uint32_t c = 0; c <= SPL_NCLS; ++c) w.ml_base[c] = m.base[c];
    w.fb_list = (uint32_t*)dc.fbl.p;
    w.n_fast_tiles = n_fast_tiles;
    w.dd_tab = tk->dedup ? (unsigned long long*)(zb + z.off_dd) : nullptr;
    w.dd_mask = tk->dedup ? (uint32_t)z.dd_slots - 1u : 0u;
    w.dup_of = (uint32_t*)dc.dupof.p;
    w.huge_pool = (uint32_t*)dc.huge.p;
    w.huge_pool_words = dc.huge_words;
    w.T = dc.d_tables;
    w.pattern = tk->host.pattern;
    w.with_special = with_special;
    w.charref = tk->charref && m.base[SPL_NCLS] < 0x40000000u;     // SPL_PV_CHARREF needs miss-list indices below 2^30
    return SPL_OK;
}

int check_special_support(spl_tokenizer* tk, uint32_t flags, bool& with_special) {
    // any set is served: where strings can contain or overlap one another the matches are those of aho-corasick's
    // Standard non-overlapping find_iter (spl_special.h, k_resolve_specials)
    with_special = (flags & SPL_ENCODE_WITH_SPECIAL) && !tk->host.sp_id.empty();
    return SPL_OK;
}

bool is_sentencepiece(const spl_tokenizer* tk) { return tk->host.pattern == SPL_PAT_SENTENCEPIECE; }

// ids an input of n bytes can produce at most: one per byte, except that SentencePiece mode turns a space into the
// three bytes of U+2581, which can stay three byte tokens
uint64_t ids_bound(const spl_tokenizer* tk, uint64_t n_bytes) { return is_sentencepiece(tk) ? 3 * n_bytes : n_bytes; }

struct EncodeArgs {
    const uint8_t* text; uint64_t N;
    const uint64_t* doc_off; uint64_t off_base; uint64_t n_docs;
    uint32_t* ids; uint64_t ids_cap; uint64_t* out_off; uint64_t* host_meta;
    const uint64_t* tok_base_in; uint64_t* tok_total_out;
};

// Enqueue the encode path for one device pass on `st`; `w` is the workspace view the pass uses (its counters are what
// the caller reads back).  SentencePiece mode (tokenizer.rs:737-795) first builds the transformed text T' -- that needs
// the size of T' on the host, i.e. one stream synchronisation in the middle -- and then runs the ordinary encode stage
// over T'.
int enqueue_encode(spl_tokenizer* tk, DevCtx& dc, cudaStream_t st, const EncodeArgs& a, bool with_special,
                   SplKernelProfile* prof, SplWork& w, int& launches) {
    int rc;
    memset(&w, 0, sizeof(w));
    if (!is_sentencepiece(tk)) {
        if ((rc = prepare_work(tk, dc, a.N, a.n_docs, with_special, st, w))) return rc;
        w.text = a.text; w.doc_off = a.doc_off; w.off_base = a.off_base;
        w.ids = a.ids; w.out_off = a.out_off; w.host_meta = a.host_meta;
        w.tok_base_in = a.tok_base_in; w.tok_total_out = a.tok_total_out;
        launches += spl_launch_encode(w, dc.num_sms, st, prof);
        return SPL_OK;
    }
    if (3 * a.N > kMaxShardBytes || a.n_docs > 0xFFFFFFF0ull) {
        tk->err = "one SentencePiece-mode device pass is limited to 4/3 GiB of text";
        return SPL_ERR_UNSUPPORTED;
    }
    // ---- bitmaps over T: counters | tinfo | hard | spec | w0 | a | rs (one memset) ----
    const size_t words = (size_t)((a.N + SPL_WIN) / 32 + 16), n_tiles = (size_t)(a.N / SPL_TILE) + 1;
    const size_t off_tinfo = 256, off_hard = align_up(off_tinfo + (n_tiles + 2) * sizeof(SplTileInfo), 256);
    const size_t bm = align_up(words * 4, 256);
    const size_t off_spec = off_hard + bm, off_w0 = off_spec + (with_special ? bm : 0), off_a = off_w0 + bm, off_rs = off_a + bm;
    const bool amb = with_special && !tk->host.specials_unambiguous;
    const size_t off_cand = off_rs + bm, sp_total = off_cand + (amb ? bm : 0);
    if ((rc = dc.sp_zero.ensure(sp_total, tk->err))) return rc;
    if ((rc = dc.sp_tiles.ensure((2 * n_tiles + 2) * 4, tk->err))) return rc;
    if ((rc = dc.sp_doc.ensure((a.n_docs + 1) * 8, tk->err))) return rc;
    CUDA_TRY(cudaMemsetAsync(dc.sp_zero.p, 0, sp_total, st), tk->err);
    uint8_t* zb = (uint8_t*)dc.sp_zero.p;
    SplWork v;
    memset(&v, 0, sizeof(v));
    v.text = a.text; v.N = (uint32_t)a.N; v.doc_off = a.doc_off; v.off_base = a.off_base; v.n_docs = (uint32_t)a.n_docs;
    v.n_tiles = (uint32_t)n_tiles;
    v.counters = (uint32_t*)zb; v.tinfo = (SplTileInfo*)(zb + off_tinfo);
    v.hard = (uint32_t*)(zb + off_hard); v.pstart = v.hard;            // the sentinel bit N is a hard bit as well
    v.spec = with_special ? (uint32_t*)(zb + off_spec) : nullptr;
    v.cand = amb ? (uint32_t*)(zb + off_cand) : nullptr;
    v.T = dc.d_tables; v.pattern = tk->host.pattern; v.with_special = with_special;
    launches += spl_launch_mark(v, dc.num_sms, st);
    SplSpWork s;
    memset(&s, 0, sizeof(s));
    s.text = a.text; s.N = v.N; s.doc_off = a.doc_off; s.off_base = a.off_base; s.n_docs = v.n_docs; s.n_tiles = v.n_tiles;
    s.hard = v.hard; s.spec = v.spec; s.tinfo = v.tinfo;
    s.w0 = (uint32_t*)(zb + off_w0); s.a = (uint32_t*)(zb + off_a); s.rs = (uint32_t*)(zb + off_rs);
    s.tile_cnt = (uint32_t*)dc.sp_tiles.p; s.tile_pref = s.tile_cnt + n_tiles;
    s.counters = v.counters; s.T = dc.d_tables;
    launches += spl_launch_sp_scan(s, st);
    uint32_t h_ctr[8];
    CUDA_TRY(cudaMemcpyAsync(h_ctr, v.counters, sizeof(h_ctr), cudaMemcpyDeviceToHost, st), tk->err);
    CUDA_TRY(cudaStreamSynchronize(st), tk->err);
    if (h_ctr[SPL_CTR_ERR] & SPL_DEVERR_OFFSETS) {
        tk->err = "document offsets are not a non-decreasing sequence from 0 to n_bytes";
        return SPL_ERR_INVALID_ARG;
    }
    const uint64_t N2 = a.N + 2ull * h_ctr[SPL_SPCTR_CONV];
    if (a.ids_cap < N2) {
        tk->err = "ids_capacity too small (SentencePiece mode: up to three ids per input byte)";
        return SPL_ERR_INVALID_ARG;
    }
    if ((rc = dc.sp_text.ensure((size_t)N2 + 64, tk->err))) return rc;
    if ((rc = prepare_work(tk, dc, N2, a.n_docs, with_special, st, w))) return rc;
    w.text = (const uint8_t*)dc.sp_text.p; w.doc_off = (const uint64_t*)dc.sp_doc.p; w.off_base = 0;
    w.ids = a.ids; w.out_off = a.out_off; w.host_meta = a.host_meta;
    w.tok_base_in = a.tok_base_in; w.tok_total_out = a.tok_total_out;
    w.pretok_done = true;
    s.text2 = (uint8_t*)dc.sp_text.p; s.pstart2 = w.pstart; s.spec2 = w.spec; s.doc_off2 = (uint64_t*)dc.sp_doc.p;
    launches += spl_launch_sp_emit(s, st);
    launches += spl_launch_encode(w, dc.num_sms, st, prof);
    return SPL_OK;
}

// The same pass as ONE graph launch: a memset node and the kernels of spl_describe_encode in a chain.  The executable
// graph is built once per device and flag set; every chunk only rewrites the node parameters (pointers, sizes, grids).
// Under saturating host-to-device traffic a kernel launch takes 25-45 us to reach the device: one launch per chunk
// instead of nine brings the first ids of a call out earlier and shortens its tail.
int enqueue_encode_graph(spl_tokenizer* tk, DevCtx& dc, cudaStream_t st, const EncodeArgs& a, bool with_special,
                         SplWork& w, int& launches) {
    int rc;
    memset(&w, 0, sizeof(w));
    size_t zero_bytes = 0;
    if ((rc = prepare_work(tk, dc, a.N, a.n_docs, with_special, st, w, &zero_bytes))) return rc;
    w.text = a.text; w.doc_off = a.doc_off; w.off_base = a.off_base;
    w.ids = a.ids; w.out_off = a.out_off; w.host_meta = a.host_meta;
    w.tok_base_in = a.tok_base_in; w.tok_total_out = a.tok_total_out;
    SplLaunchDesc d[SPL_MAX_LAUNCHES];
    const int n = spl_describe_encode(w, dc.num_sms, d);
    DevCtx::GraphCache& g = dc.gc[with_special ? 1 : 0];
    cudaMemsetParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.dst = dc.zero.p; mp.value = 0; mp.elementSize = 4; mp.width = zero_bytes / 4; mp.height = 1; mp.pitch = zero_bytes;
    uint32_t u32s[SPL_MAX_LAUNCHES];
    void* args[SPL_MAX_LAUNCHES][2];
    cudaKernelNodeParams kp[SPL_MAX_LAUNCHES];
    for (int i = 0; i < n; ++i) {
        u32s[i] = d[i].u32;
        args[i][0] = &w; args[i][1] = &u32s[i];
        memset(&kp[i], 0, sizeof(kp[i]));
        kp[i].func = const_cast<void*>(d[i].func);
        kp[i].gridDim = dim3(d[i].grid); kp[i].blockDim = dim3(d[i].block);
        kp[i].sharedMemBytes = (unsigned)d[i].smem;
        kp[i].kernelParams = args[i];
    }
    bool same = g.exec && (int)g.funcs.size() == n;
    for (int i = 0; same && i < n; ++i) same = g.funcs[i] == d[i].func;
    if (!same) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        if (g.graph) { cudaGraphDestroy(g.graph); g.graph = nullptr; }
        g.knodes.assign(n, nullptr); g.funcs.assign(n, nullptr);
        CUDA_TRY(cudaGraphCreate(&g.graph, 0), tk->err);
        CUDA_TRY(cudaGraphAddMemsetNode(&g.mnode, g.graph, nullptr, 0, &mp), tk->err);
        // A chain.  (SPL_BPE_OVERLAP=1 / 2 puts k_bpe_long beside k_bpe -- their miss lists are disjoint, everything
        // they share is atomic counters -- with k_bpe_fin waiting for both.  Both kernels size themselves for the whole
        // device, so whichever node is dispatched first keeps the other out: cfg4 gains 3 % one way and loses 13 % the
        // other, cfg5 the reverse; the chain is the default.)
        cudaGraphNode_t prev = g.mnode;
        // kernel -> kernel edges are programmatic (SPL_PDL_ENTER in spl_kernels.cuh): the next kernel's blocks are
        // scheduled into the tail of the one before
        const bool pdl = tk->use_pdl && tk->bpe_overlap == 0;
        for (int i = 0; i < n; ++i) {
            const bool fork = tk->bpe_overlap && i + 2 < n && !strcmp(d[i].name, "k_bpe") && !strcmp(d[i + 1].name, "k_bpe_long");
            if (fork) {
                if (tk->bpe_overlap == 2) {
                    kp[i].gridDim = dim3(std::min<unsigned>(kp[i].gridDim.x, 4u * (unsigned)dc.num_sms));
                    CUDA_TRY(cudaGraphAddKernelNode(&g.knodes[i], g.graph, &prev, 1, &kp[i]), tk->err);
                    CUDA_TRY(cudaGraphAddKernelNode(&g.knodes[i + 1], g.graph, &prev, 1, &kp[i + 1]), tk->err);
                } else {
                    CUDA_TRY(cudaGraphAddKernelNode(&g.knodes[i + 1], g.graph, &prev, 1, &kp[i + 1]), tk->err);
                    CUDA_TRY(cudaGraphAddKernelNode(&g.knodes[i], g.graph, &prev, 1, &kp[i]), tk->err);
                }
                cudaGraphNode_t both[2] = {g.knodes[i], g.knodes[i + 1]};
                CUDA_TRY(cudaGraphAddKernelNode(&g.knodes[i + 2], g.graph, both, 2, &kp[i + 2]), tk->err);
                g.funcs[i] = d[i].func; g.funcs[i + 1] = d[i + 1].func; g.funcs[i + 2] = d[i + 2].func;
                prev = g.knodes[i + 2];
                i += 2;
                continue;
            }
            if (pdl && i > 0) {
                CUDA_TRY(cudaGraphAddKernelNode(&g.knodes[i], g.graph, nullptr, 0, &kp[i]), tk->err);
                cudaGraphEdgeData ed;
                memset(&ed, 0, sizeof(ed));
                ed.from_port = cudaGraphKernelNodePortProgrammatic; ed.type = cudaGraphDependencyTypeProgrammatic;
                if (cudaGraphAddDependencies_v2(g.graph, &prev, &g.knodes[i], &ed, 1) != cudaSuccess) {
                    cudaGetLastError();                          // (a driver without the edge type: ordinary edges from here on)
                    tk->use_pdl = false;
                    CUDA_TRY(cudaGraphAddDependencies(g.graph, &prev, &g.knodes[i], 1), tk->err);
                }
            } else {
                CUDA_TRY(cudaGraphAddKernelNode(&g.knodes[i], g.graph, &prev, 1, &kp[i]), tk->err);
            }
            prev = g.knodes[i];
            g.funcs[i] = d[i].func;
        }
        CUDA_TRY(cudaGraphInstantiate(&g.exec, g.graph, 0), tk->err);
    } else {
        CUDA_TRY(cudaGraphExecMemsetNodeSetParams(g.exec, g.mnode, &mp), tk->err);
        for (int i = 0; i < n; ++i) {
            if (tk->bpe_overlap == 2 && !strcmp(d[i].name, "k_bpe")) kp[i].gridDim = dim3(std::min<unsigned>(kp[i].gridDim.x, 4u * (unsigned)dc.num_sms));
            CUDA_TRY(cudaGraphExecKernelNodeSetParams(g.exec, g.knodes[i], &kp[i]), tk->err);
        }
    }
    CUDA_TRY(cudaGraphLaunch(g.exec, st), tk->err);
    launches += n;
    return SPL_OK;
}

struct JlOut { uint8_t* text; size_t text_cap; uint64_t* off; size_t off_cap; };

// JSON Lines -> packed text + offsets for one device pass on `st` (two synchronisations).  `outputs(n_lines, o)` is
// called once the line count is known and names the buffers.
template <class OutFn>
int ingest_jsonl_core(spl_tokenizer* tk, DevCtx& dc, cudaStream_t st, const uint8_t* d_jsonl, size_t n_bytes,
                      const char* field, size_t flen, OutFn outputs, spl_ingest_stats* stats) {
    if (n_bytes > kMaxShardBytes) { tk->err = "one device pass is limited to 4 GiB"; return SPL_ERR_UNSUPPORTED; }
    memset(stats, 0, sizeof(*stats));
    SplJlWork w;
    memset(&w, 0, sizeof(w));
    w.text = d_jsonl; w.N = (uint32_t)n_bytes; w.n_tiles = (uint32_t)(n_bytes / SPL_TILE) + 1;
    memcpy(w.field, field, flen); w.flen = (uint32_t)flen;
    int rc;
    const size_t tiles_bytes = align_up(256 + (2 * (size_t)w.n_tiles + 2) * 4, 256);
    if ((rc = dc.jl_tiles.ensure(tiles_bytes, tk->err))) return rc;
    CUDA_TRY(cudaMemsetAsync(dc.jl_tiles.p, 0, 256, st), tk->err);
    w.counters = (uint32_t*)dc.jl_tiles.p;
    w.tile_cnt = w.counters + 64; w.tile_pref = w.tile_cnt + w.n_tiles;
    spl_launch_jsonl_count(w, st);
    uint32_t h_ctr[8];
    CUDA_TRY(cudaMemcpyAsync(h_ctr, w.counters, sizeof(h_ctr), cudaMemcpyDeviceToHost, st), tk->err);
    CUDA_TRY(cudaStreamSynchronize(st), tk->err);
    const size_t n_lines = (size_t)h_ctr[SPL_JLCTR_NEWLINES] + 1;
    w.n_lines = (uint32_t)n_lines;
    // per-line arrays: line_start | doc_idx | text_off (n_lines + 1 each) | is_doc | out_len (n_lines each) | span
    const size_t a1 = align_up((n_lines + 1) * 4, 16), a0 = align_up(n_lines * 4, 16);
    if ((rc = dc.jl_lines.ensure(3 * a1 + 2 * a0 + n_lines * 16 + 64, tk->err))) return rc;
    uint8_t* lb = (uint8_t*)dc.jl_lines.p;
    w.line_start = (uint32_t*)lb; w.doc_idx = (uint32_t*)(lb + a1); w.text_off = (uint32_t*)(lb + 2 * a1);
    w.is_doc = (uint32_t*)(lb + 3 * a1); w.out_len = (uint32_t*)(lb + 3 * a1 + a0);
    w.span = (SplJlSpan*)(lb + 3 * a1 + 2 * a0);
    JlOut o{nullptr, 0, nullptr, 0};
    if ((rc = outputs(n_lines, o))) return rc;
    w.out_text = o.text; w.text_capacity = o.text ? o.text_cap : 0;
    w.out_off = o.off; w.off_capacity = o.off ? o.off_cap : 0;
    spl_launch_jsonl_extract(w, st);
    CUDA_TRY(cudaGetLastError(), tk->err);
    CUDA_TRY(cudaMemcpyAsync(h_ctr, w.counters, sizeof(h_ctr), cudaMemcpyDeviceToHost, st), tk->err);
    CUDA_TRY(cudaStreamSynchronize(st), tk->err);
    stats->n_lines = n_lines; stats->n_docs = h_ctr[SPL_JLCTR_DOCS]; stats->n_text_bytes = h_ctr[SPL_JLCTR_TEXT];
    stats->n_missing = h_ctr[SPL_JLCTR_MISSING]; stats->n_bad = h_ctr[SPL_JLCTR_BAD]; stats->n_launches = 6;
    if (stats->n_docs + 1 > w.off_capacity || stats->n_text_bytes > w.text_capacity) {
        tk->err = "output capacity too small (needed sizes returned in the stats: n_docs + 1 offsets, n_text_bytes bytes)";
        return SPL_ERR_INVALID_ARG;
    }
    return SPL_OK;
}

PinnedBuf take_pinned(spl_tokenizer* tk, size_t bytes) {
    bytes = std::max<size_t>(bytes, 64);
    {
        std::lock_guard<std::mutex> g(tk->pool_mu);
        size_t best = (size_t)-1;
        for (size_t i = 0; i < tk->pinned_pool.size(); ++i)
            if (tk->pinned_pool[i].cap >= bytes && (best == (size_t)-1 || tk->pinned_pool[i].cap < tk->pinned_pool[best].cap)) best = i;
        if (best != (size_t)-1) {
            PinnedBuf b = tk->pinned_pool[best];
            tk->pinned_pool.erase(tk->pinned_pool.begin() + best);
            return b;
        }
    }
    PinnedBuf b{nullptr, 0};
    size_t want = bytes + bytes / 8;
    if (cudaHostAlloc(&b.p, want, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return b; }
    b.cap = want;
    return b;
}

void give_pinned(spl_tokenizer* tk, PinnedBuf b) {
    if (!b.p) return;
    std::lock_guard<std::mutex> g(tk->pool_mu);
    // keep what one call of a multi-device handle takes (result ids + offsets, chunk records, one staging buffer per
    // device): a pool that evicts makes the next call pin gigabytes again (cudaHostAlloc of 2.5 GB: ~1 s)
    if (tk->pinned_pool.size() >= 8 + 2 * tk->devs.size()) {
        size_t small = 0;
        for (size_t i = 1; i < tk->pinned_pool.size(); ++i) if (tk->pinned_pool[i].cap < tk->pinned_pool[small].cap) small = i;
        if (tk->pinned_pool[small].cap >= b.cap) { cudaFreeHost(b.p); return; }
        cudaFreeHost(tk->pinned_pool[small].p);
        tk->pinned_pool[small] = b;
        return;
    }
    tk->pinned_pool.push_back(b);
}

}  // namespace

extern "C" {

const char* spl_version(void) { return "splintr_b200 0.1.0 sm_100a"; }

int spl_create(const uint8_t* vocab, size_t vocab_len, int pattern_id, uint32_t flags,
               const char* const* special_strs, const uint32_t* special_ids, size_t n_special,
               const int* devices, int n_devices, spl_tokenizer** out) {
    if (!out) return SPL_ERR_INVALID_ARG;
    *out = nullptr;
    if (!vocab || (n_special && (!special_strs || !special_ids))) { g_create_error = "null argument"; return SPL_ERR_INVALID_ARG; }
    if (pattern_id != SPL_PATTERN_CL100K && pattern_id != SPL_PATTERN_O200K && pattern_id != SPL_PATTERN_MISTRAL_V3 &&
        pattern_id != SPL_PATTERN_SENTENCEPIECE) {
        g_create_error = "unknown pattern id";
        return SPL_ERR_INVALID_ARG;
    }
    if (((flags & SPL_CREATE_SENTENCEPIECE) != 0) != (pattern_id == SPL_PATTERN_SENTENCEPIECE) ||
        ((flags & SPL_CREATE_SENTENCEPIECE) && (flags & SPL_CREATE_BYTE_LEVEL))) {
        g_create_error = "SPL_CREATE_SENTENCEPIECE goes with SPL_PATTERN_SENTENCEPIECE (and not with SPL_CREATE_BYTE_LEVEL)";
        return SPL_ERR_INVALID_ARG;
    }
    int ndev_avail = 0;
    if (cudaGetDeviceCount(&ndev_avail) != cudaSuccess || ndev_avail == 0) {
        cudaGetLastError();
        g_create_error = "no CUDA device available (splintr_b200 has no CPU fallback)";
        return SPL_ERR_NO_DEVICE;
    }
    spl_tokenizer* tk = new (std::nothrow) spl_tokenizer();
    if (!tk) return SPL_ERR_OOM;
    if (const char* tr = getenv("SPL_TRACE")) tk->trace = tr[0] == '1';
    if (const char* tc = getenv("SPL_TRACE_CHUNK")) tk->trace_chunk = atoi(tc);
    if (const char* nd = getenv("SPL_NO_DEDUP")) tk->dedup = nd[0] == '0';
    if (const char* nc = getenv("SPL_NO_CHARREF")) tk->charref = nc[0] == '0';
    if (const char* gr = getenv("SPL_GRAPH")) tk->use_graph = gr[0] != '0';
    if (const char* bo = getenv("SPL_BPE_OVERLAP")) tk->bpe_overlap = atoi(bo);
    if (const char* pd = getenv("SPL_PDL")) tk->use_pdl = pd[0] != '0';
    if (const char* cb = getenv("SPL_CHUNK_BYTES")) tk->chunk_bytes = strtoull(cb, nullptr, 10);
    if (const char* rp = getenv("SPL_RAMP")) {
        tk->ramp_div.clear();
        for (const char* p = rp; *p;) { char* e; uint64_t v = strtoull(p, &e, 10); if (e == p) break; if (v) tk->ramp_div.push_back(v); p = *e ? e + 1 : e; }
    }
    if (const char* ch = getenv("SPL_CHUNKS_PER_DEV")) tk->chunks_per_dev = std::max<uint64_t>(strtoull(ch, nullptr, 10), 1);
    uint32_t hflags = ((flags & SPL_CREATE_BYTE_LEVEL) ? SPL_FLAG_BYTE_LEVEL : 0) |
                      ((flags & SPL_CREATE_SENTENCEPIECE) ? SPL_FLAG_SENTENCEPIECE : 0);
    if (!spl_build_tables(tk->host, vocab, vocab_len, pattern_id, hflags, special_strs, special_ids, n_special)) {
        g_create_error = tk->host.error;
        bool unsupported = tk->host.error.find("byte-level vocabulary") != std::string::npos ||
                           tk->host.error.find("not supported") != std::string::npos;
        delete tk;
        return unsupported ? SPL_ERR_UNSUPPORTED : SPL_ERR_VOCAB;
    }
    DeviceGuard guard;
    std::vector<int> devs;
    if (!devices || n_devices <= 0) devs.push_back(guard.prev >= 0 ? guard.prev : 0);
    else devs.assign(devices, devices + n_devices);
    tk->devs.resize(devs.size());
    int rc = SPL_OK;
    for (size_t i = 0; i < devs.size() && rc == SPL_OK; ++i) {
        DevCtx& dc = tk->devs[i];
        dc.device = devs[i];
        if (dc.device < 0 || dc.device >= ndev_avail) { tk->err = "device index out of range"; rc = SPL_ERR_INVALID_ARG; break; }
        auto init = [&]() -> int {
            CUDA_TRY(cudaSetDevice(dc.device), tk->err);
            CUDA_TRY(cudaDeviceGetAttribute(&dc.num_sms, cudaDevAttrMultiProcessorCount, dc.device), tk->err);
            CUDA_TRY(cudaStreamCreateWithFlags(&dc.stream, cudaStreamNonBlocking), tk->err);
            CUDA_TRY(cudaStreamCreateWithFlags(&dc.s_in, cudaStreamNonBlocking), tk->err);
            CUDA_TRY(cudaStreamCreateWithFlags(&dc.s_out, cudaStreamNonBlocking), tk->err);
            for (auto& e : dc.ev) CUDA_TRY(cudaEventCreate(&e), tk->err);
            spl_kernels_init();
            return upload_tables(tk, dc);
        };
        rc = init();
    }
    if (rc != SPL_OK) {
        g_create_error = tk->err;
        for (auto& dc : tk->devs) destroy_ctx(dc);
        delete tk;
        return rc;
    }
    *out = tk;
    return SPL_OK;
}

void spl_destroy(spl_tokenizer* tk) {
    if (!tk) return;
    DeviceGuard guard;
    for (auto& dc : tk->devs) destroy_ctx(dc);
    for (auto& b : tk->pinned_pool) cudaFreeHost(b.p);
    delete tk;
}

const char* spl_last_error(const spl_tokenizer* tk) { return tk ? tk->err.c_str() : g_create_error.c_str(); }

int spl_launches_per_call(const spl_tokenizer* tk, uint32_t flags) {
    if (!tk) return 0;
    bool ws = (flags & SPL_ENCODE_WITH_SPECIAL) && !tk->host.sp_id.empty();
    int extra = 0;
    if (ws && !tk->host.specials_unambiguous) ws = false, ++extra;      // + k_resolve_specials
    if (is_sentencepiece(tk)) return 12 + (ws ? 1 : 0) + 2 * extra;                // mark (T), 4 x scan, sp_emit, mark (T'), probe, bpe, bpe_long, bpe_fin (+ chunk scan), emit
    int pre = tk->host.pattern == SPL_PAT_MISTRAL_V3 ? 1 : 2;          // sequential rules | bit-parallel + fallback
    return 6 + pre + (ws ? 1 : 0) + 2 * extra;                         // mark_docs, probe, bpe, bpe_long, bpe_fin (+ chunk scan), emit
}

int spl_set_profiling(spl_tokenizer* tk, int enable) {
    if (!tk) return SPL_ERR_INVALID_ARG;
    tk->profiling = enable != 0;
    return SPL_OK;
}

int spl_last_kernel_times(spl_tokenizer* tk, int dev_index, const char** names, float* ms, int cap) {
    if (!tk || dev_index < 0 || (size_t)dev_index >= tk->devs.size()) return SPL_ERR_INVALID_ARG;
    DevCtx& dc = tk->devs[dev_index];
    if (!dc.prof_ready) return 0;
    int n = std::min(dc.prof.n, cap);
    for (int i = 0; i < n; ++i) {
        float t = 0;
        if (cudaEventElapsedTime(&t, dc.prof.ev[i], dc.prof.ev[i + 1]) != cudaSuccess) { cudaGetLastError(); t = -1.f; }
        if (names) names[i] = dc.prof.name[i];
        if (ms) ms[i] = t;
    }
    return n;
}

void* spl_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, std::max<size_t>(bytes, 64), cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void spl_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int spl_encode_batch_device(spl_tokenizer* tk, int dev_index, const uint8_t* d_bytes, size_t n_bytes,
                            const uint64_t* d_offsets, size_t n_docs, uint32_t flags,
                            uint32_t* d_ids, size_t ids_capacity, uint64_t* d_out_offsets,
                            void* cuda_stream, uint64_t* n_tokens_out) {
    if (!tk) return SPL_ERR_INVALID_ARG;
    if (dev_index < 0 || (size_t)dev_index >= tk->devs.size() || !d_offsets || !d_out_offsets ||
        (n_bytes && (!d_bytes || !d_ids)) || (!is_sentencepiece(tk) && ids_capacity < n_bytes) || ((uintptr_t)d_bytes & 15u)) {
        tk->err = "invalid argument (null pointer, ids_capacity < n_bytes, or d_bytes not 16-byte aligned)";
        return SPL_ERR_INVALID_ARG;
    }
    bool with_special;
    int rc = check_special_support(tk, flags, with_special);
    if (rc) return rc;
    DeviceGuard guard;
    DevCtx& dc = tk->devs[dev_index];
    CUDA_TRY(cudaSetDevice(dc.device), tk->err);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    for (int attempt = 0;; ++attempt) {
        SplWork w;
        SplKernelProfile* prof = nullptr;
        if (tk->profiling) {
            if (!dc.prof_ready) {
                for (auto& e : dc.prof.ev) CUDA_TRY(cudaEventCreate(&e), tk->err);
                dc.prof.n = 0;
                dc.prof_ready = true;
            }
            prof = &dc.prof;
        }
        int launches = 0;
        EncodeArgs ea{d_bytes, n_bytes, d_offsets, 0, n_docs, d_ids, ids_capacity, d_out_offsets, nullptr, nullptr, nullptr};
        // one graph launch (k_bpe_long beside k_bpe) unless per-kernel events are wanted
        if (tk->use_graph && !prof && !is_sentencepiece(tk)) rc = enqueue_encode_graph(tk, dc, st, ea, with_special, w, launches);
        else rc = enqueue_encode(tk, dc, st, ea, with_special, prof, w, launches);
        if (rc) return rc;
        CUDA_TRY(cudaGetLastError(), tk->err);
        if (!n_tokens_out) return SPL_OK;
        uint32_t h_counters[4];
        uint64_t total = 0;
        CUDA_TRY(cudaMemcpyAsync(h_counters, w.counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st), tk->err);
        CUDA_TRY(cudaMemcpyAsync(&total, d_out_offsets + n_docs, 8, cudaMemcpyDeviceToHost, st), tk->err);
        CUDA_TRY(cudaStreamSynchronize(st), tk->err);
        if (h_counters[1] & SPL_DEVERR_OFFSETS) { tk->err = "document offsets are not a non-decreasing sequence from 0 to n_bytes"; return SPL_ERR_INVALID_ARG; }
        if (h_counters[1] & SPL_DEVERR_HUGE_POOL) {
            if (attempt >= 3) { tk->err = "scratch pool for very long pieces exhausted"; return SPL_ERR_OOM; }
            const uint64_t need = (uint64_t)h_counters[2] | ((uint64_t)h_counters[3] << 32);      // what this pass asked for in all
            dc.huge_words = std::max<size_t>(dc.huge_words * 2, (size_t)need + 1024);
            continue;
        }
        *n_tokens_out = total;
        return SPL_OK;
    }
}

int spl_device_status(spl_tokenizer* tk, int dev_index, void* cuda_stream, uint32_t* flags_out) {
    if (!tk || !flags_out || dev_index < 0 || (size_t)dev_index >= tk->devs.size()) return SPL_ERR_INVALID_ARG;
    DeviceGuard guard;
    DevCtx& dc = tk->devs[dev_index];
    CUDA_TRY(cudaSetDevice(dc.device), tk->err);
    *flags_out = 0;
    if (!dc.zero.p) return SPL_OK;                                // no pass yet
    uint32_t h = 0;
    static_assert(SPL_DEVERR_OFFSETS == SPL_STATUS_BAD_OFFSETS && SPL_DEVERR_HUGE_POOL == SPL_STATUS_SCRATCH_EXHAUSTED, "flag values");
    CUDA_TRY(cudaMemcpyAsync(&h, (const uint32_t*)dc.zero.p + SPL_CTR_ERR, 4, cudaMemcpyDeviceToHost, (cudaStream_t)cuda_stream), tk->err);
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)cuda_stream), tk->err);
    *flags_out = h;
    return SPL_OK;
}

int spl_debug_counters(spl_tokenizer* tk, int dev_index, void* cuda_stream, uint32_t* out32) {
    if (!tk || !out32 || dev_index < 0 || (size_t)dev_index >= tk->devs.size()) return SPL_ERR_INVALID_ARG;
    DeviceGuard guard;
    DevCtx& dc = tk->devs[dev_index];
    CUDA_TRY(cudaSetDevice(dc.device), tk->err);
    memset(out32, 0, SPL_CTR_WORDS * 4);
    if (!dc.zero.p) return SPL_OK;
    CUDA_TRY(cudaMemcpyAsync(out32, dc.zero.p, SPL_CTR_WORDS * 4, cudaMemcpyDeviceToHost, (cudaStream_t)cuda_stream), tk->err);
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)cuda_stream), tk->err);
    return SPL_OK;
}

// One pipeline stage of spl_encode_batch: a contiguous range of documents of one device.
struct Chunk {
    int g;                      // device index
    size_t d0, d1;              // documents [d0, d1)
    uint64_t b0, b1;            // bytes [b0, b1) of the caller's buffer
    size_t text_off;            // 16-byte aligned offset of the chunk inside the device text buffer
    size_t ids_off;             // u32 index of the chunk's id region inside the device id buffer
    size_t ev;                  // first of its 4 timing events in DevCtx::pipe_ev (copied in*, kernels start, kernels end,
                                // ids copied out*; * = trace only); its 2 ordering events are sync_ev[ev / 2 ..] (in, done)
    size_t slot;                // index of the chunk within its device's shard (running-total slot)
    size_t doc_slot;            // first entry of the chunk in the device's doc_off / out_off buffers
    bool first, last;           // first / last chunk of its device
    volatile uint64_t* meta;    // pinned, mapped: [0] id count, [1] error flags | huge-pool need << 32
    uint64_t* d_meta;           // the same memory as the device sees it
    uint64_t n_tokens;
    uint64_t tok_base;          // position of its ids in the result
};

int spl_encode_batch(spl_tokenizer* tk, const uint8_t* bytes, const uint64_t* offsets, size_t n_docs,
                     uint32_t flags, spl_result** out) {
    if (!tk || !out) return SPL_ERR_INVALID_ARG;
    *out = nullptr;
    const auto h_t0 = std::chrono::steady_clock::now();
    auto h_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h_t0).count(); };
    double h_plan = 0, h_reserved = 0, h_enq = 0, h_drained = 0, h_synced = 0;
    if (!offsets || offsets[0] != 0) { tk->err = "offsets must start at 0"; return SPL_ERR_INVALID_ARG; }
    const uint64_t N = offsets[n_docs];
    if (N && !bytes) { tk->err = "null bytes"; return SPL_ERR_INVALID_ARG; }
    for (size_t i = 0; i < n_docs; ++i)
        if (offsets[i + 1] < offsets[i]) { tk->err = "offsets must be non-decreasing"; return SPL_ERR_INVALID_ARG; }
    bool with_special;
    int rc = check_special_support(tk, flags, with_special);
    if (rc) return rc;

    // ---- documents -> pipeline chunks -> devices ----------------------------------------------------------------
    // One device: its shard is the whole batch, cut into chunks.  Several devices: the chunks of the WHOLE batch, in
    // document order, go to the devices round robin.  (Contiguous shards per device serialised the devices: the ids of
    // device g + 1 belong behind ALL ids of device g in the result, so its copy-out could only start once device g
    // had finished -- measured 44 GB/s on two GPUs where one gives 36.  With interleaved chunks the position of a chunk's
    // ids is known as soon as the chunk in front of it -- on the neighbouring device -- has been counted, and all
    // devices copy in and out at the same time.)
    const size_t G = tk->devs.size();
    const bool rr = G > 1;
    std::vector<size_t> dlo(G + 1, 0);
    for (size_t g = 1; g <= G; ++g) dlo[g] = n_docs;               // (one device: [0, n_docs); round robin: not used)
    std::vector<Chunk> chunks;
    std::vector<size_t> text_need(G, 0), max_nb(G, 0), max_nd(G, 0), ids_need(G, 0), doc_need(G, 0), n_chunks(G, 0);
    {
        const uint64_t per_dev = N / G;
        uint64_t target = tk->chunk_bytes ? tk->chunk_bytes : std::min<uint64_t>(std::max<uint64_t>(per_dev / tk->chunks_per_dev, 4u << 20), 256u << 20);
        target = std::min<uint64_t>(target, kMaxShardBytes / (is_sentencepiece(tk) ? 6 : 2));
        // the pipeline fills with the first chunk's copy-in and drains with the last chunk's copy-out: ramp the chunk
        // size up at the start and down at the end (quarter, half, full ... full, half, quarter) unless it was pinned
        const bool ramp = !tk->chunk_bytes && per_dev >= 4 * target;
        size_t d = 0, k = 0;
        do {
            Chunk c;
            memset(&c, 0, sizeof(c));
            const size_t g = rr ? k % G : 0;
            c.g = (int)g; c.d0 = d; c.b0 = offsets[d];
            uint64_t want = target;
            const uint64_t rem = N - c.b0;
            if (ramp) {
                // every device starts small (target / ramp_div[round]) ...
                const size_t round = k / G;
                if (round < tk->ramp_div.size()) want = std::max<uint64_t>(target / tk->ramp_div[round], 1);
                // ... and ends small: the last rounds mirror the first ones
                const uint64_t last_div = tk->ramp_div.empty() ? 1 : tk->ramp_div[0];
                if (rem <= (target / last_div + target / (2 * last_div)) * G) want = std::max<uint64_t>(rem / G, 1);
                else if (rem <= target * G) want = std::max<uint64_t>((rem - target / last_div * G) / G, 1);
            }
            size_t e = std::upper_bound(offsets + d, offsets + n_docs + 1, c.b0 + want) - offsets;   // first doc end beyond the target
            e = std::min(std::max(e, d + 1), n_docs);
            if (!rr) {
                if (!ramp && rem <= target + target / 4) e = n_docs;                              // no runt at the end
                if (ramp && rem <= want + want / 8) e = n_docs;
            } else if (rem <= want + want / 8) e = n_docs;
            if (d == n_docs) e = d;                                                                // batch without documents
            c.d1 = e; c.b1 = offsets[e];
            c.text_off = text_need[g]; c.ids_off = ids_need[g]; c.doc_slot = doc_need[g]; c.slot = n_chunks[g]++;
            c.first = c.slot == 0;
            text_need[g] += align_up((size_t)(c.b1 - c.b0) + 16, 16);
            ids_need[g] += (size_t)ids_bound(tk, c.b1 - c.b0) + (rr ? 16 : 0);
            doc_need[g] += (c.d1 - c.d0) + (rr ? 1 : 0);
            max_nb[g] = std::max<size_t>(max_nb[g], (size_t)(c.b1 - c.b0));
            max_nd[g] = std::max<size_t>(max_nd[g], c.d1 - c.d0);
            chunks.push_back(c);
            d = e; ++k;
        } while (d < n_docs);
        for (size_t g = 0; g < G; ++g) { text_need[g] += 64; ids_need[g] += 16; doc_need[g] += 1; }
        std::vector<int> seen(G, 0);
        for (size_t i = chunks.size(); i-- > 0;) { chunks[i].last = !seen[chunks[i].g]; seen[chunks[i].g] = 1; }
    }
    const size_t C = chunks.size();
    h_plan = h_ms();

    DeviceGuard guard;
    spl_result* r = new (std::nothrow) spl_result();
    if (!r) return SPL_ERR_OOM;
    memset(&r->stats, 0, sizeof(r->stats));
    r->owner = tk; r->n_docs = n_docs; r->n_tokens = 0;
    r->ids_buf = PinnedBuf{nullptr, 0};
    r->off_buf = take_pinned(tk, (n_docs + 1) * 8);
    PinnedBuf meta_buf = take_pinned(tk, C * 32);
    std::vector<PinnedBuf> stage_off(G, PinnedBuf{nullptr, 0});      // round robin: every device's chunk-relative output offsets
    auto fail = [&](int code) {
        for (auto& dc : tk->devs) { cudaSetDevice(dc.device); cudaStreamSynchronize(dc.s_in); cudaStreamSynchronize(dc.stream); cudaStreamSynchronize(dc.s_out); }
        cudaGetLastError();
        give_pinned(tk, r->off_buf); give_pinned(tk, r->ids_buf); give_pinned(tk, meta_buf);
        for (auto& b : stage_off) give_pinned(tk, b);
        delete r;
        return code;
    };
    if (!r->off_buf.p || !meta_buf.p) { tk->err = "pinned host allocation failed"; return fail(SPL_ERR_OOM); }
    if (rr)
        for (size_t g = 0; g < G; ++g) {
            stage_off[g] = take_pinned(tk, (doc_need[g] + 1) * 8);
            if (!stage_off[g].p) { tk->err = "pinned host allocation failed"; return fail(SPL_ERR_OOM); }
        }
    uint64_t* res_off = (uint64_t*)r->off_buf.p;
    {
        void* d_meta_base = nullptr;
        if (cudaHostGetDevicePointer(&d_meta_base, meta_buf.p, 0) != cudaSuccess) { cudaGetLastError(); tk->err = "pinned host memory is not mapped"; return fail(SPL_ERR_CUDA); }
        for (size_t c = 0; c < C; ++c) {
            chunks[c].meta = (volatile uint64_t*)((uint8_t*)meta_buf.p + c * 32);
            chunks[c].d_meta = (uint64_t*)((uint8_t*)d_meta_base + c * 32);
        }
    }

    // ---- device buffers: whole-shard text / ids / offsets, workspace for the largest chunk ----
    for (size_t g = 0; g < G; ++g) {
        DevCtx& dc = tk->devs[g];
        auto reserve = [&]() -> int {
            CUDA_TRY(cudaSetDevice(dc.device), tk->err);
            int rc2;
            if ((rc2 = dc.text.ensure(text_need[g], tk->err))) return rc2;
            if ((rc2 = dc.doc_off.ensure((doc_need[g] + 1) * 8, tk->err))) return rc2;
            if ((rc2 = dc.ids.ensure((ids_need[g] + 16) * 4, tk->err))) return rc2;
            if ((rc2 = dc.out_off.ensure((doc_need[g] + 1) * 8, tk->err))) return rc2;
            if ((rc2 = reserve_work(tk, dc, ids_bound(tk, max_nb[g]), max_nd[g], with_special))) return rc2;
            size_t n_ev = 0, n_slot = 0;
            for (auto& c : chunks) if (c.g == (int)g) { c.ev = n_ev; n_ev += 4; c.slot = n_slot++; }
            while (dc.pipe_ev.size() < n_ev + 1) {
                cudaEvent_t e;
                CUDA_TRY(cudaEventCreate(&e), tk->err);
                dc.pipe_ev.push_back(e);
            }
            while (dc.sync_ev.size() < n_ev / 2 + 1) {
                cudaEvent_t e;
                CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), tk->err);
                dc.sync_ev.push_back(e);
            }
            if ((rc2 = dc.run_tot.ensure((n_slot + 2) * 8, tk->err))) return rc2;
            return SPL_OK;
        };
        if ((rc = reserve())) return fail(rc);
    }

    h_reserved = h_ms();
    int launches = 0;
    for (int attempt = 0;; ++attempt) {
        launches = 0;
        memset(&r->stats, 0, sizeof(r->stats));
        bool retry = false;
        int err_code = SPL_OK;
        std::vector<uint64_t> huge_need(G, 0);   // words of huge-piece scratch the chunks of each device asked for (max)
        uint64_t total = 0;                 // ids of the chunks drained so far
        size_t next_out = 0;                // chunks are drained in document order

        // copy the ids of chunk c out as soon as its count is known; grows the result buffer on demand
        auto drain = [&](size_t ci) -> int {
            Chunk& c = chunks[ci];
            DevCtx& dc = tk->devs[c.g];
            const uint32_t errbits = (uint32_t)c.meta[1];
            if (errbits & SPL_DEVERR_OFFSETS) { tk->err = "invalid document offsets"; return SPL_ERR_INVALID_ARG; }
            if (errbits & SPL_DEVERR_HUGE_POOL) {
                // remember the largest request of any chunk of this device; the pool is resized ONCE before the retry
                huge_need[c.g] = std::max<uint64_t>(huge_need[c.g], (uint64_t)c.meta[2]);
                retry = true;
                return SPL_OK;
            }
            c.n_tokens = c.meta[0];
            c.tok_base = total;
            total += c.n_tokens;
            if (retry) return SPL_OK;
            if ((total + 16) * 4 > r->ids_buf.cap) {
                // size the result by the ids-per-byte ratio seen so far (+12 %), at least what is needed now
                const uint64_t done_bytes = std::max<uint64_t>(c.b1, 1);
                uint64_t est = (uint64_t)((double)total / (double)done_bytes * (double)N * 1.125) + 4096;
                est = std::min<uint64_t>(std::max<uint64_t>(est, total + 16), ids_bound(tk, N) + 16);
                PinnedBuf nb = take_pinned(tk, est * 4);
                if (!nb.p) { tk->err = "pinned host allocation failed"; return SPL_ERR_OOM; }
                if (r->ids_buf.p) {
                    for (auto& d2 : tk->devs) { cudaSetDevice(d2.device); cudaStreamSynchronize(d2.s_out); }
                    memcpy(nb.p, r->ids_buf.p, (size_t)c.tok_base * 4);
                    give_pinned(tk, r->ids_buf);
                }
                r->ids_buf = nb;
            }
            CUDA_TRY(cudaSetDevice(dc.device), tk->err);
            if (c.n_tokens)
                CUDA_TRY(cudaMemcpyAsync((uint32_t*)r->ids_buf.p + c.tok_base, (uint32_t*)dc.ids.p + c.ids_off, c.n_tokens * 4,
                                         cudaMemcpyDeviceToHost, dc.s_out), tk->err);
            r->stats.d2h_bytes += c.n_tokens * 4;
            if (tk->trace) cudaEventRecord(dc.pipe_ev[c.ev + 3], dc.s_out);
            const size_t g = (size_t)c.g;
            if (c.last) {
                // last chunk of the device: the per-document offsets of all its chunks in one copy (one small copy per
                // chunk would cost the copy engine more than it moves).  One device: they are batch-relative (every
                // chunk's k_emit adds the ids of the chunks before it, kept on the device) and go straight to the result.
                // Round robin: chunk-relative, staged, and put in place with the chunk's base once everything has arrived.
                const size_t n_off = rr ? doc_need[g] : n_docs + 1;
                void* dst = rr ? stage_off[g].p : (void*)res_off;
                if (n_off)
                    CUDA_TRY(cudaMemcpyAsync(dst, dc.out_off.p, n_off * 8, cudaMemcpyDeviceToHost, dc.s_out), tk->err);
                r->stats.d2h_bytes += n_off * 8;
            }
            return SPL_OK;
        };

        for (size_t ci = 0; ci < C && err_code == SPL_OK; ++ci) {
            Chunk& c = chunks[ci];
            DevCtx& dc = tk->devs[c.g];
            auto enqueue = [&]() -> int {
                CUDA_TRY(cudaSetDevice(dc.device), tk->err);
                const size_t nd = c.d1 - c.d0, g = (size_t)c.g;
                const uint64_t nb = c.b1 - c.b0;
                const bool first = c.first;
                cudaEvent_t ev_in = dc.sync_ev[c.ev / 2], ev_done = dc.sync_ev[c.ev / 2 + 1];
                cudaEvent_t ev_k0 = dc.pipe_ev[c.ev + 1], ev_k1 = dc.pipe_ev[c.ev + 2];
                uint64_t* run_tot = (uint64_t*)dc.run_tot.p;
                if (first) CUDA_TRY(cudaEventRecord(dc.ev[0], dc.s_in), tk->err);
                if (first && !rr) {
                    // the document offsets of the whole batch go in with one copy, ahead of the text
                    CUDA_TRY(cudaMemcpyAsync(dc.doc_off.p, offsets, (n_docs + 1) * 8, cudaMemcpyHostToDevice, dc.s_in), tk->err);
                    CUDA_TRY(cudaMemsetAsync(run_tot, 0, 8, dc.s_in), tk->err);
                    r->stats.h2d_bytes += (n_docs + 1) * 8;
                }
                // stage in: text to its aligned slot
                uint8_t* d_text = (uint8_t*)dc.text.p + c.text_off;
                uint64_t* d_doc = (uint64_t*)dc.doc_off.p + c.doc_slot;
                if (rr) {                                          // round robin: the chunk's own slice of the document offsets
                    CUDA_TRY(cudaMemcpyAsync(d_doc, offsets + c.d0, (nd + 1) * 8, cudaMemcpyHostToDevice, dc.s_in), tk->err);
                    r->stats.h2d_bytes += (nd + 1) * 8;
                }
                if (nb) CUDA_TRY(cudaMemcpyAsync(d_text, bytes + c.b0, nb, cudaMemcpyHostToDevice, dc.s_in), tk->err);
                if (tk->trace) CUDA_TRY(cudaEventRecord(dc.pipe_ev[c.ev], dc.s_in), tk->err);
                CUDA_TRY(cudaEventRecord(ev_in, dc.s_in), tk->err);
                r->stats.h2d_bytes += nb;
                // kernels
                CUDA_TRY(cudaStreamWaitEvent(dc.stream, ev_in, 0), tk->err);
                CUDA_TRY(cudaEventRecord(ev_k0, dc.stream), tk->err);
                SplWork w;
                int rc2;
                uint64_t* d_out = (uint64_t*)dc.out_off.p + c.doc_slot;
                EncodeArgs ea{d_text, nb, d_doc, c.b0, nd, (uint32_t*)dc.ids.p + c.ids_off, ids_bound(tk, nb), d_out, c.d_meta,
                              rr ? nullptr : run_tot + c.slot, rr ? nullptr : run_tot + c.slot + 1};
                SplKernelProfile* prof = nullptr;
                if (tk->trace && (int)c.slot == tk->trace_chunk) {        // SPL_TRACE_CHUNK=k: per-kernel times of chunk k
                    if (!dc.prof_ready) {
                        for (auto& e : dc.prof.ev) CUDA_TRY(cudaEventCreate(&e), tk->err);
                        dc.prof.n = 0;
                        dc.prof_ready = true;
                    }
                    prof = &dc.prof;
                }
                if (tk->use_graph && !prof && !is_sentencepiece(tk)) rc2 = enqueue_encode_graph(tk, dc, dc.stream, ea, with_special, w, launches);
                else rc2 = enqueue_encode(tk, dc, dc.stream, ea, with_special, prof, w, launches);
                if (rc2) return rc2;
                CUDA_TRY(cudaGetLastError(), tk->err);
                CUDA_TRY(cudaEventRecord(ev_k1, dc.stream), tk->err);
                // the chunk's id count and error flags arrive in mapped host memory (written by k_emit): no copy on the
                // kernel stream, which would queue behind the previous chunk's ids on the device-to-host engine
                CUDA_TRY(cudaEventRecord(ev_done, dc.stream), tk->err);
                r->stats.d2h_bytes += 24;
                return SPL_OK;
            };
            err_code = enqueue();
            // drain what has finished, without blocking
            while (err_code == SPL_OK && next_out <= ci) {
                Chunk& o = chunks[next_out];
                cudaSetDevice(tk->devs[o.g].device);
                cudaError_t q = cudaEventQuery(tk->devs[o.g].sync_ev[o.ev / 2 + 1]);
                if (q == cudaErrorNotReady) break;
                if (q != cudaSuccess) { tk->err = std::string("encode kernels: ") + cudaGetErrorString(q); err_code = SPL_ERR_CUDA; break; }
                err_code = drain(next_out++);
            }
        }
        h_enq = h_ms();
        while (err_code == SPL_OK && next_out < C) {
            Chunk& o = chunks[next_out];
            cudaSetDevice(tk->devs[o.g].device);
            cudaError_t e = cudaEventSynchronize(tk->devs[o.g].sync_ev[o.ev / 2 + 1]);
            if (e != cudaSuccess) { tk->err = std::string("encode kernels: ") + cudaGetErrorString(e); err_code = SPL_ERR_CUDA; break; }
            err_code = drain(next_out++);
        }
        if (err_code != SPL_OK) return fail(err_code);
        h_drained = h_ms();
        float kmax = 0, tmax = 0;
        for (size_t g = 0; g < G; ++g) {
            DevCtx& dc = tk->devs[g];
            if (!n_chunks[g]) continue;                        // round robin over few chunks: this device had nothing to do
            cudaSetDevice(dc.device);
            cudaEventRecord(dc.ev[3], dc.s_out);
            cudaError_t e = cudaStreamSynchronize(dc.s_out);
            if (e == cudaSuccess) e = cudaStreamSynchronize(dc.stream);
            if (e != cudaSuccess) { tk->err = std::string("copy-out: ") + cudaGetErrorString(e); return fail(SPL_ERR_CUDA); }
            float ksum = 0, t = 0;
            for (auto& c : chunks)
                if (c.g == (int)g) { float k = 0; cudaEventElapsedTime(&k, dc.pipe_ev[c.ev + 1], dc.pipe_ev[c.ev + 2]); ksum += k; }
            cudaEventElapsedTime(&t, dc.ev[0], dc.ev[3]);
            kmax = std::max(kmax, ksum); tmax = std::max(tmax, t);
            if (tk->trace && !retry)
                for (auto& c : chunks)
                    if (c.g == (int)g) {
                        float a = 0, b = 0, d = 0, e = 0;
                        cudaEventElapsedTime(&a, dc.ev[0], dc.pipe_ev[c.ev]);
                        cudaEventElapsedTime(&b, dc.ev[0], dc.pipe_ev[c.ev + 1]);
                        cudaEventElapsedTime(&d, dc.ev[0], dc.pipe_ev[c.ev + 2]);
                        cudaEventElapsedTime(&e, dc.ev[0], dc.pipe_ev[c.ev + 3]);
                        fprintf(stderr, "[spl trace] dev %zu docs %zu..%zu bytes %llu: in %.3f  k0 %.3f  done %.3f  out %.3f ms\n",
                                g, c.d0, c.d1, (unsigned long long)(c.b1 - c.b0), a, b, d, e);
                        if ((int)c.slot == tk->trace_chunk && dc.prof_ready) {
                            fprintf(stderr, "[spl trace]   kernels of this chunk (us):");
                            for (int i = 0; i < dc.prof.n; ++i) {
                                float t2 = 0;
                                cudaEventElapsedTime(&t2, dc.prof.ev[i], dc.prof.ev[i + 1]);
                                fprintf(stderr, " %s %.0f", dc.prof.name[i], t2 * 1000.f);
                            }
                            float t0 = 0;
                            cudaEventElapsedTime(&t0, dc.pipe_ev[c.ev + 1], dc.prof.ev[0]);
                            fprintf(stderr, " | k0 -> first kernel %.0f\n", t0 * 1000.f);
                        }
                    }
        }
        h_synced = h_ms();
        if (retry) {
            if (attempt >= 3) { tk->err = "scratch pool for very long pieces exhausted"; return fail(SPL_ERR_OOM); }
            for (size_t g = 0; g < G; ++g) {
                cudaSetDevice(tk->devs[g].device);
                if (huge_need[g]) tk->devs[g].huge_words = std::max<size_t>(tk->devs[g].huge_words * 2, (size_t)huge_need[g] + 1024);
                if ((rc = reserve_work(tk, tk->devs[g], ids_bound(tk, max_nb[g]), max_nd[g], with_special))) return fail(rc);
            }
            continue;
        }
        if (!r->ids_buf.p) {
            r->ids_buf = take_pinned(tk, 64);
            if (!r->ids_buf.p) { tk->err = "pinned host allocation failed"; return fail(SPL_ERR_OOM); }
        }
        // round robin: the per-document offsets came back chunk-relative, per device: every chunk's slice gets the chunk's
        // position in the result (all copies have completed)
        if (rr)
            for (auto& c : chunks) {
                const uint64_t* src = (const uint64_t*)stage_off[c.g].p + c.doc_slot;
                for (size_t d = c.d0; d < c.d1; ++d) res_off[d] = src[d - c.d0] + c.tok_base;
            }
        res_off[n_docs] = total;
        r->n_tokens = total;
        r->stats.n_docs = n_docs; r->stats.n_bytes = N; r->stats.n_tokens = total;
        r->stats.kernel_ms = kmax; r->stats.total_ms = tmax;
        r->stats.n_devices = (int)G; r->stats.n_launches = launches;
        break;
    }
    give_pinned(tk, meta_buf);
    for (auto& b : stage_off) give_pinned(tk, b);
    *out = r;
    if (tk->trace)
        fprintf(stderr, "[spl trace] host ms: plan %.3f  reserved %.3f  enqueued %.3f  drained %.3f  synced %.3f  end %.3f  (%zu chunks)\n",
                h_plan, h_reserved, h_enq, h_drained, h_synced, h_ms(), C);
    return SPL_OK;
}


// ---- decode (row N2): ids -> bytes --------------------------------------------------------------------

int spl_decode_batch_device(spl_tokenizer* tk, int dev_index, const uint32_t* d_ids, size_t n_tokens,
                            const uint64_t* d_tok_offsets, size_t n_docs,
                            uint8_t* d_bytes_out, size_t bytes_capacity, uint64_t* d_out_offsets,
                            void* cuda_stream, uint64_t* n_bytes_out) {
    if (!tk) return SPL_ERR_INVALID_ARG;
    if (dev_index < 0 || (size_t)dev_index >= tk->devs.size() || !d_tok_offsets || !d_out_offsets || !n_bytes_out ||
        (n_tokens && !d_ids)) {
        tk->err = "invalid argument (null pointer)";
        return SPL_ERR_INVALID_ARG;
    }
    DeviceGuard guard;
    DevCtx& dc = tk->devs[dev_index];
    CUDA_TRY(cudaSetDevice(dc.device), tk->err);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    SplDecLaunch L;
    memset(&L, 0, sizeof(L));
    L.ids = d_ids; L.n_tok = n_tokens; L.tok_off = d_tok_offsets; L.n_docs = n_docs;
    L.n_tiles = (uint32_t)(n_tokens / SPL_DEC_TILE) + 1;
    if (n_tokens / SPL_DEC_TILE >= 0xFFFFFFF0ull) { tk->err = "too many ids for one device pass"; return SPL_ERR_UNSUPPORTED; }
    int rc;
    const size_t ws_sum = align_up((size_t)L.n_tiles * 4, 256);
    if ((rc = dc.dec_ws.ensure(ws_sum + ((size_t)L.n_tiles + 1) * 8, tk->err))) return rc;
    L.tile_sum = (uint32_t*)dc.dec_ws.p;
    L.tile_pref = (uint64_t*)((uint8_t*)dc.dec_ws.p + ws_sum);
    L.out = d_bytes_out; L.capacity = bytes_capacity; L.out_off = d_out_offsets; L.T = dc.d_tables;
    spl_launch_decode_count(L, st);
    uint64_t total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, L.tile_pref + L.n_tiles, 8, cudaMemcpyDeviceToHost, st), tk->err);
    CUDA_TRY(cudaStreamSynchronize(st), tk->err);
    *n_bytes_out = total;
    if (total > bytes_capacity || (total && !d_bytes_out)) {
        tk->err = "output capacity too small for the decoded bytes (needed size returned in n_bytes_out)";
        return SPL_ERR_INVALID_ARG;
    }
    spl_launch_decode_emit(L, st);
    CUDA_TRY(cudaGetLastError(), tk->err);
    CUDA_TRY(cudaStreamSynchronize(st), tk->err);
    return SPL_OK;
}

int spl_decode_batch(spl_tokenizer* tk, const uint32_t* ids, const uint64_t* offsets, size_t n_docs, spl_result** out) {
    if (!tk || !out) return SPL_ERR_INVALID_ARG;
    *out = nullptr;
    if (!offsets || offsets[0] != 0) { tk->err = "offsets must start at 0"; return SPL_ERR_INVALID_ARG; }
    for (size_t i = 0; i < n_docs; ++i)
        if (offsets[i + 1] < offsets[i]) { tk->err = "offsets must be non-decreasing"; return SPL_ERR_INVALID_ARG; }
    const uint64_t n_tok = offsets[n_docs];
    if (n_tok && !ids) { tk->err = "null ids"; return SPL_ERR_INVALID_ARG; }
    DeviceGuard guard;
    DevCtx& dc = tk->devs[0];
    CUDA_TRY(cudaSetDevice(dc.device), tk->err);
    int rc;
    if ((rc = dc.dec_ids.ensure((size_t)(n_tok + 16) * 4, tk->err))) return rc;
    if ((rc = dc.dec_off.ensure((n_docs + 1) * 8, tk->err))) return rc;
    if ((rc = dc.dec_out_off.ensure((n_docs + 1) * 8, tk->err))) return rc;
    cudaStream_t st = dc.stream;
    cudaEvent_t e0 = dc.ev[0], e1 = dc.ev[1];
    CUDA_TRY(cudaEventRecord(e0, st), tk->err);
    if (n_tok) CUDA_TRY(cudaMemcpyAsync(dc.dec_ids.p, ids, n_tok * 4, cudaMemcpyHostToDevice, st), tk->err);
    CUDA_TRY(cudaMemcpyAsync(dc.dec_off.p, offsets, (n_docs + 1) * 8, cudaMemcpyHostToDevice, st), tk->err);
    // first attempt with the capacity of the previous call (at least 4 bytes per id), retry once with the exact size
    uint64_t total = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        size_t cap = std::max<size_t>(dc.dec_out.cap, (size_t)n_tok * 4 + 256);
        if ((rc = dc.dec_out.ensure(std::max<size_t>(cap, (size_t)total + 256), tk->err))) return rc;
        rc = spl_decode_batch_device(tk, 0, (const uint32_t*)dc.dec_ids.p, (size_t)n_tok, (const uint64_t*)dc.dec_off.p, n_docs,
                                     (uint8_t*)dc.dec_out.p, dc.dec_out.cap, (uint64_t*)dc.dec_out_off.p, st, &total);
        if (rc == SPL_OK) break;
        if (rc != SPL_ERR_INVALID_ARG || total <= dc.dec_out.cap || attempt == 1) return rc;
    }
    spl_result* r = new (std::nothrow) spl_result();
    if (!r) return SPL_ERR_OOM;
    memset(&r->stats, 0, sizeof(r->stats));
    r->owner = tk; r->n_docs = n_docs; r->n_tokens = (size_t)n_tok; r->n_bytes = (size_t)total;
    r->ids_buf = take_pinned(tk, (size_t)total + 64);
    r->off_buf = take_pinned(tk, (n_docs + 1) * 8);
    if (!r->ids_buf.p || !r->off_buf.p) {
        give_pinned(tk, r->ids_buf); give_pinned(tk, r->off_buf);
        delete r;
        tk->err = "pinned host allocation failed";
        return SPL_ERR_OOM;
    }
    auto finish = [&]() -> int {
        if (total) CUDA_TRY(cudaMemcpyAsync(r->ids_buf.p, dc.dec_out.p, (size_t)total, cudaMemcpyDeviceToHost, st), tk->err);
        CUDA_TRY(cudaMemcpyAsync(r->off_buf.p, dc.dec_out_off.p, (n_docs + 1) * 8, cudaMemcpyDeviceToHost, st), tk->err);
        CUDA_TRY(cudaEventRecord(e1, st), tk->err);
        CUDA_TRY(cudaStreamSynchronize(st), tk->err);
        return SPL_OK;
    };
    if ((rc = finish())) { give_pinned(tk, r->ids_buf); give_pinned(tk, r->off_buf); delete r; return rc; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    r->stats.n_docs = n_docs; r->stats.n_tokens = n_tok; r->stats.n_bytes = total;
    r->stats.h2d_bytes = n_tok * 4 + (n_docs + 1) * 8; r->stats.d2h_bytes = total + (n_docs + 1) * 8;
    r->stats.total_ms = ms; r->stats.n_devices = 1; r->stats.n_launches = 3;
    *out = r;
    return SPL_OK;
}

// ---- ingestion (row N4): JSON Lines -> packed text + offsets ------------------------------------------------

int spl_ingest_jsonl_device(spl_tokenizer* tk, int dev_index, const uint8_t* d_jsonl, size_t n_bytes, const char* field,
                            uint8_t* d_text_out, size_t text_capacity, uint64_t* d_offsets_out, size_t offsets_capacity,
                            void* cuda_stream, spl_ingest_stats* stats) {
    if (!tk) return SPL_ERR_INVALID_ARG;
    const size_t flen = field ? strlen(field) : 0;
    if (dev_index < 0 || (size_t)dev_index >= tk->devs.size() || !stats || !field || flen == 0 || flen > 64 ||
        (n_bytes && !d_jsonl) || ((uintptr_t)d_jsonl & 15u)) {
        tk->err = "invalid argument (null pointer, member name of 0 or more than 64 bytes, or d_jsonl not 16-byte aligned)";
        return SPL_ERR_INVALID_ARG;
    }
    DeviceGuard guard;
    DevCtx& dc = tk->devs[dev_index];
    CUDA_TRY(cudaSetDevice(dc.device), tk->err);
    JlOut o{d_text_out, d_text_out ? text_capacity : 0, d_offsets_out, d_offsets_out ? offsets_capacity : 0};
    return ingest_jsonl_core(tk, dc, (cudaStream_t)cuda_stream, d_jsonl, n_bytes, field, flen,
                             [&](size_t, JlOut& out) { out = o; return SPL_OK; }, stats);
}

// File bytes in host memory -> ids in host memory: spl_encode_batch for a JSON Lines file.  Chunks end at line ends;
// every chunk is copied in as it is, ingested (spl_ingest.h) and encoded on the device, its ids copied out while the
// next chunk is being worked on.  One device (the handle's first).
int spl_encode_jsonl(spl_tokenizer* tk, const uint8_t* bytes, size_t n_bytes, const char* field, uint32_t flags,
                     spl_result** out, spl_ingest_stats* ingest_stats) {
    if (!tk || !out) return SPL_ERR_INVALID_ARG;
    *out = nullptr;
    const size_t flen = field ? strlen(field) : 0;
    if ((n_bytes && !bytes) || flen == 0 || flen > 64) { tk->err = "invalid argument (null bytes, member name of 0 or more than 64 bytes)"; return SPL_ERR_INVALID_ARG; }
    bool with_special;
    int rc = check_special_support(tk, flags, with_special);
    if (rc) return rc;
    // ---- chunks that end right behind a '\n' -------------------------------------------------------------------
    struct JChunk { size_t b0, b1, text_off; uint64_t n_docs, doc_base, n_tokens, tok_base; volatile uint64_t* meta; uint64_t* d_meta; int obuf; };
    std::vector<JChunk> chunks;
    {
        // few, large chunks: one thread parses one line, so the ingestion kernels of a chunk take about as long for
        // 12 000 lines as for 100 000, and every chunk costs two stream synchronisations (measured on a 104 MB file:
        // 8 chunks 6.8 ms, 3 chunks 5.1 ms, 1 chunk 5.4 ms)
        uint64_t target = tk->chunk_bytes ? tk->chunk_bytes : std::min<uint64_t>(std::max<uint64_t>(n_bytes / 3, 16u << 20), 256u << 20);
        target = std::min<uint64_t>(target, kMaxShardBytes / (is_sentencepiece(tk) ? 6 : 2));
        size_t b = 0, toff = 0;
        while (b < n_bytes) {
            size_t e = n_bytes;
            if (n_bytes - b > target + target / 4) {
                const void* nl = memchr(bytes + b + target, '\n', n_bytes - (b + target));
                e = nl ? (size_t)((const uint8_t*)nl - bytes) + 1 : n_bytes;
            }
            JChunk c;
            memset(&c, 0, sizeof(c));
            c.b0 = b; c.b1 = e; c.text_off = toff; c.obuf = (int)(chunks.size() & 1);
            toff += align_up(e - b + 16, 16);
            chunks.push_back(c);
            b = e;
        }
    }
    const size_t C = chunks.size();
    size_t max_nb = 0;
    for (auto& c : chunks) max_nb = std::max(max_nb, c.b1 - c.b0);
    if (max_nb > kMaxShardBytes / (is_sentencepiece(tk) ? 3 : 1)) { tk->err = "a single line exceeds what one device pass can hold"; return SPL_ERR_UNSUPPORTED; }

    DeviceGuard guard;
    DevCtx& dc = tk->devs[0];
    CUDA_TRY(cudaSetDevice(dc.device), tk->err);
    spl_result* r = new (std::nothrow) spl_result();
    if (!r) return SPL_ERR_OOM;
    memset(&r->stats, 0, sizeof(r->stats));
    r->owner = tk; r->n_docs = 0; r->n_tokens = 0;
    r->ids_buf = PinnedBuf{nullptr, 0};
    r->off_buf = take_pinned(tk, 4096 * 8);
    PinnedBuf meta_buf = take_pinned(tk, (C + 1) * 32);
    spl_ingest_stats tot;
    memset(&tot, 0, sizeof(tot));
    static thread_local int jsonl_depth = 0;
    auto fail = [&](int code) {
        cudaStreamSynchronize(dc.s_in); cudaStreamSynchronize(dc.stream); cudaStreamSynchronize(dc.s_out);
        cudaGetLastError();
        give_pinned(tk, r->off_buf); give_pinned(tk, r->ids_buf); give_pinned(tk, meta_buf);
        delete r;
        if (code == SPL_ERR_OOM + 1000) {                      // huge-piece scratch was too small and has been enlarged
            if (jsonl_depth >= 3) return (int)SPL_ERR_OOM;
            ++jsonl_depth;
            const int rc2 = spl_encode_jsonl(tk, bytes, n_bytes, field, flags, out, ingest_stats);
            --jsonl_depth;
            return rc2;
        }
        return code;
    };
    if (!r->off_buf.p || !meta_buf.p) { tk->err = "pinned host allocation failed"; return fail(SPL_ERR_OOM); }
    void* d_meta_base = nullptr;
    if (cudaHostGetDevicePointer(&d_meta_base, meta_buf.p, 0) != cudaSuccess) { cudaGetLastError(); tk->err = "pinned host memory is not mapped"; return fail(SPL_ERR_CUDA); }
    auto setup = [&]() -> int {
        int rc2;
        size_t raw_need = 64;
        for (auto& c : chunks) raw_need = std::max(raw_need, c.text_off + (c.b1 - c.b0) + 64);
        if ((rc2 = dc.text.ensure(raw_need, tk->err))) return rc2;
        if ((rc2 = dc.jl_text.ensure(max_nb + 64, tk->err))) return rc2;
        if ((rc2 = dc.ids.ensure((ids_bound(tk, n_bytes) + 16) * 4, tk->err))) return rc2;
        if ((rc2 = reserve_work(tk, dc, ids_bound(tk, max_nb), max_nb / 2 + 2, with_special))) return rc2;
        if ((rc2 = dc.run_tot.ensure((C + 2) * 8, tk->err))) return rc2;
        while (dc.sync_ev.size() < 2 * C + 2) {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), tk->err);
            dc.sync_ev.push_back(e);
        }
        return SPL_OK;
    };
    if ((rc = setup())) return fail(rc);
    uint64_t* run_tot = (uint64_t*)dc.run_tot.p;
    // every chunk of the file goes in right away, one copy each (they do not depend on anything)
    auto copy_in = [&]() -> int {
        CUDA_TRY(cudaEventRecord(dc.ev[0], dc.s_in), tk->err);
        CUDA_TRY(cudaMemsetAsync(run_tot, 0, 8, dc.s_in), tk->err);
        for (size_t k = 0; k < C; ++k) {
            JChunk& c = chunks[k];
            c.meta = (volatile uint64_t*)((uint8_t*)meta_buf.p + k * 32);
            c.d_meta = (uint64_t*)((uint8_t*)d_meta_base + k * 32);
            CUDA_TRY(cudaMemcpyAsync((uint8_t*)dc.text.p + c.text_off, bytes + c.b0, c.b1 - c.b0, cudaMemcpyHostToDevice, dc.s_in), tk->err);
            CUDA_TRY(cudaEventRecord(dc.sync_ev[2 * k], dc.s_in), tk->err);
            r->stats.h2d_bytes += c.b1 - c.b0;
        }
        return SPL_OK;
    };
    if ((rc = copy_in())) return fail(rc);

    uint64_t total = 0, docs = 0;
    int launches = 0;
    uint64_t* res_off = (uint64_t*)r->off_buf.p;
    const auto h_t0 = std::chrono::steady_clock::now();
    auto h_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h_t0).count(); };
    // ids and offsets of chunk k to the host (its kernels have finished)
    auto drain = [&](size_t k) -> int {
        JChunk& c = chunks[k];
        const uint32_t errbits = (uint32_t)c.meta[1];
        if (errbits & SPL_DEVERR_HUGE_POOL) {              // grow the pool and run the whole call again (see the end of the function)
            dc.huge_words = std::max<size_t>(dc.huge_words * 2, (size_t)c.meta[2] + 1024);
            tk->err = "scratch pool for very long pieces exhausted";
            return SPL_ERR_OOM + 1000;
        }
        if (errbits) { tk->err = "device error flags set by the encode kernels"; return SPL_ERR_CUDA; }
        c.n_tokens = c.meta[0];
        c.tok_base = total;
        total += c.n_tokens;
        if ((total + 16) * 4 > r->ids_buf.cap) {
            uint64_t est = (uint64_t)((double)total / (double)std::max<size_t>(c.b1, 1) * (double)n_bytes * 1.125) + 4096;
            est = std::min<uint64_t>(std::max<uint64_t>(est, total + 16), ids_bound(tk, n_bytes) + 16);
            PinnedBuf nb = take_pinned(tk, est * 4);
            if (!nb.p) { tk->err = "pinned host allocation failed"; return SPL_ERR_OOM; }
            if (r->ids_buf.p) { cudaStreamSynchronize(dc.s_out); memcpy(nb.p, r->ids_buf.p, (size_t)c.tok_base * 4); give_pinned(tk, r->ids_buf); }
            r->ids_buf = nb;
        }
        if ((c.doc_base + c.n_docs + 2) * 8 > r->off_buf.cap) {
            uint64_t est = (uint64_t)((double)(c.doc_base + c.n_docs) / (double)std::max<size_t>(c.b1, 1) * (double)n_bytes * 1.25) + 4096;
            PinnedBuf nb = take_pinned(tk, std::max<uint64_t>(est, c.doc_base + c.n_docs + 2) * 8);
            if (!nb.p) { tk->err = "pinned host allocation failed"; return SPL_ERR_OOM; }
            cudaStreamSynchronize(dc.s_out);
            memcpy(nb.p, r->off_buf.p, (size_t)c.doc_base * 8);
            give_pinned(tk, r->off_buf);
            r->off_buf = nb;
            res_off = (uint64_t*)nb.p;
        }
        if (c.n_tokens)
            CUDA_TRY(cudaMemcpyAsync((uint32_t*)r->ids_buf.p + c.tok_base, (uint32_t*)dc.ids.p + ids_bound(tk, c.b0), c.n_tokens * 4,
                                     cudaMemcpyDeviceToHost, dc.s_out), tk->err);
        CUDA_TRY(cudaMemcpyAsync(res_off + c.doc_base, dc.jl_out_off[c.obuf].p, (c.n_docs + 1) * 8, cudaMemcpyDeviceToHost, dc.s_out), tk->err);
        CUDA_TRY(cudaEventRecord(dc.sync_ev[2 * k + 1], dc.s_out), tk->err);
        r->stats.d2h_bytes += c.n_tokens * 4 + (c.n_docs + 1) * 8 + 16;
        return SPL_OK;
    };
    for (size_t k = 0; k < C; ++k) {
        JChunk& c = chunks[k];
        auto work = [&]() -> int {
            CUDA_TRY(cudaStreamWaitEvent(dc.stream, dc.sync_ev[2 * k], 0), tk->err);
            spl_ingest_stats st;
            int rc2 = ingest_jsonl_core(tk, dc, dc.stream, (const uint8_t*)dc.text.p + c.text_off, c.b1 - c.b0, field, flen,
                                        [&](size_t n_lines, JlOut& o) {
                                            int e = dc.jl_off.ensure((n_lines + 2) * 8, tk->err);
                                            o = JlOut{(uint8_t*)dc.jl_text.p, dc.jl_text.cap, (uint64_t*)dc.jl_off.p, n_lines + 1};
                                            return e;
                                        }, &st);
            if (rc2) return rc2;
            if (tk->trace) fprintf(stderr, "[spl trace] jsonl chunk %zu (%zu bytes, %llu docs): ingested at %.3f ms\n", k, c.b1 - c.b0, (unsigned long long)st.n_docs, h_ms());
            launches += st.n_launches;
            tot.n_lines += st.n_lines; tot.n_docs += st.n_docs; tot.n_text_bytes += st.n_text_bytes;
            tot.n_missing += st.n_missing; tot.n_bad += st.n_bad;
            c.n_docs = st.n_docs; c.doc_base = docs;
            docs += st.n_docs;
            // the stream is idle here (ingest_jsonl_core has synchronised it): the previous chunk can leave
            if (k > 0 && (rc2 = drain(k - 1))) return rc2;
            DevBuf& ob = dc.jl_out_off[c.obuf];
            if ((c.n_docs + 2) * 8 > ob.cap) {                 // may reallocate: the copy that read this buffer two chunks ago must be done
                if (k >= 2) CUDA_TRY(cudaEventSynchronize(dc.sync_ev[2 * (k - 2) + 1]), tk->err);
                if ((rc2 = ob.ensure((c.n_docs + 2) * 8, tk->err))) return rc2;
            } else if (k >= 2) {
                CUDA_TRY(cudaStreamWaitEvent(dc.stream, dc.sync_ev[2 * (k - 2) + 1], 0), tk->err);
            }
            SplWork w;
            EncodeArgs ea{(const uint8_t*)dc.jl_text.p, st.n_text_bytes, (const uint64_t*)dc.jl_off.p, 0, st.n_docs,
                          (uint32_t*)dc.ids.p + ids_bound(tk, c.b0), ids_bound(tk, c.b1 - c.b0), (uint64_t*)ob.p, c.d_meta,
                          run_tot + k, run_tot + k + 1};
            if ((rc2 = enqueue_encode(tk, dc, dc.stream, ea, with_special, nullptr, w, launches))) return rc2;
            CUDA_TRY(cudaGetLastError(), tk->err);
            if (tk->trace) fprintf(stderr, "[spl trace] jsonl chunk %zu: encode enqueued at %.3f ms\n", k, h_ms());
            return SPL_OK;
        };
        if ((rc = work())) return fail(rc);
    }
    auto finish = [&]() -> int {
        if (C) {
            CUDA_TRY(cudaStreamSynchronize(dc.stream), tk->err);
            int rc2 = drain(C - 1);
            if (rc2) return rc2;
        }
        CUDA_TRY(cudaEventRecord(dc.ev[3], dc.s_out), tk->err);
        CUDA_TRY(cudaStreamSynchronize(dc.s_out), tk->err);
        return SPL_OK;
    };
    if ((rc = finish())) return fail(rc);
    if (tk->trace) fprintf(stderr, "[spl trace] jsonl: all copied out at %.3f ms\n", h_ms());
    if (!r->ids_buf.p) {
        r->ids_buf = take_pinned(tk, 64);
        if (!r->ids_buf.p) { tk->err = "pinned host allocation failed"; return fail(SPL_ERR_OOM); }
    }
    res_off[docs] = total;
    float t = 0;
    cudaEventElapsedTime(&t, dc.ev[0], dc.ev[3]);
    r->n_docs = (size_t)docs; r->n_tokens = (size_t)total;
    r->stats.n_docs = docs; r->stats.n_bytes = n_bytes; r->stats.n_tokens = total;
    r->stats.total_ms = t; r->stats.n_devices = 1; r->stats.n_launches = launches;
    if (ingest_stats) *ingest_stats = tot;
    give_pinned(tk, meta_buf);
    *out = r;
    return SPL_OK;
}

// ---- ingestion (row N4): Parquet string column -> packed text + offsets -----------------------------------------

}  // extern "C"

namespace {

struct PqOut { uint8_t* text; size_t text_cap; uint64_t* off; size_t off_cap; };

// One batch (a run of row groups) of the planned column on `st`: column chunks host -> device as they lie in the file,
// pages -> row spans -> offsets (one synchronisation: the text size), `outputs(text_bytes, o)` names the buffers, rows ->
// packed text.  The batch's offsets start at 0.
template <class OutFn>
int ingest_parquet_batch(spl_tokenizer* tk, DevCtx& dc, cudaStream_t st, const uint8_t* file, const SplPqPlan& plan,
                         const SplPqBatch& b, OutFn outputs, uint64_t& text_bytes, int& launches, uint64_t& h2d) {
    int rc;
    const size_t n_pages = b.page1 - b.page0;
    const uint32_t n_blocks = (uint32_t)std::max<uint64_t>(1, (b.n_rows + 2047) / 2048);
    if ((rc = dc.pq_stage.ensure((size_t)b.stage_bytes + 64, tk->err))) return rc;
    if ((rc = dc.pq_scratch.ensure((size_t)b.scratch_bytes + 64, tk->err))) return rc;
    if ((rc = dc.pq_pages.ensure((n_pages + 1) * sizeof(SplPqPage), tk->err))) return rc;
    if ((rc = dc.pq_rows.ensure((size_t)(b.n_rows + 2) * 12 + 64, tk->err))) return rc;
    if ((rc = dc.pq_dict.ensure((size_t)(b.dict_entries + 2) * 12 + 64, tk->err))) return rc;
    if ((rc = dc.pq_small.ensure(256 + ((size_t)n_blocks + 2) * 8, tk->err))) return rc;
    if ((rc = dc.jl_off.ensure((size_t)(b.n_rows + 2) * 8, tk->err))) return rc;
    SplPqWork w;
    memset(&w, 0, sizeof(w));
    w.file = (const uint8_t*)dc.pq_stage.p; w.scratch = (uint8_t*)dc.pq_scratch.p;
    w.pages = (const SplPqPage*)dc.pq_pages.p;
    w.row_off = (uint64_t*)dc.pq_rows.p; w.row_len = (uint32_t*)((uint8_t*)dc.pq_rows.p + align_up((size_t)(b.n_rows + 1) * 8, 16));
    w.dict_off = (uint64_t*)dc.pq_dict.p; w.dict_len = (uint32_t*)((uint8_t*)dc.pq_dict.p + align_up((size_t)(b.dict_entries + 1) * 8, 16));
    w.n_rows = b.n_rows; w.n_blocks = n_blocks;
    w.counters = (uint32_t*)dc.pq_small.p; w.bsum = (unsigned long long*)((uint8_t*)dc.pq_small.p + 256);
    w.out_off = (uint64_t*)dc.jl_off.p;
    CUDA_TRY(cudaMemsetAsync(dc.pq_small.p, 0, 256, st), tk->err);
    for (size_t k = b.range0; k < b.range1; ++k) {
        const SplPqRange& r = plan.ranges[k];
        CUDA_TRY(cudaMemcpyAsync((uint8_t*)dc.pq_stage.p + r.stage_off, file + r.file_off, (size_t)r.len, cudaMemcpyHostToDevice, st), tk->err);
        h2d += r.len;
    }
    bool has_dict = false, has_snappy = false;
    for (size_t k = b.page0; k < b.page1; ++k) { has_dict |= plan.pages[k].kind == SPL_PQ_DICT; has_snappy |= plan.pages[k].codec == SPL_PQ_CODEC_SNAPPY; }
    if (n_pages) CUDA_TRY(cudaMemcpyAsync(dc.pq_pages.p, plan.pages.data() + b.page0, n_pages * sizeof(SplPqPage), cudaMemcpyHostToDevice, st), tk->err);
    launches += spl_launch_pq_spans(w, 0, (uint32_t)n_pages, has_dict, has_snappy, st);
    if (n_pages == 0) CUDA_TRY(cudaMemsetAsync(dc.jl_off.p, 0, 8, st), tk->err);            // no rows: offsets = {0}
    CUDA_TRY(cudaGetLastError(), tk->err);
    uint32_t h_ctr[8];
    CUDA_TRY(cudaMemcpyAsync(h_ctr, w.counters, sizeof(h_ctr), cudaMemcpyDeviceToHost, st), tk->err);
    CUDA_TRY(cudaStreamSynchronize(st), tk->err);
    if (h_ctr[SPL_PQCTR_ERR]) {
        const uint32_t e = h_ctr[SPL_PQCTR_ERR];
        tk->err = std::string("parquet: damaged page (") + ((e & SPL_PQ_ERR_SNAPPY) ? "snappy stream " : "") + ((e & SPL_PQ_ERR_LEVELS) ? "definition levels " : "") +
                  ((e & SPL_PQ_ERR_VALUES) ? "values " : "") + ((e & SPL_PQ_ERR_DICT_INDEX) ? "dictionary index " : "") + "do not decode)";
        return SPL_ERR_INVALID_ARG;
    }
    memcpy(&text_bytes, &h_ctr[SPL_PQCTR_TEXT], 8);
    PqOut o{nullptr, 0, nullptr, 0};
    if ((rc = outputs(text_bytes, o))) return rc;
    if (o.off && o.off != (uint64_t*)dc.jl_off.p)
        CUDA_TRY(cudaMemcpyAsync(o.off, dc.jl_off.p, (size_t)(b.n_rows + 1) * 8, cudaMemcpyDeviceToDevice, st), tk->err);
    if (o.text) {
        w.out_text = o.text;
        launches += spl_launch_pq_copy(w, text_bytes, st);
        CUDA_TRY(cudaGetLastError(), tk->err);
    }
    return SPL_OK;
}

int plan_parquet(spl_tokenizer* tk, const uint8_t* bytes, size_t n_bytes, const char* column, uint64_t batch_bytes, SplPqPlan& plan) {
    if (!spl_pq_plan(bytes, n_bytes, column, batch_bytes, plan)) {
        tk->err = plan.err;
        return plan.unsupported ? SPL_ERR_UNSUPPORTED : SPL_ERR_INVALID_ARG;
    }
    return SPL_OK;
}

}  // namespace

extern "C" {

int spl_ingest_parquet(spl_tokenizer* tk, int dev_index, const uint8_t* bytes, size_t n_bytes, const char* column,
                       uint8_t* d_text_out, size_t text_capacity, uint64_t* d_offsets_out, size_t offsets_capacity,
                       void* cuda_stream, spl_ingest_stats* stats) {
    if (!tk) return SPL_ERR_INVALID_ARG;
    if (dev_index < 0 || (size_t)dev_index >= tk->devs.size() || !stats || !column || !column[0] || (n_bytes && !bytes)) {
        tk->err = "invalid argument (null pointer, empty column name, or device index out of range)";
        return SPL_ERR_INVALID_ARG;
    }
    memset(stats, 0, sizeof(*stats));
    SplPqPlan plan;
    int rc = plan_parquet(tk, bytes, n_bytes, column, ~0ull, plan);          // one batch, whatever its size
    if (rc) return rc;
    if (plan.batches.size() > 1) { tk->err = "parquet: the column does not fit one device pass (spl_encode_parquet works batch by batch)"; return SPL_ERR_UNSUPPORTED; }
    DeviceGuard guard;
    DevCtx& dc = tk->devs[dev_index];
    CUDA_TRY(cudaSetDevice(dc.device), tk->err);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    stats->n_lines = stats->n_docs = plan.n_rows;
    SplPqBatch empty;
    memset(&empty, 0, sizeof(empty));
    const SplPqBatch& b = plan.batches.empty() ? empty : plan.batches[0];
    uint64_t text_bytes = 0, h2d = 0;
    int launches = 0;
    bool small = false;
    rc = ingest_parquet_batch(tk, dc, st, bytes, plan, b, [&](uint64_t tb, PqOut& o) {
        stats->n_text_bytes = tb;
        small = (d_text_out && tb > text_capacity) || (d_offsets_out && plan.n_rows + 1 > offsets_capacity);
        if (!small) o = PqOut{d_text_out, text_capacity, d_offsets_out, offsets_capacity};
        return SPL_OK;
    }, text_bytes, launches, h2d);
    if (rc) return rc;
    stats->n_launches = launches;
    CUDA_TRY(cudaStreamSynchronize(st), tk->err);
    if (small) { tk->err = "output capacity too small (the needed sizes are in the stats)"; return SPL_ERR_INVALID_ARG; }
    return SPL_OK;
}

// File bytes in host memory -> ids in host memory: spl_encode_batch for one string column of a Parquet file, one
// document per row.  Batch by batch (runs of row groups): column chunks in, pages decoded and rows packed on the
// device, encoded there, ids and offsets out.  One device (the handle's first).
int spl_encode_parquet(spl_tokenizer* tk, const uint8_t* bytes, size_t n_bytes, const char* column, uint32_t flags,
                       spl_result** out, spl_ingest_stats* ingest_stats) {
    if (!tk || !out) return SPL_ERR_INVALID_ARG;
    *out = nullptr;
    if ((n_bytes && !bytes) || !column || !column[0]) { tk->err = "invalid argument (null bytes or empty column name)"; return SPL_ERR_INVALID_ARG; }
    bool with_special;
    int rc = check_special_support(tk, flags, with_special);
    if (rc) return rc;
    SplPqPlan plan;
    const uint64_t batch_target = tk->chunk_bytes ? tk->chunk_bytes : (1ull << 30) / (is_sentencepiece(tk) ? 3 : 1);
    if ((rc = plan_parquet(tk, bytes, n_bytes, column, batch_target, plan))) return rc;

    DeviceGuard guard;
    DevCtx& dc = tk->devs[0];
    CUDA_TRY(cudaSetDevice(dc.device), tk->err);
    spl_result* r = new (std::nothrow) spl_result();
    if (!r) return SPL_ERR_OOM;
    memset(&r->stats, 0, sizeof(r->stats));
    r->owner = tk; r->n_docs = 0; r->n_tokens = 0;
    r->ids_buf = PinnedBuf{nullptr, 0};
    r->off_buf = take_pinned(tk, (size_t)(plan.n_rows + 2) * 8);
    PinnedBuf meta_buf = take_pinned(tk, 64);
    static thread_local int pq_depth = 0;
    auto fail = [&](int code) {
        cudaStreamSynchronize(dc.stream);
        cudaGetLastError();
        give_pinned(tk, r->off_buf); give_pinned(tk, r->ids_buf); give_pinned(tk, meta_buf);
        delete r;
        if (code == SPL_ERR_OOM + 1000) {                      // huge-piece scratch was too small and has been enlarged
            if (pq_depth >= 3) return (int)SPL_ERR_OOM;
            ++pq_depth;
            const int rc2 = spl_encode_parquet(tk, bytes, n_bytes, column, flags, out, ingest_stats);
            --pq_depth;
            return rc2;
        }
        return code;
    };
    if (!r->off_buf.p || !meta_buf.p) { tk->err = "pinned host allocation failed"; return fail(SPL_ERR_OOM); }
    void* d_meta = nullptr;
    if (cudaHostGetDevicePointer(&d_meta, meta_buf.p, 0) != cudaSuccess) { cudaGetLastError(); tk->err = "pinned host memory is not mapped"; return fail(SPL_ERR_CUDA); }
    volatile uint64_t* meta = (volatile uint64_t*)meta_buf.p;
    uint64_t* res_off = (uint64_t*)r->off_buf.p;
    uint64_t total = 0, docs = 0, text_total = 0, h2d = 0;
    int launches = 0;
    auto run = [&]() -> int {
        int rc2;
        if ((rc2 = dc.run_tot.ensure(16, tk->err))) return rc2;
        uint64_t* run_tot = (uint64_t*)dc.run_tot.p;
        CUDA_TRY(cudaEventRecord(dc.ev[0], dc.stream), tk->err);
        for (const SplPqBatch& b : plan.batches) {
            uint64_t tb = 0;
            rc2 = ingest_parquet_batch(tk, dc, dc.stream, bytes, plan, b, [&](uint64_t text_bytes, PqOut& o) {
                if (text_bytes > kMaxShardBytes / (is_sentencepiece(tk) ? 3 : 1)) {
                    tk->err = "parquet: a batch of row groups expands to more text than one device pass takes (dictionary-encoded column): write smaller row groups";
                    return (int)SPL_ERR_UNSUPPORTED;
                }
                int e = dc.jl_text.ensure((size_t)text_bytes + 64, tk->err);
                o = PqOut{(uint8_t*)dc.jl_text.p, dc.jl_text.cap, (uint64_t*)dc.jl_off.p, (size_t)b.n_rows + 1};
                return e;
            }, tb, launches, h2d);
            if (rc2) return rc2;
            text_total += tb;
            if (tb == 0) {                                         // rows, but no text: every document of the batch is empty
                for (uint64_t i = 0; i <= b.n_rows; ++i) res_off[docs + i] = total;
                docs += b.n_rows;
                continue;
            }
            const uint64_t cap = ids_bound(tk, tb);
            if ((rc2 = dc.ids.ensure((size_t)(cap + 16) * 4, tk->err))) return rc2;
            if ((rc2 = dc.jl_out_off[0].ensure((size_t)(b.n_rows + 2) * 8, tk->err))) return rc2;
            CUDA_TRY(cudaMemcpyAsync(run_tot, &total, 8, cudaMemcpyHostToDevice, dc.stream), tk->err);   // ids of the batches in front
            SplWork w;
            EncodeArgs ea{(const uint8_t*)dc.jl_text.p, tb, (const uint64_t*)dc.jl_off.p, 0, b.n_rows,
                          (uint32_t*)dc.ids.p, cap, (uint64_t*)dc.jl_out_off[0].p, (uint64_t*)d_meta, run_tot, run_tot + 1};
            meta[0] = meta[1] = meta[2] = 0;
            if ((rc2 = enqueue_encode(tk, dc, dc.stream, ea, with_special, nullptr, w, launches))) return rc2;
            CUDA_TRY(cudaGetLastError(), tk->err);
            CUDA_TRY(cudaStreamSynchronize(dc.stream), tk->err);
            const uint32_t errbits = (uint32_t)meta[1];
            if (errbits & SPL_DEVERR_HUGE_POOL) {
                dc.huge_words = std::max<size_t>(dc.huge_words * 2, (size_t)meta[2] + 1024);
                tk->err = "scratch pool for very long pieces exhausted";
                return SPL_ERR_OOM + 1000;
            }
            if (errbits) { tk->err = "device error flags set by the encode kernels"; return SPL_ERR_CUDA; }
            const uint64_t n_tok = meta[0];
            if ((total + n_tok + 16) * 4 > r->ids_buf.cap) {
                uint64_t est = plan.batches.size() > 1 ? (uint64_t)((double)(total + n_tok) * 1.25 * (double)plan.n_rows / (double)std::max<uint64_t>(docs + b.n_rows, 1)) : 0;
                PinnedBuf nb = take_pinned(tk, (size_t)(std::max<uint64_t>(est, total + n_tok) + 16) * 4);
                if (!nb.p) { tk->err = "pinned host allocation failed"; return SPL_ERR_OOM; }
                if (r->ids_buf.p) { memcpy(nb.p, r->ids_buf.p, (size_t)total * 4); give_pinned(tk, r->ids_buf); }
                r->ids_buf = nb;
            }
            if (n_tok) CUDA_TRY(cudaMemcpyAsync((uint32_t*)r->ids_buf.p + total, dc.ids.p, (size_t)n_tok * 4, cudaMemcpyDeviceToHost, dc.stream), tk->err);
            CUDA_TRY(cudaMemcpyAsync(res_off + docs, dc.jl_out_off[0].p, (size_t)(b.n_rows + 1) * 8, cudaMemcpyDeviceToHost, dc.stream), tk->err);
            CUDA_TRY(cudaStreamSynchronize(dc.stream), tk->err);
            r->stats.d2h_bytes += n_tok * 4 + (b.n_rows + 1) * 8 + 24;
            total += n_tok; docs += b.n_rows;
        }
        CUDA_TRY(cudaEventRecord(dc.ev[3], dc.stream), tk->err);
        CUDA_TRY(cudaStreamSynchronize(dc.stream), tk->err);
        return SPL_OK;
    };
    if ((rc = run())) return fail(rc);
    if (!r->ids_buf.p) {
        r->ids_buf = take_pinned(tk, 64);
        if (!r->ids_buf.p) { tk->err = "pinned host allocation failed"; return fail(SPL_ERR_OOM); }
    }
    res_off[docs] = total;
    float t = 0;
    cudaEventElapsedTime(&t, dc.ev[0], dc.ev[3]);
    r->n_docs = (size_t)docs; r->n_tokens = (size_t)total;
    r->stats.n_docs = docs; r->stats.n_bytes = text_total; r->stats.n_tokens = total; r->stats.h2d_bytes = h2d;
    r->stats.total_ms = t; r->stats.n_devices = 1; r->stats.n_launches = launches;
    if (ingest_stats) {
        memset(ingest_stats, 0, sizeof(*ingest_stats));
        ingest_stats->n_lines = ingest_stats->n_docs = docs; ingest_stats->n_text_bytes = text_total; ingest_stats->n_launches = launches;
    }
    give_pinned(tk, meta_buf);
    *out = r;
    return SPL_OK;
}

const uint8_t* spl_result_bytes(const spl_result* r) { return r ? (const uint8_t*)r->ids_buf.p : nullptr; }
size_t spl_result_n_bytes(const spl_result* r) { return r ? r->n_bytes : 0; }

const uint32_t* spl_result_ids(const spl_result* r) { return r ? (const uint32_t*)r->ids_buf.p : nullptr; }
const uint64_t* spl_result_offsets(const spl_result* r) { return r ? (const uint64_t*)r->off_buf.p : nullptr; }
size_t spl_result_n_docs(const spl_result* r) { return r ? r->n_docs : 0; }
size_t spl_result_n_tokens(const spl_result* r) { return r ? r->n_tokens : 0; }
void spl_result_stats(const spl_result* r, spl_stats* out) { if (r && out) *out = r->stats; }
void spl_result_free(spl_result* r) {
    if (!r) return;
    give_pinned(r->owner, r->ids_buf);
    give_pinned(r->owner, r->off_buf);
    delete r;
}

}  // extern "C"
