// Parquet ingestion: the pages of ONE string column -> packed text + offsets.
// Page decoding is __host__ __device__ so that the exact device logic runs on the CPU against pyarrow
// (tests/test_parquet_host.py).
//
// Like the JSON Lines path this replaces nothing inside the reference: it is the loop a splintr user runs in front of
// `encode_batch`,
//     texts = pyarrow.parquet.read_table(path, columns=[column])[column].to_pylist()
// followed by the packing of `texts`.  The host reads the footer and the page headers (Thrift compact protocol,
// spl_parquet_meta.cpp: a few hundred bytes per page); the page BODIES go to the device as they lie in the file and
// are decompressed and decoded there.
//
// Supported (what pyarrow / parquet-mr / arrow-rs write for a text column unless told otherwise):
//   physical type BYTE_ARRAY, not repeated (no lists), any number of OPTIONAL ancestors (a null at any level is an
//   EMPTY document: one document per row, always); codecs UNCOMPRESSED and SNAPPY; data pages V1 and V2; encodings
//   PLAIN and PLAIN_DICTIONARY / RLE_DICTIONARY (dictionary page PLAIN), also mixed within a column chunk (the writer's
//   fall-back when a dictionary grows too large); definition levels RLE / bit-packed hybrid.
// Refused with SPL_ERR_UNSUPPORTED and a message that says what to rewrite: other codecs (gzip, zstd, lz4, brotli),
// DELTA_* encodings, repeated columns, encrypted files.
//
// Formats (Apache Parquet format specification, parquet.thrift + Encodings.md; restated, no code taken):
//   PLAIN BYTE_ARRAY       u32 little-endian length + bytes, back to back
//   RLE / bit-packed hybrid  runs: varint header h; h & 1 == 0: RLE run of h >> 1 values, the value in ceil(bw / 8)
//                          bytes; h & 1 == 1: (h >> 1) groups of 8 values, bw bits each, packed LSB first
//   dictionary indices     one byte bit width, then hybrid runs
//   data page V1           [u32 length + hybrid definition levels, if the column has any] values; all of it compressed
//   data page V2           repetition levels, definition levels (hybrid, lengths in the header), values; only the
//                          values are compressed
//   snappy (raw)           varint uncompressed length, then elements: tag & 3 == 0 literal, 1 / 2 / 3 copy with 1- / 2-
//                          / 4-byte offset
#pragma once
#include "spl_common.h"

enum : uint8_t { SPL_PQ_DATA_V1 = 0, SPL_PQ_DICT = 2, SPL_PQ_DATA_V2 = 3 };        // PageType of parquet.thrift
enum : uint8_t { SPL_PQ_CODEC_NONE = 0, SPL_PQ_CODEC_SNAPPY = 1 };                 // CompressionCodec
enum : uint8_t { SPL_PQ_ENC_PLAIN = 0, SPL_PQ_ENC_DICT = 1 };                      // (ours) how a page's values are stored
enum : uint32_t {                                                                 // error bits a page can raise
    SPL_PQ_ERR_SNAPPY = 1u, SPL_PQ_ERR_LEVELS = 2u, SPL_PQ_ERR_VALUES = 4u, SPL_PQ_ERR_DICT_INDEX = 8u
};

#define SPL_PQ_IN_SCRATCH (1ull << 63)               // span source: decompressed bytes (else: the staged file bytes)

// One page of the column (built on the host from the page header, consumed by one lane group on the device).
struct SplPqPage {
    uint64_t src;            // its body (behind the header) in the staged file bytes
    uint64_t scratch;        // where its decompressed bytes go (compressed pages)
    uint64_t first_row;      // data page: row of its first value (rows of the batch)
    uint64_t dict_base;      // first entry of the column chunk's dictionary in the entry arrays
    uint32_t comp_size, uncomp_size;
    uint32_t num_values;     // data page: rows (nulls included); dictionary page: entries
    uint32_t dict_count;     // entries of the chunk's dictionary (data pages check their indices against it)
    uint32_t def_bytes, rep_bytes;   // V2: bytes of the level sections in front of the values
    uint8_t kind, codec, encoding, max_def;
    uint8_t v2_compressed, pad_[3];
};

// Pages are decoded in two launches: dictionary pages and the data pages that need no dictionary (PLAIN) first, then
// the dictionary-encoded data pages (they read the dictionary's spans).
SPL_HD bool spl_pq_first_pass(const SplPqPage& pg) { return pg.kind == SPL_PQ_DICT || pg.encoding == SPL_PQ_ENC_PLAIN; }

// A span of source bytes: what a row (or a dictionary entry) is.
struct SplPqSpans { uint64_t* off; uint32_t* len; };

// The lanes that decode one page together.  Everything but the byte copies is computed by every lane alike.
// win / inbuf: fast memory for the snappy decoder (shared memory on the device), or null.
#define SPL_SNAPPY_WIN 65536u            // bytes of output kept at hand: every offset the snappy format's 64 KiB blocks produce
#define SPL_SNAPPY_INBUF 8192u           // two 4 KiB slots of input
struct SplPqOneLane {
    static constexpr uint32_t NL = 1;
    uint32_t lane = 0;
    uint8_t* win = nullptr; uint8_t* inbuf = nullptr;
    SPL_HD void sync() const {}
};

struct alignas(16) SplU128 { uint64_t a, b; };

// ---- snappy --------------------------------------------------------------------------------------------------
// src[0, n) -> dst[0, cap); true iff the stream is well formed and yields exactly cap bytes.  A copy reads bytes an
// earlier element wrote, possibly by another lane: sync() orders them.
template <class G>
SPL_HD bool spl_snappy_decode(const G& g, const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t cap) {
    uint32_t ip = 0, total = 0, shift = 0;
    for (;;) {
        if (ip >= n || shift > 28u) return false;
        const uint32_t b = src[ip++];
        total |= (b & 0x7Fu) << shift;
        if (!(b & 0x80u)) break;
        shift += 7u;
    }
    if (total != cap) return false;
    uint32_t op = 0;
    while (ip < n) {
        const uint32_t tag = src[ip++];
        if ((tag & 3u) == 0u) {                                  // literal
            uint32_t l = tag >> 2;
            if (l >= 60u) {
                const uint32_t nb = l - 59u;                     // 1..4 length bytes
                if (nb > n - ip) return false;
                l = 0;
                for (uint32_t k = 0; k < nb; ++k) l |= (uint32_t)src[ip + k] << (8u * k);
                ip += nb;
                if (l == 0xFFFFFFFFu) return false;
            }
            l += 1u;
            if (l > n - ip || l > cap - op) return false;
            for (uint32_t i = g.lane; i < l; i += G::NL) dst[op + i] = src[ip + i];
            ip += l; op += l;
        } else {                                                 // copy of l bytes from `off` bytes back
            uint32_t l, off;
            if ((tag & 3u) == 1u) {
                if (ip >= n) return false;
                l = ((tag >> 2) & 7u) + 4u; off = ((tag >> 5) << 8) | src[ip]; ip += 1u;
            } else if ((tag & 3u) == 2u) {
                if (2u > n - ip) return false;
                l = (tag >> 2) + 1u; off = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8); ip += 2u;
            } else {
                if (4u > n - ip) return false;
                l = (tag >> 2) + 1u;
                off = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16) | ((uint32_t)src[ip + 3] << 24);
                ip += 4u;
            }
            if (off == 0u || off > op || l > cap - op) return false;
            const uint8_t* from = dst + (op - off);              // all of it written before this element
            for (uint32_t i = g.lane; i < l; i += G::NL) dst[op + i] = from[off >= l ? i : i % off];
            op += l;
        }
        g.sync();
    }
    return op == cap;
}

// The same stream decoded out of fast memory: input comes in 4 KiB slots (aligned 16-byte loads, two slots = a
// look-ahead of 4 KiB), the last 64 KiB of output live in a ring (g.win) that serves the copies and is written to dst
// in 16-byte units -- per element the lanes touch no global memory.  (The plain version above costs 4-5 dependent global
// round trips per element: 130 ms for a 1 MiB page of text.)  src must be readable from its 16-byte-aligned start to its
// 16-byte-aligned end, dst must be 16-byte aligned (both hold for the staged file bytes and the scratch buffer).
template <class G>
SPL_HD bool spl_snappy_decode_staged(const G& g, const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t cap) {
    uint8_t* const win = g.win; uint8_t* const inbuf = g.inbuf;
    const uint32_t lead = (uint32_t)((uintptr_t)src & 15u);
    const uint8_t* a0 = src - lead;                            // input positions are relative to a0
    const uint32_t in_end = lead + n, in_end16 = (in_end + 15u) & ~15u;
    uint32_t loaded = 0;                                       // 4 KiB chunks [0, loaded) have been brought in
    auto need = [&](uint32_t pos) {                            // the chunk of pos and the next one are present
        const uint32_t want = (pos >> 12) + 2u;
        if (loaded >= want) return;
        while (loaded < want) {
            const uint32_t c0 = loaded << 12;
            if (c0 < in_end16) {
                const uint32_t units = (in_end16 - c0 < 4096u ? in_end16 - c0 : 4096u) >> 4;
                const SplU128* s4 = reinterpret_cast<const SplU128*>(a0 + c0);
                SplU128* d4 = reinterpret_cast<SplU128*>(inbuf + ((loaded & 1u) << 12));
                for (uint32_t k = g.lane; k < units; k += G::NL) d4[k] = s4[k];
            }
            ++loaded;
        }
        g.sync();
    };
    auto in = [&](uint32_t pos) -> uint32_t { return inbuf[pos & (SPL_SNAPPY_INBUF - 1u)]; };
    uint32_t ip = lead, op = 0, flushed = 0;                   // flushed: multiple of 16, output [0, flushed) is in dst
    auto flush_to = [&](uint32_t limit16) {
        const uint32_t units = (limit16 - flushed) >> 4;
        SplU128* d4 = reinterpret_cast<SplU128*>(dst + flushed);
        for (uint32_t k = g.lane; k < units; k += G::NL)
            d4[k] = *reinterpret_cast<const SplU128*>(win + ((flushed + (k << 4)) & (SPL_SNAPPY_WIN - 1u)));
        flushed = limit16;
    };
    auto flush_all = [&]() {                                   // (the odd bytes at the end are written again later)
        flush_to(op & ~15u);
        const uint32_t t = op & 15u;
        for (uint32_t k = g.lane; k < t; k += G::NL) dst[flushed + k] = win[(flushed + k) & (SPL_SNAPPY_WIN - 1u)];
    };
    need(ip);
    uint32_t total = 0, shift = 0;
    for (;;) {
        if (ip >= in_end || shift > 28u) return false;
        const uint32_t b = in(ip++);
        total |= (b & 0x7Fu) << shift;
        if (!(b & 0x80u)) break;
        shift += 7u;
    }
    if (total != cap) return false;
    while (ip < in_end) {
        need(ip);
        const uint32_t tag = in(ip++);
        if ((tag & 3u) == 0u) {
            uint32_t l = tag >> 2;
            if (l >= 60u) {
                const uint32_t nb = l - 59u;
                if (nb > in_end - ip) return false;
                l = 0;
                for (uint32_t k = 0; k < nb; ++k) l |= in(ip + k) << (8u * k);
                ip += nb;
                if (l == 0xFFFFFFFFu) return false;
            }
            l += 1u;
            if (l > in_end - ip || l > cap - op) return false;
            while (l) {
                need(ip);
                const uint32_t seg = l < 4096u ? l : 4096u;
                for (uint32_t i = g.lane; i < seg; i += G::NL) win[(op + i) & (SPL_SNAPPY_WIN - 1u)] = (uint8_t)in(ip + i);
                ip += seg; op += seg; l -= seg;
                g.sync();
                if (op - flushed >= 4096u) flush_to(op & ~15u);
            }
        } else {
            uint32_t l, off;
            if ((tag & 3u) == 1u) {
                if (ip >= in_end) return false;
                l = ((tag >> 2) & 7u) + 4u; off = ((tag >> 5) << 8) | in(ip); ip += 1u;
            } else if ((tag & 3u) == 2u) {
                if (2u > in_end - ip) return false;
                l = (tag >> 2) + 1u; off = in(ip) | (in(ip + 1) << 8); ip += 2u;
            } else {
                if (4u > in_end - ip) return false;
                l = (tag >> 2) + 1u; off = in(ip) | (in(ip + 1) << 8) | (in(ip + 2) << 16) | (in(ip + 3) << 24); ip += 4u;
            }
            if (off == 0u || off > op || l > cap - op) return false;
            if (off <= SPL_SNAPPY_WIN - 64u) {
                const uint32_t from = op - off;
                for (uint32_t i = g.lane; i < l; i += G::NL)
                    win[(op + i) & (SPL_SNAPPY_WIN - 1u)] = win[(from + (off >= l ? i : i % off)) & (SPL_SNAPPY_WIN - 1u)];
            } else {                                             // beyond the ring: through dst
                flush_all();
                g.sync();
                const uint8_t* from = dst + (op - off);
                for (uint32_t i = g.lane; i < l; i += G::NL) win[(op + i) & (SPL_SNAPPY_WIN - 1u)] = from[i];       // (off > l here)
            }
            op += l;
            g.sync();
            if (op - flushed >= 4096u) flush_to(op & ~15u);
        }
    }
    flush_all();
    g.sync();
    return op == cap;
}

// The warp version: elements are short on text (cfg2 through pyarrow's snappy: 485 000 elements per 2 MB page, 4.1
// bytes each), and one element at a time leaves a warp waiting on its own instruction latencies (~530 cycles per
// element, with or without shared memory).  Here the 32 lanes parse the next 32 INPUT bytes speculatively -- lane i
// assumes an element starts at byte i --, the real chain of element starts is found by pointer doubling (5 shuffle
// rounds), output positions by a warp scan, and the batch (~11 elements) is copied at once: short literals and copies
// whose source lies in front of the batch by their own lanes, the few copies that read what the batch itself produces
// one after the other, a long literal (always the batch's last element) by the whole warp.
// G: 32 lanes with shfl(value, source lane) / ballot(predicate) / sync(), all called by every lane alike.
template <class G>
SPL_HD bool spl_snappy_decode_warp(const G& g, const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t cap) {
    constexpr uint32_t M = SPL_SNAPPY_WIN - 1u;
    constexpr uint32_t NEAR = SPL_SNAPPY_WIN - 4096u;          // offsets the ring serves while a batch (< 2.1 KiB without its long literal) is written
    constexpr uint32_t SHORT = 32u;                            // literals up to this long are copied by their own lane
    uint8_t* const win = g.win; uint8_t* const inbuf = g.inbuf;
    const uint32_t lane = g.lane;
    const uint32_t lead = (uint32_t)((uintptr_t)src & 15u);
    const uint8_t* a0 = src - lead;
    const uint32_t in_end = lead + n, in_end16 = (in_end + 15u) & ~15u;
    uint32_t loaded = 0;
    auto need = [&](uint32_t pos) {
        const uint32_t want = (pos >> 12) + 2u;
        if (loaded >= want) return;
        while (loaded < want) {
            const uint32_t c0 = loaded << 12;
            if (c0 < in_end16) {
                const uint32_t units = (in_end16 - c0 < 4096u ? in_end16 - c0 : 4096u) >> 4;
                const SplU128* s4 = reinterpret_cast<const SplU128*>(a0 + c0);
                SplU128* d4 = reinterpret_cast<SplU128*>(inbuf + ((loaded & 1u) << 12));
                for (uint32_t k = lane; k < units; k += 32u) d4[k] = s4[k];
            }
            ++loaded;
        }
        g.sync();
    };
    auto in = [&](uint32_t pos) -> uint32_t { return inbuf[pos & (SPL_SNAPPY_INBUF - 1u)]; };
    uint32_t ip = lead, op = 0, flushed = 0;
    auto flush_to = [&](uint32_t limit16) {
        const uint32_t units = (limit16 - flushed) >> 4;
        SplU128* d4 = reinterpret_cast<SplU128*>(dst + flushed);
        for (uint32_t k = lane; k < units; k += 32u) d4[k] = *reinterpret_cast<const SplU128*>(win + ((flushed + (k << 4)) & M));
        flushed = limit16;
    };
    auto flush_all = [&](uint32_t limit) {
        flush_to(limit & ~15u);
        if (lane < (limit & 15u)) dst[flushed + lane] = win[(flushed + lane) & M];
    };
    need(ip);
    uint32_t total_len = 0, shift = 0;
    for (;;) {
        if (ip >= in_end || shift > 28u) return false;
        const uint32_t b = in(ip++);
        total_len |= (b & 0x7Fu) << shift;
        if (!(b & 0x80u)) break;
        shift += 7u;
    }
    if (total_len != cap) return false;
    while (ip < in_end) {
        need(ip);
        // ---- 1. what an element starting at byte ip + lane would be (no branches: the lanes' tags differ) ---------
        const uint32_t s = ip + lane;
        const bool inside = s < in_end;
        uint32_t hs = 1, pl = 0, l = 0, off = 0;
        bool lit = false, bad = false;
        {
            // the eight bytes from s on: two aligned words of the input slots
            const uint32_t* in32 = reinterpret_cast<const uint32_t*>(inbuf);
            const uint32_t wi = (s & (SPL_SNAPPY_INBUF - 1u)) >> 2, sh = (s & 3u) * 8u;
            const uint32_t w0 = in32[wi], w1 = in32[(wi + 1u) & (SPL_SNAPPY_INBUF / 4u - 1u)], w2 = in32[(wi + 2u) & (SPL_SNAPPY_INBUF / 4u - 1u)];
            const uint32_t lo = sh ? (w0 >> sh) | (w1 << (32u - sh)) : w0, hi = sh ? (w1 >> sh) | (w2 << (32u - sh)) : w1;
            const uint32_t tag = lo & 0xFFu, t = tag & 3u, n6 = tag >> 2;
            const uint32_t w = (lo >> 8) | (hi << 24);                   // the four bytes behind the tag
            const uint32_t avail = inside ? in_end - s - 1u : 0u;
            lit = t == 0u;
            const uint32_t nb = lit ? (n6 >= 60u ? n6 - 59u : 0u) : (t == 1u ? 1u : (t == 2u ? 2u : 4u));   // bytes behind the tag that belong to the header
            hs = 1u + nb;
            const uint32_t field = nb >= 4u ? w : (w & ((1u << (8u * nb)) - 1u));
            if (lit) {
                l = nb ? field : n6;
                bad = nb > avail || l == 0xFFFFFFFFu;
                l += 1u;
                pl = l;
                bad = bad || l > in_end - s - hs;                        // (evaluated only where it matters: mine && bad)
            } else {
                l = t == 1u ? ((n6 & 7u) + 4u) : n6 + 1u;
                off = t == 1u ? (((tag >> 5) << 8) | field) : field;
                bad = nb > avail;
            }
            if (!inside) { bad = true; lit = false; pl = 0; l = 0; }
        }
        // ---- 2. the chain of element starts from lane 0 (pointer doubling) ----------------------------------------
        uint32_t far = 32u;                                      // where one step from here lands (32: beyond the window)
        if (inside && !bad) { const uint32_t st = pl > 64u ? 64u : hs + pl; far = lane + st > 32u ? 32u : lane + st; }
        uint32_t reach = 1u << lane;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const uint32_t r2 = g.shfl(reach, far & 31u), f2 = g.shfl(far, far & 31u);
            if (far < 32u) { reach |= r2; far = f2; }
        }
        const uint32_t starts = g.shfl(reach, 0u);
        const bool mine = inside && ((starts >> lane) & 1u);
        if (g.ballot(mine && bad)) return false;
        // ---- 3. output positions (saturating scan: a literal can claim up to 4 GiB) --------------------------------
        const uint32_t x = mine ? l : 0u;
        uint32_t incl = x;
#pragma unroll
        for (uint32_t o = 1; o < 32u; o <<= 1) {
            const uint32_t t = g.shfl(incl, (lane - o) & 31u);
            if (lane >= o) incl = incl + t < incl ? 0xFFFFFFFFu : incl + t;
        }
        const uint32_t total = g.shfl(incl, 31u);
        if (total > cap - op) return false;
        const uint32_t oe = op + (incl - x);                     // where my element's output starts
        const bool cp = mine && !lit;
        if (g.ballot(cp && (off == 0u || off > oe))) return false;
        // ---- 4. copies ------------------------------------------------------------------------------------------
        if (mine && lit && pl <= SHORT) {
            uint32_t ps = (s + hs) & (SPL_SNAPPY_INBUF - 1u), pd = oe & M;
            for (uint32_t i = 0; i < pl; ++i) { win[pd] = inbuf[ps]; ps = (ps + 1u) & (SPL_SNAPPY_INBUF - 1u); pd = (pd + 1u) & M; }
        }
        const uint32_t srcb = oe - off, srce = srcb + (l < off ? l : off);
        const bool indep = cp && srce <= op && off <= NEAR;      // reads nothing this batch writes
        g.sync();
        if (indep) {
            uint32_t pd = oe & M;
            if (off >= l) {
                uint32_t ps = srcb & M;
                for (uint32_t i = 0; i < l; ++i) { win[pd] = win[ps]; ps = (ps + 1u) & M; pd = (pd + 1u) & M; }
            } else {                                             // the pattern of `off` bytes repeats
                for (uint32_t i = 0, j = 0; i < l; ++i) { win[pd] = win[(srcb + j) & M]; pd = (pd + 1u) & M; if (++j == off) j = 0; }
            }
        }
        g.sync();
        uint32_t deps = g.ballot(cp && !indep);
        while (deps) {                                           // in stream order, each by the whole warp
#if defined(__CUDA_ARCH__)
            const uint32_t e = __ffs(deps) - 1u;
#else
            const uint32_t e = (uint32_t)__builtin_ctz(deps);
#endif
            deps &= deps - 1u;
            const uint32_t eo = g.shfl(oe, e), eoff = g.shfl(off, e), el = g.shfl(l, e);
            if (eoff <= NEAR) {
                for (uint32_t i = lane; i < el; i += 32u) win[(eo + i) & M] = win[(eo - eoff + (eoff >= el ? i : i % eoff)) & M];
            } else {                                             // further back than the ring reaches: through dst
                flush_all(eo);
                g.sync();
                const uint8_t* from = dst + (eo - eoff);
                for (uint32_t i = lane; i < el; i += 32u) win[(eo + i) & M] = from[i];              // (eoff > el here)
            }
            g.sync();
        }
        // the last element of the chain: where the next batch starts; a long literal is copied by the whole warp
#if defined(__CUDA_ARCH__)
        const uint32_t last = 31u - (uint32_t)__clz((int)(starts & g.ballot(inside)));
#else
        const uint32_t last = 31u - (uint32_t)__builtin_clz(starts & g.ballot(inside));
#endif
        const uint32_t l_pos = g.shfl(s + hs, last), l_pl = g.shfl(pl, last), l_out = g.shfl(oe, last);
        if (l_pl > SHORT) {
            uint32_t pos = l_pos, o2 = l_out, left = l_pl;
            while (left) {
                need(pos);
                const uint32_t seg = left < 4096u ? left : 4096u;
                for (uint32_t i = lane; i < seg; i += 32u) win[(o2 + i) & M] = (uint8_t)in(pos + i);
                pos += seg; o2 += seg; left -= seg;
                g.sync();
                if (o2 - flushed >= 4096u) flush_to(o2 & ~15u);
            }
        }
        ip = l_pos + l_pl;
        op += total;
        if (op - flushed >= 4096u) flush_to(op & ~15u);
        g.sync();
    }
    flush_all(op);
    g.sync();
    return op == cap;
}

// ---- RLE / bit-packed hybrid -----------------------------------------------------------------------------------
struct SplPqHybrid {
    const uint8_t* p; const uint8_t* end;
    uint32_t bw;                 // bits per value, 0..32
    uint32_t left;               // values left in the current run
    uint32_t rle_val;            // RLE run: its value
    const uint8_t* run;          // bit-packed run: its first byte
    const uint8_t* run_end;      //                 and the end of its bytes
    uint32_t idx;                // bit-packed run: values taken
    bool packed;

    SPL_HD void init(const uint8_t* b, const uint8_t* e, uint32_t width) { p = b; end = e; bw = width; left = 0; rle_val = 0; run = b; run_end = b; idx = 0; packed = false; }

    SPL_HD bool next(uint32_t& v) {
        while (left == 0u) {
            uint32_t h = 0, shift = 0;
            for (;;) {
                if (p >= end || shift > 28u) return false;
                const uint32_t b = *p++;
                h |= (b & 0x7Fu) << shift;
                if (!(b & 0x80u)) break;
                shift += 7u;
            }
            if (h & 1u) {
                const uint32_t groups = h >> 1;
                const uint64_t bytes = (uint64_t)groups * bw, avail = (uint64_t)(end - p);
                const uint64_t take = bytes < avail ? bytes : avail;     // (a writer may cut the padding of the last group)
                packed = true; run = p; run_end = p + take; idx = 0;
                left = groups > 0x1FFFFFFFu ? 0xFFFFFFFFu : groups * 8u;
                p += take;
            } else {
                const uint32_t nb = (bw + 7u) >> 3;
                if (nb > (uint32_t)(end - p)) return false;
                rle_val = 0;
                for (uint32_t k = 0; k < nb; ++k) rle_val |= (uint32_t)p[k] << (8u * k);
                p += nb;
                packed = false; left = h >> 1;
            }
        }
        --left;
        if (!packed) { v = rle_val; return true; }
        const uint64_t bit = (uint64_t)idx * bw;
        ++idx;
        const uint8_t* q = run + (bit >> 3);
        uint64_t wv = 0;
        if (bw && q + ((((uint32_t)bit & 7u) + bw + 7u) >> 3) > run_end) return false;  // the value lies beyond the run's bytes
        for (uint32_t k = 0; k < 5u; ++k) if (q + k < run_end) wv |= (uint64_t)q[k] << (8u * k);
        v = (uint32_t)((wv >> (bit & 7u)) & (bw >= 32u ? 0xFFFFFFFFull : ((1ull << bw) - 1ull)));
        return true;
    }
};

SPL_HD uint32_t spl_pq_bit_width(uint32_t max_value) {
    uint32_t w = 0;
    while (max_value) { ++w; max_value >>= 1; }
    return w;
}

SPL_HD uint32_t spl_pq_le32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// ---- one page ----------------------------------------------------------------------------------------------------
// file = the staged file bytes, scratch = the decompression buffer.  Data pages fill rows[first_row ..], dictionary
// pages fill dict[dict_base ..] (a data page of a dictionary-encoded chunk reads them: dictionary pages are decoded
// in an earlier launch).  Returns the error bits (0: fine); on an error the page's remaining rows are empty.
template <class G>
SPL_HD uint32_t spl_pq_decode_page(const G& g, const SplPqPage& pg, const uint8_t* file, uint8_t* scratch,
                                   const SplPqSpans& rows, const SplPqSpans& dict) {
    uint32_t err = 0;
    const bool is_dict = pg.kind == SPL_PQ_DICT, v2 = pg.kind == SPL_PQ_DATA_V2;
    const uint32_t lvl = v2 ? pg.def_bytes + pg.rep_bytes : 0u;          // V2: levels lie uncompressed in front
    const uint8_t* body = file + pg.src + lvl;
    uint64_t base = pg.src + lvl;                                        // span offset of body[0]
    uint32_t body_len = pg.uncomp_size - lvl;
    bool ok = lvl <= pg.uncomp_size && lvl <= pg.comp_size;
    if (ok && pg.codec == SPL_PQ_CODEC_SNAPPY && (!v2 || pg.v2_compressed)) {
        uint8_t* d = scratch + pg.scratch;
        if (body_len == 0u) ok = true;                                   // (writers emit no stream for an empty body)
        else if (!g.win) ok = spl_snappy_decode(g, body, pg.comp_size - lvl, d, body_len);
        else if constexpr (G::NL == 32u) ok = spl_snappy_decode_warp(g, body, pg.comp_size - lvl, d, body_len);
        else ok = spl_snappy_decode_staged(g, body, pg.comp_size - lvl, d, body_len);
        body = d; base = SPL_PQ_IN_SCRATCH | pg.scratch;
        if (!ok) err |= SPL_PQ_ERR_SNAPPY;
    } else if (ok && pg.comp_size != pg.uncomp_size) {
        ok = false; err |= SPL_PQ_ERR_VALUES;
    }
    const uint8_t* vend = body + body_len;
    const uint8_t* vp = body;
    SplPqHybrid defs;
    bool has_defs = false;
    if (ok && !is_dict && pg.max_def) {
        const uint32_t bw = spl_pq_bit_width(pg.max_def);
        if (v2) {
            const uint8_t* l0 = file + pg.src + pg.rep_bytes;
            defs.init(l0, l0 + pg.def_bytes, bw);
            has_defs = pg.def_bytes != 0u;                               // (no section: every value is present)
        } else {
            if (body_len < 4u) { ok = false; err |= SPL_PQ_ERR_LEVELS; }
            else {
                const uint32_t L = spl_pq_le32(body);
                if (L > body_len - 4u) { ok = false; err |= SPL_PQ_ERR_LEVELS; }
                else { defs.init(body + 4, body + 4 + L, bw); vp = body + 4 + L; has_defs = true; }
            }
        }
    }
    SplPqHybrid idx;
    if (ok && !is_dict && pg.encoding == SPL_PQ_ENC_DICT) {
        if (vp >= vend) {
            idx.init(vend, vend, 0);                                     // a page of nulls only may have no index section
        } else {
            const uint32_t bw = *vp++;
            if (bw > 32u) { ok = false; err |= SPL_PQ_ERR_VALUES; }
            else idx.init(vp, vend, bw);
        }
    }
    const SplPqSpans& out = is_dict ? dict : rows;
    const uint64_t o0 = is_dict ? pg.dict_base : pg.first_row;
    for (uint32_t r = 0; r < pg.num_values; ++r) {
        uint64_t so = 0; uint32_t sl = 0;
        if (ok) {
            bool present = true;
            if (has_defs) {
                uint32_t d;
                if (!defs.next(d)) { ok = false; err |= SPL_PQ_ERR_LEVELS; present = false; }
                else present = d == pg.max_def;
            }
            if (ok && present) {
                if (is_dict || pg.encoding == SPL_PQ_ENC_PLAIN) {
                    if ((uint64_t)(vend - vp) < 4u) { ok = false; err |= SPL_PQ_ERR_VALUES; }
                    else {
                        const uint32_t l = spl_pq_le32(vp);
                        vp += 4;
                        if (l > (uint64_t)(vend - vp)) { ok = false; err |= SPL_PQ_ERR_VALUES; }
                        else { so = base + (uint64_t)(vp - body); sl = l; vp += l; }
                    }
                } else {
                    uint32_t k;
                    if (!idx.next(k)) { ok = false; err |= SPL_PQ_ERR_VALUES; }
                    else if (k >= pg.dict_count) { ok = false; err |= SPL_PQ_ERR_DICT_INDEX; }
                    else { so = dict.off[pg.dict_base + k]; sl = dict.len[pg.dict_base + k]; }
                }
            }
        }
        if (g.lane == 0u) { out.off[o0 + r] = so; out.len[o0 + r] = sl; }
    }
    return err;
}
