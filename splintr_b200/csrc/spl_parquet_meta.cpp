// Footer and page headers of a Parquet file (Thrift compact protocol) -> the page descriptors spl_parquet.h decodes.
// Written from the format specification (parquet.thrift, thrift compact protocol spec); host only; every read is
// bounds-checked -- the file is untrusted input.
#include "spl_parquet_meta.h"

#include <cstring>

namespace {

enum : uint8_t { T_STOP = 0, T_TRUE = 1, T_FALSE = 2, T_BYTE = 3, T_I16 = 4, T_I32 = 5, T_I64 = 6, T_DOUBLE = 7,
                 T_BINARY = 8, T_LIST = 9, T_SET = 10, T_MAP = 11, T_STRUCT = 12 };

struct TReader {
    const uint8_t* p; const uint8_t* end;
    bool ok = true;

    TReader(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
    uint8_t byte() { if (p >= end) { ok = false; return 0; } return *p++; }
    uint64_t varint() {
        uint64_t v = 0;
        for (uint32_t shift = 0; shift < 64; shift += 7) {
            const uint8_t b = byte();
            if (!ok) return 0;
            v |= (uint64_t)(b & 0x7F) << shift;
            if (!(b & 0x80)) return v;
        }
        ok = false;
        return 0;
    }
    int64_t zigzag() { const uint64_t u = varint(); return (int64_t)(u >> 1) ^ -(int64_t)(u & 1); }
    // next field of the current struct: false at its end.  `last` carries the previous field id.
    bool field(int& id, uint8_t& type, int& last) {
        const uint8_t h = byte();
        if (!ok || h == T_STOP) return false;
        type = h & 0x0F;
        const int delta = h >> 4;
        id = delta ? last + delta : (int)zigzag();
        last = id;
        return ok;
    }
    bool list(uint32_t& size, uint8_t& type) {
        const uint8_t h = byte();
        type = h & 0x0F;
        size = h >> 4;
        if (size == 15) { const uint64_t s = varint(); if (s > 0x7FFFFFFFull) ok = false; size = (uint32_t)s; }
        // every element takes at least... nothing (empty structs are one byte): a size beyond the bytes left is malformed
        if ((uint64_t)size > (uint64_t)(end - p) + 1) ok = false;
        return ok;
    }
    bool binary(const uint8_t*& b, uint32_t& n) {
        const uint64_t l = varint();
        if (!ok || l > (uint64_t)(end - p)) { ok = false; return false; }
        b = p; n = (uint32_t)l; p += l;
        return true;
    }
    std::string str() { const uint8_t* b; uint32_t n; return binary(b, n) ? std::string((const char*)b, n) : std::string(); }
    void skip(uint8_t type, bool in_container, int depth = 0) {
        if (!ok || depth > 48) { ok = false; return; }
        switch (type) {
            case T_TRUE: case T_FALSE: if (in_container) byte(); break;
            case T_BYTE: byte(); break;
            case T_I16: case T_I32: case T_I64: varint(); break;
            case T_DOUBLE: if ((size_t)(end - p) < 8) ok = false; else p += 8; break;
            case T_BINARY: { const uint8_t* b; uint32_t n; binary(b, n); break; }
            case T_LIST: case T_SET: {
                uint32_t n; uint8_t et;
                if (!list(n, et)) return;
                for (uint32_t i = 0; i < n && ok; ++i) skip(et, true, depth + 1);
                break;
            }
            case T_MAP: {
                const uint64_t n = varint();
                if (!ok || n == 0) return;
                if (n > (uint64_t)(end - p)) { ok = false; return; }
                const uint8_t kv = byte();
                for (uint64_t i = 0; i < n && ok; ++i) { skip(kv >> 4, true, depth + 1); skip(kv & 0x0F, true, depth + 1); }
                break;
            }
            case T_STRUCT: {
                int id, last = 0; uint8_t t;
                while (field(id, t, last)) skip(t, false, depth + 1);
                break;
            }
            default: ok = false;
        }
    }
};

struct SchemaEl { int type = -1, repetition = 0, num_children = 0; std::string name; };
struct ChunkMeta {
    int type = -1, codec = 0;
    int64_t num_values = 0, total_compressed = 0, total_uncompressed = 0, data_page_offset = 0, dict_page_offset = 0;
    bool has_dict_offset = false, external = false, encrypted = false, has_meta = false;
};
struct RowGroupMeta { std::vector<ChunkMeta> cols; int64_t num_rows = 0; };

void parse_schema_el(TReader& r, SchemaEl& s) {
    int id, last = 0; uint8_t t;
    while (r.field(id, t, last)) {
        if (id == 1 && t == T_I32) s.type = (int)r.zigzag();
        else if (id == 3 && t == T_I32) s.repetition = (int)r.zigzag();
        else if (id == 4 && t == T_BINARY) s.name = r.str();
        else if (id == 5 && t == T_I32) s.num_children = (int)r.zigzag();
        else r.skip(t, false);
    }
}

void parse_col_meta(TReader& r, ChunkMeta& c) {
    int id, last = 0; uint8_t t;
    c.has_meta = true;
    while (r.field(id, t, last)) {
        if (id == 1 && t == T_I32) c.type = (int)r.zigzag();
        else if (id == 4 && t == T_I32) c.codec = (int)r.zigzag();
        else if (id == 5 && t == T_I64) c.num_values = r.zigzag();
        else if (id == 6 && t == T_I64) c.total_uncompressed = r.zigzag();
        else if (id == 7 && t == T_I64) c.total_compressed = r.zigzag();
        else if (id == 9 && t == T_I64) c.data_page_offset = r.zigzag();
        else if (id == 11 && t == T_I64) { c.dict_page_offset = r.zigzag(); c.has_dict_offset = true; }
        else r.skip(t, false);
    }
}

void parse_chunk(TReader& r, ChunkMeta& c) {
    int id, last = 0; uint8_t t;
    while (r.field(id, t, last)) {
        if (id == 1 && t == T_BINARY) { c.external = !r.str().empty(); }
        else if (id == 3 && t == T_STRUCT) parse_col_meta(r, c);
        else if (id == 8 && t == T_STRUCT) { c.encrypted = true; r.skip(t, false); }
        else r.skip(t, false);
    }
}

void parse_row_group(TReader& r, RowGroupMeta& g) {
    int id, last = 0; uint8_t t;
    while (r.field(id, t, last)) {
        if (id == 1 && t == T_LIST) {
            uint32_t n; uint8_t et;
            if (!r.list(n, et) || et != T_STRUCT) { r.ok = false; return; }
            g.cols.resize(n);
            for (uint32_t i = 0; i < n && r.ok; ++i) parse_chunk(r, g.cols[i]);
        } else if (id == 3 && t == T_I64) g.num_rows = r.zigzag();
        else r.skip(t, false);
    }
}

struct PageHdr {
    int type = -1;
    int64_t uncomp = -1, comp = -1;
    int64_t num_values = -1, encoding = -1, def_enc = 3, def_bytes = 0, rep_bytes = 0;
    bool v2_compressed = true;
};

void parse_page_sub(TReader& r, PageHdr& h, int kind) {       // kind: 0 DataPageHeader, 1 DictionaryPageHeader, 2 DataPageHeaderV2
    int id, last = 0; uint8_t t;
    while (r.field(id, t, last)) {
        if (id == 1 && t == T_I32) h.num_values = r.zigzag();
        else if (kind != 2 && id == 2 && t == T_I32) h.encoding = r.zigzag();
        else if (kind == 0 && id == 3 && t == T_I32) h.def_enc = r.zigzag();
        else if (kind == 2 && id == 4 && t == T_I32) h.encoding = r.zigzag();
        else if (kind == 2 && id == 5 && t == T_I32) h.def_bytes = r.zigzag();
        else if (kind == 2 && id == 6 && t == T_I32) h.rep_bytes = r.zigzag();
        else if (kind == 2 && id == 7 && (t == T_TRUE || t == T_FALSE)) h.v2_compressed = t == T_TRUE;
        else r.skip(t, false);
    }
}

void parse_page_header(TReader& r, PageHdr& h) {
    int id, last = 0; uint8_t t;
    while (r.field(id, t, last)) {
        if (id == 1 && t == T_I32) h.type = (int)r.zigzag();
        else if (id == 2 && t == T_I32) h.uncomp = r.zigzag();
        else if (id == 3 && t == T_I32) h.comp = r.zigzag();
        else if (id == 5 && t == T_STRUCT) parse_page_sub(r, h, 0);
        else if (id == 7 && t == T_STRUCT) parse_page_sub(r, h, 1);
        else if (id == 8 && t == T_STRUCT) parse_page_sub(r, h, 2);
        else r.skip(t, false);
    }
}

uint64_t align16(uint64_t x) { return (x + 15u) & ~15ull; }

const char* codec_name(int c) {
    static const char* n[] = {"UNCOMPRESSED", "SNAPPY", "GZIP", "LZO", "BROTLI", "LZ4", "ZSTD", "LZ4_RAW"};
    return c >= 0 && c < 8 ? n[c] : "unknown";
}

}  // namespace

bool spl_pq_plan(const uint8_t* file, size_t n, const char* column, uint64_t batch_bytes, SplPqPlan& plan) {
    plan = SplPqPlan();
    auto fail = [&](const std::string& m, bool unsupported = false) { plan.err = "parquet: " + m; plan.unsupported = unsupported; return false; };
    if (!file || n < 12) return fail("not a Parquet file (too short)");
    if (memcmp(file + n - 4, "PARE", 4) == 0) return fail("encrypted footer: not supported", true);
    if (memcmp(file, "PAR1", 4) != 0 || memcmp(file + n - 4, "PAR1", 4) != 0) return fail("not a Parquet file (magic bytes)");
    const uint32_t meta_len = spl_pq_le32(file + n - 8);
    if ((uint64_t)meta_len + 12 > n) return fail("footer length beyond the file");
    TReader r(file + n - 8 - meta_len, file + n - 8);

    std::vector<SchemaEl> schema;
    std::vector<RowGroupMeta> groups;
    {
        int id, last = 0; uint8_t t;
        while (r.field(id, t, last)) {
            if (id == 2 && t == T_LIST) {
                uint32_t cnt; uint8_t et;
                if (!r.list(cnt, et) || et != T_STRUCT) return fail("malformed footer (schema)");
                schema.resize(cnt);
                for (uint32_t i = 0; i < cnt && r.ok; ++i) parse_schema_el(r, schema[i]);
            } else if (id == 4 && t == T_LIST) {
                uint32_t cnt; uint8_t et;
                if (!r.list(cnt, et) || (cnt && et != T_STRUCT)) return fail("malformed footer (row groups)");
                groups.resize(cnt);
                for (uint32_t i = 0; i < cnt && r.ok; ++i) parse_row_group(r, groups[i]);
            } else if (id == 8 && t == T_STRUCT) {
                return fail("encrypted columns: not supported", true);
            } else r.skip(t, false);
        }
        if (!r.ok) return fail("malformed footer");
    }
    if (schema.empty()) return fail("malformed footer (no schema)");

    // ---- the leaf: index among the leaves (= index of its chunk in every row group), levels along its path -----------
    int leaf_index = -1, leaf_type = -1;
    uint32_t max_def = 0, max_rep = 0;
    {
        const std::string want = column ? column : "";
        struct Frame { int left; std::string path; uint32_t def, rep; };
        std::vector<Frame> stack;
        stack.push_back(Frame{schema[0].num_children, "", 0, 0});
        int leaves = 0;
        bool found = false;
        for (size_t i = 1; i < schema.size(); ++i) {
            while (!stack.empty() && stack.back().left == 0) stack.pop_back();
            if (stack.empty()) return fail("malformed footer (schema tree)");
            Frame& top = stack.back();
            --top.left;
            const SchemaEl& e = schema[i];
            const std::string path = top.path.empty() ? e.name : top.path + "." + e.name;
            const uint32_t def = top.def + (e.repetition != 0 ? 1u : 0u), rep = top.rep + (e.repetition == 2 ? 1u : 0u);
            if (e.num_children > 0) {
                if (stack.size() > 64) return fail("schema nested too deeply");
                stack.push_back(Frame{e.num_children, path, def, rep});
            } else {
                if (!found && (path == want || (e.name == want && stack.size() == 1))) {
                    found = true; leaf_index = leaves; leaf_type = e.type; max_def = def; max_rep = rep;
                }
                ++leaves;
            }
        }
        if (!found) return fail("no column named '" + want + "'");
    }
    if (leaf_type != 6) return fail("the column is not a BYTE_ARRAY (string / binary) column", true);
    if (max_rep != 0) return fail("the column is repeated (a list): one document per row needs a flat column", true);
    if (max_def > 255) return fail("definition level beyond 255", true);
    plan.max_def = max_def;

    // ---- pages, row group by row group --------------------------------------------------------------------------------
    if (batch_bytes == 0) batch_bytes = 1ull << 30;
    SplPqBatch cur;
    memset(&cur, 0, sizeof(cur));
    auto close_batch = [&]() {
        if (cur.page1 > cur.page0 || cur.n_rows) plan.batches.push_back(cur);
        const size_t p1 = cur.page1, r1 = cur.range1;
        memset(&cur, 0, sizeof(cur));
        cur.page0 = cur.page1 = p1; cur.range0 = cur.range1 = r1;
    };
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const RowGroupMeta& g = groups[gi];
        if ((size_t)leaf_index >= g.cols.size()) return fail("row group without the column's chunk");
        const ChunkMeta& c = g.cols[leaf_index];
        if (c.external) return fail("column chunk in another file: not supported", true);
        if (c.encrypted) return fail("encrypted column: not supported", true);
        if (!c.has_meta) return fail("column chunk without metadata");
        if (c.type != 6) return fail("column chunk type differs from the schema");
        if (c.codec != SPL_PQ_CODEC_NONE && c.codec != SPL_PQ_CODEC_SNAPPY)
            return fail(std::string("codec ") + codec_name(c.codec) + " is not supported (UNCOMPRESSED and SNAPPY are): rewrite the file, "
                        "e.g. pyarrow.parquet.write_table(table, path, compression='snappy')", true);
        if (c.num_values < 0 || g.num_rows < 0 || c.num_values != g.num_rows) return fail("value count differs from the row count (flat column expected)");
        if (c.num_values == 0) continue;                                 // (writers emit an empty row group for an empty table)
        if (c.total_compressed < 0 || c.data_page_offset < 0) return fail("malformed column metadata");
        uint64_t start = (uint64_t)c.data_page_offset;
        if (c.has_dict_offset && c.dict_page_offset > 0 && (uint64_t)c.dict_page_offset < start) start = (uint64_t)c.dict_page_offset;
        const uint64_t stop = start + (uint64_t)c.total_compressed;
        if (start < 4 || stop > n - 8 || stop < start) return fail("column chunk beyond the file");
        if ((uint64_t)c.total_uncompressed > 0xE0000000ull || (uint64_t)g.num_rows > 0xFFFFFFF0ull)
            return fail("a row group of more than 3.5 GiB / 2^32 rows in this column: write smaller row groups", true);
        if (cur.n_rows && cur.text_bound + (uint64_t)c.total_uncompressed > batch_bytes) close_batch();
        if (cur.text_bound + (uint64_t)c.total_uncompressed > 0xE0000000ull) close_batch();

        SplPqRange rg{start, stop - start, align16(cur.stage_bytes)};
        const uint64_t dict_base = cur.dict_entries;
        uint32_t dict_count = 0;
        bool have_dict = false;
        uint64_t pos = start, seen = 0;
        while (seen < (uint64_t)c.num_values) {
            if (pos >= stop) return fail("column chunk ends before all its values");
            TReader pr(file + pos, file + stop);
            PageHdr h;
            parse_page_header(pr, h);
            if (!pr.ok || h.comp < 0 || h.uncomp < 0 || h.type < 0) return fail("malformed page header");
            const uint64_t body = (uint64_t)(pr.p - file);
            if ((uint64_t)h.comp > stop - body) return fail("page beyond its column chunk");
            pos = body + (uint64_t)h.comp;
            if (h.type == 1) continue;                                   // index page
            if (h.type != SPL_PQ_DATA_V1 && h.type != SPL_PQ_DICT && h.type != SPL_PQ_DATA_V2) return fail("unknown page type");
            if (h.num_values < 0 || h.num_values > 0x7FFFFFFF) return fail("malformed page header (value count)");
            SplPqPage pg;
            memset(&pg, 0, sizeof(pg));
            pg.kind = (uint8_t)h.type; pg.codec = (uint8_t)c.codec; pg.max_def = (uint8_t)max_def;
            pg.src = rg.stage_off + (body - start);
            pg.comp_size = (uint32_t)h.comp; pg.uncomp_size = (uint32_t)h.uncomp;
            pg.num_values = (uint32_t)h.num_values;
            pg.v2_compressed = h.v2_compressed ? 1 : 0;
            if (h.type == SPL_PQ_DICT) {
                if (have_dict) return fail("two dictionary pages in one column chunk");
                if (h.encoding != 0 && h.encoding != 2) return fail("dictionary page encoding other than PLAIN", true);
                have_dict = true; dict_count = pg.num_values;
                pg.dict_base = dict_base; pg.dict_count = dict_count;
                cur.dict_entries += dict_count;
            } else {
                if (h.encoding == 0) pg.encoding = SPL_PQ_ENC_PLAIN;
                else if (h.encoding == 2 || h.encoding == 8) {
                    if (!have_dict) return fail("dictionary-encoded page without a dictionary page");
                    pg.encoding = SPL_PQ_ENC_DICT;
                } else {
                    static const char* en[] = {"PLAIN", "?", "PLAIN_DICTIONARY", "RLE", "BIT_PACKED", "DELTA_BINARY_PACKED",
                                               "DELTA_LENGTH_BYTE_ARRAY", "DELTA_BYTE_ARRAY", "RLE_DICTIONARY", "BYTE_STREAM_SPLIT"};
                    return fail(std::string("page encoding ") + (h.encoding >= 0 && h.encoding < 10 ? en[h.encoding] : "unknown") +
                                " is not supported (PLAIN and dictionary encodings are): rewrite the column without column_encoding", true);
                }
                if (h.type == SPL_PQ_DATA_V1 && max_def && h.def_enc != 3) return fail("definition levels not RLE-encoded (BIT_PACKED is deprecated)", true);
                if (h.type == SPL_PQ_DATA_V2) {
                    if (h.def_bytes < 0 || h.rep_bytes < 0 || (uint64_t)h.def_bytes + (uint64_t)h.rep_bytes > (uint64_t)h.comp ||
                        (uint64_t)h.def_bytes + (uint64_t)h.rep_bytes > (uint64_t)h.uncomp) return fail("malformed V2 page header (level lengths)");
                    pg.def_bytes = (uint32_t)h.def_bytes; pg.rep_bytes = (uint32_t)h.rep_bytes;
                }
                pg.dict_base = dict_base; pg.dict_count = dict_count;
                pg.first_row = cur.n_rows + seen;
                seen += pg.num_values;
                if (seen > (uint64_t)c.num_values) return fail("pages hold more values than the column chunk declares");
            }
            const bool compressed = c.codec == SPL_PQ_CODEC_SNAPPY && (h.type != SPL_PQ_DATA_V2 || h.v2_compressed);
            if (compressed) {
                pg.scratch = align16(cur.scratch_bytes);
                cur.scratch_bytes = pg.scratch + pg.uncomp_size;
            } else if (pg.comp_size != pg.uncomp_size) return fail("uncompressed page with two different sizes");
            plan.pages.push_back(pg);
            ++cur.page1;
        }
        plan.ranges.push_back(rg);
        ++cur.range1;
        cur.stage_bytes = rg.stage_off + rg.len;
        cur.n_rows += (uint64_t)g.num_rows;
        cur.text_bound += (uint64_t)c.total_uncompressed;
        plan.n_rows += (uint64_t)g.num_rows;
    }
    close_batch();
    return true;
}
