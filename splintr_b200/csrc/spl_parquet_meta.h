// Host side of the Parquet ingestion (spl_parquet.h): footer + page headers of one column -> page descriptors.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "spl_parquet.h"

// The column's pages of a run of row groups that go through the device together.
struct SplPqBatch {
    size_t page0, page1;             // pages[page0, page1): every dictionary page ahead of the data pages that use it
    uint64_t n_rows;
    uint64_t stage_bytes;            // staged file bytes (the column chunks, each 16-byte aligned)
    uint64_t scratch_bytes;          // decompressed bytes
    uint64_t dict_entries;
    uint64_t text_bound;             // no more text bytes than this
    size_t range0, range1;           // ranges[range0, range1)
};

struct SplPqRange { uint64_t file_off, len, stage_off; };   // a column chunk: where it lies in the file / in the staged bytes

struct SplPqPlan {
    std::vector<SplPqPage> pages;    // src / scratch / first_row / dict_base relative to their batch
    std::vector<SplPqRange> ranges;
    std::vector<SplPqBatch> batches;
    uint64_t n_rows = 0;
    uint32_t max_def = 0;
    std::string err;
    bool unsupported = false;        // err describes something the format allows and this reader does not take
};

// file[0, n): a whole Parquet file.  column: the leaf's name ("text") or dotted path ("meta.body").  A batch ends at a
// row group boundary once it holds batch_bytes of (uncompressed) column data.  Returns false with plan.err set.
bool spl_pq_plan(const uint8_t* file, size_t n, const char* column, uint64_t batch_bytes, SplPqPlan& plan);
