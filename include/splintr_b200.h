/* splintr_b200 -- C ABI of the B200-native batch BPE encoder.
 *
 * This is the drop-in boundary for ONE path of ml-rust/splintr: Tokenizer::encode_batch
 * (and encode / encode_with_special / encode_batch_with_special, which are the same path
 * with one text / with the special-token scan).  A host (the reference's PyO3 class, or
 * the ctypes host in splintr_b200/tokenizer.py) packs its texts into ONE contiguous UTF-8
 * byte buffer plus n_docs+1 offsets and calls spl_encode_batch; everything between the
 * packed bytes and the token ids runs on the GPU.  There is no CPU implementation behind
 * these entry points: without a CUDA device spl_create fails with SPL_ERR_NO_DEVICE.
 *
 * Reference interfaces replaced (paths under the reference repository):
 *   spl_create            <- Tokenizer::from_bytes / from_bytes_byte_level        src/core/tokenizer.rs:552-569
 *                            (+ with_full_options: regex build, Aho-Corasick build  src/core/tokenizer.rs:410-456)
 *                            Python: Tokenizer.from_pretrained / from_bytes / __new__  src/python/bindings.rs:70-187
 *   spl_encode_batch      <- Tokenizer::encode_batch                              src/core/tokenizer.rs:932-934
 *                            Tokenizer::encode_batch_with_special (flag)           src/core/tokenizer.rs:937-942
 *                            Tokenizer::encode / encode_with_special (n_docs = 1)  src/core/tokenizer.rs:729-808, 842-874
 *                            Python: bindings.rs:254-288, 337-350
 *   spl_encode_batch_device  (same path, device-resident buffers; no reference counterpart)
 *   spl_destroy           <- Drop for Tokenizer
 *   spl_last_error        <- TokenizerError Display                               src/core/tokenizer.rs:19-36
 *
 * Conventions: every function returns SPL_OK (0) or a negative SPL_ERR_* code and never
 * throws across the ABI.  Inputs are owned by the caller and only read during the call.
 * Results are owned by the library (pinned host memory) until spl_result_free.  One call
 * in flight per tokenizer handle; distinct handles may be used from distinct threads.
 */
#ifndef SPLINTR_B200_H
#define SPLINTR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spl_tokenizer spl_tokenizer;
typedef struct spl_result spl_result;

enum {
    SPL_OK = 0,
    SPL_ERR_INVALID_ARG = -1,   /* null pointer, bad offsets, unknown pattern id ...        */
    SPL_ERR_VOCAB = -2,         /* vocabulary data does not parse (reference: VocabError)   */
    SPL_ERR_CUDA = -3,          /* a CUDA runtime call or kernel failed                     */
    SPL_ERR_OOM = -4,           /* host or device allocation failed                         */
    SPL_ERR_UNSUPPORTED = -5,   /* valid request this implementation cannot serve exactly   */
    SPL_ERR_NO_DEVICE = -6      /* no usable CUDA device (there is no CPU fallback)         */
};

/* pattern_id: which of the reference's split patterns the pre-tokenizer reproduces */
#define SPL_PATTERN_CL100K      0   /* CL100K_BASE_PATTERN  src/core/tokenizer.rs:39            */
#define SPL_PATTERN_O200K       1   /* O200K_BASE_PATTERN = LLAMA3_PATTERN  tokenizer.rs:42,45  */
#define SPL_PATTERN_MISTRAL_V3  2   /* MISTRAL_V3_PATTERN  src/core/tokenizer.rs:64             */
#define SPL_PATTERN_SENTENCEPIECE 3 /* SENTENCEPIECE_PATTERN tokenizer.rs:56, with SPL_CREATE_SENTENCEPIECE */

/* spl_create flags */
#define SPL_CREATE_BYTE_LEVEL   1u  /* vocabulary keys are GPT-2 byte-level strings (from_bytes_byte_level) */
#define SPL_CREATE_SENTENCEPIECE 2u /* SentencePiece mode (Tokenizer::from_bytes_sentencepiece, tokenizer.rs:589-640):
                                     * spaces become U+2581 prefixes of the following word, other ASCII whitespace is
                                     * encoded byte by byte (the encode branch at tokenizer.rs:737-795); of duplicated
                                     * vocabulary keys the FIRST id encodes and every id decodes (vocab.rs:101-143).
                                     * Requires SPL_PATTERN_SENTENCEPIECE.  decode returns the raw bytes; turning
                                     * U+2581 back into a space (tokenizer.rs:923-930) is string-level work of the host. */

/* spl_encode_batch flags */
#define SPL_ENCODE_WITH_SPECIAL 1u  /* recognise special-token strings (encode_batch_with_special) */

typedef struct spl_stats {
    uint64_t n_docs;
    uint64_t n_bytes;          /* input bytes                                   */
    uint64_t n_tokens;         /* output ids                                    */
    uint64_t h2d_bytes;        /* bytes copied host -> device for this call     */
    uint64_t d2h_bytes;        /* bytes copied device -> host for this call     */
    float kernel_ms;           /* device time of the kernels (max over devices) */
    float total_ms;            /* device time incl. copies (max over devices)   */
    int n_devices;
    int n_launches;            /* kernels launched (all devices)                */
} spl_stats;

/* Build a tokenizer: parse `vocab` (tiktoken format: "base64 SP rank LF" lines), build the
 * device tables on every listed device.  `devices` may be NULL (=> device 0 only).
 * On failure *out is NULL and spl_last_error(NULL) describes the error. */
int spl_create(const uint8_t* vocab, size_t vocab_len, int pattern_id, uint32_t flags,
               const char* const* special_strs, const uint32_t* special_ids, size_t n_special,
               const int* devices, int n_devices, spl_tokenizer** out);

void spl_destroy(spl_tokenizer* tok);

/* Last error message of `tok` (or of the last failed spl_create when tok is NULL). */
const char* spl_last_error(const spl_tokenizer* tok);

/* Encode n_docs documents.  `bytes` holds the concatenated UTF-8 texts, document i is
 * bytes[offsets[i] .. offsets[i+1]); offsets has n_docs+1 entries, offsets[0] == 0.
 * Documents are sharded over the handle's devices by cumulative bytes.  Host buffers may be
 * pageable or pinned (spl_alloc_pinned). */
int spl_encode_batch(spl_tokenizer* tok, const uint8_t* bytes, const uint64_t* offsets, size_t n_docs,
                     uint32_t flags, spl_result** out);

const uint32_t* spl_result_ids(const spl_result* r);       /* n_tokens ids, document order        */
const uint64_t* spl_result_offsets(const spl_result* r);   /* n_docs+1 offsets into ids            */
size_t spl_result_n_docs(const spl_result* r);
size_t spl_result_n_tokens(const spl_result* r);
void spl_result_stats(const spl_result* r, spl_stats* out);
void spl_result_free(spl_result* r);

/* Device-resident variant on the handle's device number `dev_index` (index into the
 * `devices` list of spl_create): all four buffers are device pointers on that device;
 * d_bytes must be 16-byte aligned and readable up to n_bytes rounded up to 16;
 * d_ids must hold ids_capacity >= n_bytes entries (worst case: one id per byte; for a SentencePiece-mode
 * handle n_bytes + 2 * (number of spaces) <= 3 * n_bytes, because a space becomes the three bytes of U+2581).
 * Work is enqueued on `cuda_stream` (a cudaStream_t, NULL = legacy default stream).
 * If n_tokens_out is non-NULL the call synchronises the stream, checks the device-side error flags (invalid
 * offsets -> SPL_ERR_INVALID_ARG; scratch for pieces beyond 1 KiB exhausted -> the pool is enlarged and the pass
 * repeated) and stores the id count; otherwise it returns SPL_OK after enqueueing (d_out_offsets[n_docs] holds the
 * count once the stream has run) and the flags are the caller's to collect: see spl_device_status. */
int spl_encode_batch_device(spl_tokenizer* tok, int dev_index,
                            const uint8_t* d_bytes, size_t n_bytes,
                            const uint64_t* d_offsets, size_t n_docs, uint32_t flags,
                            uint32_t* d_ids, size_t ids_capacity, uint64_t* d_out_offsets,
                            void* cuda_stream, uint64_t* n_tokens_out);

/* Error flags of the LAST pass enqueued on device `dev_index` by spl_encode_batch_device with n_tokens_out == NULL:
 * synchronises `cuda_stream` and stores the flags.  0 = the pass is good.  With SPL_STATUS_BAD_OFFSETS no kernel used
 * the offsets as indices and the output buffers hold nothing; with SPL_STATUS_SCRATCH_EXHAUSTED pieces beyond 1 KiB
 * produced no ids (call again with n_tokens_out != NULL, which enlarges the pool and repeats the pass). */
#define SPL_STATUS_BAD_OFFSETS        1u
#define SPL_STATUS_SCRATCH_EXHAUSTED  2u
int spl_device_status(spl_tokenizer* tok, int dev_index, void* cuda_stream, uint32_t* flags_out);

/* Diagnostics: the 32 device counters of the last pass on device `dev_index` (synchronises `cuda_stream`):
 * [1] error flags, [4] tiles the bit-parallel pre-tokenizer handed to the sequential rules, [5] long pieces that repeated
 * an earlier one (their merge loop was skipped), [6] tiles that went through the refining pass of the probe, [7] single
 * characters of two or three ids that the probe settled from its table,
 * [8 .. 15] pieces filed for the merge loop by length class. */
int spl_debug_counters(spl_tokenizer* tok, int dev_index, void* cuda_stream, uint32_t* out32);

/* ---- decode (the step on the other side of the path) ------------------------------
 * Replaces, for batches, Tokenizer::decode_bytes / decode_batch (src/core/tokenizer.rs:877-897, 945-958;
 * Python: bindings.rs:300-379): every id is looked up in the vocabulary (byte-level keys are returned as raw
 * bytes, tokenizer.rs:882-887), then among the special tokens; unknown ids contribute nothing.  UTF-8
 * validation of the result (decode vs decode_lossy) stays with the caller.
 * `ids` holds the concatenated token ids, document i is ids[offsets[i] .. offsets[i+1]).  The result carries
 * the concatenated bytes (spl_result_bytes / spl_result_n_bytes) and n_docs+1 byte offsets (spl_result_offsets). */
int spl_decode_batch(spl_tokenizer* tok, const uint32_t* ids, const uint64_t* offsets, size_t n_docs, spl_result** out);
const uint8_t* spl_result_bytes(const spl_result* r);
size_t spl_result_n_bytes(const spl_result* r);

/* Device-resident variant: all buffers are device pointers on the handle's device `dev_index`.  The call
 * synchronises `cuda_stream`; *n_bytes_out receives the decoded size.  If that exceeds bytes_capacity nothing
 * is written and SPL_ERR_INVALID_ARG is returned (retry with a buffer of *n_bytes_out bytes). */
int spl_decode_batch_device(spl_tokenizer* tok, int dev_index, const uint32_t* d_ids, size_t n_tokens,
                            const uint64_t* d_tok_offsets, size_t n_docs,
                            uint8_t* d_bytes_out, size_t bytes_capacity, uint64_t* d_out_offsets,
                            void* cuda_stream, uint64_t* n_bytes_out);

/* ---- ingestion: JSON Lines -> packed text + offsets, on the device ------------------
 * Not a reference function: it replaces the loop a splintr user runs in front of Tokenizer.encode_batch
 * (python/splintr/__init__.py documents `encode_batch(texts)` over a list the user built),
 *     texts = [json.loads(line)[field] for line in open(path) if line.strip()]
 * and the packing of `texts`.  d_jsonl = the file's bytes on the device (16-byte aligned, readable up to n_bytes
 * rounded up to 16).  Every non-blank line is one JSON object and yields one document: the unescaped string value of
 * its LAST top-level member named `field` (1..64 bytes); a line without it, with a non-string there, or that does not
 * parse yields an empty document and is counted in the stats.  Outputs (device memory of the caller): d_text_out
 * (text_capacity >= n_bytes always suffices; 16-byte aligned and padded if it is to be fed to
 * spl_encode_batch_device) and d_offsets_out (offsets_capacity entries; n_docs + 1 are written; the number of '\n'
 * bytes + 2 always suffices).  Too small a capacity: SPL_ERR_INVALID_ARG with the needed sizes in *stats (either
 * output pointer may be NULL to ask for the sizes only).  Synchronises the stream twice. */
typedef struct spl_ingest_stats {
    uint64_t n_lines;          /* '\n'-separated lines, blank ones included                 */
    uint64_t n_docs;           /* non-blank lines = documents                                */
    uint64_t n_text_bytes;     /* bytes of packed text                                       */
    uint64_t n_missing;        /* documents left empty: member absent or not a string        */
    uint64_t n_bad;            /* documents left empty: the line is not a JSON object        */
    int n_launches;
} spl_ingest_stats;
int spl_ingest_jsonl_device(spl_tokenizer* tok, int dev_index, const uint8_t* d_jsonl, size_t n_bytes, const char* field,
                            uint8_t* d_text_out, size_t text_capacity, uint64_t* d_offsets_out, size_t offsets_capacity,
                            void* cuda_stream, spl_ingest_stats* stats);

/* spl_encode_batch for a JSON Lines file held in host memory (pinned memory makes the copies asynchronous): the
 * file is cut into chunks at line ends; each chunk is copied in as it is, ingested and encoded on the handle's first
 * device, and its ids are copied out while the next chunk is worked on.  The result is the same object
 * spl_encode_batch returns (one document per non-blank line); `ingest_stats` (optional) receives the totals. */
int spl_encode_jsonl(spl_tokenizer* tok, const uint8_t* bytes, size_t n_bytes, const char* field, uint32_t flags,
                     spl_result** out, spl_ingest_stats* ingest_stats);

/* ---- ingestion, Parquet: one string column of a Parquet file -> packed text + offsets ---
 * Replaces the loop a splintr user runs in front of Tokenizer.encode_batch,
 *     texts = pyarrow.parquet.read_table(path, columns=[column])[column].to_pylist()
 * and the packing of `texts`.  `bytes` = the whole file in HOST memory: the footer and the page headers are read on the
 * host (a few hundred bytes per page), the column chunks go to the device as they lie in the file, and pages are
 * decompressed (snappy) and decoded (PLAIN, dictionary, V1 / V2 data pages, definition levels) there.  `column` =
 * the leaf's name or dotted path ("text", "meta.body").  One document per row; a null at any level is an EMPTY document.
 * Supported: BYTE_ARRAY (string / binary) columns that are not repeated; UNCOMPRESSED and SNAPPY.  Anything else the
 * format allows (other codecs, DELTA_* encodings, lists, encryption) is refused with SPL_ERR_UNSUPPORTED and a message
 * that names what to rewrite; a damaged file gives SPL_ERR_INVALID_ARG.
 *
 * spl_ingest_parquet: outputs in device memory of the caller (d_text_out 16-byte aligned and 16 bytes longer than the
 * text if it is to be fed to spl_encode_batch_device; d_offsets_out: rows + 1 entries).  The sizes are not known in
 * advance (a dictionary page can expand): with too small a capacity -- or NULL outputs -- nothing is written, the needed
 * sizes are in *stats (n_docs rows, n_text_bytes) and SPL_ERR_INVALID_ARG is returned (NULL outputs: SPL_OK).  The
 * column must fit one device pass (4 GiB of text).  stats: n_lines = n_docs = rows. */
int spl_ingest_parquet(spl_tokenizer* tok, int dev_index, const uint8_t* bytes, size_t n_bytes, const char* column,
                       uint8_t* d_text_out, size_t text_capacity, uint64_t* d_offsets_out, size_t offsets_capacity,
                       void* cuda_stream, spl_ingest_stats* stats);

/* spl_encode_batch for one string column of a Parquet file held in host memory: runs of row groups (about 1 GiB of
 * column data each) are ingested and encoded on the handle's first device, their ids copied out.  The result is the
 * object spl_encode_batch returns (one document per row). */
int spl_encode_parquet(spl_tokenizer* tok, const uint8_t* bytes, size_t n_bytes, const char* column, uint32_t flags,
                       spl_result** out, spl_ingest_stats* ingest_stats);

/* number of kernels one spl_encode_batch_device call launches for these flags */
int spl_launches_per_call(const spl_tokenizer* tok, uint32_t flags);

/* Per-kernel device timing of spl_encode_batch_device (measurement aid, off by default): when
 * enabled, CUDA events are recorded on the caller's stream around every kernel of the path.
 * After the caller has synchronised that stream, spl_last_kernel_times stores up to `cap`
 * kernel names / durations (ms) of the most recent device call and returns how many. */
int spl_set_profiling(spl_tokenizer* tok, int enable);
int spl_last_kernel_times(spl_tokenizer* tok, int dev_index, const char** names, float* ms, int cap);

/* pinned host memory helpers for callers that want full-rate PCIe copies */
void* spl_alloc_pinned(size_t bytes);
void spl_free_pinned(void* p);

/* library / build identification, e.g. "splintr_b200 0.1.0 sm_100a" */
const char* spl_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SPLINTR_B200_H */
