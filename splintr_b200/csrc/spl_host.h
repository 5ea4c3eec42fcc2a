// Host-side table construction for the device encode path (no CUDA in this file's API).
//
// Replaces, for the GPU path, what `Tokenizer::with_full_options`
// (/root/reference/src/core/tokenizer.rs:410-456) builds at construction time:
//   * the byte-keyed encoder map (vocab.rs:57-89)      -> whole-piece tables T8/T16/TL
//   * the rank lookups of bpe.rs:99-111                 -> (left,right)->merged pair table
//   * the Aho-Corasick matcher (tokenizer.rs:429-434)   -> special-string list + first-byte set
//   * byte_level_encode (byte_level.rs:46-74,105-107)   -> folded into the keys (translated to raw bytes)
#pragma once
#include <string>
#include <vector>
#include <unordered_map>
#include "spl_common.h"

enum : uint32_t { SPL_FLAG_BYTE_LEVEL = 1u,
                  SPL_FLAG_SENTENCEPIECE = 2u };   // vocab.rs:101-143: first id of a duplicated key encodes, every id decodes

struct SplHostTables {
    int pattern = 0;
    uint32_t flags = 0;
    std::unordered_map<std::string, uint32_t> encoder;   // RAW-byte keys -> id (byte-level keys translated)
    std::vector<SplKey8>  t8;   uint32_t t8_log2 = 0;
    std::vector<SplKey16> t16;  uint32_t t16_log2 = 0;
    std::vector<SplKeyL>  tl;   uint32_t tl_log2 = 0;
    std::vector<uint8_t>  tok_bytes;
    std::vector<uint32_t> tok_off;
    uint32_t n_ids = 0, max_key_len = 0;
    std::vector<uint32_t> pair; uint32_t pair_log2 = 0; size_t n_pairs = 0;      // buckets of SPL_PAIR_WORDS words
    std::vector<uint32_t> bpair;                                                  // [65536] dense byte x byte corner of it
    std::vector<uint32_t> seg_irr, seg_h2, char_tok, char_ids;  uint32_t seg_h2_log2 = 16; size_t seg_pairs = 0;   // spl_segment.h
    size_t t8_displaced = 0, pair_displaced = 0;         // keys that are not in their home bucket
    uint32_t byte_sym[256];
    // decode (tokenizer.rs:877-897): id -> bytes for every vocabulary id (byte-level keys translated back to raw
    // bytes, untranslatable ones kept as they are) and, where the vocabulary has no entry, the special-token string
    std::vector<uint8_t>  dec_bytes;
    std::vector<uint32_t> dec_off;                        // [n_dec + 1]
    std::vector<uint8_t>  sp_bytes;
    std::vector<uint32_t> sp_off, sp_id;
    uint32_t sp_first[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool specials_unambiguous = true;    // no two special strings can overlap in any text
    std::string error;
};

// Parse `base64 SP rank LF` lines (vocab.rs:57-89).  Later duplicates overwrite.
bool spl_parse_tiktoken(const uint8_t* data, size_t len,
                        std::vector<std::pair<std::string, uint32_t>>& out, std::string& err);

// Build every table.  Returns false and sets t.error on failure.
bool spl_build_tables(SplHostTables& t, const uint8_t* vocab, size_t vocab_len, int pattern, uint32_t flags,
                      const char* const* special_strs, const uint32_t* special_ids, size_t n_special);

// Host probes that mirror the device probes exactly (same hash, same layout).
uint32_t spl_host_lookup_piece(const SplHostTables& t, const uint8_t* p, uint32_t len);   // SPL_RANK_NONE if absent
uint32_t spl_host_lookup_pair(const SplHostTables& t, uint32_t l, uint32_t r);
uint64_t spl_host_hashL(const uint8_t* p, uint32_t len);
// merge loop of bpe.rs:83-194 without the whole-piece probe (ids appended to out)
void spl_host_merge_loop(const SplHostTables& t, const uint8_t* p, uint32_t n, std::vector<uint32_t>& out);
