// Host-side model: what HuggingFaceTokenizer::from_tokenizer_json_with_config builds
// (reference src/huggingface/mod.rs:247-334), reduced to the tables the hot path needs.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "regex_dfa.hpp"

namespace ctk {

constexpr uint32_t kNoId = 0xFFFFFFFFu;

struct AddedTok {
    std::string content;        // as written in tokenizer.json (compared against byte-MAPPED words)
    std::vector<uint8_t> bytes; // content translated back through the byte map (only if may_match)
    uint32_t id = 0;
    bool special = false, single_word = false, lstrip = false, rstrip = false;
    bool may_match = false;     // can occur inside one pre-token (see loader.cpp: analyse_added)
};

struct PairEntry { uint32_t a, b, rank, new_id; };

struct HostModel {
    std::unordered_map<std::string, uint32_t> vocab;      // model.vocab
    std::vector<std::string> id_to_token;                  // dense, index = id
    std::vector<uint8_t> id_present;
    std::vector<PairEntry> pairs;                          // merge_ranks with new_id resolved (bpe.rs:52-79,:141)
    std::vector<AddedTok> added;                           // de-duplicated by content (HashMap semantics)
    std::vector<std::pair<std::string, uint32_t>> specials;// special_tokens map (mod.rs:290-291)
    bool nfc = true;                                       // parsing.rs:89
    bool add_prefix_space = false;                         // parsing.rs:99-107
    // Metaspace pipelines (pretokenizers.rs:188-200, decoders.rs:121-131): the LAST pre-tokenizer stage is Metaspace
    // instead of ByteLevel (words are runs of non-white-space characters, BPE symbols are characters), and / or the decoder is
    bool metaspace = false; uint32_t meta_replacement = 0x2581; bool meta_prefix = true;       // parsing.rs:108-123
    bool dec_metaspace = false; uint32_t dec_meta_replacement = 0x2581; bool dec_meta_strip = true;   // parsing.rs:279-293
    std::vector<uint32_t> meta_empty_ids;                  // ids of the word that consists of the replacement alone (what an EMPTY text encodes to when add_prefix_space is set)
    std::vector<std::pair<uint32_t, uint32_t>> char_ids;   // (code point, id) of every single-character vocabulary entry (bpe.rs:94-97 on characters)
    std::vector<SplitStage> split_stages;                  // Split stages in front of the ByteLevel stage (parsing.rs:145-167), compiled at load
    uint32_t byte_init_id[256];                            // byte -> id of its mapped char, kNoId if absent (bpe.rs:94-97)
    // decode side (vocab.rs:47-51 + decoders.rs:94-116 folded per token at load)
    std::vector<uint8_t> dec_blob;
    std::vector<uint32_t> dec_off;                         // max_id+2 entries; absent ids have empty ranges
    std::vector<uint8_t> dec_special;                      // 1 if the token string is in special_tokens
    size_t dec_max_bytes = 0;
    bool any_added_may_match = false;
    // Round-parallel merging of very long pre-tokens (encode_xlong.cuh) is exact only for MONOTONE tables:
    // every pair that contains a merged token ranks after every merge producing that token.
    bool merges_monotone = false;
    uint32_t max_token_span = 1;                           // longest merged token, in initial symbols
    // Per-token reach (in initial symbols): how far a token that ENDS with this one extends to its left, and how far
    // a token that STARTS with it extends to its right.  Packed left | right << 16, index = id.  Only filled when
    // `round_parallel` holds: monotone table, every merge product is the concatenation of its parts, ids unique.
    std::vector<uint32_t> reach;
    bool round_parallel = false;
    // ---- rich `Encoding` outputs (loader.cpp: parse_post_processor)
    // The post-processor reduced to what `process(ids, None)` does (postprocessors.rs:34-55 with pair_ids = None,
    // which is the only way mod.rs:372-375 calls it): an ordered list of items, -1 = "the ids" ($A), else a literal id.
    std::vector<int64_t> pp_items;                         // {-1} when the file has no (parsable) post_processor
    bool has_post_processor = false;
    std::vector<uint32_t> token_str_len;                   // byte length of the vocabulary string of each id (0: absent), mod.rs:421-425
    uint32_t pad_id = 0;                                   // special_tokens["[PAD]"] or ["<pad>"] or 0 (mod.rs:504-507)
    std::string pad_token;                                 // id_to_token(pad_id) or "<pad>"
};

// Returns a CTK_* code; on failure `err` holds the message.
int load_model(const uint8_t* json, size_t len, HostModel& m, std::string& err);

// GPT-2 byte <-> code point map (pretokenizers.rs:130-153 / decoders.rs:70-91)
void byte_map(uint32_t byte_to_cp[256]);

}  // namespace ctk
