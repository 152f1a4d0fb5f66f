// Split pre-tokenizer patterns (reference src/pretokenizers.rs:298-433, src/huggingface/parsing.rs:145-167).
//
// The reference hands the pattern to the `regex` crate and walks find_iter.  Here the pattern is compiled ON THE HOST, at
// load time, into a table-driven automaton the device walks (csrc/split.cu):
//
//   pattern --parse--> AST --Thompson--> NFA with priorities --ordered subsets--> leftmost-first DFA over code-point CLASSES
//
// * leftmost-first: a DFA state is an ORDERED list of NFA threads; everything after the first thread that has reached the
//   end of the pattern is dropped, so "the last accepting position seen before the walk dies" is exactly the match the
//   crate's (Perl-like) priority order selects.
// * code-point classes: the pattern's sets partition U+0000..U+10FFFF into a few equivalence classes; the device maps a
//   decoded code point to its class through a two-stage table and indexes trans[state][class].
//
// Three outcomes, never a guess: compiled / the crate certainly rejects the pattern (look-around, back-references: the
// reference then passes text through, pretokenizers.rs:299-302) / unsupported (CTK_ERR_UNSUPPORTED at load).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ctk {

enum SplitBehavior : int { SPLIT_REMOVED = 0, SPLIT_ISOLATED = 1, SPLIT_MERGED_PREV = 2, SPLIT_MERGED_NEXT = 3, SPLIT_CONTIGUOUS = 4 };

struct SplitDfa {
    uint32_t n_states = 0, n_classes = 0, start = 0;     // state 0 is the dead state
    std::vector<uint16_t> trans;                          // [n_states * n_classes]: next state | 0x8000 if that state accepts
    std::vector<uint8_t> ascii_class;                     // [128]
    std::vector<uint16_t> stage1;                         // [0x1100] cp >> 8 -> block
    std::vector<uint8_t> blocks;                          // [n_blocks * 256] class of every code point of the block
    uint32_t neutral[4] = {0, 0, 0, 0};                   // bit per ASCII byte that belongs to NONE of the pattern's sets: no match can contain it
    // bit (c1 * 128 + c2): no match can contain the ASCII byte c1 immediately followed by c2 -- find_iter is fresh at such a c2
    std::vector<uint32_t> pair_impossible;                // [512]
    std::vector<uint16_t> trans_ascii;                    // [n_states * 128] trans[state][ascii_class[byte]] (empty if n_states > 512)
};

struct SplitStage {
    std::string pattern;
    int behavior = SPLIT_REMOVED;                         // parsing.rs:156-166 (default Removed)
    bool invert = false;
    SplitDfa dfa;
};

// 0 = compiled, 1 = unsupported (err says why).  Patterns the crate certainly rejects never get here (loader.cpp).
int compile_split_regex(const std::string& pattern, SplitDfa& out, std::string& err);

}  // namespace ctk
