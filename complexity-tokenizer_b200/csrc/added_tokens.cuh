// In-word added-token matching (reference: HuggingFaceTokenizer::encode, src/huggingface/mod.rs:566-610,
// find_next_added_token_in_word :616-634, find_added_token :637-675).
//
// The reference searches added tokens INSIDE each already-split, byte-mapped pre-token, not in the raw
// text.  The byte map is a bijection on bytes, so matching the mapped strings char by char equals matching
// the raw bytes; `token.len()` comparisons ("longest at position 0") only ever compare a token with one of
// its own prefixes, so byte lengths order them the same way.  Mapped characters are never whitespace, so
// `lstrip` only allows position 0 and `rstrip` only a match that ends the remaining text.
// The whole procedure is a pure function of the pre-token's bytes: it runs on the cache-miss path only,
// and its result (BPE ids of the chunks interleaved with added-token ids) is what the cache stores.
#pragma once
#include "device_common.cuh"

namespace ctk {

#if defined(__CUDACC__)
// find_added_token for token `ti` in rem[0..n): position of its FIRST occurrence if the flags accept it, else -1
__device__ __forceinline__ int added_find(const DevTables& t, uint32_t ti, const uint8_t* rem, int n) {
    const uint4 m = __ldg(t.added_meta + ti);
    const int L = (int)m.y;
    if (L > n) return -1;
    const uint8_t* tok = t.added_blob + m.x;
    int pos = -1;
    for (int s = 0; s + L <= n; ++s) {
        int k = 0;
        while (k < L && rem[s + k] == __ldg(tok + k)) ++k;
        if (k == L) { pos = s; break; }
    }
    if (pos < 0) return -1;
    const int end = pos + L;
    if (m.w & 1u) {                                           // single_word (mod.rs:641-656)
        bool before_ok = pos == 0 || !__ldg(t.mapped_alnum + rem[pos - 1]);
        bool after_ok = end >= n || !__ldg(t.mapped_alnum + rem[end]);
        if (!before_ok || !after_ok) return -1;
    }
    if ((m.w & 2u) && pos > 0) return -1;                     // lstrip: previous char would have to be whitespace
    if ((m.w & 4u) && end < n) return -1;                     // rstrip: next char would have to be whitespace
    return pos;
}

// One step of the reference's `while !remaining.is_empty()` loop, warp-cooperative (lane i checks tokens i, i+32, ...).
// Returns the piece length; *id = the added token's id if the piece is an added token, kNone if it is a BPE chunk.
__device__ __forceinline__ int added_next_piece(const DevTables& t, const uint8_t* rem, int n, int lane, uint32_t* id) {
    const unsigned full = 0xFFFFFFFFu;
    uint32_t best = 0;                                        // (length << 8 | lane-local flag) of the longest match at 0
    uint32_t best_id = 0;
    uint32_t nxt = (uint32_t)n;
    for (uint32_t ti = (uint32_t)lane; ti < t.n_added; ti += 32) {
        int pos = added_find(t, ti, rem, n);
        if (pos == 0) {
            uint32_t L = __ldg(t.added_meta + ti).y;
            if (L > best) { best = L; best_id = __ldg(t.added_meta + ti).z; }
        } else if (pos > 0 && (uint32_t)pos < nxt) nxt = (uint32_t)pos;
    }
    const uint32_t gbest = __reduce_max_sync(full, best);
    if (gbest) {                                              // mod.rs:590-594
        int owner = __ffs(__ballot_sync(full, best == gbest)) - 1;
        *id = __shfl_sync(full, best_id, owner);
        return (int)gbest;
    }
    *id = kNone;                                              // mod.rs:596-607: BPE up to the next added token
    return (int)__reduce_min_sync(full, nxt);
}
#endif

}  // namespace ctk
