// Device-side tables and the primitives shared by the encode kernels.
//
// What each piece restates (reference paths relative to /root/reference):
//   cp_class / is_pretoken_start   the pre-token pattern of src/pretokenizers.rs:13, as the bounded-window
//                                  local rule derived in SURVEY.md section 3.2(i)
//   PairSlot / pair_lookup         BpeTokenizer::merge_ranks + merges[rank].new_id (src/bpe.rs:52-79, :141)
//   bpe_warp32                     BpeTokenizer::encode_with_dropout at dropout 0 (src/bpe.rs:88-153):
//                                  one merge per iteration, lowest rank, leftmost on ties
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CTK_HD __host__ __device__ __forceinline__
#define CTK_D __device__ __forceinline__
#else
#define CTK_HD inline
#define CTK_D inline
#endif

namespace ctk {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int CLS_O = 0, CLS_L = 1, CLS_N = 2, CLS_W = 3;

struct PairSlot { uint32_t a, b, rank, new_id; };      // 16 B; a == kNone marks an empty slot

struct CacheSlot {                                       // 32 B = one L2 sector; see encode_fused.cu
    uint64_t k0, k1;                                     // pre-token bytes, zero padded
    uint32_t meta;                                       // kNone = empty ; else len | ntok<<8 | READY<<31
    uint32_t tok[3];
};

struct DevTables {
    const PairSlot* pairs;      // open addressing, linear probing, capacity = pair_mask + 1 (power of two, load <= 1/4)
    uint32_t pair_mask;
    const uint32_t* byte_init;  // [256] byte -> initial id (kNone: mapped char not in vocab, symbol is dropped)
    const uint8_t* trie_index;  // [0x1100] cp >> 8 -> block
    const uint8_t* trie_blocks; // 128 B per block: 4 bits per cp (bits 0-1 class, bit 2 NFC-suspect)
    // added tokens that can occur inside one pre-token (mod.rs:566-675 searches words, not raw text)
    const uint8_t* added_blob;  // their contents translated back to raw bytes
    const uint4* added_meta;    // {offset, length, id, flags: 1 single_word, 2 lstrip, 4 rstrip}
    uint32_t n_added;
    const uint8_t* mapped_alnum;// [256] char::is_alphanumeric of the byte-mapped character of each byte
    const uint32_t* reach;      // [n ids] token reach left | right << 16 (model.hpp); only if round_parallel
    uint32_t round_parallel;    // the round-parallel merge rule is exact for this table (loader.cpp: token_reach)
};

// ------------------------------------------------------------------------------------------------
// code point classes
CTK_HD int ascii_class(uint32_t c) {
    uint32_t l = (c | 0x20u) - 'a';
    if (l < 26u) return CLS_L;
    if (c - '0' < 10u) return CLS_N;
    if (c == ' ' || c - 9u < 5u) return CLS_W;
    return CLS_O;
}
CTK_HD uint32_t trie_nibble(const uint8_t* idx, const uint8_t* blocks, uint32_t cp) {
    if (cp >= 0x110000u) return 0;
    uint32_t blk = idx[cp >> 8];
    uint32_t byte = blocks[blk * 128u + ((cp & 255u) >> 1)];
    return (cp & 1u) ? (byte >> 4) : (byte & 15u);
}

// ------------------------------------------------------------------------------------------------
// Text view used by the scalar (reference-shaped) start predicate.  `ds` is a bitmap with one bit
// per byte position: set where a document starts.  Position n (end of text) behaves as a start.
struct TextView {
    const uint8_t* text;
    uint64_t n;
    const uint32_t* ds;
    const uint8_t* trie_index;
    const uint8_t* trie_blocks;
    uint64_t dlo, dhi;          // when ds == nullptr: the view is one document [dlo, dhi)
    CTK_HD bool docstart(uint64_t i) const {
        if (i >= n) return true;
        if (ds) return (ds[i >> 5] >> (i & 31)) & 1u;
        return i <= dlo || i >= dhi;
    }
    CTK_HD uint32_t byte(uint64_t i) const { return i < n ? text[i] : 0u; }
    // class of the code point that contains byte i (continuation bytes inherit their lead's class)
    CTK_HD int cls(uint64_t i) const {
        uint32_t b = byte(i);
        if (b < 0x80u) return ascii_class(b);
        uint64_t l = i;
        for (int k = 0; k < 3 && (text[l] & 0xC0u) == 0x80u && l > 0 && !docstart(l); ++k) --l;
        uint32_t c = text[l], cp;
        if (c < 0xE0u) cp = ((c & 0x1Fu) << 6) | (byte(l + 1) & 63u);
        else if (c < 0xF0u) cp = ((c & 0x0Fu) << 12) | ((byte(l + 1) & 63u) << 6) | (byte(l + 2) & 63u);
        else cp = ((c & 7u) << 18) | ((byte(l + 1) & 63u) << 12) | ((byte(l + 2) & 63u) << 6) | (byte(l + 3) & 63u);
        return (int)(trie_nibble(trie_index, trie_blocks, cp) & 3u);
    }
    // " ?" of the pattern: a U+0020 that gets glued to the next pre-token: single space (previous
    // char is not whitespace, or the space opens the document) followed by a non-whitespace char.
    CTK_HD bool glued_space(uint64_t j) const {
        if (byte(j) != 0x20u || j >= n) return false;
        if (!docstart(j) && cls(j - 1) == CLS_W) return false;
        if (docstart(j + 1)) return false;
        return cls(j + 1) != CLS_W;
    }
    // 's|'t|'re|'ve|'m|'ll|'d starting at j, and j is a position where the regex tries a new match.
    // Returns the contraction's byte length (2 or 3) or 0.
    CTK_HD int contraction_at(uint64_t j) const {
        if (byte(j) != '\'' || j >= n) return 0;
        if (!docstart(j)) {
            int p = cls(j - 1);
            bool ms = p == CLS_L || p == CLS_N || (p == CLS_W && !glued_space(j - 1));
            if (!ms) return 0;
        }
        if (docstart(j + 1)) return 0;
        uint32_t c1 = byte(j + 1);
        if (c1 == 's' || c1 == 't' || c1 == 'm' || c1 == 'd') return 2;
        if (docstart(j + 2)) return 0;
        uint32_t c2 = byte(j + 2);
        if ((c1 == 'r' && c2 == 'e') || (c1 == 'v' && c2 == 'e') || (c1 == 'l' && c2 == 'l')) return 3;
        return 0;
    }
    // Does a pre-token (a match of the pattern) start at byte i?
    CTK_HD bool is_start(uint64_t i) const {
        if (docstart(i)) return true;
        uint32_t b = byte(i);
        if ((b & 0xC0u) == 0x80u) return false;                 // continuation byte
        int c = cls(i), p1 = cls(i - 1);
        if (c == CLS_W) return p1 != CLS_W;
        // non-whitespace
        if (i >= 2 && !docstart(i - 1) && contraction_at(i - 2) == 2) return true;        // right after 's 't 'm 'd
        if (i >= 3 && !docstart(i - 1) && !docstart(i - 2) && contraction_at(i - 3) == 3) return true;   // after 're 've 'll
        if (contraction_at(i - 1) != 0) return false;                                       // first letter of a contraction
        if (i >= 2 && !docstart(i - 1) && contraction_at(i - 2) == 3) return false;       // second letter
        if (p1 == CLS_W) return !glued_space(i - 1);
        return p1 != c;
    }
};

// ------------------------------------------------------------------------------------------------
CTK_HD uint32_t pair_hash(uint32_t a, uint32_t b) {
    uint32_t h = a * 0x9E3779B1u + b * 0x85EBCA6Bu;
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    return h;
}

#if defined(__CUDACC__)
// (rank, new_id) of the pair, rank == kNone if the pair has no merge.
// (Buckets of four slots loaded together were measured too: fewer dependent round trips per lookup, but more
// instructions and L2 sectors per probe; the encode kernel lost 2 % and the round kernels 10 %.)
CTK_D uint2 pair_lookup_from(const DevTables& t, uint32_t a, uint32_t b, uint32_t h) {
    for (;;) {
        uint4 s = __ldg(reinterpret_cast<const uint4*>(t.pairs + h));
        if (s.x == a && s.y == b) return make_uint2(s.z, s.w);
        if (s.x == kNone) return make_uint2(kNone, 0u);
        h = (h + 1u) & t.pair_mask;
    }
}
CTK_D uint2 pair_lookup(const DevTables& t, uint32_t a, uint32_t b) {
    return pair_lookup_from(t, a, b, pair_hash(a, b) & t.pair_mask);
}
// two independent lookups whose first probes are in flight together (one round trip instead of two, usually)
CTK_D void pair_lookup2(const DevTables& t, uint32_t a1, uint32_t b1, uint32_t a2, uint32_t b2, uint2& r1, uint2& r2) {
    const uint32_t h1 = pair_hash(a1, b1) & t.pair_mask, h2 = pair_hash(a2, b2) & t.pair_mask;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(t.pairs + h1));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(t.pairs + h2));
    if (u.x == a1 && u.y == b1) r1 = make_uint2(u.z, u.w);
    else if (u.x == kNone) r1 = make_uint2(kNone, 0u);
    else r1 = pair_lookup_from(t, a1, b1, (h1 + 1u) & t.pair_mask);
    if (v.x == a2 && v.y == b2) r2 = make_uint2(v.z, v.w);
    else if (v.x == kNone) r2 = make_uint2(kNone, 0u);
    else r2 = pair_lookup_from(t, a2, b2, (h2 + 1u) & t.pair_mask);
}

// Warp-cooperative BPE of one pre-token of n <= 32 symbols.  Lane i holds symbol i (valid for i < n).
// On return lanes [0, m) hold the final ids, m is returned (same value on every lane).
// Each iteration applies exactly ONE merge: the lowest rank, leftmost on ties (bpe.rs:127-153).
CTK_D int bpe_warp32(const DevTables& t, uint32_t& sym, int n) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    uint32_t rank = kNone, nid = 0;
    {
        uint32_t nxt = __shfl_down_sync(full, sym, 1);
        if (lane + 1 < n) { uint2 r = pair_lookup(t, sym, nxt); rank = r.x; nid = r.y; }
    }
    while (n > 1) {
        // ranks are < 2^27 (merge line index), so rank*32+lane is an order-preserving u32 key
        uint32_t key = rank == kNone ? kNone : (rank << 5) | (uint32_t)lane;
        uint32_t best = __reduce_min_sync(full, key);
        if (best == kNone) break;
        int idx = (int)(best & 31u);
        uint32_t new_id = __shfl_sync(full, nid, idx);
        if (lane == idx) sym = new_id;
        // close the gap: lanes > idx take their right neighbour's symbol and cached pair
        uint32_t s_up = __shfl_down_sync(full, sym, 1);
        uint32_t r_up = __shfl_down_sync(full, rank, 1);
        uint32_t n_up = __shfl_down_sync(full, nid, 1);
        if (lane > idx) { sym = s_up; rank = r_up; nid = n_up; }
        --n;
        // only the pairs touching the new symbol changed: (idx-1, idx) and (idx, idx+1)
        uint32_t nxt = __shfl_down_sync(full, sym, 1);
        if (lane == idx || lane == idx - 1) {
            rank = kNone; nid = 0;
            if (lane + 1 < n) { uint2 r = pair_lookup(t, sym, nxt); rank = r.x; nid = r.y; }
        }
        if (lane >= n - 1) rank = kNone;
    }
    return n;
}
// Pre-tokens longer than 32 symbols: same one-merge-per-iteration order, symbols kept compacted in
// global memory (the tmp_ids slice of this pre-token), every pair re-probed each iteration.
// O(n^2/32) probes per lane: exact but slow; long pre-tokens are rare outside config 4.
CTK_D int bpe_warp_long(const DevTables& t, uint32_t* sym, int n) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    while (n > 1) {
        uint32_t best = kNone, best_new = 0;
        uint32_t best_pos = 0;
        for (int i = lane; i + 1 < n; i += 32) {
            uint2 r = pair_lookup(t, sym[i], sym[i + 1]);
            if (r.x < best) { best = r.x; best_new = r.y; best_pos = (uint32_t)i; }   // strict <: leftmost within the lane
        }
        // lowest rank, then leftmost position, across lanes
        unsigned long long key = best == kNone ? ~0ull : ((unsigned long long)best << 32) | best_pos;
        unsigned long long k2 = key;
        for (int o = 16; o; o >>= 1) { unsigned long long v = __shfl_xor_sync(full, k2, o); k2 = v < k2 ? v : k2; }
        if (k2 == ~0ull) break;
        int idx = (int)(uint32_t)k2;
        int owner = __ffs(__ballot_sync(full, key == k2)) - 1;
        uint32_t new_id = __shfl_sync(full, best_new, owner);
        __syncwarp();
        // shift left by one beyond idx, in rounds of 32 so reads happen before writes
        for (int base = idx + 1; base < n - 1; base += 32) {
            int i = base + lane;
            uint32_t v = (i < n - 1) ? sym[i + 1] : 0u;
            __syncwarp();
            if (i < n - 1) sym[i] = v;
            __syncwarp();
        }
        if (lane == 0) sym[idx] = new_id;
        __syncwarp();
        --n;
    }
    return n;
}

#endif

}  // namespace ctk
