// Rich `Encoding` outputs for a whole batch on the GPU.
//
// Reference (paths under /root/reference/src):
//   huggingface/mod.rs:357-395   encode_to_encoding_impl: single (+pair) -> post-processor -> masks -> mark specials -> truncate
//   huggingface/mod.rs:397-444   encode_single_to_encoding: NFC, pre-tokenise, BPE per word (NO added-token scan), offsets, word ids
//   huggingface/mod.rs:447-478   pre_tokenize_with_offsets: str::find of the byte-mapped word from a running position
//   encoding.rs:44-267           from_ids / mark_special_tokens / pad / truncate / merge
//   postprocessors.rs:34-188     process(ids, None)
//   bindings/tokenizer.rs:33-201 __call__: truncate + pad of the batch
//
// Pipeline (everything stays in HBM; the ids come from the same fused encode kernel as encode_batch):
//   ids            nfc_stage -> prefix_space_stage -> encode_fused (added-token matching switched off: mod.rs:407 calls bpe.encode on
//                  whole words) ; add_special_tokens = 0 uses encode_device as encode_batch does (bindings/tokenizer.rs:88-96)
//   words          starts_bitmap (1 bit per byte) -> scan -> k_rank_list (rank per 32-byte word + sorted list of word starts)
//   token -> word  one scan over the tokens of {decoded byte length, vocabulary-string length}: a token's byte position in the
//                  normalised text is the prefix sum of the lengths before it (every byte has a symbol, merges concatenate), its
//                  word is the rank of that position in the bitmap.  k_tok_check verifies the alignment at every text boundary.
//   word spans     k_word_spans, one warp per text: the reference's sequential find chain, searched with 32 candidate positions
//                  per step (a word that holds a byte-mapped two-byte char cannot occur in a text without bytes C2..C5: no search)
//   offsets        k_tok_first / k_tok_offsets: start = min(word_start + strlen of the word's earlier tokens, word_end)
//   rows           k_row_len -> k_row_final -> scan -> k_rows_fill: template, masks, type ids, truncation, padding
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <cstring>

#include "device_common.cuh"
#include "engine.hpp"

namespace ctk {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;

// ---- words -------------------------------------------------------------------------------------------------------
// one thread per 32-byte word of the start bitmap: rank of the word's first byte + the sorted list of word starts
__global__ void __launch_bounds__(256) k_rank_list(const uint32_t* __restrict__ start_bits, uint64_t n_words32,
                                                   const uint32_t* __restrict__ block_base, uint32_t* __restrict__ wrank,
                                                   uint32_t* __restrict__ starts) {
    const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t bits = w < n_words32 ? start_bits[w] : 0u;
    const int c = __popc(bits), lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
    __shared__ int s[8];
    if (lane == 31) s[wid] = incl;
    __syncthreads();
    int wbase = 0;
    for (int k = 0; k < wid; ++k) wbase += s[k];
    uint32_t o = block_base[blockIdx.x] + (uint32_t)(wbase + incl - c);
    if (w < n_words32) wrank[w] = o;
    while (bits) {
        const int k = __ffs(bits) - 1;
        bits &= bits - 1;
        starts[o++] = (uint32_t)(w * 32 + k);
    }
}

struct WordIndex {
    const uint32_t* bits;
    const uint32_t* wrank;
    uint64_t n_bytes;
    uint32_t n_words;
    __device__ uint32_t before(uint64_t p) const {        // words that start before byte p
        if (p >= n_bytes) return n_words;
        return wrank[p >> 5] + __popc(bits[p >> 5] & ((1u << (p & 31)) - 1u));
    }
    __device__ bool starts_at(uint64_t p) const { return p < n_bytes && ((bits[p >> 5] >> (p & 31)) & 1u); }
};

// ---- token -> position -------------------------------------------------------------------------------------------
struct TokLen {
    const uint32_t* info;
    uint32_t n_ids;
    __host__ __device__ uint64_t operator()(uint32_t id) const {
        const uint32_t v = id < n_ids ? info[id] : 0u;
        return (uint64_t)(v & 0xFFFFu) | ((uint64_t)(v >> 16) << 32);
    }
};

// pos[i] = {bytes, string bytes} of the tokens before token i (exclusive scan over the whole batch).  The texts are contiguous and
// tokens decode to exactly their bytes, so the low half IS the token's byte position in the normalised batch: checked here at every
// text boundary (it fails only for a vocabulary without all 256 byte symbols, or with two tokens under one id).
__global__ void k_tok_check(const uint64_t* __restrict__ pos, const uint32_t* __restrict__ ids, TokLen tl, const uint64_t* __restrict__ tok_off,
                            const uint64_t* __restrict__ n_off, uint64_t n_docs, uint64_t n_tokens, uint64_t n_bytes,
                            uint32_t* __restrict__ err) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    if (d == n_docs) {
        if (n_tokens) {
            const uint64_t end = (uint32_t)pos[n_tokens - 1] + (uint32_t)tl(ids[n_tokens - 1]);
            if (end != n_bytes) atomicOr(err, ERRF_ALIGN);
        } else if (n_bytes) atomicOr(err, ERRF_ALIGN);
        return;
    }
    const uint64_t t = tok_off[d];
    if (t < n_tokens && (uint32_t)pos[t] != (uint32_t)n_off[d] && tok_off[d + 1] > t) atomicOr(err, ERRF_ALIGN);
}

// ---- word spans: the find chain of mod.rs:447-478 ------------------------------------------------------------------
// Per word, data-parallel: {length, leading U+0020 count, bytes outside 0x21..0x7E (each becomes a two-byte mapped char),
// first byte of the mapped string that will be searched for}.  One lane per word; words longer than 64 bytes are
// scanned by the whole warp.
__global__ void __launch_bounds__(256) k_word_meta(const uint8_t* __restrict__ N, uint64_t n_bytes, const uint32_t* __restrict__ starts,
                                                   uint32_t n_words, const uint16_t* __restrict__ map2, uint4* __restrict__ meta) {
    const uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = k < n_words;
    const uint32_t ws = valid ? starts[k] : 0u;
    const uint32_t we = valid ? (k + 1 < n_words ? starts[k + 1] : (uint32_t)n_bytes) : 0u;
    const uint32_t wlen = we - ws;
    uint32_t nsp = 0, np = 0;
    constexpr uint32_t kShort = 64;
    if (valid && wlen <= kShort) {
        bool lead = true;
        for (uint32_t i = 0; i < wlen; ++i) {
            const uint32_t b = N[ws + i];
            lead = lead && b == 0x20u;
            nsp += lead ? 1u : 0u;
            np += (b < 0x21u || b > 0x7Eu) ? 1u : 0u;
        }
    }
    unsigned longm = __ballot_sync(kFull, valid && wlen > kShort);
    while (longm) {
        const int src = __ffs(longm) - 1;
        longm &= longm - 1;
        const uint32_t lws = __shfl_sync(kFull, ws, src), lwlen = __shfl_sync(kFull, wlen, src);
        uint32_t tn = 0, tp = 0;
        bool lead = true;
        for (uint32_t base = 0; base < lwlen; base += 32) {
            const uint32_t i = base + lane;
            const uint32_t b = i < lwlen ? N[lws + i] : (uint32_t)'a';         // padding lanes: printable, not a space
            const unsigned sp = __ballot_sync(kFull, b == 0x20u);
            tp += __popc(__ballot_sync(kFull, b < 0x21u || b > 0x7Eu));
            if (lead) {
                const uint32_t t = ~sp ? (uint32_t)(__ffs(~sp) - 1) : 32u;
                tn += t;
                lead = t == 32u;
            }
        }
        if (lane == src) { nsp = tn; np = tp; }
    }
    if (valid) {
        const bool all_space = nsp == wlen;                                      // trim_start_matches('Ġ') left nothing: the word itself is searched
        const uint32_t ts = all_space ? 0u : nsp;
        const uint32_t np_trim = all_space ? np : np - nsp;
        // first two bytes of the needle (the mapped string of the word from ts on); second byte 0 when the needle has one byte
        const uint32_t v0 = map2[N[ws + ts]];
        uint32_t b1 = v0 >> 8;
        if (!b1 && wlen - ts > 1) b1 = map2[N[ws + ts + 1]] & 0xFFu;
        meta[k] = make_uint4(wlen, nsp, np, (v0 & 0xFFu) | (b1 << 8));
    }
}

// has_hi[d] = text d holds a byte C2..C5 (the only lead bytes byte-mapped two-byte chars have).  One warp per text.
__global__ void __launch_bounds__(256) k_doc_has_hi(const uint8_t* __restrict__ O, const uint64_t* __restrict__ o_off, uint64_t n_docs,
                                                    uint8_t* __restrict__ has_hi) {
    const uint64_t d = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (d >= n_docs) return;
    const int lane = threadIdx.x & 31;
    const uint64_t lo = o_off[d], hi = o_off[d + 1];
    bool hh = false;
    for (uint64_t i = lo + lane; i < hi; i += 32) hh |= (uint32_t)(O[i] - 0xC2u) <= 3u;
    hh = __any_sync(kFull, hh);
    if (lane == 0) has_hi[d] = hh ? 1 : 0;
}

struct SpanArgs {
    const uint8_t* O; const uint64_t* o_off;         // the original batch (what str::find searches)
    const uint8_t* N; const uint64_t* n_off;         // the normalised batch (what the words are cut from)
    uint64_t n_docs;
    WordIndex wi;
    const uint32_t* starts;
    const uint4* meta;
    const uint16_t* map2;
    const uint8_t* has_hi;
    uint4* wspan;                                    // {start, end, first word of the text, -}
    uint32_t* err;                                   // err[0] flags, err[8] first panicking text
    int same;                                        // N is O (nothing was normalised or prepended)
};

// One warp per text walks its words in order (the position each search starts from is the previous result).
//  * on the rails: the text is its own normalised form, the word is printable ASCII after its leading spaces and the
//    running position lies inside those spaces: str::find returns the word's own position (the bytes in between are
//    spaces, the needle does not start with one), and the next word is entered exactly at its start.  A RUN of such words
//    is settled 32 at a time, one lane per word, without reading a byte of the text.
//  * saturated: once the position has reached the end of the text nothing can be found any more: the remaining words all get
//    the empty span at the end, 32 at a time.
//  * a needle with a two-byte mapped char cannot occur in a text that has no byte C2..C5: not found, no search.
//  * otherwise the warp searches: 16 candidate positions per lane and step.
__global__ void __launch_bounds__(128) k_word_spans(SpanArgs a) {
    const uint64_t d = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (d >= a.n_docs) return;
    const int lane = threadIdx.x & 31;
    const uint64_t nb = a.n_off[d], ne = a.n_off[d + 1];
    if (ne == nb) return;
    const uint8_t* __restrict__ od = a.O + a.o_off[d];
    const uint32_t olen = (uint32_t)(a.o_off[d + 1] - a.o_off[d]);
    const uint32_t k0 = a.wi.before(nb), k1 = a.wi.before(ne);
    const bool has_hi = a.has_hi[d] != 0;
    uint32_t ss = 0;
    bool boundary = true;                                                        // ss is known to be a char boundary of the text
    uint32_t k = k0;
    while (k < k1) {
        if (ss == olen) {                                                        // saturated (mod.rs:470-474 with nothing left to search)
            for (uint32_t kk = k + lane; kk < k1; kk += 32) a.wspan[kk] = make_uint4(olen, olen, k0, 0u);
            break;
        }
        // lane j looks at word k + j
        const uint32_t kj = k + lane;
        const bool valid = kj < k1;
        const uint4 mj = valid ? a.meta[kj] : make_uint4(1u, 0u, 1u, 0u);
        const uint32_t wsj = valid ? a.starts[kj] : 0u;
        if (a.same) {
            const bool all_sp = mj.y == mj.x;
            const bool pure_j = valid && !all_sp && mj.z == mj.y;                // every byte after the leading spaces is printable ASCII
            const uint64_t rel0 = (uint64_t)__shfl_sync(kFull, wsj, 0) - nb;
            const uint32_t ts0 = __shfl_sync(kFull, mj.y, 0);
            const unsigned pm = __ballot_sync(kFull, pure_j);
            const int run = (ss >= rel0 && ss <= rel0 + ts0) ? (~pm ? __ffs(~pm) - 1 : 32) : 0;
            if (run > 0) {
                const uint32_t relj = (uint32_t)(wsj - nb);
                if (lane < run) a.wspan[kj] = make_uint4(relj + mj.y, relj + mj.x, k0, 0u);
                ss = __shfl_sync(kFull, relj + mj.x, run - 1);
                boundary = true;
                k += run;
                continue;
            }
        }
        // one word, sequentially: word k (lane 0's)
        const uint4 me = make_uint4(__shfl_sync(kFull, mj.x, 0), __shfl_sync(kFull, mj.y, 0), __shfl_sync(kFull, mj.z, 0), __shfl_sync(kFull, mj.w, 0));
        const uint64_t ws = __shfl_sync(kFull, wsj, 0);
        const uint32_t wlen = me.x, nsp = me.y, np = me.z, fb = me.w & 0xFFu, sb = me.w >> 8;
        const bool all_space = nsp == wlen;
        const uint32_t ts = all_space ? 0u : nsp;
        const uint32_t np_trim = all_space ? np : np - nsp;
        const uint32_t tl = wlen - ts;
        const uint64_t m = (uint64_t)tl + np_trim;                               // bytes of the mapped string searched for
        const uint64_t mword = (uint64_t)wlen + np;                              // word.len()
        const bool pure = np_trim == 0;
        const uint64_t rel = ws - nb;                                            // the word's own position when N is O
        int64_t found = -1;
        if (a.same && pure && ss >= rel && ss <= rel + ts) {
            found = (int64_t)(rel + ts);
        } else {
            if (!boundary && ss < olen && (od[ss] & 0xC0u) == 0x80u) {           // original[search_start..] panics
                if (lane == 0) { atomicOr(a.err, ERRF_PANIC); atomicMin(a.err + 8, (uint32_t)(d < 0xFFFFFFFFull ? d : 0xFFFFFFFFull)); }
                return;
            }
            if (m <= (uint64_t)(olen - ss) && (pure || has_hi)) {
                // 16 candidate positions per lane and step: one aligned 16-byte load (+4 bytes for the second-byte test),
                // byte-wise compares of the needle's first two bytes, survivors verified in ascending order
                const uint8_t* __restrict__ tp = a.N + ws + ts;
                const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(od) & 15u);
                const uint8_t* __restrict__ ab = od - mis;                       // aligned; the bytes before od belong to the same buffer
                const int64_t last = (int64_t)olen - (int64_t)m;                 // last position a match can start at
                const uint32_t f4 = fb * 0x01010101u, s4 = sb * 0x01010101u;
                for (int64_t c = (int64_t)((mis + ss) >> 4) + lane;; c += 32) {
                    const int64_t cs = c * 16 - (int64_t)mis;                    // position of the chunk's first byte in the text
                    const int64_t cs0 = cs - 16 * lane;                          // lane 0's chunk: uniform loop exit
                    if (cs0 > last) break;
                    uint32_t cand = 0;
                    if (cs <= last) {
                        const uint4 v = *reinterpret_cast<const uint4*>(ab + c * 16);
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                        uint32_t e1 = 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) e1 |= (((__vcmpeq4(w[q], f4) & 0x01010101u) * 0x10204080u) >> 28) << (4 * q);
                        cand = e1;
                        if (m > 1) {
                            const uint32_t nx = *reinterpret_cast<const uint32_t*>(ab + c * 16 + 16);
                            uint32_t e2 = 0;
#pragma unroll
                            for (int q = 0; q < 4; ++q) e2 |= (((__vcmpeq4(w[q], s4) & 0x01010101u) * 0x10204080u) >> 28) << (4 * q);
                            e2 |= (((__vcmpeq4(nx, s4) & 0x01010101u) * 0x10204080u) >> 28) << 16;
                            cand &= e2 >> 1;
                        }
                        const int64_t lo = (int64_t)ss - cs, hi = last - cs;     // valid bit range [lo, hi]
                        if (lo > 0) cand &= lo >= 16 ? 0u : ~((1u << lo) - 1u);
                        if (hi < 15) cand &= (2u << hi) - 1u;
                    }
                    int64_t hit = -1;
                    while (cand) {
                        const int j = __ffs(cand) - 1;
                        cand &= cand - 1;
                        const uint64_t p = (uint64_t)(cs + j);
                        bool ok = true;
                        if (pure) {
                            for (uint32_t i = 2; i < tl; ++i) if (od[p + i] != tp[i]) { ok = false; break; }
                        } else {
                            uint64_t q = p;
                            for (uint32_t i = 0; i < tl; ++i) {
                                const uint32_t mv = a.map2[tp[i]];
                                if (od[q] != (mv & 0xFFu)) { ok = false; break; }
                                ++q;
                                if (mv >> 8) { if (od[q] != (mv >> 8)) { ok = false; break; } ++q; }
                            }
                        }
                        if (ok) { hit = (int64_t)p; break; }
                    }
                    const unsigned mask = __ballot_sync(kFull, hit >= 0);
                    if (mask) { found = __shfl_sync(kFull, hit, __ffs(mask) - 1); break; }
                }
            }
        }
        uint32_t start, end;
        if (found >= 0) { start = (uint32_t)found; end = (uint32_t)(found + (int64_t)m); boundary = true; }
        else { start = ss; const uint64_t e = (uint64_t)ss + mword; end = e < olen ? (uint32_t)e : olen; boundary = end == olen; }
        if (lane == 0) a.wspan[k] = make_uint4(start, end, k0, 0u);
        ss = end;
        ++k;
    }
}

// ---- token offsets ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tok_first(const uint64_t* __restrict__ pos, uint64_t n_tokens, WordIndex wi,
                                                   uint32_t* __restrict__ wfirst) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_tokens) return;
    const uint64_t v = pos[i];
    const uint64_t g = (uint32_t)v;
    if (wi.starts_at(g)) wfirst[wi.before(g)] = (uint32_t)(v >> 32);
}

__global__ void __launch_bounds__(256) k_tok_offsets(const uint64_t* __restrict__ pos, const uint32_t* __restrict__ ids, TokLen tl,
                                                     uint64_t n_tokens, WordIndex wi, const uint32_t* __restrict__ wfirst,
                                                     const uint4* __restrict__ wspan,
                                                     uint2* __restrict__ offsets, uint32_t* __restrict__ word_ids) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_tokens) return;
    const uint64_t v = pos[i];
    const uint64_t g = (uint32_t)v;
    const uint32_t w = wi.before(g + 1) - 1u;                                   // the word that holds byte g
    const uint4 sp = wspan[w];
    const uint64_t before = (uint32_t)((uint32_t)(v >> 32) - wfirst[w]);         // string bytes of the word's earlier tokens
    const uint64_t slen = (uint32_t)(tl(ids[i]) >> 32);
    const uint64_t s = (uint64_t)sp.x + before, e = s + slen;                    // mod.rs:421-428: end = min(off + len, word_end)
    offsets[i] = make_uint2((uint32_t)(s < sp.y ? s : sp.y), (uint32_t)(e < sp.y ? e : sp.y));
    word_ids[i] = w - sp.z;
}

// ---- rows ----------------------------------------------------------------------------------------------------------------
constexpr int kMaxItems = 16;
struct RowArgs {
    const uint64_t* tok_off;
    const uint32_t* raw_ids;
    uint64_t n_rows;
    int g;                                  // texts per row
    int n_items;
    int64_t items[kMaxItems];               // -1 = the ids, else a literal id
    uint64_t n_lit, n_a;
    int truncation; uint64_t max_length;
    int padding; uint64_t pad_to;
    int pad_left;
    uint32_t pad_id;
    int mark;                               // encode_to_encoding path: added positions and special ids get special_tokens_mask = 1
    const uint32_t* special_bits; uint32_t n_special_words;
};

__global__ void __launch_bounds__(256) k_row_len(RowArgs a, uint64_t* __restrict__ cut, uint64_t* __restrict__ full,
                                                 unsigned long long* __restrict__ maxcut) {
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    unsigned long long c = 0;
    if (r < a.n_rows) {
        const uint64_t len = a.tok_off[(r + 1) * a.g] - a.tok_off[r * a.g];
        const uint64_t f = a.n_lit + a.n_a * len;
        c = (a.truncation && f > a.max_length) ? a.max_length : f;
        cut[r] = c;
        full[r] = f;
    }
    for (int o = 16; o; o >>= 1) { const unsigned long long v = __shfl_xor_sync(kFull, c, o); c = v > c ? v : c; }
    if ((threadIdx.x & 31) == 0 && c) atomicMax(maxcut, c);
}

__global__ void __launch_bounds__(256) k_row_final(const uint64_t* __restrict__ cut, uint64_t n_rows, int padding, uint64_t pad_to,
                                                   const unsigned long long* __restrict__ maxcut, uint64_t* __restrict__ fin) {
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    if (r == n_rows) { fin[r] = 0; return; }
    const uint64_t target = padding == 1 ? (uint64_t)*maxcut : padding == 2 ? pad_to : 0ull;
    const uint64_t c = cut[r];
    fin[r] = c < target ? target : c;                                            // encoding.rs:87: pad only when shorter
}

__global__ void __launch_bounds__(256) k_rows_fill(RowArgs a, const uint64_t* __restrict__ cut, const uint64_t* __restrict__ row_off,
                                                   uint32_t* __restrict__ out_ids, uint8_t* __restrict__ out_attn,
                                                   uint8_t* __restrict__ out_type, uint8_t* __restrict__ out_spec) {
    for (uint64_t r = blockIdx.x; r < a.n_rows; r += gridDim.x) {
        const uint64_t a0 = a.tok_off[r * a.g], a1 = a.tok_off[(r + 1) * a.g];
        const uint64_t len = a1 - a0;
        const uint64_t len_a = a.g == 2 ? a.tok_off[r * a.g + 1] - a0 : len;
        const uint64_t base = row_off[r], fin = row_off[r + 1] - base, c = cut[r];
        const uint64_t npad = fin - c;
        for (uint64_t j = threadIdx.x; j < fin; j += blockDim.x) {
            uint64_t jj = j;
            bool is_pad;
            if (a.pad_left) { is_pad = j < npad; jj = j - npad; } else is_pad = j >= c;
            uint32_t id = a.pad_id;
            uint8_t attn = 0, ty = 0, sp = 1;                                    // encoding.rs:97-127
            if (!is_pad) {
                uint64_t p = 0;
                id = 0;
                for (int k = 0; k < a.n_items; ++k) {
                    if (a.items[k] < 0) {
                        if (jj >= p && jj < p + len) { id = a.raw_ids[a0 + (jj - p)]; break; }
                        p += len;
                    } else {
                        if (jj == p) { id = (uint32_t)a.items[k]; break; }
                        p += 1;
                    }
                }
                attn = 1;
                ty = (jj >= len_a && jj < len) ? 1 : 0;                          // merge(): positional, then zeros for what was added
                sp = 0;
                if (a.mark) {                                                   // mod.rs:380-386
                    const uint32_t wd = id >> 5;
                    sp = (jj >= len || (wd < a.n_special_words && ((a.special_bits[wd] >> (id & 31)) & 1u))) ? 1 : 0;
                }
            }
            out_ids[base + j] = id;
            out_attn[base + j] = attn;
            out_type[base + j] = ty;
            out_spec[base + j] = sp;
        }
    }
}

}  // namespace

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return eng.cuda_fail(e_, #x); } while (0)

// slots 48.. of the workspace belong to this file
enum { S_RAW = 48, S_TOKOFF, S_DS, S_SB, S_BC, S_BB, S_ERR, S_WRANK, S_STARTS, S_POS, S_CUB, S_WSPAN, S_WDOC0, S_WFIRST, S_OFFS, S_WIDS,
       S_CUT, S_FIN, S_ROWOFF, S_FULL, S_OIDS, S_OATTN, S_OTYPE, S_OSPEC, S_HASHI };

int encode_rich_device(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_texts, uint64_t n_bytes,
                       const ctk_encoding_options& opt, RichOut* out, cudaStream_t st) {
    *out = RichOut{};
    if (opt.pair && (n_texts & 1)) return eng.fail(CTK_ERR_ARG, "pair mode needs an even number of texts (text, text_pair, text, ...)");
    if (n_bytes >= 0xFFFFFFF0ull) return eng.fail(CTK_ERR_ARG, "one call handles less than 4 GiB of text");
    const HostModel& m = eng.model;
    if (eng.model.metaspace) return eng.fail(CTK_ERR_UNSUPPORTED, "Encoding outputs are not built for Metaspace pipelines; encode / encode_batch / decode are");
    if (!eng.split_dev.empty()) return eng.fail(CTK_ERR_UNSUPPORTED, "Encoding outputs (offsets, word ids, padded rows) are not built for tokenizers with Split stages; encode / encode_batch are");
    RowArgs ra{};
    ra.mark = opt.add_special_tokens ? 1 : 0;
    if (opt.add_special_tokens) {
        if ((int)m.pp_items.size() > kMaxItems) return eng.fail(CTK_ERR_UNSUPPORTED, "post-processor template with more than 16 items");
        for (int64_t it : m.pp_items) { ra.items[ra.n_items++] = it; if (it < 0) ++ra.n_a; else ++ra.n_lit; }
        if (ra.n_a == 0) return eng.fail(CTK_ERR_PANIC, "post-processor template without $A: the reference underflows at mod.rs:377");
    } else {
        ra.items[0] = -1; ra.n_items = 1; ra.n_a = 1;
    }
    Workspace& ws = eng.ws;
    uint32_t* err;
    CK(ws.get(S_ERR, 256, (void**)&err));
    CK(cudaMemsetAsync(err, 0, 256, st));
    CK(cudaMemsetAsync(err + 8, 0xFF, 4, st));

    // ---- 1. ids ------------------------------------------------------------------------------------------------
    uint32_t* raw;
    uint64_t* tok_off;
    const uint64_t ids_cap = n_bytes + 64;
    CK(ws.get(S_RAW, ids_cap * 4, (void**)&raw));
    CK(ws.get(S_TOKOFF, (n_texts + 2) * 8, (void**)&tok_off));
    const uint8_t* nt = d_text; const uint64_t* noff = d_off; uint64_t nbytes = n_bytes;
    uint64_t n_tokens = 0;
    int rc;
    if (opt.add_special_tokens) {
        rc = nfc_stage(eng, d_text, d_off, n_texts, n_bytes, &nt, &noff, &nbytes, st);
        if (rc != CTK_OK) return rc;
        rc = prefix_space_stage(eng, nt, noff, n_texts, nbytes, &nt, &noff, &nbytes, st);
        if (rc != CTK_OK) return rc;
        const uint32_t saved = eng.tables.n_added;                               // mod.rs:407: bpe.encode(word), nothing else
        if (saved) { eng.tables.n_added = 0; eng.cache_valid = false; }
        rc = encode_fused(eng, nt, noff, n_texts, nbytes, raw, ids_cap, tok_off, &n_tokens, st);
        if (saved) { eng.tables.n_added = saved; eng.cache_valid = false; }
    } else {
        rc = encode_device(eng, d_text, d_off, n_texts, n_bytes, raw, ids_cap, tok_off, &n_tokens, st);
    }
    if (rc != CTK_OK) return rc;
    out->n_tokens = n_tokens; out->tok_off = tok_off; out->raw_ids = raw;

    // ---- 2. words, offsets -------------------------------------------------------------------------------------
    const bool want = opt.want_offsets && opt.add_special_tokens;
    if (want && n_tokens) {
        for (int b = 0; b < 256; ++b) if (m.byte_init_id[b] == kNoId)
            return eng.fail(CTK_ERR_UNSUPPORTED, "offsets need a vocabulary with all 256 byte symbols (tokens must cover the text)");
        const uint64_t n_w32 = (nbytes + 31) / 32;
        const uint32_t n_blocks = (uint32_t)((n_w32 + 255) / 256);
        uint32_t *ds, *sb, *bc, *bb;
        CK(ws.get(S_DS, (n_w32 + 1) * 4, (void**)&ds));
        CK(ws.get(S_SB, (n_w32 + 1) * 4, (void**)&sb));
        CK(ws.get(S_BC, ((uint64_t)n_blocks + 1) * 4, (void**)&bc));
        CK(ws.get(S_BB, ((uint64_t)n_blocks + 1) * 4, (void**)&bb));
        eng.mark(nullptr, st);
        rc = starts_bitmap(eng, nt, noff, n_texts, nbytes, ds, sb, bc, err, st);
        if (rc != CTK_OK) return rc;
        eng.mark("rich:starts_bitmap", st);
        size_t cub_bytes = 0, cub2 = 0;
        void* cub_tmp;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, bc, bb, n_blocks + 1, st));
        TokLen tl{eng.rich.tok_info, eng.dec.n_ids};
        cub::TransformInputIterator<uint64_t, TokLen, const uint32_t*> lens(raw, tl);
        uint64_t* pos;
        CK(ws.get(S_POS, (n_tokens + 1) * 8, (void**)&pos));
        CK(cub::DeviceScan::ExclusiveSum(nullptr, cub2, lens, pos, n_tokens, st));
        if (cub2 > cub_bytes) cub_bytes = cub2;
        CK(ws.get(S_CUB, cub_bytes + 16, &cub_tmp));
        CK(cudaMemsetAsync(bc + n_blocks, 0, 4, st));
        CK(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, bc, bb, n_blocks + 1, st));
        uint32_t n_words = 0;
        CK(cudaMemcpyAsync(&n_words, bb + n_blocks, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        uint32_t *wrank, *starts, *wfirst, *word_ids;
        uint4 *wspan, *wmeta;
        uint2* offs;
        CK(ws.get(S_WRANK, (n_w32 + 1) * 4, (void**)&wrank));
        CK(ws.get(S_STARTS, ((uint64_t)n_words + 2) * 4, (void**)&starts));
        CK(ws.get(S_WSPAN, ((uint64_t)n_words + 1) * 16, (void**)&wspan));
        CK(ws.get(S_WDOC0, ((uint64_t)n_words + 1) * 16, (void**)&wmeta));
        CK(ws.get(S_WFIRST, ((uint64_t)n_words + 1) * 4, (void**)&wfirst));
        CK(ws.get(S_OFFS, (n_tokens + 1) * 8, (void**)&offs));
        CK(ws.get(S_WIDS, (n_tokens + 1) * 4, (void**)&word_ids));
        eng.mark(nullptr, st);
        k_rank_list<<<n_blocks, 256, 0, st>>>(sb, n_w32, bb, wrank, starts);
        eng.mark("rich:k_rank_list", st);
        CK(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, lens, pos, n_tokens, st));
        eng.mark("rich:scan_token_lengths", st);
        k_tok_check<<<(unsigned)((n_texts + 1 + 255) / 256), 256, 0, st>>>(pos, raw, tl, tok_off, noff, n_texts, n_tokens, nbytes, err);
        WordIndex wi{sb, wrank, nbytes, n_words};
        k_word_meta<<<(n_words + 255) / 256, 256, 0, st>>>(nt, nbytes, starts, n_words, eng.rich.byte_map2, wmeta);
        eng.mark("rich:k_word_meta", st);
        uint8_t* has_hi;
        CK(ws.get(S_HASHI, n_texts + 16, (void**)&has_hi));
        k_doc_has_hi<<<(unsigned)((n_texts * 32 + 255) / 256), 256, 0, st>>>(d_text, d_off, n_texts, has_hi);
        SpanArgs sa{d_text, d_off, nt, noff, n_texts, wi, starts, wmeta, eng.rich.byte_map2, has_hi, wspan, err, nt == d_text ? 1 : 0};
        k_word_spans<<<(unsigned)((n_texts * 32 + 127) / 128), 128, 0, st>>>(sa);
        eng.mark("rich:k_word_spans", st);
        const unsigned tg = (unsigned)((n_tokens + 255) / 256);
        k_tok_first<<<tg, 256, 0, st>>>(pos, n_tokens, wi, wfirst);
        k_tok_offsets<<<tg, 256, 0, st>>>(pos, raw, tl, n_tokens, wi, wfirst, wspan, offs, word_ids);
        eng.mark("rich:k_tok_offsets", st);
        eng.launched(9);
        CK(cudaGetLastError());
        out->offsets = offs; out->word_ids = word_ids;
    }

    // ---- 3. rows ----------------------------------------------------------------------------------------------
    const uint64_t n_rows = opt.pair ? n_texts / 2 : n_texts;
    ra.tok_off = tok_off; ra.raw_ids = raw; ra.n_rows = n_rows; ra.g = opt.pair ? 2 : 1;
    ra.truncation = opt.truncation; ra.max_length = opt.max_length;
    ra.padding = opt.padding; ra.pad_to = opt.pad_to; ra.pad_left = opt.pad_left; ra.pad_id = m.pad_id;
    ra.special_bits = eng.rich.special_bits; ra.n_special_words = eng.rich.n_special_words;
    uint64_t *cut, *fin, *row_off, *full;
    CK(ws.get(S_CUT, (n_rows + 2) * 8, (void**)&cut));
    CK(ws.get(S_FIN, (n_rows + 2) * 8, (void**)&fin));
    CK(ws.get(S_ROWOFF, (n_rows + 2) * 8, (void**)&row_off));
    CK(ws.get(S_FULL, (n_rows + 2) * 8, (void**)&full));
    unsigned long long* maxcut = reinterpret_cast<unsigned long long*>(err + 16);
    const unsigned rg = (unsigned)((n_rows + 1 + 255) / 256);
    eng.mark(nullptr, st);
    k_row_len<<<rg, 256, 0, st>>>(ra, cut, full, maxcut);
    k_row_final<<<rg, 256, 0, st>>>(cut, n_rows, opt.padding, opt.pad_to, maxcut, fin);
    {
        size_t cub_bytes = 0;
        void* cub_tmp;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, fin, row_off, n_rows + 1, st));
        CK(ws.get(S_CUB, cub_bytes + 16, &cub_tmp));
        CK(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, fin, row_off, n_rows + 1, st));
    }
    eng.launched(3);
    CK(eng.publish({{err, 1, 0}, {err + 8, 1, 1}, {row_off + n_rows, 2, 2}, {err + 16, 2, 4}}, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t f = eng.h_flags[0];
    if (f & ERRF_PANIC) {
        char msg[160];
        snprintf(msg, sizeof msg, "the reference panics on text %u: byte index is not a char boundary (mod.rs:461, original[search_start..])", eng.h_flags[1]);
        return eng.fail(CTK_ERR_PANIC, msg);
    }
    if (f & ERRF_ALIGN) return eng.fail(CTK_ERR_UNSUPPORTED, "tokens do not cover the text byte for byte (two tokens under one id?): offsets unavailable");
    if (f & ERRF_OFFSETS) return eng.fail(CTK_ERR_ARG, "offsets must start at 0, be non-decreasing and end at the buffer length");
    uint64_t total = 0, mx = 0;
    memcpy(&total, eng.h_flags + 2, 8);
    memcpy(&mx, eng.h_flags + 4, 8);
    uint32_t* oids; uint8_t *oattn, *otype, *ospec;
    CK(ws.get(S_OIDS, (total + 16) * 4, (void**)&oids));
    CK(ws.get(S_OATTN, total + 16, (void**)&oattn));
    CK(ws.get(S_OTYPE, total + 16, (void**)&otype));
    CK(ws.get(S_OSPEC, total + 16, (void**)&ospec));
    if (n_rows && total) {
        const unsigned grid = (unsigned)(n_rows < (1u << 20) ? n_rows : (1u << 20));
        k_rows_fill<<<grid, 256, 0, st>>>(ra, cut, row_off, oids, oattn, otype, ospec);
        eng.launched(1);
        CK(cudaGetLastError());
    }
    eng.mark("rich:rows", st);
    CK(cudaStreamSynchronize(st));
    eng.collect_marks();
    out->n_rows = n_rows; out->total = total; out->max_row = mx; out->row_off = row_off; out->row_full = full;
    out->ids = oids; out->attention = oattn; out->type_ids = otype; out->special = ospec;
    return CTK_OK;
}

// ---- host-buffer entry point -----------------------------------------------------------------------------------------
struct Encodings {                          // ONE page-locked block from the pool, carved into the arrays
    size_t n_rows = 0, n_texts = 0;
    void* block = nullptr; size_t cap = 0;
    const void *row_off = nullptr, *row_full = nullptr, *ids = nullptr, *attn = nullptr, *type = nullptr, *spec = nullptr,
               *tok_off = nullptr, *raw = nullptr, *offs = nullptr, *wids = nullptr;
    ~Encodings() { pinned_put(block, cap); }
};

}  // namespace ctk

using namespace ctk;

extern "C" {

int ctk_encode_batch_to_encoding(const ctk_tokenizer* tok, const uint8_t* text, const uint64_t* text_off, size_t n_texts,
                                 const ctk_encoding_options* opt, ctk_encodings** res) {
    if (!tok || !text_off || !opt || !res) { set_last_error("NULL argument"); return CTK_ERR_ARG; }
    *res = nullptr;
    Engine& eng = *const_cast<Engine*>(reinterpret_cast<const Engine*>(tok));
    std::lock_guard<std::mutex> lk(eng.mu);
    CK(cudaSetDevice(eng.device));
    const uint64_t n_bytes = text_off[n_texts];
    if (text_off[0] != 0) return eng.fail(CTK_ERR_ARG, "offsets must start at 0, be non-decreasing and end at the buffer length");
    if (n_bytes && !text) { set_last_error("NULL argument"); return CTK_ERR_ARG; }
    cudaStream_t st = eng.st_comp;
    uint8_t* d_text; uint64_t* d_off;
    CK(eng.ws.get(78, n_bytes + 128, (void**)&d_text));
    CK(eng.ws.get(79, (n_texts + 2) * 8, (void**)&d_off));
    if (n_bytes) CK(cudaMemcpyAsync(d_text, text, n_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_text + n_bytes, 0, 64, st));
    CK(cudaMemcpyAsync(d_off, text_off, (n_texts + 1) * 8, cudaMemcpyHostToDevice, st));
    RichOut o;
    int rc = encode_rich_device(eng, d_text, d_off, n_texts, n_bytes, *opt, &o, st);
    if (rc != CTK_OK) return rc;
    Encodings* e = new (std::nothrow) Encodings();
    if (!e) return eng.fail(CTK_ERR_CUDA, "out of memory");
    e->n_rows = o.n_rows; e->n_texts = n_texts;
    struct Part { const void** dst; const void* src; size_t bytes; };
    const Part parts[] = {
        {&e->row_off, o.row_off, (o.n_rows + 1) * 8}, {&e->row_full, o.row_full, o.n_rows * 8}, {&e->tok_off, o.tok_off, (n_texts + 1) * 8},
        {&e->ids, o.ids, o.total * 4}, {&e->raw, o.raw_ids, o.n_tokens * 4},
        {&e->offs, o.offsets, o.offsets ? o.n_tokens * 8 : 0}, {&e->wids, o.word_ids, o.word_ids ? o.n_tokens * 4 : 0},
        {&e->attn, o.attention, o.total}, {&e->type, o.type_ids, o.total}, {&e->spec, o.special, o.total}};
    size_t need = 0;
    for (const Part& p : parts) need += (p.bytes + 63) & ~(size_t)63;
    cudaError_t ce = pinned_get(need + 64, &e->block, &e->cap);
    uint64_t d2h = 0;
    size_t at = 0;
    for (const Part& p : parts) {
        if (ce != cudaSuccess) break;
        if (!p.src) continue;                                                    // offsets / word ids not computed: accessor returns NULL
        *p.dst = static_cast<uint8_t*>(e->block) + at;
        if (p.bytes) ce = cudaMemcpyAsync(static_cast<uint8_t*>(e->block) + at, p.src, p.bytes, cudaMemcpyDeviceToHost, st);
        at += (p.bytes + 63) & ~(size_t)63;
        d2h += p.bytes;
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { delete e; return eng.cuda_fail(ce, "copy of the encodings to the host"); }
    eng.last_h2d_bytes = n_bytes + (n_texts + 1) * 8;
    eng.last_d2h_bytes = d2h;
    *res = reinterpret_cast<ctk_encodings*>(e);
    return CTK_OK;
}

#define ENC(res) reinterpret_cast<const Encodings*>(res)
size_t ctk_encodings_rows(const ctk_encodings* res) { return ENC(res)->n_rows; }
const uint64_t* ctk_encodings_row_offsets(const ctk_encodings* res) { return (const uint64_t*)ENC(res)->row_off; }
const uint64_t* ctk_encodings_row_full_lengths(const ctk_encodings* res) { return (const uint64_t*)ENC(res)->row_full; }
const uint32_t* ctk_encodings_input_ids(const ctk_encodings* res) { return (const uint32_t*)ENC(res)->ids; }
const uint8_t* ctk_encodings_attention_mask(const ctk_encodings* res) { return (const uint8_t*)ENC(res)->attn; }
const uint8_t* ctk_encodings_type_ids(const ctk_encodings* res) { return (const uint8_t*)ENC(res)->type; }
const uint8_t* ctk_encodings_special_tokens_mask(const ctk_encodings* res) { return (const uint8_t*)ENC(res)->spec; }
const uint64_t* ctk_encodings_token_offsets(const ctk_encodings* res) { return (const uint64_t*)ENC(res)->tok_off; }
const uint32_t* ctk_encodings_token_ids(const ctk_encodings* res) { return (const uint32_t*)ENC(res)->raw; }
const uint32_t* ctk_encodings_offsets(const ctk_encodings* res) { return (const uint32_t*)ENC(res)->offs; }
const uint32_t* ctk_encodings_word_ids(const ctk_encodings* res) { return (const uint32_t*)ENC(res)->wids; }
void ctk_encodings_free(ctk_encodings* res) { delete reinterpret_cast<Encodings*>(res); }

// debug/test hook: the pre-token start bitmap of a batch as the device computes it for the Encoding path (one bit per byte;
// out_bits: (n + 31) / 32 words).  CTK_SCALAR_STARTS=1 selects the scalar predicate instead of the bit-parallel kernel.
int ctk_debug_starts_device(const ctk_tokenizer* tok, const uint8_t* text, uint64_t n, const uint64_t* off, size_t n_docs, uint32_t* out_bits) {
    Engine& eng = *const_cast<Engine*>(reinterpret_cast<const Engine*>(tok));
    std::lock_guard<std::mutex> lk(eng.mu);
    CK(cudaSetDevice(eng.device));
    cudaStream_t st = eng.st_comp;
    const uint64_t n_w32 = (n + 31) / 32;
    const uint32_t n_blocks = (uint32_t)((n_w32 + 255) / 256);
    uint8_t* d_text; uint64_t* d_off; uint32_t *ds, *sb, *bc, *err;
    CK(eng.ws.get(78, n + 128, (void**)&d_text));
    CK(eng.ws.get(79, (n_docs + 2) * 8, (void**)&d_off));
    CK(eng.ws.get(S_DS, (n_w32 + 1) * 4, (void**)&ds));
    CK(eng.ws.get(S_SB, (n_w32 + 1) * 4, (void**)&sb));
    CK(eng.ws.get(S_BC, ((uint64_t)n_blocks + 1) * 4, (void**)&bc));
    CK(eng.ws.get(S_ERR, 256, (void**)&err));
    CK(cudaMemsetAsync(d_text, 0, n + 128, st));
    if (n) CK(cudaMemcpyAsync(d_text, text, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_off, off, (n_docs + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(err, 0, 256, st));
    int rc = n ? starts_bitmap(eng, d_text, d_off, n_docs, n, ds, sb, bc, err, st) : CTK_OK;
    if (rc != CTK_OK) return rc;
    if (n) CK(cudaMemcpyAsync(out_bits, sb, n_w32 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return CTK_OK;
}

size_t ctk_post_processor_items(const ctk_tokenizer* tok, int64_t* items, size_t cap) {
    const HostModel& m = reinterpret_cast<const Engine*>(tok)->model;
    for (size_t i = 0; i < m.pp_items.size() && i < cap; ++i) items[i] = m.pp_items[i];
    return m.pp_items.size();
}

uint32_t ctk_pad_token(const ctk_tokenizer* tok, const uint8_t** token, size_t* len) {
    const HostModel& m = reinterpret_cast<const Engine*>(tok)->model;
    if (token) *token = (const uint8_t*)m.pad_token.data();
    if (len) *len = m.pad_token.size();
    return m.pad_id;
}

}  // extern "C"
