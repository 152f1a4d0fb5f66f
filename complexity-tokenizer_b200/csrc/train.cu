// BPE training on the GPU.  Replaces BpeTrainer::train (src/bpe_trainer.rs:100-228) behind the
// C ABI ctk_train_bpe; the reference's two hash-order decisions are fixed as oracle/py_trainer.py states them
// (best pair: highest count, then smallest (left index, right index); characters of equal frequency: by code point).
//
// Stage W -- word histogram (bpe_trainer.rs:241-275), data-parallel over the text bytes, HBM-bound:
//   k_mark_breaks     one bit per text start (words do not span texts)
//   k_word_bits       thread per 32 bytes: White_Space chars -> bitmaps S (a word starts) and E (a word continues)
//   k_expand_starts   S -> list of word starts
//   k_word_insert     four words per thread (length from the E bits, 64-bit hash); equal hashes of a CTA's 1024 words are merged in
//                     shared memory, then inserted into an open-addressing table {hash, count, one occurrence}
//   k_word_verify     one thread per word: bytes == bytes of the slot's first occurrence (a 64-bit collision is
//                     detected, never trusted; the call then repeats with another hash seed)
//   k_unique_*        table -> packed unique words (bytes, offsets, counts) for the host
// The host (train_host below) does what the reference does once per unique word: initial vocabulary
// (:278-320), split into symbols (:323-338).  Stage M never leaves the device:
// Stage M -- the merge loop (:141-183), three launches per merge, no host round trip inside a batch of merges:
//   k_count_all       the reference's full recount (:341-376), one thread per symbol: builds the pair hash table once
//                     (and again when the table has to grow); afterwards the same sums are kept up to date
//   k_detect          one thread per symbol: which words contain the pair -> list
//   k_apply           warp per listed word: merge in place, left to right (:379-401); pairs that disappear are
//                     subtracted from the table and the new ones added (sums mod 2^32, as the reference's u32)
//   k_best_pair       scan the pair table, lexicographic (count, pair) reduction; the last CTA decides:
//                     stop (:147-165), or name the merged symbol -- symbols are STRINGS in the reference, so a
//                     merged string that already exists must get the existing symbol.  The device keeps two
//                     64-bit polynomial hashes per symbol (hash(ab) = hash(a) * p^len(b) + hash(b)) and a symbol
//                     map; the host replays every merge on real strings afterwards and fails the call if a
//                     single decision differs, so the result is exact, not probabilistic.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ctk.h"
#include "engine.hpp"

namespace cg = cooperative_groups;

namespace ctk {
namespace {

constexpr uint32_t INVALID = 0xFFFFFFFFu;
constexpr uint64_t EMPTY64 = ~0ull;
constexpr uint64_t P1 = 0x100000001B3ull, P2 = 0x9E3779B97F4A7C15ull;
constexpr int BATCH = 256;             // merges per host round trip

// ------------------------------------------------------------------------------------------------ stage W
// ---- word bitmaps.  One thread per 32 bytes decides, for each byte, whether it belongs to a White_Space char (what
// str::split_whitespace splits on: 25 code points, 1..3 bytes each; valid UTF-8 assumed) and writes two bits per byte:
//   S = a word starts here        E = this byte continues the word of the previous byte
// Words do not span texts (brk has a bit at every text start).  Everything after this kernel works on the bitmaps:
// a word's length is a run of E bits, no byte of the text is classified twice.
__global__ void __launch_bounds__(256) k_word_bits(const uint8_t* __restrict__ text, uint64_t n, const uint32_t* __restrict__ brk,
                                                   uint32_t n_groups, uint32_t* __restrict__ S, uint32_t* __restrict__ E,
                                                   unsigned long long* total) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t starts = 0;
    if (g < n_groups) {
        // bytes [32g - 4, 32g + 36): window byte q is position 32g + q - 4 (the text is padded with zero bytes)
        uint32_t w[10];
        const uint4 a = reinterpret_cast<const uint4*>(text)[2 * (size_t)g], b = reinterpret_cast<const uint4*>(text)[2 * (size_t)g + 1];
        w[0] = g ? reinterpret_cast<const uint32_t*>(text)[8 * (size_t)g - 1] : 0u;
        w[1] = a.x; w[2] = a.y; w[3] = a.z; w[4] = a.w; w[5] = b.x; w[6] = b.y; w[7] = b.z; w[8] = b.w;
        w[9] = reinterpret_cast<const uint32_t*>(text)[8 * (size_t)g + 8];
        uint32_t hi = 0;
#pragma unroll
        for (int k = 0; k < 10; ++k) hi |= w[k];
        unsigned long long A = 0, L2 = 0, L3 = 0;
        #define WB(q) ((w[(q) >> 2] >> (((q) & 3) * 8)) & 0xFFu)
        if (hi & 0x80808080u) {                                         // some non-ASCII byte: the multi-byte White_Space chars
#pragma unroll
            for (int q = 1; q <= 35; ++q) {
                const uint32_t c0 = WB(q), c1 = WB(q + 1), c2 = WB(q + 2);
                const bool as = c0 == 0x20 || (c0 - 9u) < 5u;
                const bool l2 = c0 == 0xC2 && (c1 == 0x85 || c1 == 0xA0);
                const bool l3 = (c0 == 0xE1 && c1 == 0x9A && c2 == 0x80) ||
                                (c0 == 0xE2 && ((c1 == 0x80 && ((c2 - 0x80u) <= 0x0Au || c2 == 0xA8 || c2 == 0xA9 || c2 == 0xAF)) || (c1 == 0x81 && c2 == 0x9F))) ||
                                (c0 == 0xE3 && c1 == 0x80 && c2 == 0x80);
                A |= (unsigned long long)as << q; L2 |= (unsigned long long)l2 << q; L3 |= (unsigned long long)l3 << q;
            }
        } else {
#pragma unroll
            for (int q = 3; q <= 35; ++q) { const uint32_t c0 = WB(q); A |= (unsigned long long)(c0 == 0x20 || (c0 - 9u) < 5u) << q; }
        }
        #undef WB
        const unsigned long long ws = A | L2 | (L2 << 1) | L3 | (L3 << 1) | (L3 << 2);
        const uint64_t lo = 32ull * g;
        const uint32_t vm = lo + 32 <= n ? 0xFFFFFFFFu : ((1u << (uint32_t)(n - lo)) - 1u);
        const uint32_t W = ~(uint32_t)(ws >> 4) & vm;
        const uint32_t Wprev = g ? (uint32_t)(~(ws >> 3) & 1ull) : 0u;
        const uint32_t e = W & ((W << 1) | Wprev) & ~brk[g];
        const uint32_t st = W & ~e;
        S[g] = st; E[g] = e;
        starts = __popc(st);
    }
    starts = __reduce_add_sync(0xFFFFFFFFu, starts);
    __shared__ uint32_t sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = starts;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int k = 0; k < 8; ++k) t += sh[k];
        if (t) atomicAdd(total, (unsigned long long)t);
    }
}

// S bitmap -> list of word-start positions (the order of the list does not matter: counts are sums)
__global__ void __launch_bounds__(256) k_expand_starts(const uint32_t* __restrict__ S, uint32_t n_groups, uint32_t* __restrict__ starts,
                                                       unsigned int* cursor) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    uint32_t bits = g < n_groups ? S[g] : 0u;
    const uint32_t c = __popc(bits);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((int)lane >= o) incl += u; }
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t base;
    if (lane == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int k = 0; k < 8; ++k) { uint32_t v = wsum[k]; wsum[k] = t; t += v; }
        base = t ? atomicAdd(cursor, t) : 0u;
    }
    __syncthreads();
    uint32_t o = base + wsum[threadIdx.x >> 5] + incl - c;
    while (bits) { starts[o++] = 32u * g + (uint32_t)__ffs(bits) - 1u; bits &= bits - 1; }
}

// length of the word that starts at s: 1 + the run of E bits after it (E is zero past the end of the text)
__device__ __forceinline__ uint32_t word_len(const uint32_t* __restrict__ E, uint64_t s) {
    uint64_t p = s + 1;
    uint32_t len = 1;
    for (;;) {
        const uint32_t sh = (uint32_t)p & 31u, avail = 32u - sh;
        const uint32_t stop = ~(E[p >> 5] >> sh);                       // the shifted-in top bits read as "stop"
        const uint32_t run = (uint32_t)__ffs(stop) - 1u;                // stop != 0 whenever sh > 0
        if (stop != 0 && run < avail) return len + run;
        len += avail; p += avail;
    }
}

// four text bytes starting at byte position p (aligned loads; the text is padded)
struct Bytes4 {
    const uint32_t* t32; uint32_t idx, sh, cur;
    __device__ __forceinline__ Bytes4(const uint8_t* text, uint64_t p) : t32(reinterpret_cast<const uint32_t*>(text)), idx((uint32_t)(p >> 2)), sh(((uint32_t)p & 3u) * 8u) { cur = t32[idx]; }
    __device__ __forceinline__ uint32_t next() { const uint32_t nx = t32[++idx]; const uint32_t v = __funnelshift_r(cur, nx, sh); cur = nx; return v; }
};

__global__ void k_mark_breaks(const uint64_t* off, size_t n_texts, uint32_t* brk) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_texts) return;
    uint64_t o = off[i];
    atomicOr(&brk[o >> 5], 1u << (o & 31));
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

struct __align__(16) WordEntry { unsigned long long key; uint32_t count, rep; };   // one sector per distinct word: hash, occurrences, one occurrence's position
struct WordTable { WordEntry* e; uint32_t mask; };

__global__ void k_word_table_clear(WordTable tab) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= tab.mask) *reinterpret_cast<uint4*>(tab.e + i) = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0xFFFFFFFFu);
}

// A CTA takes 256 consecutive words: each thread walks and hashes one, equal hashes are merged in a shared-memory
// table first (the most frequent word of a text is ~5 % of all words: without this every one of its occurrences is an
// atomic on the same global address), then one thread per distinct hash updates the global table.
constexpr int WI_THREADS = 256, WI_PER = 4, WI_WORDS = WI_THREADS * WI_PER, WI_SLOTS = 2048;
struct WordFlags { uint32_t collision, overflow, fill; };

// hash of the word [s, s + len): four bytes per step, the tail masked
__device__ __forceinline__ uint64_t word_hash_loop(const uint8_t* __restrict__ t, uint64_t s, uint32_t len, uint64_t seed) {
    uint64_t h = seed;
    Bytes4 rd(t, s);
    for (uint32_t o = 0; o < len; o += 4) {
        uint32_t v = rd.next();
        if (len - o < 4) v &= (1u << (8 * (len - o))) - 1u;
        h = (h ^ v) * P1; h ^= h >> 29;
    }
    return h;
}

// A CTA takes 1024 words, four per thread.  The kernel is a chain of dependent memory round trips (start -> E bits ->
// text -> shared table -> global table), so every thread keeps the loads of its four words in flight together: for
// words of at most 16 bytes (the rest take the loop) the addresses depend on the start position only.
__global__ void __launch_bounds__(WI_THREADS) k_word_insert(const uint8_t* __restrict__ t, const uint32_t* __restrict__ E, const uint32_t* __restrict__ starts,
                                                            uint32_t n_words, uint64_t seed, WordTable tab, uint32_t* slot_of, WordFlags* fl) {
    __shared__ unsigned long long s_key[WI_SLOTS];
    __shared__ uint32_t s_count[WI_SLOTS], s_rep[WI_SLOTS], s_gslot[WI_SLOTS];
    for (int k = threadIdx.x; k < WI_SLOTS; k += WI_THREADS) { s_key[k] = EMPTY64; s_count[k] = 0; s_rep[k] = INVALID; }
    const uint32_t w0 = blockIdx.x * WI_WORDS + threadIdx.x;
    const uint32_t* t32 = reinterpret_cast<const uint32_t*>(t);
    uint32_t st[WI_PER], e0[WI_PER], e1[WI_PER], x[WI_PER][5];
#pragma unroll
    for (int k = 0; k < WI_PER; ++k) { const uint32_t w = w0 + k * WI_THREADS; st[k] = w < n_words ? starts[w] : 0u; }
#pragma unroll
    for (int k = 0; k < WI_PER; ++k) {
        const uint32_t p = st[k] + 1;                                    // < 2^32: the call takes less than 4 GiB
        e0[k] = E[p >> 5]; e1[k] = E[(p >> 5) + 1];
#pragma unroll
        for (int j = 0; j < 5; ++j) x[k][j] = t32[(st[k] >> 2) + j];
    }
    __syncthreads();
    uint32_t mine[WI_PER];
#pragma unroll
    for (int k = 0; k < WI_PER; ++k) {
        mine[k] = INVALID;
        if (w0 + k * WI_THREADS >= n_words) continue;
        const uint32_t p = st[k] + 1, sh = p & 31u;
        const unsigned long long cont = (((unsigned long long)e1[k] << 32) | e0[k]) >> sh;    // E bits from p on, at least 33 of them
        const uint32_t run = (uint32_t)__ffsll((long long)~cont) - 1u;
        uint32_t len;
        uint64_t h;
        if (run < 16) {                                                  // the word is 1 + run <= 16 bytes: registers only
            len = 1 + run;
            h = seed;
            const uint32_t bs = (st[k] & 3u) * 8u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if ((uint32_t)(4 * j) < len) {
                    uint32_t v = __funnelshift_r(x[k][j], x[k][j + 1], bs);
                    if (len - 4 * j < 4) v &= (1u << (8 * (len - 4 * j))) - 1u;
                    h = (h ^ v) * P1; h ^= h >> 29;
                }
            }
        } else {
            len = word_len(E, st[k]);
            h = word_hash_loop(t, st[k], len, seed);
        }
        h = mix64(h ^ (len * P2));
        if (h == EMPTY64) h = 0;
        uint32_t slot = (uint32_t)(h >> 40) & (WI_SLOTS - 1);
        for (;;) {                                                       // 1024 words, 2048 slots: always ends
            unsigned long long q = s_key[slot];
            if (q == EMPTY64) {
                q = atomicCAS(&s_key[slot], EMPTY64, h);
                if (q == EMPTY64) { s_rep[slot] = st[k]; q = h; }        // any occurrence can stand for the word: all are compared with it
            }
            if (q == h) break;
            slot = (slot + 1) & (WI_SLOTS - 1);
        }
        atomicAdd(&s_count[slot], 1u);
        mine[k] = slot;
    }
    __syncthreads();
    // one thread per distinct hash of the CTA updates the global table; the first probes of a thread's slots go out together
    constexpr int FL = WI_SLOTS / WI_THREADS;
    unsigned long long hk[FL], g0[FL];
#pragma unroll
    for (int j = 0; j < FL; ++j) {
        hk[j] = s_key[threadIdx.x + j * WI_THREADS];
        g0[j] = hk[j] != EMPTY64 ? tab.e[(uint32_t)(hk[j] >> 20) & tab.mask].key : 0ull;
    }
#pragma unroll
    for (int j = 0; j < FL; ++j) {
        const uint64_t h = hk[j];
        if (h == EMPTY64) continue;
        const int k = threadIdx.x + j * WI_THREADS;
        uint32_t slot = (uint32_t)(h >> 20) & tab.mask, probes = 0;
        uint64_t g = g0[j];
        for (;; ++probes) {
            if (probes > 256u || probes > tab.mask) { fl->overflow = 1; slot = 0; break; }   // a probe this long means the table is too full: the host grows it
            if (g == EMPTY64) {
                g = atomicCAS(&tab.e[slot].key, EMPTY64, h);
                if (g == EMPTY64) { atomicAdd(&fl->fill, 1u); tab.e[slot].rep = s_rep[k]; g = h; }
            }
            if (g == h) break;
            slot = (slot + 1) & tab.mask;
            g = tab.e[slot].key;
        }
        atomicAdd(&tab.e[slot].count, s_count[k]);
        s_gslot[k] = slot;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WI_PER; ++k) if (mine[k] != INVALID) slot_of[w0 + k * WI_THREADS] = s_gslot[mine[k]];
}

__global__ void k_word_verify(const uint8_t* __restrict__ t, const uint32_t* __restrict__ E, const uint32_t* __restrict__ starts, uint32_t n_words,
                              WordTable tab, const uint32_t* __restrict__ slot_of, WordFlags* fl) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const uint64_t s = starts[w], r = tab.e[slot_of[w]].rep;
    if (r == s) return;
    const uint32_t len = word_len(E, s);
    bool ok = word_len(E, r) == len;
    if (ok) {
        Bytes4 a(t, s), b(t, r);
        for (uint32_t o = 0; o < len; o += 4) {
            uint32_t d = a.next() ^ b.next();
            if (len - o < 4) d &= (1u << (8 * (len - o))) - 1u;
            if (d) { ok = false; break; }
        }
    }
    if (!ok) fl->collision = 1;
}

struct U32ToU64 { __device__ uint64_t operator()(uint32_t v) const { return v; } };

struct SlotUsed {
    const WordEntry* e;
    __device__ bool operator()(uint32_t i) const { return e[i].key != EMPTY64; }
};

__global__ void k_unique_len(const uint32_t* __restrict__ E, WordTable tab, const uint32_t* uslot,
                             uint32_t n_unique, uint32_t* ulen, uint32_t* ucount, uint32_t* urep) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_unique) return;
    uint32_t slot = uslot[u], r = tab.e[slot].rep;
    ulen[u] = word_len(E, r);
    ucount[u] = tab.e[slot].count;
    urep[u] = r;
}

__global__ void k_unique_gather(const uint8_t* t, const uint32_t* urep, const uint32_t* ulen, const uint64_t* uoff,
                                uint32_t n_unique, uint8_t* out) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_unique) return;
    const uint8_t* s = t + urep[u];
    uint8_t* d = out + uoff[u];
    for (uint32_t k = 0, l = ulen[u]; k < l; ++k) d[k] = s[k];
}

// ------------------------------------------------------------------------------------------------ stage M
struct TrainState {
    uint32_t done, reason;               // reason: 1 no pairs, 2 below min_frequency, 3 vocabulary full, 4 symbol table full
    uint32_t pause;                      // the pair table needs to grow (or overflowed): the host rebuilds it from the words
    uint32_t overflow, fill;             // pair table: an insert found no slot / number of keys in it
    uint32_t n_symbols, sym_cap;
    uint32_t vocab_len, vocab_size, min_freq;
    uint32_t cur_l, cur_r, cur_m;        // the merge k_detect / k_apply carry out
    uint32_t n_log, ticket;
    uint32_t iter, n_dirty;              // merge number; words that contain the current pair
    unsigned long long t_phase[4];       // cluster kernel: nanoseconds in detect / apply / best pair / barriers (CTA 0's view)
};

struct __align__(16) TrainPair { unsigned long long key; uint32_t val, pad; };   // one 16-byte access reads key and count
struct Best { uint32_t count, dirty; uint64_t key; };    // (dirty: only used by the per-segment cache)
// The table is cut into segments of SEG slots; seg_best caches each segment's best pair and seg_dirty says whether a count of
// the segment changed since it was computed (every change goes through pair_add).  A merge touches a handful of pairs, so the
// best-pair step re-reads only the few segments that changed instead of the whole table.
constexpr uint32_t SEG = 128;
struct PairTable { TrainPair* e; uint32_t mask; Best* seg_best; };

// count[(a, b)] += f (mod 2^32, like the reference's u32 sums; f may be "negative")
__device__ __forceinline__ void pair_add(const PairTable& pt, TrainState* st, uint32_t a, uint32_t b, uint32_t f) {
    uint64_t key = ((uint64_t)a << 32) | b;
    uint32_t slot = (uint32_t)(mix64(key) >> 24) & pt.mask;
    for (uint32_t probes = 0; probes <= pt.mask && probes <= 512u; ++probes) {   // a longer probe = too full: overflow, the host rebuilds
        uint64_t k = pt.e[slot].key;
        if (k == EMPTY64) {
            k = atomicCAS(&pt.e[slot].key, EMPTY64, key);
            if (k == EMPTY64) { atomicAdd(&st->fill, 1u); k = key; }
        }
        if (k == key) { atomicAdd(&pt.e[slot].val, f); pt.seg_best[slot / SEG].dirty = 1u; return; }
        slot = (slot + 1) & pt.mask;
    }
    st->overflow = 1;
}

// Loads of data another CTA may have written earlier in the SAME launch (cluster loop) must bypass the per-SM L1;
// across kernel boundaries (the default three-kernel loop) plain loads are right and measurably faster (22.6 vs 26.7 us).
template <bool CG, class T> __device__ __forceinline__ T ld(const T* p) { if (CG) return __ldcg(p); else return *p; }
template <bool CG> __device__ __forceinline__ uint32_t lds(const uint32_t* p) { if (CG) return *(const volatile uint32_t*)p; else return *p; }   // state words

struct Words {                           // every word keeps its slot range [woff, woff + initial length); wlen shrinks
    uint32_t* sym; const uint32_t* slot_word; const uint32_t* woff; uint32_t* wlen; const uint32_t* wfreq;
    uint32_t* dirty_stamp; uint4* dirty_list; uint32_t n_slots, n_words;   // list entry: word, first slot, length, count
};

// The reference's full recount (bpe_trainer.rs:341-376), one thread per symbol slot.  Runs before the first merge and
// whenever the pair table is rebuilt; between merges the counts are kept up to date by k_apply (same sums).
__global__ void k_table_clear(PairTable pt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= pt.mask) *reinterpret_cast<uint4*>(pt.e + i) = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
    if (i <= pt.mask / SEG) pt.seg_best[i] = Best{0u, 0u, EMPTY64};
}

__global__ void __launch_bounds__(256) k_count_all(TrainState* st, Words W, PairTable pt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W.n_slots) return;
    uint32_t w = W.slot_word[i], k = i - W.woff[w];
    if (k + 1 >= W.wlen[w]) return;
    pair_add(pt, st, W.sym[i], W.sym[i + 1], W.wfreq[w]);
}

// Which words contain the pair (cur_l, cur_r)?  One thread per symbol slot .
template <bool CG> __device__ __forceinline__ void detect_hit(TrainState* st, const Words& W, uint32_t i, uint32_t w, uint32_t stamp) {
    const uint32_t o = W.woff[w], len = ld<CG>(W.wlen + w), f = W.wfreq[w];
    if (i - o + 1 >= len) return;                                               // a dead slot, or the word's last symbol
    if (atomicExch(&W.dirty_stamp[w], stamp) != stamp) W.dirty_list[atomicAdd(&st->n_dirty, 1u)] = make_uint4(w, o, len, f);
}

// four slots per thread, all loads issued before the first compare (the step is latency-bound, not bandwidth-bound)
template <bool CG> __device__ __forceinline__ void detect_range(TrainState* st, const Words& W, uint32_t tid, uint32_t n_threads) {
    for (uint32_t b = tid * 4; b < W.n_slots; b += n_threads * 4) {
        const uint4 v = ld<CG>(reinterpret_cast<const uint4*>(W.sym + b));      // .cg: another CTA of the cluster may have rewritten the word
        const uint32_t v4 = ld<CG>(W.sym + b + 4);
        const uint4 ws = *reinterpret_cast<const uint4*>(W.slot_word + b);
        const uint32_t l = lds<CG>(&st->cur_l), r = lds<CG>(&st->cur_r), stamp = lds<CG>(&st->iter) + 1;
        if (l == INVALID) return;
        if (v.x == l && v.y == r) detect_hit<CG>(st, W, b, ws.x, stamp);
        if (v.y == l && v.z == r && b + 1 < W.n_slots) detect_hit<CG>(st, W, b + 1, ws.y, stamp);
        if (v.z == l && v.w == r && b + 2 < W.n_slots) detect_hit<CG>(st, W, b + 2, ws.z, stamp);
        if (v.w == l && v4 == r && b + 3 < W.n_slots) detect_hit<CG>(st, W, b + 3, ws.w, stamp);
    }
}

// Programmatic dependent launch: the three kernels of a merge are launched with programmatic stream serialisation, so a
// kernel's CTAs are scheduled while the previous kernel drains and wait HERE until its memory is visible.  The chain of
// three dependent launches per merge is launch-latency-bound (each kernel does microseconds of work); this removes most of
// the gap between them.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__global__ void __launch_bounds__(256) k_detect(TrainState* st, Words W) {
    pdl_enter();
    if (st->done || st->pause) return;
    detect_range<false>(st, W, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// Apply the merge to the listed words (bpe_trainer.rs:379-401: left to right, so "aaa" -> "aa a"), a warp per word,
// 32 symbols at a time, compacted in place; pairs that disappear are subtracted from the table, new ones added.
template <bool CG> __device__ __forceinline__ void apply_range(TrainState* st, const Words& W, const PairTable& pt, uint32_t warp, uint32_t n_warps) {
    const uint4 first = ld<CG>(W.dirty_list + (warp < W.n_words ? warp : 0));   // issued together with the state loads
    const uint32_t l = lds<CG>(&st->cur_l);
    if (l == INVALID) return;
    const uint32_t r = lds<CG>(&st->cur_r), m = lds<CG>(&st->cur_m), n_dirty = lds<CG>(&st->n_dirty), lane = threadIdx.x & 31;
    for (uint32_t d = warp; d < n_dirty; d += n_warps) {
        const uint4 ent = d == warp ? first : ld<CG>(W.dirty_list + d);
        const uint32_t w = ent.x, len = ent.z, f = ent.w;
        uint32_t* s = W.sym + ent.y;
        uint32_t out_base = 0, carry = 0, new_last = INVALID, new_last_m = 0, old_last = INVALID, old_last_inv = 0;
        for (uint32_t c = 0; c < len; c += 32) {
            const uint32_t i = c + lane;
            const bool valid = i < len;
            const uint32_t o = valid ? ld<CG>(s + i) : INVALID;
            const uint32_t ahead = (c + 32 < len) ? ld<CG>(s + c + 32) : INVALID;
            uint32_t nx = __shfl_down_sync(0xFFFFFFFFu, o, 1);
            if (lane == 31) nx = ahead;
            const uint32_t M = __ballot_sync(0xFFFFFFFFu, valid && o == l && nx == r && nx != INVALID);
            uint32_t merges = M;
            if (l == r) {                                               // overlapping candidates: leftmost first
                merges = 0;
                uint32_t skip = carry;
                for (int b = 0; b < 32; ++b) {
                    if (skip) { skip = 0; continue; }
                    if ((M >> b) & 1u) { merges |= 1u << b; skip = 1; }
                }
            }
            const uint32_t consumed = (merges << 1) | carry;
            carry = merges >> 31;
            const uint32_t validm = __ballot_sync(0xFFFFFFFFu, valid);
            const uint32_t kept = validm & ~consumed, inv = merges | consumed;
            // pairs of the old sequence that disappear: (old[i-1], old[i]) with either end merged or consumed
            uint32_t op = __shfl_up_sync(0xFFFFFFFFu, o, 1);
            bool op_inv = (inv >> ((lane + 31) & 31)) & 1u;
            if (lane == 0) { op = old_last; op_inv = old_last_inv; }
            {
                const bool has = valid && op != INVALID && (((inv >> lane) & 1u) || op_inv);
                const uint64_t key = has ? (((uint64_t)op << 32) | o) : EMPTY64 - 1 - lane;
                const uint32_t same = __match_any_sync(0xFFFFFFFFu, key);
                if (has && (uint32_t)(__ffs(same) - 1) == lane) pair_add(pt, st, op, o, 0u - f * (uint32_t)__popc(same));
            }
            old_last = __shfl_sync(0xFFFFFFFFu, o, 31); old_last_inv = inv >> 31;   // only read again if the word goes on
            // the new sequence
            const bool keep = (kept >> lane) & 1u, mg = (merges >> lane) & 1u;
            const uint32_t v = mg ? m : o;
            const uint32_t below = kept & ((1u << lane) - 1u);
            const int src = below ? 31 - __clz(below) : 0;
            uint32_t pv = __shfl_sync(0xFFFFFFFFu, v, src);
            bool pm = (merges >> src) & 1u;
            if (!below) { pv = new_last; pm = new_last_m; }
            if (keep && (mg || out_base + __popc(below) != i)) s[out_base + __popc(below)] = v;
            {
                const bool has = keep && pv != INVALID && (mg || pm);
                const uint64_t key = has ? (((uint64_t)pv << 32) | v) : EMPTY64 - 1 - lane;
                const uint32_t same = __match_any_sync(0xFFFFFFFFu, key);
                if (has && (uint32_t)(__ffs(same) - 1) == lane) pair_add(pt, st, pv, v, f * (uint32_t)__popc(same));
            }
            if (kept) {
                const int top = 31 - __clz(kept);
                new_last = __shfl_sync(0xFFFFFFFFu, v, top); new_last_m = (merges >> top) & 1u;
            }
            out_base += __popc(kept);
            __syncwarp();
        }
        if (lane == 0) W.wlen[w] = out_base;
    }
}

__global__ void __launch_bounds__(128) k_apply(TrainState* st, Words W, PairTable pt) {
    pdl_enter();
    if (st->done || st->pause) return;
    apply_range<false>(st, W, pt, (blockIdx.x * blockDim.x + threadIdx.x) >> 5, (gridDim.x * blockDim.x) >> 5);
}

// Detect and apply in ONE launch, a thread per word (words are short: 7.5 symbols on average for config 1): does the word
// contain the pair?  Then its old pairs leave the table, the merge is applied left to right in place (bpe_trainer.rs:379-401),
// and its new pairs enter -- the same u32 sums as the incremental update of k_apply, two launches per merge instead of three.
// Chosen by the host when no word is longer than FUSED_MAX_WORD symbols (a thread walks its word alone).
constexpr uint32_t FUSED_MAX_WORD = 256;
__global__ void __launch_bounds__(256) k_word_merge(TrainState* st, Words W, PairTable pt) {
    pdl_enter();
    if (st->done || st->pause) return;
    const uint32_t l = st->cur_l;
    if (l == INVALID) return;
    const uint32_t r = st->cur_r, m = st->cur_m;
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W.n_words) return;
    const uint32_t len = W.wlen[w];
    if (len < 2) return;
    uint32_t* const s = W.sym + W.woff[w];
    bool hit = false;
    uint32_t prev = INVALID;
    for (uint32_t c = 0; c < len && !hit; c += 8) {                  // eight loads in flight: the scan is a chain of L2 round trips otherwise
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = c + k < len ? s[c + k] : INVALID;
#pragma unroll
        for (int k = 0; k < 8; ++k) { hit = hit || (prev == l && v[k] == r); prev = v[k]; }
    }
    if (!hit) return;
    const uint32_t f = W.wfreq[w];
    prev = s[0];
    for (uint32_t i = 1; i < len; ++i) { const uint32_t cur = s[i]; pair_add(pt, st, prev, cur, 0u - f); prev = cur; }
    uint32_t out = 0;
    for (uint32_t i = 0; i < len;) {
        const uint32_t a = s[i];
        if (i + 1 < len && a == l && s[i + 1] == r) { s[out++] = m; i += 2; }
        else { s[out++] = a; ++i; }
    }
    W.wlen[w] = out;
    prev = s[0];
    for (uint32_t i = 1; i < out; ++i) { const uint32_t cur = s[i]; pair_add(pt, st, prev, cur, f); prev = cur; }
}

struct SymTab {                          // device copy of what the host knows about every symbol
    uint64_t *h1, *h2, *pw1, *pw2;       // polynomial hashes of the symbol's string and p^length
    uint8_t* in_vocab;
    uint64_t* map_key; uint64_t* map_h2; uint32_t* map_id; uint32_t map_mask;   // hash -> symbol
};

__device__ __forceinline__ bool better(const Best& a, const Best& b) { return a.count > b.count || (a.count == b.count && a.key < b.key); }

// Best pair of the table (bpe_trainer.rs:152-155 with the tie rule of oracle/py_trainer.py); the last CTA decides.
template <bool CG> __device__ __forceinline__ void best_phase(TrainState* st, const PairTable& pt, Best* block_best, const SymTab& sy, uint4* log) {
    Best best{0u, 0u, EMPTY64};
    const uint32_t cap = pt.mask + 1;
    {   // a warp per segment: the cached best of a clean segment, a fresh scan of a dirty one (which also refreshes the cache)
        const uint32_t n_seg = cap / SEG, lane = threadIdx.x & 31;
        const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps_all = (gridDim.x * blockDim.x) >> 5;
        for (uint32_t sg = warp; sg < n_seg; sg += n_warps_all) {
            Best b{0u, 0u, EMPTY64};
            const uint4 cached = ld<CG>(reinterpret_cast<const uint4*>(pt.seg_best + sg));      // {count, dirty, key lo, key hi}: one access
            if (cached.y) {                                            // (warp-uniform)
#pragma unroll
                for (uint32_t q = 0; q < SEG / 32; ++q) {
                    const uint4 e = ld<CG>(reinterpret_cast<const uint4*>(pt.e + sg * SEG + q * 32 + lane));
                    if (e.z == 0) continue;                            // empty slot, or a pair that no longer occurs
                    Best c{e.z, 0u, ((uint64_t)e.y << 32) | e.x};
                    if (better(c, b)) b = c;
                }
                for (int d = 16; d > 0; d >>= 1) {
                    Best o{__shfl_down_sync(0xFFFFFFFFu, b.count, d), 0u, __shfl_down_sync(0xFFFFFFFFu, b.key, d)};
                    if (better(o, b)) b = o;
                }
                if (lane == 0) pt.seg_best[sg] = b;                    // (dirty = 0 again)
            } else b = Best{cached.x, 0u, ((uint64_t)cached.w << 32) | cached.z};
            if (lane == 0 && better(b, best)) best = b;
        }
    }
    __shared__ Best sb[32];
    __shared__ bool s_last;
    const int n_warps = blockDim.x >> 5;
    for (int d = 16; d > 0; d >>= 1) {
        Best o{__shfl_down_sync(0xFFFFFFFFu, best.count, d), 0u, __shfl_down_sync(0xFFFFFFFFu, best.key, d)};
        if (better(o, best)) best = o;
    }
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < n_warps; ++k) if (better(sb[k], best)) best = sb[k];
        block_best[blockIdx.x] = best;
        __threadfence();
        s_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    best = Best{0u, 0u, EMPTY64};
    for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
        const volatile Best* p = block_best + i;
        Best c{p->count, 0u, p->key};
        if (better(c, best)) best = c;
    }
    for (int d = 16; d > 0; d >>= 1) {
        Best o{__shfl_down_sync(0xFFFFFFFFu, best.count, d), 0u, __shfl_down_sync(0xFFFFFFFFu, best.key, d)};
        if (better(o, best)) best = o;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int k = 1; k < n_warps; ++k) if (better(sb[k], best)) best = sb[k];
    st->ticket = 0;
    st->cur_l = INVALID;
    st->n_dirty = 0;
    if (lds<CG>(&st->overflow) || lds<CG>(&st->fill) > (cap >> 1) + (cap >> 3)) { st->pause = 1; return; }        // host rebuilds, then this step repeats
    if (best.count == 0) { st->done = 1; st->reason = 1; return; }                            // bpe_trainer.rs:147-149
    if (best.count < lds<CG>(&st->min_freq)) { st->done = 1; st->reason = 2; return; }                  // :162-165
    uint32_t l = (uint32_t)(best.key >> 32), r = (uint32_t)best.key;
    uint64_t h1 = ld<CG>(sy.h1 + l) * ld<CG>(sy.pw1 + r) + ld<CG>(sy.h1 + r), h2 = ld<CG>(sy.h2 + l) * ld<CG>(sy.pw2 + r) + ld<CG>(sy.h2 + r);
    uint64_t mk = h1 == EMPTY64 ? 0 : h1;
    uint32_t slot = (uint32_t)(mix64(mk) >> 24) & sy.map_mask, id = INVALID;
    for (;;) {
        uint64_t k = ld<CG>(sy.map_key + slot);
        if (k == EMPTY64) break;
        if (k == mk && ld<CG>(sy.map_h2 + slot) == h2) { id = ld<CG>(sy.map_id + slot); break; }
        slot = (slot + 1) & sy.map_mask;
    }
    if (id == INVALID) {                                                                      // a new string
        if (lds<CG>(&st->n_symbols) >= lds<CG>(&st->sym_cap)) { st->done = 1; st->reason = 4; return; }
        id = lds<CG>(&st->n_symbols); st->n_symbols = id + 1;
        sy.map_key[slot] = mk; sy.map_h2[slot] = h2; sy.map_id[slot] = id;
        sy.h1[id] = h1; sy.h2[id] = h2; sy.pw1[id] = ld<CG>(sy.pw1 + l) * ld<CG>(sy.pw1 + r); sy.pw2[id] = ld<CG>(sy.pw2 + l) * ld<CG>(sy.pw2 + r);
        sy.in_vocab[id] = 0;
    }
    if (!ld<CG>(sy.in_vocab + id)) { sy.in_vocab[id] = 1; st->vocab_len = lds<CG>(&st->vocab_len) + 1; }                           // :168-169
    { const uint32_t nl = lds<CG>(&st->n_log); log[nl] = make_uint4(l, r, id, best.count); st->n_log = nl + 1; }
    st->cur_l = l; st->cur_r = r; st->cur_m = id;
    st->iter = lds<CG>(&st->iter) + 1;
    if (lds<CG>(&st->vocab_len) >= lds<CG>(&st->vocab_size)) { st->done = 1; st->reason = 3; }                    // :141
}

__global__ void __launch_bounds__(256) k_best_pair(TrainState* st, PairTable pt, Best* block_best, SymTab sy, uint4* log) {
    pdl_enter();
    if (st->done || st->pause) return;
    best_phase<false>(st, pt, block_best, sy, log);
}

// The same three phases for up to `iters` merges inside ONE thread-block cluster: the hardware cluster barrier replaces
// the kernel boundaries.  Everything one CTA writes and another reads afterwards is read with .cg / volatile loads
// (L1 is per SM and not coherent); __threadfence() before each barrier orders the writes.
__global__ void __launch_bounds__(1024, 1) k_train_cluster(TrainState* st, Words W, PairTable pt, Best* block_best, SymTab sy, uint4* log, int iters) {
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    const volatile TrainState* vs = st;
    auto now = []() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; };
    unsigned long long acc[4] = {0, 0, 0, 0}, t0 = now(), t1;
    #define PHASE(k) t1 = now(); acc[k] += t1 - t0; t0 = t1;
    for (int it = 0; it < iters; ++it) {
        if (vs->cur_l != INVALID) {                                     // uniform: written before the last barrier
            detect_range<true>(st, W, tid, n_threads);
            PHASE(0) __threadfence(); cluster.sync(); PHASE(3)
            apply_range<true>(st, W, pt, tid >> 5, n_threads >> 5);
            PHASE(1) __threadfence(); cluster.sync(); PHASE(3)
        }
        best_phase<true>(st, pt, block_best, sy, log);
        PHASE(2) __threadfence(); cluster.sync(); PHASE(3)
        if (vs->done || vs->pause) break;
    }
    #undef PHASE
    if (tid == 0) for (int k = 0; k < 4; ++k) st->t_phase[k] += acc[k];
}

// ------------------------------------------------------------------------------------------------ host side
struct DevBuf {
    std::vector<void*> all;
    template <class T> cudaError_t get(T** p, size_t n) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) { all.push_back(q); *p = (T*)q; }
        return e;
    }
    ~DevBuf() { for (void* q : all) cudaFree(q); }
};

uint32_t pow2_at_least(uint64_t x) { uint64_t p = 1024; while (p < x) p <<= 1; return (uint32_t)std::min<uint64_t>(p, 1ull << 31); }

void hash_string(const std::string& s, uint64_t* h1, uint64_t* h2, uint64_t* pw1, uint64_t* pw2) {
    uint64_t a = 0, b = 0, pa = 1, pb = 1;
    for (unsigned char c : s) { a = a * P1 + c + 1; b = b * P2 + c + 1; pa *= P1; pb *= P2; }
    *h1 = a; *h2 = b; *pw1 = pa; *pw2 = pb;
}
uint64_t mix64_host(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

void put_utf8(std::string& s, uint32_t cp) {
    if (cp < 0x80) s += (char)cp;
    else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 63)); }
    else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
    else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 63)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
}

}  // namespace

struct Trained {
    std::vector<std::string> symbols;        // every symbol string, by symbol index
    std::vector<int64_t> vocab_id;           // per symbol: id in the vocabulary map, -1 = not in it
    std::vector<uint32_t> merges;            // 2 per merge: symbol indices
    std::vector<uint32_t> merge_counts;      // the pair's count when it was chosen
    // packed views for the C ABI
    std::vector<uint8_t> sym_bytes; std::vector<uint64_t> sym_off;
    double ms_words = 0, ms_words_kernels = 0, ms_merges = 0, ms_host = 0;
    uint64_t n_words = 0, n_unique = 0, n_bytes = 0, n_symbols0 = 0, kernels = 0;
    uint32_t stop_reason = 0, rebuilds = 0; int cluster = 0;
};

struct PhaseTrace {                       // CTK_TRAIN_TRACE=1: host-timed phases (each ends with the stream idle) on stderr
    bool on = getenv("CTK_TRAIN_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void mark(const char* what, cudaStream_t st) {
        if (!on) return;
        cudaStreamSynchronize(st);
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[ctk train] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

struct SegTimer {                         // device time of the kernel segments between the host's allocations and read-backs
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> v;
    void begin(cudaStream_t s) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); v.push_back({a, b}); }
    void end(cudaStream_t s) { cudaEventRecord(v.back().second, s); }
    double total() {                                                      // after a synchronise
        double t = 0;
        for (auto& p : v) { float ms = 0; if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) t += ms; cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
        v.clear();
        return t;
    }
    ~SegTimer() { total(); }
};

#define TCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_last_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call); return CTK_ERR_CUDA; } } while (0)

static int train_impl(const ctk_bpe_trainer_config& cfg, int device, const uint8_t* text, const uint64_t* off, size_t n_texts, Trained& out) {
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev == 0) {
        set_last_error(std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e0));
        return CTK_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_last_error("bad device index"); return CTK_ERR_ARG; }
    int prev_device = 0;
    cudaGetDevice(&prev_device);
    TCK(cudaSetDevice(device));
    struct DeviceGuard { int d; ~DeviceGuard() { cudaSetDevice(d); } } dguard{prev_device};   // the caller's current device is left as it was
    const uint64_t n = n_texts ? off[n_texts] : 0;
    if (n >= (1ull << 32) - 64) { set_last_error("ctk_train_bpe: one call takes less than 4 GiB of text"); return CTK_ERR_UNSUPPORTED; }
    for (size_t i = 0; i < n_texts; ++i) if (off[i] > off[i + 1]) { set_last_error("text offsets are not monotone"); return CTK_ERR_ARG; }
    out.n_bytes = n;
    cudaStream_t st;
    TCK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{st};
    cudaEvent_t ev[4];
    for (auto& e : ev) TCK(cudaEventCreate(&e));
    struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int i = 0; i < 4; ++i) cudaEventDestroy(e[i]); } } eg{ev};
    uint64_t launches = 0;

    // ---------------- stage W
    std::vector<uint8_t> ubytes; std::vector<uint64_t> uoff_h; std::vector<uint32_t> ucount_h;
    uint32_t n_words = 0, n_unique = 0;
    {
        DevBuf db;
        uint8_t* d_text; uint64_t* d_off; uint32_t* d_brk; uint32_t* d_starts; uint32_t* d_num;
        TCK(db.get(&d_text, n + 64)); TCK(db.get(&d_off, n_texts + 1)); TCK(db.get(&d_brk, (n >> 5) + 2));
        TCK(db.get(&d_num, 4));
        TCK(cudaMemcpyAsync(d_text, text, n, cudaMemcpyDefault, st));                // host or device text (ctk_train_new_from_texts hands over device text)
        if (n_texts) TCK(cudaMemcpyAsync(d_off, off, (n_texts + 1) * 8, cudaMemcpyHostToDevice, st));
        TCK(cudaEventRecord(ev[0], st));
        PhaseTrace tr; tr.mark("alloc + copy in", st);
        SegTimer seg;
        uint32_t *d_S = nullptr, *d_E = nullptr;
        const uint32_t n_groups = (uint32_t)((n + 31) / 32);
        if (n > 0) {
            unsigned long long* d_total; unsigned long long total = 0;
            unsigned int* d_cursor;
            TCK(db.get(&d_total, 1)); TCK(db.get(&d_cursor, 1)); TCK(db.get(&d_S, (size_t)n_groups + 2)); TCK(db.get(&d_E, (size_t)n_groups + 2));
            seg.begin(st);
            TCK(cudaMemsetAsync(d_text + n, 0, 64, st));                 // the kernels read whole words past the end
            TCK(cudaMemsetAsync(d_brk, 0, ((n >> 5) + 2) * 4, st));
            TCK(cudaMemsetAsync(d_E + n_groups, 0, 8, st));
            TCK(cudaMemsetAsync(d_total, 0, 8, st));
            TCK(cudaMemsetAsync(d_cursor, 0, 4, st));
            k_mark_breaks<<<(unsigned)((n_texts + 256) / 256), 256, 0, st>>>(d_off, n_texts, d_brk);
            k_word_bits<<<(n_groups + 255) / 256, 256, 0, st>>>(d_text, n, d_brk, n_groups, d_S, d_E, d_total); launches += 2;
            seg.end(st);
            TCK(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, st));
            TCK(cudaStreamSynchronize(st));
            n_words = (uint32_t)total;
            TCK(db.get(&d_starts, (size_t)total + 1));
            seg.begin(st);
            k_expand_starts<<<(n_groups + 255) / 256, 256, 0, st>>>(d_S, n_groups, d_starts, d_cursor); ++launches;
            seg.end(st);
            tr.mark("word bitmaps + starts", st);
        }
        if (n_words > 0) {
            // The table starts small (distinct words are a few per cent of the words of a text) and is rebuilt four times
            // larger when it fills beyond 5/8 -- at most up to 2 slots per word; a 64-bit collision repeats with a new seed.
            WordTable tab{}; uint32_t* d_slot; uint32_t* d_uslot; WordFlags* d_fl;
            const uint32_t cap_max = pow2_at_least(2ull * n_words);
            const uint32_t div = getenv("CTK_TRAIN_WORD_TABLE_DIV") ? std::max(1, atoi(getenv("CTK_TRAIN_WORD_TABLE_DIV"))) : 16;
            uint32_t cap = std::min<uint32_t>(cap_max, std::max<uint32_t>(1u << 16, pow2_at_least(n_words / div)));
            TCK(db.get(&d_slot, n_words)); TCK(db.get(&d_fl, 1));
            void* tab_mem = nullptr;
            struct MemGuard { void** p; ~MemGuard() { if (*p) cudaFree(*p); } } mg{&tab_mem};
            WordFlags fl{1, 0, 0};
            for (int attempt = 0, seed_no = 0; attempt < 24; ++attempt) {
                if (!tab_mem) {
                    TCK(cudaMalloc(&tab_mem, (size_t)cap * 16));
                    tab.e = (WordEntry*)tab_mem; tab.mask = cap - 1;
                }
                seg.begin(st);
                k_word_table_clear<<<(cap + 255) / 256, 256, 0, st>>>(tab); ++launches;
                TCK(cudaMemsetAsync(d_fl, 0, sizeof(WordFlags), st));
                k_word_insert<<<(n_words + WI_WORDS - 1) / WI_WORDS, WI_THREADS, 0, st>>>(d_text, d_E, d_starts, n_words,
                                                                                           0x9E37ull + 0x51ED27ull * seed_no, tab, d_slot, d_fl);
                k_word_verify<<<(n_words + 127) / 128, 128, 0, st>>>(d_text, d_E, d_starts, n_words, tab, d_slot, d_fl); launches += 2;
                seg.end(st);
                TCK(cudaMemcpyAsync(&fl, d_fl, sizeof fl, cudaMemcpyDeviceToHost, st));
                TCK(cudaStreamSynchronize(st));
                tr.mark("table: insert + verify", st);
                if (fl.overflow || (fl.fill > cap / 8 * 5 && cap < cap_max)) {      // grow and repeat
                    cudaFree(tab_mem); tab_mem = nullptr;
                    cap = cap < cap_max / 4 ? cap * 4 : cap_max;
                    fl.collision = 1;
                    continue;
                }
                if (!fl.collision) break;
                ++seed_no;
            }
            if (fl.collision) { set_last_error("ctk_train_bpe: word hash collisions under every seed"); return CTK_ERR_UNSUPPORTED; }
            TCK(db.get(&d_uslot, n_words));
            cub::CountingInputIterator<uint32_t> it(0);
            SlotUsed used{tab.e};
            size_t tmp_bytes = 0;
            TCK(cub::DeviceSelect::If(nullptr, tmp_bytes, it, d_uslot, d_num, (int64_t)cap, used, st));
            uint8_t* d_tmp; TCK(db.get(&d_tmp, tmp_bytes));
            seg.begin(st);
            TCK(cub::DeviceSelect::If(d_tmp, tmp_bytes, it, d_uslot, d_num, (int64_t)cap, used, st)); launches += 2;
            seg.end(st);
            TCK(cudaMemcpyAsync(&n_unique, d_num, 4, cudaMemcpyDeviceToHost, st));
            TCK(cudaStreamSynchronize(st));
            tr.mark("unique slots", st);
            uint32_t *d_ulen, *d_ucount, *d_urep; uint64_t* d_uoff; uint8_t* d_ubytes;
            TCK(db.get(&d_ulen, n_unique + 1)); TCK(db.get(&d_ucount, n_unique)); TCK(db.get(&d_urep, n_unique)); TCK(db.get(&d_uoff, n_unique + 1));
            TCK(cudaMemsetAsync(d_ulen + n_unique, 0, 4, st));
            unsigned g = (n_unique + 127) / 128;
            cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> lens(d_ulen, U32ToU64());
            size_t tb2 = 0;
            TCK(cub::DeviceScan::ExclusiveSum(nullptr, tb2, lens, d_uoff, (int64_t)n_unique + 1, st));
            uint8_t* d_tmp2; TCK(db.get(&d_tmp2, tb2));
            seg.begin(st);
            k_unique_len<<<g, 128, 0, st>>>(d_E, tab, d_uslot, n_unique, d_ulen, d_ucount, d_urep); ++launches;
            TCK(cub::DeviceScan::ExclusiveSum(d_tmp2, tb2, lens, d_uoff, (int64_t)n_unique + 1, st)); launches += 2;
            seg.end(st);
            uoff_h.resize(n_unique + 1); ucount_h.resize(n_unique);
            TCK(cudaMemcpyAsync(uoff_h.data(), d_uoff, (n_unique + 1) * 8ull, cudaMemcpyDeviceToHost, st));
            TCK(cudaStreamSynchronize(st));
            ubytes.resize(uoff_h[n_unique]);
            TCK(db.get(&d_ubytes, ubytes.size()));
            seg.begin(st);
            k_unique_gather<<<g, 128, 0, st>>>(d_text, d_urep, d_ulen, d_uoff, n_unique, d_ubytes); ++launches;
            seg.end(st);
            TCK(cudaEventRecord(ev[1], st));
            TCK(cudaMemcpyAsync(ubytes.data(), d_ubytes, ubytes.size(), cudaMemcpyDeviceToHost, st));
            TCK(cudaMemcpyAsync(ucount_h.data(), d_ucount, n_unique * 4ull, cudaMemcpyDeviceToHost, st));
            TCK(cudaStreamSynchronize(st));
            tr.mark("unique words out", st);
        } else {
            TCK(cudaEventRecord(ev[1], st));
            TCK(cudaStreamSynchronize(st));
            uoff_h.assign(1, 0);
        }
        float ms = 0; cudaEventElapsedTime(&ms, ev[0], ev[1]); out.ms_words = ms;
        out.ms_words_kernels = seg.total();
    }
    out.n_words = n_words; out.n_unique = n_unique;

    // ---------------- host: initial vocabulary and symbol sequences (once per unique word)
    std::string suffix = cfg.end_of_word_suffix ? std::string((const char*)cfg.end_of_word_suffix, cfg.suffix_len) : std::string();
    std::string prefix = cfg.continuing_subword_prefix ? std::string((const char*)cfg.continuing_subword_prefix, cfg.prefix_len) : std::string();
    const bool has_prefix = cfg.continuing_subword_prefix != nullptr;
    std::unordered_map<std::string, uint32_t> index;                   // symbol string -> symbol index
    std::vector<std::string>& symbols = out.symbols;
    std::vector<int64_t>& vocab_id = out.vocab_id;
    auto sym = [&](const std::string& s) -> uint32_t {
        auto it = index.find(s);
        if (it != index.end()) return it->second;
        uint32_t id = (uint32_t)symbols.size();
        index.emplace(s, id); symbols.push_back(s); vocab_id.push_back(-1);
        return id;
    };
    uint64_t next_id = 0, vocab_len = 0;
    auto vocab_insert = [&](uint32_t s, uint64_t id) { if (vocab_id[s] < 0) ++vocab_len; vocab_id[s] = (int64_t)id; };
    for (size_t i = 0; i < cfg.n_special; ++i) {                        // bpe_trainer.rs:283-286
        std::string t((const char*)cfg.special_tokens + cfg.special_off[i], cfg.special_off[i + 1] - cfg.special_off[i]);
        vocab_insert(sym(t), next_id++);
    }
    if (cfg.initial_alphabet) {                                         // :289-297
        for (size_t i = 0; i < cfg.n_alphabet; ++i) {
            std::string c; put_utf8(c, cfg.initial_alphabet[i]);
            uint32_t s = sym(c);
            if (vocab_id[s] < 0) vocab_insert(s, next_id++);
        }
    }
    // chars of the data with their frequencies (:300-306); a word is its bytes + the suffix
    std::unordered_map<uint32_t, uint32_t> char_freq;
    auto for_chars = [](const uint8_t* p, size_t len, auto&& fn) {
        for (size_t i = 0; i < len;) {
            uint32_t c = p[i], cp; size_t l;
            if (c < 0x80) { cp = c; l = 1; }
            else if (c < 0xE0) { cp = c & 31; l = 2; }
            else if (c < 0xF0) { cp = c & 15; l = 3; }
            else { cp = c & 7; l = 4; }
            for (size_t k = 1; k < l && i + k < len; ++k) cp = (cp << 6) | (p[i + k] & 63);
            fn(cp);
            i += l;
        }
    };
    for (uint32_t u = 0; u < n_unique; ++u) {
        uint32_t f = ucount_h[u];
        for_chars(ubytes.data() + uoff_h[u], uoff_h[u + 1] - uoff_h[u], [&](uint32_t cp) { char_freq[cp] += f; });
        for_chars((const uint8_t*)suffix.data(), suffix.size(), [&](uint32_t cp) { char_freq[cp] += f; });
    }
    std::vector<std::pair<uint32_t, uint32_t>> chars(char_freq.begin(), char_freq.end());
    std::sort(chars.begin(), chars.end(), [](auto& a, auto& b) { return a.second != b.second ? a.second > b.second : a.first < b.first; });
    const uint64_t limit = cfg.limit_alphabet >= 0 ? (uint64_t)cfg.limit_alphabet : chars.size();
    std::unordered_map<uint32_t, uint32_t> char_sym, char_sym_cont;
    for (size_t k = 0; k < chars.size(); ++k) {                         // :309-317
        std::string c; put_utf8(c, chars[k].first);
        uint32_t s = sym(c);
        if (k < limit && vocab_id[s] < 0) vocab_insert(s, next_id++);
        char_sym[chars[k].first] = s;
    }
    if (has_prefix) for (auto& ch : chars) { std::string c = prefix; put_utf8(c, ch.first); char_sym_cont[ch.first] = sym(c); }
    // symbol sequences (:323-338); words of one symbol have no pairs and are left out
    struct W { uint32_t off, len, freq; };
    std::vector<uint32_t> seq; std::vector<W> ws;
    seq.reserve(ubytes.size() + 16);
    for (uint32_t u = 0; u < n_unique; ++u) {
        uint32_t o = (uint32_t)seq.size();
        auto push = [&](uint32_t cp) { seq.push_back(cp); };
        for_chars(ubytes.data() + uoff_h[u], uoff_h[u + 1] - uoff_h[u], push);
        for_chars((const uint8_t*)suffix.data(), suffix.size(), push);
        uint32_t len = (uint32_t)seq.size() - o;
        for (uint32_t k = 0; k < len; ++k) seq[o + k] = (has_prefix && len > 1 && k > 0) ? char_sym_cont[seq[o + k]] : char_sym[seq[o + k]];
        if (len < 2 || ucount_h[u] == 0) { seq.resize(o); continue; }
        ws.push_back({o, len, ucount_h[u]});
    }
    uint64_t n_pairs = 0;
    for (auto& w : ws) n_pairs += w.len - 1;
    out.n_symbols0 = seq.size();
    const uint32_t n_sym0 = (uint32_t)symbols.size();

    // ---------------- stage M
    const uint64_t vocab_size = cfg.vocab_size;
    if (vocab_len < vocab_size && !ws.empty()) {
        DevBuf db;
        const uint64_t grow = std::min<uint64_t>(vocab_size - vocab_len, n_pairs);   // every new symbol takes a vocabulary slot
        const uint32_t sym_cap = (uint32_t)std::min<uint64_t>(n_sym0 + grow + 1, 0xFFFFFFF0ull);
        TrainState hs{}; hs.n_symbols = n_sym0; hs.sym_cap = sym_cap; hs.vocab_len = (uint32_t)vocab_len;
        hs.vocab_size = (uint32_t)std::min<uint64_t>(vocab_size, 0xFFFFFFFFull); hs.min_freq = cfg.min_frequency;
        hs.cur_l = hs.cur_r = hs.cur_m = INVALID;
        TrainState* d_st; Best* d_bb; uint4* d_log; uint32_t *d_sym, *d_slot_word, *d_woff, *d_wlen, *d_wfreq, *d_stamp; uint4* d_list;
        PairTable pt{}; SymTab sy;
        const uint32_t n_slots = (uint32_t)seq.size(), nw = (uint32_t)ws.size();
        const uint32_t pcap_max = pow2_at_least(4ull * n_slots + 1024);
        uint32_t pcap = std::min<uint32_t>(pcap_max, std::max<uint32_t>(1u << 16, pow2_at_least(n_pairs / 8)));
        const uint32_t mcap = pow2_at_least(2ull * sym_cap); sy.map_mask = mcap - 1;
        TCK(db.get(&d_st, 1)); TCK(db.get(&d_sym, (size_t)n_slots + 8)); TCK(db.get(&d_slot_word, (size_t)n_slots + 4)); TCK(db.get(&d_woff, nw));
        TCK(db.get(&d_wlen, nw)); TCK(db.get(&d_wfreq, nw)); TCK(db.get(&d_stamp, nw)); TCK(db.get(&d_list, nw));
        TCK(db.get(&d_bb, 2048)); TCK(db.get(&d_log, BATCH));
        TCK(db.get(&sy.h1, sym_cap)); TCK(db.get(&sy.h2, sym_cap)); TCK(db.get(&sy.pw1, sym_cap)); TCK(db.get(&sy.pw2, sym_cap));
        TCK(db.get(&sy.in_vocab, sym_cap)); TCK(db.get(&sy.map_key, mcap)); TCK(db.get(&sy.map_h2, mcap)); TCK(db.get(&sy.map_id, mcap));
        {
            std::vector<uint64_t> h1(n_sym0), h2(n_sym0), p1(n_sym0), p2(n_sym0), mk(mcap, EMPTY64), mh(mcap, 0);
            std::vector<uint32_t> mi(mcap, 0); std::vector<uint8_t> iv(n_sym0);
            for (uint32_t s = 0; s < n_sym0; ++s) {
                hash_string(symbols[s], &h1[s], &h2[s], &p1[s], &p2[s]);
                iv[s] = vocab_id[s] >= 0;
                uint64_t k = h1[s] == EMPTY64 ? 0 : h1[s];
                uint32_t slot = (uint32_t)(mix64_host(k) >> 24) & sy.map_mask;
                while (mk[slot] != EMPTY64) {
                    if (mk[slot] == k && mh[slot] == h2[s]) { set_last_error("ctk_train_bpe: symbol hash collision"); return CTK_ERR_UNSUPPORTED; }
                    slot = (slot + 1) & sy.map_mask;
                }
                mk[slot] = k; mh[slot] = h2[s]; mi[slot] = s;
            }
            std::vector<uint32_t> woff(nw), wlen(nw), wfreq(nw), slot_word(n_slots);
            for (uint32_t i = 0; i < nw; ++i) {
                woff[i] = ws[i].off; wlen[i] = ws[i].len; wfreq[i] = ws[i].freq;
                for (uint32_t k = 0; k < ws[i].len; ++k) slot_word[ws[i].off + k] = i;
            }
            TCK(cudaMemcpyAsync(sy.h1, h1.data(), n_sym0 * 8ull, cudaMemcpyHostToDevice, st));
            TCK(cudaMemcpyAsync(sy.h2, h2.data(), n_sym0 * 8ull, cudaMemcpyHostToDevice, st));
            TCK(cudaMemcpyAsync(sy.pw1, p1.data(), n_sym0 * 8ull, cudaMemcpyHostToDevice, st));
            TCK(cudaMemcpyAsync(sy.pw2, p2.data(), n_sym0 * 8ull, cudaMemcpyHostToDevice, st));
            TCK(cudaMemcpyAsync(sy.in_vocab, iv.data(), n_sym0, cudaMemcpyHostToDevice, st));
            TCK(cudaMemcpyAsync(sy.map_key, mk.data(), mcap * 8ull, cudaMemcpyHostToDevice, st));
            TCK(cudaMemcpyAsync(sy.map_h2, mh.data(), mcap * 8ull, cudaMemcpyHostToDevice, st));
            TCK(cudaMemcpyAsync(sy.map_id, mi.data(), mcap * 4ull, cudaMemcpyHostToDevice, st));
            TCK(cudaMemcpyAsync(d_sym, seq.data(), n_slots * 4ull, cudaMemcpyHostToDevice, st));
            TCK(cudaMemcpyAsync(d_slot_word, slot_word.data(), n_slots * 4ull, cudaMemcpyHostToDevice, st));
            TCK(cudaMemcpyAsync(d_woff, woff.data(), nw * 4ull, cudaMemcpyHostToDevice, st));
            TCK(cudaMemcpyAsync(d_wlen, wlen.data(), nw * 4ull, cudaMemcpyHostToDevice, st));
            TCK(cudaMemcpyAsync(d_wfreq, wfreq.data(), nw * 4ull, cudaMemcpyHostToDevice, st));
            TCK(cudaMemsetAsync(d_stamp, 0, nw * 4ull, st));
            TCK(cudaMemcpyAsync(d_st, &hs, sizeof hs, cudaMemcpyHostToDevice, st));
            TCK(cudaStreamSynchronize(st));
        }
        Words W{d_sym, d_slot_word, d_woff, d_wlen, d_wfreq, d_stamp, d_list, n_slots, nw};
        TCK(cudaEventRecord(ev[2], st));
        const unsigned g_slots = (n_slots + 255) / 256, g_detect = (n_slots / 4 + 256) / 256;
        int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        const unsigned g_apply = (unsigned)std::min<uint64_t>((nw + 3) / 4, (uint64_t)sms * 2);
        std::vector<uint4> log(BATCH);
        TrainState back{};
        void* tab_mem = nullptr;
        // (re)build the pair table from the words: a full recount into a fresh table of `pcap` slots
        auto rebuild = [&]() -> int {
            if (tab_mem) { cudaFree(tab_mem); tab_mem = nullptr; }
            const size_t n_seg = pcap / SEG + 1;
            TCK(cudaMalloc(&tab_mem, pcap * sizeof(TrainPair) + n_seg * sizeof(Best)));
            pt.e = (TrainPair*)tab_mem; pt.mask = pcap - 1;
            pt.seg_best = reinterpret_cast<Best*>(pt.e + pcap);
            k_table_clear<<<(pcap + 255) / 256, 256, 0, st>>>(pt); ++launches;
            TCK(cudaMemsetAsync(&d_st->pause, 0, 12, st));               // pause, overflow, fill
            k_count_all<<<g_slots, 256, 0, st>>>(d_st, W, pt); ++launches;
            return CTK_OK;
        };
        struct TableGuard { void** a; ~TableGuard() { if (*a) cudaFree(*a); } } tg{&tab_mem};
        { int rc = rebuild(); if (rc != CTK_OK) return rc; }
        // Default: three kernels per merge.  CTK_TRAIN_CLUSTER=16 (or 8) runs a batch of merges inside ONE thread-block
        // cluster instead (hardware cluster barrier between the phases, no kernel boundaries).  Measured on the config-1
        // sample: 22-28 us per merge stepwise, 32-34 us in a cluster of 16 (each phase then has 16 SMs, not 148, for its
        // ~9 MB of L2 traffic), 25.8 us as one cooperative grid, 41.7 us as a CUDA graph of the same 768 kernel nodes.
        int cluster = getenv("CTK_TRAIN_CLUSTER") ? atoi(getenv("CTK_TRAIN_CLUSTER")) : 0;
        if (cluster != 8 && cluster != 16) cluster = 0;
        if (cluster) {
            cudaFuncSetAttribute(k_train_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaLaunchConfig_t qc{}; cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = cluster; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            qc.gridDim = dim3(cluster); qc.blockDim = dim3(1024); qc.attrs = qa; qc.numAttrs = 1;
            int n_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&n_clusters, k_train_cluster, &qc) != cudaSuccess || n_clusters < 1) {
                cluster = 8; qa[0].val.clusterDim.x = cluster; qc.gridDim = dim3(cluster);
                if (cudaOccupancyMaxActiveClusters(&n_clusters, k_train_cluster, &qc) != cudaSuccess || n_clusters < 1) cluster = 0;
            }
            cudaGetLastError();
        }
        out.cluster = cluster;
        uint32_t max_word = 0;
        for (uint32_t i = 0; i < nw; ++i) max_word = std::max(max_word, ws[i].len);
        const bool fused = !cluster && max_word <= FUSED_MAX_WORD && !getenv("CTK_TRAIN_THREE_KERNELS");
        for (;;) {
            const unsigned g_best = std::max(1u, std::min((unsigned)sms * 2, pcap / 1024u));   // few CTAs: one ticket atomic and one candidate each
            if (cluster) {
                cudaLaunchConfig_t lc{}; cudaLaunchAttribute la[1];
                la[0].id = cudaLaunchAttributeClusterDimension; la[0].val.clusterDim.x = cluster; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
                lc.gridDim = dim3(cluster); lc.blockDim = dim3(1024); lc.stream = st; lc.attrs = la; lc.numAttrs = 1;
                TCK(cudaLaunchKernelEx(&lc, k_train_cluster, d_st, W, pt, d_bb, sy, d_log, (int)BATCH));
                launches += 1;
            } else {
                static const bool pdl = !getenv("CTK_TRAIN_NO_PDL");
                cudaLaunchAttribute pa[1];
                pa[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                pa[0].val.programmaticStreamSerializationAllowed = 1;
                auto cfg_of = [&](unsigned grid, unsigned block) {
                    cudaLaunchConfig_t c{};
                    c.gridDim = dim3(grid); c.blockDim = dim3(block); c.stream = st; c.attrs = pa; c.numAttrs = pdl ? 1 : 0;
                    return c;
                };
                const cudaLaunchConfig_t c_detect = cfg_of(g_detect, 256), c_apply = cfg_of(g_apply, 128), c_best = cfg_of(g_best, 256),
                                         c_fused = cfg_of((nw + 255) / 256, 256);
                for (int it = 0; it < BATCH; ++it) {
                    if (fused) TCK(cudaLaunchKernelEx(&c_fused, k_word_merge, d_st, W, pt));
                    else {
                        TCK(cudaLaunchKernelEx(&c_detect, k_detect, d_st, W));
                        TCK(cudaLaunchKernelEx(&c_apply, k_apply, d_st, W, pt));
                    }
                    TCK(cudaLaunchKernelEx(&c_best, k_best_pair, d_st, pt, d_bb, sy, d_log));
                }
                launches += (fused ? 2 : 3) * BATCH;
            }
            TCK(cudaMemcpyAsync(&back, d_st, sizeof back, cudaMemcpyDeviceToHost, st));
            TCK(cudaMemcpyAsync(log.data(), d_log, BATCH * sizeof(uint4), cudaMemcpyDeviceToHost, st));
            TCK(cudaMemsetAsync(&d_st->n_log, 0, 4, st));
            TCK(cudaStreamSynchronize(st));
            // replay on strings: every decision of the device must be the one the reference's string logic takes
            for (uint32_t i = 0; i < back.n_log; ++i) {
                uint32_t l = log[i].x, r = log[i].y, m = log[i].z;
                if (l >= symbols.size() || r >= symbols.size()) { set_last_error("ctk_train_bpe: corrupt merge log"); return CTK_ERR_CUDA; }
                std::string merged = symbols[l] + symbols[r];
                auto f = index.find(merged);
                uint32_t expect = f != index.end() ? f->second : (uint32_t)symbols.size();
                if (expect != m) { set_last_error("ctk_train_bpe: symbol hash collision (device and string replay disagree)"); return CTK_ERR_UNSUPPORTED; }
                uint32_t s = sym(merged);
                uint64_t id = vocab_len;                                    // :168-169: id = vocab.len() before the insert
                vocab_insert(s, id);
                out.merges.push_back(l); out.merges.push_back(r); out.merge_counts.push_back(log[i].w);
            }
            if (back.done) break;
            if (back.pause) {
                // At the largest size a recount alone makes room: live pairs <= slots <= pcap_max / 4.
                pcap = pcap < pcap_max / 4 ? pcap * 4 : pcap_max;
                ++out.rebuilds;
                int rc = rebuild(); if (rc != CTK_OK) return rc;
            }
        }
        if (getenv("CTK_TRAIN_TRACE") && cluster)
            fprintf(stderr, "[ctk train] cluster of %d, thread 0: detect %.1f ms, apply %.1f ms, best pair %.1f ms, barriers (incl. waiting for the slowest CTA) %.1f ms\n", cluster,
                    back.t_phase[0] * 1e-6, back.t_phase[1] * 1e-6, back.t_phase[2] * 1e-6, back.t_phase[3] * 1e-6);
        if (back.reason == 4) { set_last_error("ctk_train_bpe: symbol table full"); return CTK_ERR_CUDA; }
        out.stop_reason = back.reason;
        TCK(cudaEventRecord(ev[3], st));
        TCK(cudaStreamSynchronize(st));
        float ms = 0; cudaEventElapsedTime(&ms, ev[2], ev[3]); out.ms_merges = ms;
    }
    out.kernels = launches;
    g_kernel_launches.fetch_add(launches, std::memory_order_relaxed);
    out.sym_off.assign(1, 0);
    for (auto& s : symbols) { out.sym_bytes.insert(out.sym_bytes.end(), s.begin(), s.end()); out.sym_off.push_back(out.sym_bytes.size()); }
    return CTK_OK;
}

}  // namespace ctk

// ---- train_new_from_iterator (src/huggingface/mod.rs:1231-1275): the texts go through THIS tokenizer's normaliser and
// pre-tokenizer, and the trainer sees the pre-tokens -- for a ByteLevel pipeline: byte-mapped strings without any white space,
// so each pre-token is exactly one of the trainer's words (bpe_trainer.rs:248 splits on white space).  Here the pre-tokens
// never visit the host: the encode path's start bitmap says where they begin, one kernel writes their byte-mapped characters
// (pretokenizers.rs:130-153) with a U+0020 in front of each, and stage W above takes that text where it lies.
namespace ctk {
namespace {
// thread per 32 bytes (one bitmap word): bytes of output this group produces
__global__ void k_mapped_len(const uint8_t* __restrict__ text, uint64_t n, const uint32_t* __restrict__ start_bits, const uint16_t* __restrict__ map2,
                             uint64_t n_groups, uint32_t* __restrict__ glen) {
    const uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g > n_groups) return;
    if (g == n_groups) { glen[g] = 0; return; }
    const uint32_t sb = start_bits[g];
    uint32_t len = __popc(sb);
    for (int k = 0; k < 32; ++k) {
        const uint64_t i = g * 32 + k;
        if (i < n) len += (map2[text[i]] >> 8) ? 2u : 1u;
    }
    glen[g] = len;
}
__global__ void k_mapped_write(const uint8_t* __restrict__ text, uint64_t n, const uint32_t* __restrict__ start_bits, const uint16_t* __restrict__ map2,
                               uint64_t n_groups, const uint64_t* __restrict__ gpos, uint8_t* __restrict__ out) {
    const uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint32_t sb = start_bits[g];
    uint64_t o = gpos[g];
    for (int k = 0; k < 32; ++k) {
        const uint64_t i = g * 32 + k;
        if (i >= n) break;
        if ((sb >> k) & 1u) out[o++] = ' ';
        const uint32_t m = map2[text[i]];
        out[o++] = (uint8_t)m;
        if (m >> 8) out[o++] = (uint8_t)(m >> 8);
    }
}
struct U32To64 { __host__ __device__ uint64_t operator()(uint32_t v) const { return v; } };
}  // namespace

// normalise + pre-tokenise on the device -> one text of byte-mapped pre-tokens separated by spaces (device memory of eng.ws)
static int pretokens_as_text(Engine& eng, const uint8_t* h_text, const uint64_t* h_off, size_t n_texts, const uint8_t** d_out, uint64_t* out_bytes) {
#define CKE(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return eng.cuda_fail(e_, #x); } while (0)
    cudaStream_t st = eng.st_comp;
    const uint64_t n = n_texts ? h_off[n_texts] : 0;
    *d_out = nullptr; *out_bytes = 0;
    if (eng.model.metaspace) return eng.fail(CTK_ERR_UNSUPPORTED, "train_new_from_iterator is not built for Metaspace pipelines");
    if (n == 0) return CTK_OK;
    if (n >= 0x7FFFFFF0ull) return eng.fail(CTK_ERR_UNSUPPORTED, "train_new_from_iterator: one call takes less than 2 GiB of text");
    Workspace& ws = eng.ws;
    uint8_t* d_text; uint64_t* d_off;
    CKE(ws.get(64, n + 128, (void**)&d_text));
    CKE(ws.get(65, (n_texts + 2) * 8, (void**)&d_off));
    CKE(cudaMemcpyAsync(d_text, h_text, n, cudaMemcpyHostToDevice, st));
    CKE(cudaMemsetAsync(d_text + n, 0, 64, st));
    CKE(cudaMemcpyAsync(d_off, h_off, (n_texts + 1) * 8, cudaMemcpyHostToDevice, st));
    const uint8_t* t; const uint64_t* o; uint64_t b; size_t np = n_texts;
    int rc = nfc_stage(eng, d_text, d_off, n_texts, n, &t, &o, &b, st);
    if (rc != CTK_OK) return rc;
    const uint64_t* first_piece;
    rc = split_stages(eng, t, o, n_texts, b, &t, &o, &np, &b, &first_piece, st);
    if (rc != CTK_OK) return rc;
    rc = prefix_space_stage(eng, t, o, np, b, &t, &o, &b, st);
    if (rc != CTK_OK) return rc;
    if (b == 0) return CTK_OK;
    const uint64_t n_groups = (b + 31) / 32;
    const uint32_t n_blocks = (uint32_t)((n_groups + 255) / 256);
    uint32_t *ds, *sb, *bc, *err, *glen; uint64_t* gpos; uint8_t* out; void* tmp;
    CKE(ws.get(66, (n_groups + 2) * 4, (void**)&ds));
    CKE(ws.get(67, (n_groups + 2) * 4, (void**)&sb));
    CKE(ws.get(68, ((uint64_t)n_blocks + 1) * 4, (void**)&bc));
    CKE(ws.get(69, 256, (void**)&err));
    CKE(ws.get(70, (n_groups + 2) * 4, (void**)&glen));
    CKE(ws.get(71, (n_groups + 2) * 8, (void**)&gpos));
    CKE(cudaMemsetAsync(err, 0, 256, st));
    rc = starts_bitmap(eng, t, o, np, b, ds, sb, bc, err, st);
    if (rc != CTK_OK) return rc;
    k_mapped_len<<<(unsigned)((n_groups + 1 + 255) / 256), 256, 0, st>>>(t, b, sb, eng.rich.byte_map2, n_groups, glen);
    cub::TransformInputIterator<uint64_t, U32To64, const uint32_t*> it(glen, U32To64());
    size_t tb = 0;
    CKE(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, gpos, n_groups + 1, st));
    CKE(ws.get(5, tb + 16, &tmp));
    CKE(cub::DeviceScan::ExclusiveSum(tmp, tb, it, gpos, n_groups + 1, st));
    CKE(eng.publish({{err, 1, 0}, {gpos + n_groups, 2, 2}}, st));
    CKE(cudaStreamSynchronize(st));
    if (eng.h_flags[0] & ERRF_OFFSETS) return eng.fail(CTK_ERR_ARG, "offsets must start at 0, be non-decreasing and end at the buffer length");
    uint64_t total;
    memcpy(&total, eng.h_flags + 2, 8);
    CKE(ws.get(72, total + 128, (void**)&out));
    k_mapped_write<<<(unsigned)((n_groups + 255) / 256), 256, 0, st>>>(t, b, sb, eng.rich.byte_map2, n_groups, gpos, out);
    eng.launched(5);
    CKE(cudaGetLastError());
    CKE(cudaStreamSynchronize(st));
    *d_out = out; *out_bytes = total;
    return CTK_OK;
#undef CKE
}
}  // namespace ctk

extern "C" {

struct ctk_trained { ctk::Trained t; };

int ctk_train_bpe(const ctk_bpe_trainer_config* cfg, int device, const uint8_t* text, const uint64_t* text_off, size_t n_texts,
                  ctk_trained** out) {
    if (!cfg || !out || (n_texts && (!text_off || (!text && text_off[n_texts])))) { ctk::set_last_error("null argument"); return CTK_ERR_ARG; }
    if (cfg->n_special && (!cfg->special_tokens || !cfg->special_off)) { ctk::set_last_error("special tokens missing"); return CTK_ERR_ARG; }
    *out = nullptr;
    ctk_trained* res = new (std::nothrow) ctk_trained();
    if (!res) { ctk::set_last_error("out of memory"); return CTK_ERR_CUDA; }
    int rc;
    try { rc = ctk::train_impl(*cfg, device, text, text_off, n_texts, res->t); }
    catch (const std::bad_alloc&) { ctk::set_last_error("out of host memory"); rc = CTK_ERR_CUDA; }
    if (rc != CTK_OK) { delete res; return rc; }
    *out = res;
    return CTK_OK;
}

size_t ctk_trained_symbols(const ctk_trained* t, const uint8_t** bytes, const uint64_t** off, const int64_t** vocab_id) {
    if (bytes) *bytes = t->t.sym_bytes.data();
    if (off) *off = t->t.sym_off.data();
    if (vocab_id) *vocab_id = t->t.vocab_id.data();
    return t->t.symbols.size();
}

const uint32_t* ctk_trained_merge_counts(const ctk_trained* t) { return t->t.merge_counts.data(); }

size_t ctk_trained_merges(const ctk_trained* t, const uint32_t** pairs) {
    if (pairs) *pairs = t->t.merges.data();
    return t->t.merges.size() / 2;
}

void ctk_trained_stats(const ctk_trained* t, ctk_train_stats* s) {
    s->n_bytes = t->t.n_bytes; s->n_words = t->t.n_words; s->n_unique_words = t->t.n_unique; s->n_symbols = t->t.n_symbols0;
    s->n_merges = t->t.merges.size() / 2; s->kernel_launches = t->t.kernels; s->stop_reason = t->t.stop_reason; s->table_rebuilds = t->t.rebuilds; s->cluster_size = (uint32_t)t->t.cluster;
    s->ms_words = t->t.ms_words; s->ms_words_kernels = t->t.ms_words_kernels; s->ms_merges = t->t.ms_merges;
}

void ctk_trained_free(ctk_trained* t) { delete t; }

}  // extern "C"

extern "C" int ctk_train_new_from_texts(const ctk_tokenizer* tok, const ctk_bpe_trainer_config* cfg, const uint8_t* text, const uint64_t* text_off,
                                        size_t n_texts, ctk_trained** out) {
    if (!tok || !cfg || !out || (n_texts && (!text_off || (!text && text_off[n_texts])))) { ctk::set_last_error("null argument"); return CTK_ERR_ARG; }
    if (cfg->n_special && (!cfg->special_tokens || !cfg->special_off)) { ctk::set_last_error("special tokens missing"); return CTK_ERR_ARG; }
    *out = nullptr;
    ctk::Engine* eng = const_cast<ctk::Engine*>(reinterpret_cast<const ctk::Engine*>(tok));
    std::lock_guard<std::mutex> lk(eng->mu);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(eng->device);
    if (e != cudaSuccess) return eng->cuda_fail(e, "cudaSetDevice");
    struct DeviceGuard { int d; ~DeviceGuard() { cudaSetDevice(d); } } dguard{prev};
    const uint8_t* d_words = nullptr;
    uint64_t n_bytes = 0;
    int rc = ctk::pretokens_as_text(*eng, text, text_off, n_texts, &d_words, &n_bytes);
    if (rc != CTK_OK) return rc;
    ctk_trained* res = new (std::nothrow) ctk_trained();
    if (!res) { ctk::set_last_error("out of memory"); return CTK_ERR_CUDA; }
    const uint64_t one_off[2] = {0, n_bytes};
    try { rc = ctk::train_impl(*cfg, eng->device, d_words, one_off, n_bytes ? 1 : 0, res->t); }
    catch (const std::bad_alloc&) { ctk::set_last_error("out of host memory"); rc = CTK_ERR_CUDA; }
    if (rc != CTK_OK) { delete res; return rc; }
    *out = res;
    return CTK_OK;
}

