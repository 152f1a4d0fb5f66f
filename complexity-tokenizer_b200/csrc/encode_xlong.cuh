// Block-level / grid-level path for VERY long pre-tokens (> XL_MIN bytes: whitespace or punctuation runs,
// separator-free blobs; BASELINE config 4 has single pre-tokens of 2^20 symbols).  Included by encode_fused.cu.
//
// The reference merges one pair at a time, lowest rank first, leftmost on ties (bpe.rs:104-153): O(n^2).
// For a MONOTONE merge table (loader.cpp; every table a BPE trainer emits) the same result is reached in
// rounds that merge many pairs at once.  Pair j = (sym[j], sym[j+1]), rank r[j], W = longest token in
// initial symbols:
//     blocked(j)   some pair within W symbols of j has a strictly lower rank
//     run          r[j-1] == r[j]  (then sym[j-1] == sym[j] == sym[j+1]);  s = first pair of the run
//     selected(j)  outside a run: !blocked(j);   in a run: (j - s) even and no pair of s..j is blocked
// Why it is exact: let F be the merge forest of the sequential run.  Every executed merge is in F (induction).
// If x = sym[j] were consumed in F by anything but (x, y), the consuming token spans <= W symbols around x and its
// subtree holds an unexecuted merge whose operands exist now, with a lower rank than r[j] (monotone tables run
// in non-decreasing rank order, and a token's inner merges rank below every pair that uses it): pair j would be
// blocked.  Equal ranks are one pair type (a, a); the sequential order pairs a run greedily from its left end,
// which is the parity rule.  tools/verify_window_merge.py fuzzes the claim (0 / 30 000 mismatches; it breaks
// as expected for non-monotone tables, which therefore keep the sequential path of encode_long.cuh).
//
// All very long pre-tokens of a call are laid out in ONE symbol array X (each followed by a separator that
// carries its list index) and every round is a handful of grid-wide kernels:
//     k_xl_rank    pair ranks (merge-table probes), window test in shared memory -> blocked, run heads (1 byte per pair)
//     scan 1       segmented scan of those bytes: parity of the run's start and "blocked so far in the run"
//     scan 2       new positions (the kept flag is computed on the fly);  k_xl_scatter: the next X
// until a round selects nothing.  k_xl_count then gives every region's id count to its slice, and after the
// scan of the slice counts k_xl_place writes the ids straight to their place in the packed output.
#pragma once
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

namespace ctk {

constexpr uint32_t XL_MAX_WINDOW = 1024;         // longest token (in symbols) the tile kernel's halo supports
constexpr uint32_t XL_SEP = 0x80000000u;         // separator symbol: XL_SEP | index into the xl list
constexpr int XL_THREADS = 256, XL_ITEMS = 8, XL_TILE = XL_THREADS * XL_ITEMS;

__device__ __forceinline__ bool xl_is_sym(uint32_t s) { return (s & XL_SEP) == 0; }

// X[xoff .. xoff+len) = initial ids of the pre-token (kNone where the byte has no vocab entry: dropped by the
// first round, bpe.rs:94-97), X[xoff+len] = separator
__global__ void __launch_bounds__(256) k_xl_init(const FusedParams p, const XlEntry* __restrict__ list, uint32_t n_list,
                                                 uint32_t* __restrict__ xs, uint32_t* __restrict__ holes) {
    __shared__ uint32_t s_init[256];
    s_init[threadIdx.x] = __ldg(p.t.byte_init + threadIdx.x);
    __syncthreads();
    bool hole = false;
    for (uint32_t e = blockIdx.y; e < n_list; e += gridDim.y) {
        const LongDesc dd = p.desc[list[e].desc];
        const uint8_t* src = p.text + dd.gstart;
        uint32_t* dst = xs + list[e].xoff;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= dd.len; i += (uint64_t)gridDim.x * blockDim.x) {
            uint32_t v = i < dd.len ? s_init[__ldg(src + i)] : (XL_SEP | e);
            hole = hole || v == kNone;
            dst[i] = v;
        }
    }
    if (hole) atomicOr(holes, 1u);
}

// per pair: rank, new id, and the one-byte scan element: bit 2 = run head, bit 1 = parity of the head's index, bit 0 = blocked
// shared memory: two arrays of XL_TILE + 2 W ranks (tile + halo of W pairs on each side), ping-ponged by the
// window-minimum doubling
__global__ void __launch_bounds__(XL_THREADS) k_xl_rank(const DevTables t, const uint32_t* __restrict__ xs, uint32_t n, uint32_t W,
                                                        int no_merge, uint32_t* __restrict__ rank, uint32_t* __restrict__ newid,
                                                        uint8_t* __restrict__ scanv) {
    extern __shared__ uint32_t sm[];
    const int iW = (int)W, span = XL_TILE + 2 * iW, tid = threadIdx.x;
    uint32_t* A = sm;
    uint32_t* B = sm + span;
    const long long t0 = (long long)blockIdx.x * XL_TILE, base = t0 - iW;
    auto probe = [&](long long i, uint32_t& v) -> uint32_t {   // no pair across a separator / a hole / the array ends
        v = 0;
        if (no_merge || i < 0 || i + 1 >= (long long)n) return kNone;
        const uint32_t a = xs[i], b = xs[i + 1];
        if (!xl_is_sym(a) || !xl_is_sym(b)) return kNone;
        const uint2 q = pair_lookup(t, a, b);
        v = q.y;
        return q.x;
    };
    uint32_t my_r[XL_ITEMS], my_v[XL_ITEMS], left_r[XL_ITEMS];
#pragma unroll
    for (int q = 0; q < XL_ITEMS; ++q) {
        const int own = q * XL_THREADS + tid;
        my_r[q] = probe(t0 + own, my_v[q]);
        A[iW + own] = my_r[q];
    }
    for (int j = tid; j < iW; j += XL_THREADS) {
        uint32_t v;
        A[j] = probe(base + j, v);
        A[iW + XL_TILE + j] = probe(t0 + XL_TILE + j, v);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < XL_ITEMS; ++q) left_r[q] = A[iW + q * XL_THREADS + tid - 1];     // W >= 1
    if (t.round_parallel) {
        // per-token reach (encode_long.cuh): pair (x, y) looks left as far as a token ending with x can extend and right
        // as far as a token starting with y can extend, counted in SYMBOLS here (a symbol is at least one initial
        // symbol long, so this window contains the exact one).  Three times fewer rounds than the uniform window W.
#pragma unroll
        for (int q = 0; q < XL_ITEMS; ++q) {
            const int own = q * XL_THREADS + tid;
            const long long i = t0 + own;
            if (i >= (long long)n) continue;
            const uint32_t r = my_r[q];
            bool blocked = false;
            if (r != kNone) {
                const uint32_t wl = min(__ldg(t.reach + xs[i]) & 0xFFFFu, W), wr = min(__ldg(t.reach + xs[i + 1]) >> 16, W);
                const int j = own + iW;
                for (uint32_t d = 1; d <= max(wl, wr); ++d) {
                    const uint32_t a = d <= wl ? A[j - (int)d] : kNone, b = d <= wr ? A[j + (int)d] : kNone;
                    if (min(a, b) < r) { blocked = true; break; }
                }
            }
            const bool head = r == kNone || left_r[q] != r;
            rank[i] = r; newid[i] = my_v[q];
            scanv[i] = (uint8_t)((head ? 4u | (((uint32_t)i & 1u) << 1) : 0u) | (blocked ? 1u : 0u));
        }
        return;
    }
    // uniform window W: left-aligned window minima by doubling: with width w done, A[j] = min of pairs j .. j + w - 1
    uint32_t w = 1;
    while (w * 2 <= 2 * W + 1) {
        for (int j = tid; j < span; j += XL_THREADS) {
            const int k = j + (int)w;
            B[j] = min(A[j], k < span ? A[k] : kNone);
        }
        __syncthreads();
        uint32_t* T = A; A = B; B = T;
        w *= 2;
    }
#pragma unroll
    for (int q = 0; q < XL_ITEMS; ++q) {
        const int own = q * XL_THREADS + tid;
        const long long i = t0 + own;
        if (i >= (long long)n) continue;
        const int j = own + iW;                                // its window = pairs j - W .. j + W of the array
        const uint32_t wmin = min(A[j - iW], A[j + iW - (int)w + 1]);
        const uint32_t r = my_r[q];
        const bool blocked = wmin < r;
        const bool head = r == kNone || left_r[q] != r;
        rank[i] = r; newid[i] = my_v[q];
        scanv[i] = (uint8_t)((head ? 4u | (((uint32_t)i & 1u) << 1) : 0u) | (blocked ? 1u : 0u));
    }
}

struct XlSegOp {                                               // segmented scan: (head seen, parity of the run's first index, blocked so far)
    __device__ __forceinline__ uint8_t operator()(uint8_t a, uint8_t b) const {
        return (b & 4u) ? b : (uint8_t)((a & 6u) | ((a | b) & 1u));
    }
};

__device__ __forceinline__ bool xl_selected(uint32_t i, uint32_t r, uint32_t sv) {
    if (r == kNone || (sv & 1u)) return false;
    return ((i ^ (sv >> 1)) & 1u) == 0;                        // even distance from the first pair of its run
}

// symbol i survives the round unless the pair to its left was selected (or it is a hole); computed on the fly for the
// position scan, so that no per-symbol flag array goes through HBM
struct XlKeepFn {
    const uint32_t* xs; const uint32_t* rank; const uint8_t* sv;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        if (xs[i] == kNone) return 0u;
        return (i > 0 && xl_selected(i - 1, rank[i - 1], sv[i - 1])) ? 0u : 1u;
    }
};

__global__ void __launch_bounds__(256) k_xl_scatter(const uint32_t* __restrict__ xs, const uint32_t* __restrict__ rank,
                                                    const uint32_t* __restrict__ newid, const uint8_t* __restrict__ sv,
                                                    const uint32_t* __restrict__ pos, uint32_t n,
                                                    uint32_t* __restrict__ out, uint32_t* __restrict__ n_out) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    const uint32_t x = xs[i];
    const bool keep = x != kNone && !(i > 0 && xl_selected(i - 1, rank[i - 1], sv[i - 1]));
    if (keep) out[pos[i]] = xl_selected(i, rank[i], sv[i]) ? newid[i] : x;
    if (i == n - 1) *n_out = pos[i] + (keep ? 1u : 0u);
}

// after the last round: separator positions (via their rank among separators), then every id to the long pool
__global__ void __launch_bounds__(256) k_xl_seps(const uint32_t* __restrict__ xs, uint32_t n, uint32_t* __restrict__ sep_pos) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = xs[i];
    if (!xl_is_sym(s) && s != kNone) sep_pos[s & ~XL_SEP] = i;   // regions keep their order: list index e <-> e-th separator
}

// region e = symbols between separator e-1 and separator e: its id count goes to its LongDesc and its slice
__global__ void __launch_bounds__(256) k_xl_count(const FusedParams p, const XlEntry* __restrict__ list, uint32_t n_list,
                                                  const uint32_t* __restrict__ sep_pos) {
    const uint32_t e = blockIdx.x * 256u + threadIdx.x;
    if (e >= n_list) return;
    const uint32_t start = e ? sep_pos[e - 1] + 1 : 0, cnt = sep_pos[e] - start;
    LongDesc* dd = p.desc + list[e].desc;
    dd->cnt = cnt;
    atomicAdd(p.slice_cnt + dd->slice, cnt);
}

// where region e's ids go in the packed output: base of its slice + ids of the slice's run before the pre-token
// + ids of the slice's earlier long pre-tokens (same arithmetic as k_compact)
__global__ void __launch_bounds__(256) k_xl_dst(const FusedParams p, const XlEntry* __restrict__ list, uint32_t n_list,
                                                const uint32_t* __restrict__ slice_base, uint32_t* __restrict__ region_dst) {
    const uint32_t e = blockIdx.x * 256u + threadIdx.x;
    if (e >= n_list) return;
    const uint32_t di = list[e].desc;
    const LongDesc* dd = p.desc + di;
    uint32_t before = 0;
    for (uint32_t q = p.slice_desc[dd->slice]; q < di; ++q) before += p.desc[q].cnt;
    region_dst[e] = slice_base[dd->slice] + (dd->k_at & 0xFFFFu) + before;
}

// every id of the final X straight to its place in the packed output (k_compact leaves these ranges alone)
template <typename out_t>
__global__ void __launch_bounds__(256) k_xl_place(const uint32_t* __restrict__ xs, uint32_t n, uint32_t n_list,
                                                  const uint32_t* __restrict__ sep_pos, const uint32_t* __restrict__ region_dst,
                                                  out_t* __restrict__ out, uint64_t out_cap) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    uint32_t lo = 0, hi = n_list - 1;                          // region of i: first separator at or after i
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sep_pos[mid] >= i) hi = mid; else lo = mid + 1; }
    if (sep_pos[lo] == i) return;
    const uint32_t start = lo ? sep_pos[lo - 1] + 1 : 0;
    const uint64_t dst = (uint64_t)region_dst[lo] + (i - start);
    if (dst < out_cap) out[dst] = (out_t)xs[i];                // a too small output was flagged by k_compact
}

}  // namespace ctk
