// decode_batch on the GPU.
//
//   k_dec_tile_sums / scan / k_dec_write   ids -> token byte strings -> concatenation per document
//                                      (mod.rs:717-735 filter + lookup, decoders.rs:94-116 with the
//                                      char->byte map folded into the per-token blob at load)
//   k_dec_valid                       is every document's byte string valid UTF-8?  (fast path: the
//                                      gather output already IS String::from_utf8_lossy's result)
//   k_dec_post                        per document: from_utf8_lossy (decoders.rs:118) and
//                                      clean_up_tokenization_spaces (mod.rs:749-769), exact sequential
//                                      semantics (15 ordered str::replace, split_whitespace().join(" "))
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "engine.hpp"

namespace ctk {

struct U32ToU64 { __host__ __device__ uint64_t operator()(uint32_t v) const { return v; } };

// ---- ids -> bytes in two passes over the ids, no per-token offsets in global memory -------------------------------
// A tile = DT consecutive tokens = one CTA.  Pass 1 sums the tile's token lengths; a scan of the (few) tile sums gives
// every tile its place in the output.  Pass 2 re-derives the lengths, scans them inside the CTA, lays the tile's bytes
// out in shared memory and writes them with aligned 16-byte stores; the documents that start inside the tile get their
// byte offsets on the way.  (The first version wrote a uint64 offset per token -- 2 bytes of scratch traffic per
// output byte -- and one thread copied each token's bytes to global memory.)
constexpr int DT = 2048, DTH = 256, DPT = DT / DTH;
constexpr int DSTAGE = 16 * 1024;                        // bytes of a tile staged in shared memory (a tile averages ~9 KB); the less shared memory, the more L1 for the token table
static_assert(DPT == 8, "a thread takes its eight ids with two 16-byte loads");

__device__ __forceinline__ uint32_t dec_tok_len(const DecodeTables& t, uint32_t id, int skip_special) {
    if (id >= t.n_ids) return 0;                                       // unknown ids are dropped (mod.rs:717-735)
    const uint32_t L = __ldg((skip_special ? t.len8_skip : t.len8) + id);
    return L < 255 ? L : __ldg(t.off + id + 1) - __ldg(t.off + id);    // (a special token is never that long)
}

// ids of the thread's DPT consecutive tokens: two aligned 16-byte loads (ids is 16-byte aligned when it comes from this
// library or from cudaMalloc; otherwise, and in the tile that holds the end, one by one); 0xFFFFFFFF beyond the end
__device__ __forceinline__ void dec_load_ids(const uint32_t* __restrict__ ids, uint64_t n, uint64_t j0, bool vec_ok, uint32_t (&id)[DPT]) {
    if (vec_ok && j0 + DPT <= n) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(ids + j0)), b = __ldg(reinterpret_cast<const uint4*>(ids + j0) + 1);
        id[0] = a.x; id[1] = a.y; id[2] = a.z; id[3] = a.w; id[4] = b.x; id[5] = b.y; id[6] = b.z; id[7] = b.w;
    } else {
#pragma unroll
        for (int q = 0; q < DPT; ++q) id[q] = j0 + q < n ? __ldg(ids + j0 + q) : 0xFFFFFFFFu;
    }
}

// pass 1: 128 threads per tile, 16 ids each (four 16-byte loads and sixteen 1-byte gathers in flight per thread)
constexpr int DTH1 = DT / (2 * DPT);
__global__ void __launch_bounds__(DTH1) k_dec_tile_sums(DecodeTables t, const uint32_t* __restrict__ ids, uint64_t n, int skip_special,
                                                        uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t s_part[DTH1 / 32];
    const uint64_t t0 = (uint64_t)blockIdx.x * DT;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(ids) & 15) == 0;
    uint32_t ida[DPT], idb[DPT], sum = 0;
    dec_load_ids(ids, n, t0 + (uint64_t)threadIdx.x * 2 * DPT, vec_ok, ida);
    dec_load_ids(ids, n, t0 + (uint64_t)threadIdx.x * 2 * DPT + DPT, vec_ok, idb);
    uint32_t la[DPT], lb[DPT];
#pragma unroll
    for (int q = 0; q < DPT; ++q) { la[q] = dec_tok_len(t, ida[q], skip_special); lb[q] = dec_tok_len(t, idb[q], skip_special); }   // (0xFFFFFFFF is no id: length 0)
#pragma unroll
    for (int q = 0; q < DPT; ++q) sum += la[q] + lb[q];
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < DTH1 / 32; ++w) tot += s_part[w];
        tile_sum[blockIdx.x] = tot;
    }
}

// pass 1 with the 1-byte length table in SHARED memory: a gather from global memory costs the load/store unit one wavefront
// per distinct cache line (about twenty per warp for these ids), a gather from shared memory three or four.  Persistent CTAs
// (the table is loaded once per CTA), one tile of DT ids per iteration.
__global__ void __launch_bounds__(DTH) k_dec_tile_sums_smem(DecodeTables t, const uint32_t* __restrict__ ids, uint64_t n, int skip_special,
                                                            uint32_t* __restrict__ tile_sum, uint32_t n_tiles) {
    extern __shared__ __align__(16) uint8_t s_len[];                   // t.n_ids bytes, rounded up to 16
    __shared__ uint32_t s_part[2][DTH / 32];
    {
        const uint4* src = reinterpret_cast<const uint4*>(skip_special ? t.len8_skip : t.len8);   // (padded to a multiple of 16 at upload)
        for (uint32_t v = threadIdx.x; 16 * v < t.n_ids; v += DTH) reinterpret_cast<uint4*>(s_len)[v] = __ldg(src + v);
    }
    __syncthreads();
    const bool vec_ok = (reinterpret_cast<uintptr_t>(ids) & 15) == 0;
    int par = 0;
    uint32_t nxt[DPT];                                                 // the next tile's ids are in flight while this tile's lengths are summed
    if (blockIdx.x < n_tiles) dec_load_ids(ids, n, (uint64_t)blockIdx.x * DT + (uint64_t)threadIdx.x * DPT, vec_ok, nxt);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, par ^= 1) {
        uint32_t id[DPT], sum = 0;
#pragma unroll
        for (int q = 0; q < DPT; ++q) id[q] = nxt[q];
        if (tile + gridDim.x < n_tiles) dec_load_ids(ids, n, (uint64_t)(tile + gridDim.x) * DT + (uint64_t)threadIdx.x * DPT, vec_ok, nxt);
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
            uint32_t L = id[q] < t.n_ids ? s_len[id[q]] : 0u;
            if (L == 255u) L = __ldg(t.off + id[q] + 1) - __ldg(t.off + id[q]);
            sum += L;
        }
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        if ((threadIdx.x & 31) == 0) s_part[par][threadIdx.x >> 5] = sum;
        __syncthreads();                                               // (two buffers: the next tile's partial sums do not wait for this read)
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < DTH / 32; ++w) tot += s_part[par][w];
            tile_sum[tile] = tot;
        }
    }
}

// first_doc[tile] = first document d with ids_off[d] >= tile * DT (one thread per document: k_dec_write then finds the
// documents that start inside its tile with one load instead of a binary search of 20 dependent loads per tile)
__global__ void k_dec_first_doc(const uint64_t* __restrict__ ids_off, uint64_t n_docs, uint64_t n_tiles, uint32_t* __restrict__ first_doc) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    const uint64_t hi = ids_off[d] / DT;                                           // tiles [lo, hi] start at or before ids_off[d] ...
    const uint64_t lo = d ? ids_off[d - 1] / DT + 1 : 0;                           // ... and after ids_off[d - 1]
    for (uint64_t tl = lo; tl <= hi && tl <= n_tiles; ++tl) first_doc[tl] = (uint32_t)d;
}

// Persistent CTAs, one tile of DT tokens per iteration.  The records of the first `n_hot` ids live in shared memory (a
// byte-level BPE vocabulary is ordered by frequency: single bytes, then merges in the order they were learnt, so the low ids
// are most of the occurrences): a gather from shared memory costs the load/store unit a third of what a gather from global
// memory does (one wavefront per distinct cache line), and that unit is what bounds this kernel.
__global__ void __launch_bounds__(DTH) k_dec_write(DecodeTables t, const uint32_t* __restrict__ ids, uint64_t n, int skip_special,
                                                   const uint64_t* __restrict__ tile_base, const uint64_t* __restrict__ ids_off,
                                                   const uint32_t* __restrict__ first_doc, uint32_t n_tiles, uint32_t n_hot,
                                                   uint64_t n_docs, uint8_t* __restrict__ out, uint64_t out_cap,
                                                   uint64_t* __restrict__ raw_off, uint32_t* __restrict__ err, uint32_t* __restrict__ non_ascii) {
    extern __shared__ __align__(16) uint4 s_hot[];        // records of ids 0 .. n_hot - 1
    __shared__ uint32_t s_toff[DTH + 1];                  // byte offset inside the tile of every thread's first token
    __shared__ uint32_t s_warp[DTH / 32];
    __shared__ __align__(16) uint8_t s_stage[DSTAGE + 48];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (uint32_t v = tid; v < n_hot; v += DTH) s_hot[v] = __ldg(t.rec + v);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();                                                   // the hot records are loaded / the previous tile is done with the shared arrays
    const uint64_t t0 = (uint64_t)tile * DT, base = tile_base[tile];
    const uint32_t total = (uint32_t)(tile_base[tile + 1] - base);
    if (base + total > out_cap) { if (tid == 0) atomicOr(err, ERRF_CAPACITY); continue; }
    const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(out) + base) & 15u);
    const bool staged = total + shift <= (uint32_t)DSTAGE;
    // the stage is filled by OR-ing whole words into it: it starts at zero (a token's last word may reach 15 bytes past the tile)
    if (staged) for (uint32_t v = tid; 16 * v < total + shift + 20; v += DTH) *reinterpret_cast<uint4*>(s_stage + 16 * v) = make_uint4(0, 0, 0, 0);
    // One 16-byte record per token {first 12 bytes, length}: ONE gather gives this pass both the length and the bytes.
    // The thread's DPT consecutive tokens, then an exclusive scan of their lengths over the CTA.
    uint32_t id[DPT], len[DPT], mine = 0;
    uint4 rec[DPT];
    dec_load_ids(ids, n, t0 + (uint64_t)tid * DPT, (reinterpret_cast<uintptr_t>(ids) & 15) == 0, id);
#pragma unroll
    for (int q = 0; q < DPT; ++q)                                     // unknown ids are dropped (mod.rs:717-735)
        rec[q] = id[q] < n_hot ? s_hot[id[q]] : (id[q] < t.n_ids ? __ldg(t.rec + id[q]) : make_uint4(0, 0, 0, 0));
    if (skip_special) {
#pragma unroll
        for (int q = 0; q < DPT; ++q) if (id[q] < t.n_ids && __ldg(t.special + id[q])) rec[q] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int q = 0; q < DPT; ++q) { len[q] = rec[q].w; mine += len[q]; }
    uint32_t incl = mine;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t my_off = incl - mine;
    for (int w = 0; w < wid; ++w) my_off += s_warp[w];
    s_toff[tid] = my_off;
    if (tid == DTH - 1) s_toff[DTH] = my_off + mine;
    if (staged) {
        // A record is zero beyond its token's length, so it can be laid over the zeroed stage as it is: shifted to the token's
        // byte position it spans four aligned words, each OR-ed in with one shared-memory reduction (only where it is not
        // zero: most tokens touch two words).  Neighbouring tokens only overlap in bytes that are zero in one of them.
        // A dozen instructions per token, none of them a branch on the data.
        const uint32_t stage_a = (uint32_t)__cvta_generic_to_shared(s_stage);
        uint32_t o = shift + my_off;                                   // byte position of the token in the stage
        bool any_long = false;
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
            const uint32_t sh = 8u * (o & 3u), wa = stage_a + (o & ~3u);
            const uint32_t w0 = rec[q].x << sh, w1 = __funnelshift_l(rec[q].x, rec[q].y, sh), w2 = __funnelshift_l(rec[q].y, rec[q].z, sh),
                           w3 = __funnelshift_l(rec[q].z, 0u, sh);
            asm volatile("{ .reg .pred p0, p1, p2, p3;\n\t"
                         "setp.ne.u32 p0, %1, 0; setp.ne.u32 p1, %2, 0; setp.ne.u32 p2, %3, 0; setp.ne.u32 p3, %4, 0;\n\t"
                         "@p0 red.shared.or.b32 [%0], %1; @p1 red.shared.or.b32 [%0+4], %2; @p2 red.shared.or.b32 [%0+8], %3; @p3 red.shared.or.b32 [%0+12], %4; }"
                         :: "r"(wa), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
            any_long = any_long || len[q] > 12;
            o += len[q];
        }
        if (any_long) {                                                // bytes 12.. of longer tokens (rare): one by one from the blob
            o = shift + my_off;
            for (int q = 0; q < DPT; ++q) {
                const uint32_t L = len[q];
                if (L > 12) {
                    const uint8_t* src = t.blob + __ldg(t.off + id[q]);
                    for (uint32_t k = 12; k < L; ++k)
                        atomicOr(reinterpret_cast<uint32_t*>(s_stage) + ((o + k) >> 2), (uint32_t)__ldg(src + k) << (8u * ((o + k) & 3u)));
                }
                o += L;
            }
        }
    } else {
        // a tile of unusually long tokens: every thread writes its own tokens byte by byte, straight to global memory
        uint8_t* d = out + base + my_off;
        for (int q = 0; q < DPT; ++q) {
            const uint32_t L = len[q];
            if (L == 0) continue;
            const uint8_t* src = t.blob + __ldg(t.off + id[q]);
            for (uint32_t k = 0; k < L; ++k) d[k] = __ldg(src + k);
            d += L;
        }
        if (tid == 0 && total) atomicOr(non_ascii, 1u);                 // not looked at here: let the validation kernel decide
    }
    __syncthreads();                                                   // s_toff and the stage are complete
    if (staged && total) {
        uint8_t* const g = out + base;
        const uint32_t head = min(total, (16u - shift) & 15u);            // bytes before the first aligned 16-byte group
        uint32_t hib = 0;
        if ((uint32_t)tid < head) { const uint8_t c = s_stage[shift + tid]; g[tid] = c; hib |= c; }
        const uint32_t body = (total - head) >> 4;
        for (uint32_t v = tid; v < body; v += DTH) {
            const uint4 x = *reinterpret_cast<const uint4*>(s_stage + shift + head + 16 * v);
            *reinterpret_cast<uint4*>(g + head + 16 * v) = x;
            hib |= (x.x | x.y) | (x.z | x.w);
        }
        const uint32_t done = head + 16 * body;
        if ((uint32_t)tid < total - done) { const uint8_t c = s_stage[shift + done + tid]; g[done + tid] = c; hib |= c; }
        // pure ASCII output is valid UTF-8 whatever the document cuts: only a tile with a byte >= 0x80 asks for the validation pass
        if (__any_sync(0xFFFFFFFFu, (hib & 0x80808080u) != 0) && lane == 0) atomicOr(non_ascii, 1u);
    }
    // byte offsets of the documents whose first token lies in this tile (and, in the last tile, of those at the very end):
    // the offset of the thread that holds the token, plus the lengths of that thread's tokens before it (looked up again:
    // there are one or two documents per tile)
    const uint64_t t1 = t0 + DT < n ? t0 + DT : n;
    const bool last = t0 + DT >= n;
    for (uint64_t d = (uint64_t)first_doc[tile] + tid; d <= n_docs; d += DTH) {
        const uint64_t j = ids_off[d];
        if (j < t1 || (last && j == n)) {
            const uint32_t r = (uint32_t)(j - t0);
            uint32_t o = s_toff[r / DPT];
            for (uint32_t q = 0; q < r % DPT; ++q) o += dec_tok_len(t, __ldg(ids + t0 + (r / DPT) * DPT + q), skip_special);
            raw_off[d] = base + o;
        } else break;
    }
    }
}

// length of a well-formed UTF-8 sequence at p[i] (maximal-subpart rules), 0 if ill-formed;
// *bad = number of bytes from_utf8_lossy replaces by one U+FFFD when ill-formed
__device__ __forceinline__ int utf8_seq(const uint8_t* p, uint64_t i, uint64_t n, int* bad) {
    uint8_t c = p[i];
    if (c < 0x80) return 1;
    int need = 0; uint8_t lo = 0x80, hi = 0xBF;
    if (c >= 0xC2 && c <= 0xDF) need = 1;
    else if (c == 0xE0) { need = 2; lo = 0xA0; }
    else if (c >= 0xE1 && c <= 0xEC) need = 2;
    else if (c == 0xED) { need = 2; hi = 0x9F; }
    else if (c >= 0xEE && c <= 0xEF) need = 2;
    else if (c == 0xF0) { need = 3; lo = 0x90; }
    else if (c >= 0xF1 && c <= 0xF3) need = 3;
    else if (c == 0xF4) { need = 3; hi = 0x8F; }
    if (!need) { *bad = 1; return 0; }
    for (int k = 1; k <= need; ++k) {
        if (i + k >= n) { *bad = k; return 0; }
        uint8_t d = p[i + k], l = k == 1 ? lo : 0x80, h = k == 1 ? hi : 0xBF;
        if (d < l || d > h) { *bad = k; return 0; }
    }
    return need + 1;
}

// Is the concatenation valid UTF-8?  16 bytes per thread; a sequence that starts in the group may read up
// to 3 bytes of the next one.  Document edges are checked separately (k_dec_valid_docs).
__global__ void __launch_bounds__(256) k_dec_valid_bytes(const uint8_t* __restrict__ raw, uint64_t n,
                                                         uint32_t* __restrict__ any_invalid) {
    uint64_t base = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 16;
    if (base >= n) return;
    if (base + 16 <= n) {
        uint4 v = *reinterpret_cast<const uint4*>(raw + base);            // raw is 16-byte aligned
        if (((v.x | v.y | v.z | v.w) & 0x80808080u) == 0) return;          // all ASCII
    }
    uint64_t end = base + 16 < n ? base + 16 : n;
    for (uint64_t i = base; i < end; ++i) {
        uint8_t c = raw[i];
        if (c < 0x80) continue;
        if ((c & 0xC0) == 0x80) {
            // continuation: some lead within the previous 3 bytes must cover it
            bool ok = false;
            for (int k = 1; k <= 3 && (uint64_t)k <= i; ++k) {
                uint8_t p = raw[i - k];
                if ((p & 0xC0) == 0x80) continue;
                int need = p >= 0xF0 ? 3 : (p >= 0xE0 ? 2 : (p >= 0xC0 ? 1 : 0));
                ok = need >= k;
                break;
            }
            if (!ok) { atomicOr(any_invalid, 1u); return; }
        } else {
            int bad, L = utf8_seq(raw, i, n, &bad);
            if (!L) { atomicOr(any_invalid, 1u); return; }
        }
    }
}
// a document may not begin with a continuation byte nor end inside a sequence
__global__ void k_dec_valid_docs(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ raw_off, uint64_t n_docs,
                                 uint32_t* __restrict__ any_invalid) {
    uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    uint64_t lo = raw_off[d], hi = raw_off[d + 1];
    if (lo == hi) return;
    bool bad = (raw[lo] & 0xC0) == 0x80;
    for (uint64_t k = 1; k <= 3 && hi - k >= lo && hi >= k; ++k) {
        uint8_t p = raw[hi - k];
        if ((p & 0xC0) == 0x80) continue;
        int need = p >= 0xF0 ? 3 : (p >= 0xE0 ? 2 : (p >= 0xC0 ? 1 : 0));
        if ((uint64_t)need >= k) bad = true;                                // its tail would lie in the next document
        break;
    }
    if (bad) atomicOr(any_invalid, 1u);
}

// White_Space (Rust char::is_whitespace) at p[i] in valid UTF-8: byte length or 0
__device__ __forceinline__ int ws_at(const uint8_t* p, uint64_t i, uint64_t n) {
    uint8_t c = p[i];
    if (c == 0x20 || (c >= 9 && c <= 13)) return 1;
    if (c == 0xC2 && i + 1 < n && (p[i + 1] == 0x85 || p[i + 1] == 0xA0)) return 2;
    if (i + 2 < n) {
        uint8_t d = p[i + 1], e = p[i + 2];
        if (c == 0xE1 && d == 0x9A && e == 0x80) return 3;
        if (c == 0xE2 && d == 0x80 && ((e >= 0x80 && e <= 0x8A) || e == 0xA8 || e == 0xA9 || e == 0xAF)) return 3;
        if (c == 0xE2 && d == 0x81 && e == 0x9F) return 3;
        if (c == 0xE3 && d == 0x80 && e == 0x80) return 3;
    }
    return 0;
}

__device__ uint64_t replace_pass(const uint8_t* src, uint64_t n, uint8_t* dst, const char* pat, int pl, char rep) {
    uint64_t i = 0, o = 0;
    while (i < n) {
        bool m = i + pl <= n;
        for (int k = 0; m && k < pl; ++k) m = src[i + k] == (uint8_t)pat[k];
        if (m) { dst[o++] = (uint8_t)rep; i += pl; }
        else dst[o++] = src[i++];
    }
    return o;
}

// one thread per document; scratch A/B hold 3*raw_len+4 bytes per document each
__global__ void k_dec_post(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ raw_off, uint64_t n_docs,
                           int cleanup, uint8_t* bufA, uint8_t* bufB, uint64_t* __restrict__ out_len,
                           uint8_t* __restrict__ which) {
    uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    const uint8_t* p = raw + raw_off[d];
    uint64_t n = raw_off[d + 1] - raw_off[d];
    uint64_t so = 3 * raw_off[d] + 4 * d;
    uint8_t *a = bufA + so, *b = bufB + so;
    // String::from_utf8_lossy
    uint64_t m = 0;
    for (uint64_t i = 0; i < n;) {
        int bad = 0, L = utf8_seq(p, i, n, &bad);
        if (L) { for (int k = 0; k < L; ++k) a[m++] = p[i + k]; i += L; }
        else { a[m++] = 0xEF; a[m++] = 0xBF; a[m++] = 0xBD; i += bad; }
    }
    int cur = 0;
    if (cleanup) {
        const char* pats[15] = {" .", " ,", " !", " ?", " :", " ;", "\" ", " \"", "' ", " '", "( ", " )", "[ ", " ]", " - "};
        const char reps[15] = {'.', ',', '!', '?', ':', ';', '"', '"', '\'', '\'', '(', ')', '[', ']', '-'};
        const int lens[15] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3};
        for (int k = 0; k < 15; ++k) {
            m = replace_pass(cur ? b : a, m, cur ? a : b, pats[k], lens[k], reps[k]);
            cur ^= 1;
        }
        // split_whitespace().join(" ")
        const uint8_t* s = cur ? b : a;
        uint8_t* o = cur ? a : b;
        uint64_t w = 0;
        bool pending = false, any = false;
        for (uint64_t i = 0; i < m;) {
            int wl = ws_at(s, i, m);
            if (wl) { pending = true; i += wl; continue; }
            if (pending && any) o[w++] = ' ';
            pending = false; any = true;
            o[w++] = s[i++];
        }
        m = w;
        cur ^= 1;
    }
    out_len[d] = m;
    which[d] = (uint8_t)cur;
}

__global__ void k_dec_copy_out(const uint8_t* __restrict__ bufA, const uint8_t* __restrict__ bufB,
                               const uint64_t* __restrict__ raw_off, const uint8_t* __restrict__ which,
                               const uint64_t* __restrict__ out_off, uint64_t n_docs, uint8_t* __restrict__ out,
                               uint64_t out_cap, uint32_t* __restrict__ err) {
    // one warp per document
    uint64_t d = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (d >= n_docs) return;
    int lane = threadIdx.x & 31;
    uint64_t so = 3 * raw_off[d] + 4 * d;
    const uint8_t* s = (which[d] ? bufB : bufA) + so;
    uint64_t o = out_off[d], n = out_off[d + 1] - o;
    if (o + n > out_cap) { if (lane == 0) atomicOr(err, ERRF_CAPACITY); return; }
    for (uint64_t i = lane; i < n; i += 32) out[o + i] = s[i];
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return eng.cuda_fail(e_, #x); } while (0)

// Metaspace decoder (decoders.rs:121-131): after join + replacement -> ' ' (folded into the per-token blob at load), ONE leading
// space of every text is dropped when add_prefix_space is set.
__global__ void k_strip_len(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ off, uint64_t n, uint64_t* __restrict__ new_len) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n) return;
    if (d == n) { new_len[d] = 0; return; }
    const uint64_t lo = off[d], hi = off[d + 1];
    new_len[d] = (hi - lo) - ((hi > lo && raw[lo] == 0x20) ? 1 : 0);
}
__global__ void __launch_bounds__(256) k_strip_copy(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ off, uint64_t n,
                                                    const uint64_t* __restrict__ new_off, uint8_t* __restrict__ out) {
    const uint64_t d = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (d >= n) return;
    const int lane = threadIdx.x & 31;
    const uint64_t len = new_off[d + 1] - new_off[d], src = off[d + 1] - len;
    for (uint64_t i = lane; i < len; i += 32) out[new_off[d] + i] = raw[src + i];
}

int decode_device(Engine& eng, const uint32_t* d_ids, const uint64_t* d_ids_off, size_t n_docs, uint64_t T,
                  int skip_special, int cleanup, uint8_t* d_out, uint64_t out_cap, uint64_t* d_out_off,
                  uint64_t* n_bytes_host, cudaStream_t st) {
    Workspace& ws = eng.ws;
    uint32_t *tile_sum, *err, *first_doc;
    uint64_t *tile_base, *raw_off;
    const uint64_t n_tiles = (T + DT - 1) / DT;
    if (n_docs >= 0xFFFFFFF0ull) return eng.fail(CTK_ERR_ARG, "too many documents in one decode call");
    CK(ws.get(4, 256, (void**)&err));
    CK(cudaMemsetAsync(err, 0, 256, st));
    CK(ws.get(10, (n_tiles + 2) * 4, (void**)&tile_sum));
    CK(ws.get(11, (n_tiles + 2) * 8, (void**)&tile_base));
    CK(ws.get(12, (n_docs + 2) * 8, (void**)&raw_off));
    CK(ws.get(37, (n_tiles + 2) * 4, (void**)&first_doc));
    eng.mark(nullptr, st);
    if (n_tiles) {
        const size_t table = ((size_t)eng.dec.n_ids + 15) / 16 * 16;
        if (eng.dec_sums_grid == 0) {                                   // once: does the length table fit in shared memory, and how many CTAs per SM
            eng.dec_sums_grid = -1;
            int dev_smem = 0, sms = 0, occ = 0;
            cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, eng.device);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, eng.device);
            if (table + 1024 <= (size_t)dev_smem && !getenv("CTK_DEC_SUMS_GLOBAL") &&
                cudaFuncSetAttribute(k_dec_tile_sums_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dec_tile_sums_smem, DTH, table) == cudaSuccess && occ > 0)
                eng.dec_sums_grid = sms * occ;
            cudaGetLastError();
        }
        if (eng.dec_sums_grid > 0 && n_tiles >= 64 && n_tiles < 0xFFFFFFFFull)
            k_dec_tile_sums_smem<<<(unsigned)std::min<uint64_t>(n_tiles, (uint64_t)eng.dec_sums_grid), DTH, table, st>>>(eng.dec, d_ids, T, skip_special,
                                                                                                                          tile_sum, (uint32_t)n_tiles);
        else
            k_dec_tile_sums<<<(unsigned)n_tiles, DTH1, 0, st>>>(eng.dec, d_ids, T, skip_special, tile_sum);
        k_dec_first_doc<<<(unsigned)((n_docs + 1 + 255) / 256), 256, 0, st>>>(d_ids_off, n_docs, n_tiles, first_doc);
        eng.launched(2);
    }
    CK(cudaMemsetAsync(tile_sum + n_tiles, 0, 4, st));
    eng.mark("k_dec_tile_sums", st);
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(tile_sum, U32ToU64());
    size_t cub_bytes = 0;
    void* cub_tmp;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, it, tile_base, n_tiles + 1, st));
    CK(ws.get(5, cub_bytes + 16, &cub_tmp));
    CK(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, it, tile_base, n_tiles + 1, st));
    eng.launched(1); eng.mark("scan(tile sums)", st);
    uint64_t raw_total = 0;
    CK(eng.publish({{tile_base + n_tiles, 2, 12}}, st));
    CK(cudaStreamSynchronize(st));
    memcpy(&raw_total, eng.h_flags + 12, 8);
    // Without clean-up the gathered bytes are the result (if they are valid UTF-8): write them where they belong.
    const bool strip = eng.model.dec_metaspace && eng.model.dec_meta_strip;
    const bool direct = !cleanup && !strip && d_out && raw_total <= out_cap && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0;
    uint8_t* raw;
    if (direct) raw = d_out;
    else CK(ws.get(13, raw_total + 16, (void**)&raw));
    eng.mark(nullptr, st);
    if (n_tiles) {
        if (eng.dec_write_grid == 0) {                                  // once: how many hot records fit beside two CTAs per SM
            int dev_smem = 0, sms = 0, occ = 0;
            cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, eng.device);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, eng.device);
            uint32_t hot = 1024;                                        // measured on 1 GiB: 0 -> 1.42, 1024 -> 1.39, 2048 -> 1.46, 4096 -> 2.25 ms (occupancy)
            if (const char* e = getenv("CTK_DEC_HOT")) hot = (uint32_t)atoi(e);
            if (hot > eng.dec.n_ids) hot = eng.dec.n_ids;
            while (hot && (size_t)hot * 16 + 24 * 1024 > (size_t)dev_smem) hot /= 2;
            if (cudaFuncSetAttribute(k_dec_write, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(hot * 16)) != cudaSuccess) { cudaGetLastError(); hot = 0; }
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dec_write, DTH, (size_t)hot * 16) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
            eng.dec_hot = hot;
            eng.dec_write_grid = sms * occ;
        }
        if (n_tiles >= 0xFFFFFFFFull) return eng.fail(CTK_ERR_ARG, "too many ids in one decode call");
        k_dec_write<<<(unsigned)std::min<uint64_t>(n_tiles, (uint64_t)eng.dec_write_grid), DTH, (size_t)eng.dec_hot * 16, st>>>(
            eng.dec, d_ids, T, skip_special, tile_base, d_ids_off, first_doc, (uint32_t)n_tiles, eng.dec_hot, n_docs, raw,
            direct ? out_cap : raw_total + 16, raw_off, err, err + 2);
        eng.launched(1);
    } else CK(cudaMemsetAsync(raw_off, 0, (n_docs + 1) * 8, st));
    eng.mark("k_dec_write", st);
    if (strip && n_docs) {                                               // (raw, raw_off) -> the same texts without their one leading space
        uint64_t *s_len, *s_off; uint8_t* s_raw;
        CK(ws.get(80, (n_docs + 2) * 8, (void**)&s_len));
        CK(ws.get(81, (n_docs + 2) * 8, (void**)&s_off));
        CK(ws.get(82, raw_total + 16, (void**)&s_raw));
        k_strip_len<<<(unsigned)((n_docs + 1 + 255) / 256), 256, 0, st>>>(raw, raw_off, n_docs, s_len);
        CK(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, s_len, s_off, n_docs + 1, st));
        CK(ws.get(5, cub_bytes + 16, &cub_tmp));
        CK(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, s_len, s_off, n_docs + 1, st));
        k_strip_copy<<<(unsigned)((n_docs * 32 + 255) / 256), 256, 0, st>>>(raw, raw_off, n_docs, s_off, s_raw);
        eng.launched(3);
        CK(eng.publish({{s_off + n_docs, 2, 12}}, st));
        CK(cudaStreamSynchronize(st));
        memcpy(&raw_total, eng.h_flags + 12, 8);
        raw = s_raw; raw_off = s_off;
        eng.mark("k_strip", st);
    }
    // Is the gathered byte string already valid UTF-8?  Then String::from_utf8_lossy is the identity.  Pure ASCII is
    // (k_dec_write looked at every byte on its way out); anything else is checked sequence by sequence.
    bool invalid = false, has_high = false;
    if (n_docs && raw_total) {
        CK(eng.publish({{err + 2, 1, 15}}, st));
        CK(cudaStreamSynchronize(st));
        has_high = eng.h_flags[15] != 0;
    }
    if (has_high) {
        eng.mark(nullptr, st);
        k_dec_valid_bytes<<<(unsigned)(((raw_total + 15) / 16 + 255) / 256), 256, 0, st>>>(raw, raw_total, err + 1);
        k_dec_valid_docs<<<(unsigned)((n_docs + 255) / 256), 256, 0, st>>>(raw, raw_off, n_docs, err + 1);
        eng.launched(2); eng.mark("k_dec_valid", st);
        uint32_t inv = 0;
        CK(eng.publish({{err + 1, 1, 14}}, st));
        CK(cudaStreamSynchronize(st));
        inv = eng.h_flags[14];
        invalid = inv != 0;
    }
    const bool need_post = invalid;                        // sequential per-document path only for invalid UTF-8
    if (need_post && direct) {                             // rare: the post-processing reads `raw` and writes d_out
        uint8_t* scratch;
        CK(ws.get(13, raw_total + 16, (void**)&scratch));
        CK(cudaMemcpyAsync(scratch, raw, raw_total, cudaMemcpyDeviceToDevice, st));
        raw = scratch;
    }
    if (!d_out) {                                          // host-buffer entry point: output lives in the workspace
        out_cap = need_post ? 3 * raw_total + 16 : raw_total + 16;
        CK(ws.get(25, out_cap, (void**)&d_out));
        eng.last_decode_out = d_out;
    }
    if (!need_post && cleanup) {
        int rc = clean_parallel(eng, raw, raw_off, n_docs, raw_total, d_out, out_cap, d_out_off, err, st);
        if (rc != CTK_OK) return rc;
    } else if (!need_post) {
        if (raw_total > out_cap) return eng.fail(CTK_ERR_ARG, "decode output capacity too small");
        if (raw_total && raw != d_out) CK(cudaMemcpyAsync(d_out, raw, raw_total, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(d_out_off, raw_off, (n_docs + 1) * 8, cudaMemcpyDeviceToDevice, st));
    } else {
        uint8_t *bufA, *bufB, *which;
        uint64_t *out_len;
        uint64_t sz = 3 * raw_total + 4 * (n_docs + 1) + 16;
        CK(ws.get(14, sz, (void**)&bufA));
        CK(ws.get(15, sz, (void**)&bufB));
        CK(ws.get(16, (n_docs + 2) * 8, (void**)&out_len));
        CK(ws.get(17, n_docs + 16, (void**)&which));
        if (n_docs) {
            eng.mark(nullptr, st);
            k_dec_post<<<(unsigned)((n_docs + 127) / 128), 128, 0, st>>>(raw, raw_off, n_docs, cleanup, bufA, bufB, out_len, which);
            eng.launched(1); eng.mark("k_dec_post", st);
        }
        CK(cudaMemsetAsync(out_len + n_docs, 0, 8, st));
        CK(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, out_len, d_out_off, n_docs + 1, st));
        CK(ws.get(5, cub_bytes + 16, &cub_tmp));
        CK(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, out_len, d_out_off, n_docs + 1, st));
        eng.launched(1);
        if (n_docs) {
            k_dec_copy_out<<<(unsigned)((n_docs * 32 + 255) / 256), 256, 0, st>>>(bufA, bufB, raw_off, which, d_out_off, n_docs,
                                                                                  d_out, out_cap, err);
            eng.launched(1);
        }
    }
    CK(cudaGetLastError());
    return eng.finish(err, d_out_off, n_docs, n_bytes_host, st);
}

}  // namespace ctk
