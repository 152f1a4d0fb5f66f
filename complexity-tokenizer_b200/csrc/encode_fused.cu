// Fused encode: text in, packed ids + per-document offsets out.
//
//   k_first_doc      per 448-byte slice: first document that can start in it (documents are independent,
//                    mod.rs:694-696, so every slice must know where they start)
//   k_encode_slices  ALL of normalised-text -> ids: classes, pre-token boundaries, pre-token cache, BPE.
//                    One warp per slice, warps fully independent (no barrier, no inter-warp order).
//                    Writes the slice's ids to a fixed-stride scratch run and its count.
//   k_long_prep      sorts the pre-tokens longer than 32 bytes into work lists       (encode_long.cuh)
//   k_encode_mid<N>  33..128 bytes: one LANE per pre-token, the reference's loop verbatim
//   k_encode_long    129..256 bytes: one warp per pre-token (round-parallel, or registers)
//   k_xl_*           > 256 bytes: grid-wide round-parallel merging, only when one occurs (encode_xlong.cuh)
//   (device scan of the per-slice counts)
//   k_compact        moves every slice's run to its final place in the packed output
//   k_doc_fixup      ids_off[d] (written slice-relative by k_encode_slices) += base of its slice
//
// A single-pass variant (decoupled look-back, ids staged in shared memory) was measured first
// (profiles/r1_v1_fused_full_summary.txt, r1_v2_lookback_summary.txt): with variable work per tile the
// look-back chain stalled workers ~50% of the time, and the kernel is issue-bound, not HBM-bound, so
// the extra 8T bytes of scratch traffic are the cheaper price.
//
// Inside a warp (reference lines in brackets); lane l holds bytes [16 l, 16 l + 16) of a 512-byte chunk
// = 16 bytes of left context + 448 OWNED bytes + 48 bytes of right context:
//   1 classify: SWAR class masks per lane, trie for non-ASCII   [pretokenizers.rs:13 \p{L} \p{N} \s]
//   2 boundaries: 32-bit window logic + warp shuffles           [pretokenizers.rs:13, :158-185]
//   3 compaction: prefix sum of per-lane popcounts -> list of pre-token starts
//   4 per pre-token, 32 at a time, straight-line: extract <= 16 key bytes, hash, ONE 256-bit load of the
//     batch's pre-token cache slot (BPE of a pre-token is a pure function of its bytes, so each distinct
//     pre-token of a batch is merged once and every other occurrence copies the ids), unpack the ids.
//     Everything else (cache miss, 17..32 bytes, ids that do not fit inline, > 32 bytes) takes the
//     warp-cooperative slow path: bpe_warp32 [bpe.rs:104-153]; pre-tokens > 32 bytes are only DESCRIBED here
//     (LongDesc) and merged by the kernels above.  NFC-suspect code points are reported, not handled (engine.hpp).
//   5 ids leave in pre-token order (mod.rs:562-612 `result.extend`)
#include <cub/device/device_scan.cuh>

#include <cstdlib>
#include <cstring>

#include "engine.hpp"
#include "added_tokens.cuh"
#include "start_window.cuh"

namespace ctk {

#ifndef CTK_FW
#define CTK_FW 8
#endif
constexpr int FW = CTK_FW;       // warps per CTA (9 warps x 4 CTAs = 36 warps per SM at 56 registers was measured: see DESIGN 6.1)
constexpr int SLICE = 448;       // bytes owned by a warp
constexpr int CHUNK = 512;       // bytes a warp looks at
constexpr int LCTX = 16;         // left context
constexpr int STAGE = 480;       // ids of one slice from pre-tokens of <= 32 bytes (they cover < 480 bytes)
constexpr int MAXLONG = 14;      // pre-tokens longer than 32 bytes that can start in 448 bytes
constexpr int MAXINLINE = 6;     // ids the straight-line path unpacks
constexpr uint32_t META_EMPTY = 0xFFFFFFFFu, META_BUSY = 0xFFFFFFFEu;
constexpr int PROBES = 4;
constexpr int CACHE_LOG2 = 22;        // slots of the per-batch pre-token cache (32 B each)
constexpr uint32_t END_UNKNOWN = 0xFFFFu;

// A pre-token longer than 32 bytes: found by k_encode_slices, merged by k_encode_long.
// Its `cnt` ids at long_pool[pool..] go before position `at` of the slice's run; `k` is its pre-token index.
struct LongDesc { uint64_t gstart; uint32_t len, slice, k_at, pool, cnt, chunk_end; };
// A VERY long pre-token (encode_xlong.cuh): its LongDesc and where its symbols start in the round array X
struct XlEntry { uint32_t desc, pad; unsigned long long xoff; };
constexpr int XL_IDX_SHIFT = 40;                 // one atomic hands out (list index << 40 | X offset)

struct FusedParams {
    DevTables t;
    const uint8_t* text; uint64_t n_bytes;
    const uint64_t* off; uint64_t n_docs;
    const uint32_t* first_doc; uint64_t n_slices; uint32_t n_tiles;
    CacheSlot* cache; uint32_t cache_mask, cache_shift;                 // slot = hash >> cache_shift (top bits), probes wrap with cache_mask
    uint32_t id_bits, n_inline;                          // ids packed inline in a cache slot: n_inline x id_bits <= 96
    uint32_t* ovf_pool; uint32_t ovf_cap; uint32_t* ovf_cursor;
    uint32_t* long_pool; unsigned long long long_cap; unsigned long long* long_cursor;
    LongDesc* desc; uint32_t desc_cap; uint32_t* desc_cursor;
    void* runs;                                          // n_slices x STAGE ids of run_width bytes each
    int run_width;                                       // 2: every id fits 16 bits (also 16 bits per id in a cache slot), else 4
    uint32_t* slice_cnt;                                 // ids of the slice (short + long)
    uint32_t* slice_info;                                // staged count | n_long << 16
    uint32_t* slice_desc;                                // first LongDesc of the slice (if n_long > 0)
    uint16_t* slice_first;                               // chunk-relative position of the slice's first owned start, 0xFFFF if none
    int xl_enabled;                                      // monotone table, no in-word added tokens, synchronous call
    int check_nfc;                                       // optimistic call: report NFC-suspect code points (engine.hpp)
    int no_rounds;                                       // debug (CTK_NO_ROUNDS): sequential merging in k_encode_long
    int mid_enabled;                                     // 33..128-byte pre-tokens go to k_encode_mid (one lane each)
    uint32_t* work_list;                                 // 3 lists of LongDesc indices, desc_cap entries each (k_long_prep)
    uint32_t* work_count;                                // their lengths
    unsigned long long* xl_cursor;                       // list index << XL_IDX_SHIFT | symbols handed out
    XlEntry* xl_list;
    uint64_t* ids_off_rel;                               // slice-relative document offsets (k_doc_fixup makes them absolute)
    uint64_t* ids_off; uint32_t* err;
    int ablate;                                          // debug (CTK_ABLATE): 1 = stop after boundaries, 2 = no slow path, 3 = no probe, 4 = pre-token cache off
};

struct __align__(16) ChunkBuf { uint8_t pad0[16]; uint8_t chunk[CHUNK]; uint8_t pad1[16]; };
struct __align__(16) WarpSmem {
    ChunkBuf buf[2];                                     // this slice's chunk and the next one's (in flight)
    unsigned long long bar[2];                           // CTK_BULK_CHUNK: one mbarrier per buffer
    uint32_t ds[32];
    uint16_t list[SLICE + 40];                           // owned starts + a round of sentinels
    uint16_t l_at[MAXLONG + 2], l_k[MAXLONG + 2], l_pos[MAXLONG + 2], l_len[MAXLONG + 2];
    uint32_t l_cnt[MAXLONG + 2];
};

// first_doc[s] = smallest d with off[d] >= s*SLICE - LCTX (bits 0..30); also validates the offsets
__global__ void k_first_doc(const uint64_t* __restrict__ off, uint64_t n_docs, uint64_t n_bytes, uint64_t n_slices,
                            uint32_t* __restrict__ first_doc, uint32_t* __restrict__ err) {
    uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    uint64_t p = off[d];
    if ((d == 0 && p != 0) || (d == n_docs && p != n_bytes) || (d < n_docs && off[d + 1] < p) || p > n_bytes) {
        atomicOr(err, ERRF_OFFSETS);
        return;
    }
    uint64_t s_lo = d ? (off[d - 1] + LCTX) / SLICE + 1 : 0;
    uint64_t s_hi = (p + LCTX) / SLICE;
    if (s_hi >= n_slices) s_hi = n_slices - 1;
    // bit 31: a document start (or the end of the text) lies inside the slice's 512-byte chunk
    for (uint64_t s = s_lo; s <= s_hi; ++s)
        first_doc[s] = (uint32_t)d | ((long long)p < (long long)(s * SLICE) - LCTX + CHUNK ? 0x80000000u : 0u);
}

// ids_off[d] was written relative to the slice that owns position off[d]; add that slice's base
// (k_encode_slices left run position | pre-token index << 32; long pre-tokens before that index add their ids)
__global__ void k_doc_fixup(const uint64_t* __restrict__ off, uint64_t n_docs, uint64_t n_slices,
                            const uint32_t* __restrict__ slice_base, const uint32_t* __restrict__ slice_info,
                            const uint32_t* __restrict__ slice_desc, const LongDesc* __restrict__ desc,
                            const uint64_t* __restrict__ ids_off_rel, uint64_t* __restrict__ ids_off) {
    uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    uint64_t s = off[d] / SLICE;
    if (s >= n_slices) s = n_slices - 1;
    const uint64_t v = ids_off_rel[d];
    const uint32_t rel = (uint32_t)v, k = (uint32_t)(v >> 32), n_long = slice_info[s] >> 16;
    uint64_t extra = 0;
    if (n_long) {
        const LongDesc* dd = desc + slice_desc[s];
        for (uint32_t q = 0; q < n_long; ++q) if ((dd[q].k_at >> 16) < k) extra += dd[q].cnt;
    }
    ids_off[d] = (uint64_t)slice_base[s] + rel + extra;
}

// one L2 sector, one instruction: {key[4], meta, ids[3]}
__device__ __forceinline__ void load_slot(const CacheSlot* p, uint32_t& k0, uint32_t& k1, uint32_t& k2, uint32_t& k3,
                                          uint32_t& meta, uint32_t& t0, uint32_t& t1, uint32_t& t2) {
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(k0), "=r"(k1), "=r"(k2), "=r"(k3), "=r"(meta), "=r"(t0), "=r"(t1), "=r"(t2) : "l"(p));
}

// the slot index is the TOP bits of this (h >> cache_shift): the last step is a multiply, whose high bits are the mixed ones
__device__ __forceinline__ uint32_t key_hash(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t len) {
    uint32_t h = (x0 * 0x9E3779B1u + x1 * 0x85EBCA77u) ^ (x2 * 0xC2B2AE3Du + x3 * 0x27D4EB2Fu + len * 0x165667B1u);
    h ^= h >> 15;
    return h * 0x2C1B3C6Du;
}

// SHFL.UP sets a predicate where the source lane exists: two instructions per step (shuffle, predicated add)
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t& total, int lane) {
    const unsigned full = 0xFFFFFFFFu;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
        asm volatile("{ .reg .u32 t; .reg .pred q; shfl.sync.up.b32 t|q, %0, %1, 0, 0xffffffff; @q add.u32 %0, %0, t; }" : "+r"(incl) : "r"(o));
    total = __shfl_sync(full, incl, 31);
    return incl - v;
}

// initial ids of `len` bytes at text[...] into lanes (<= 32), unknown bytes dropped (bpe.rs:94-97); returns count
__device__ __forceinline__ int init_symbols32(const uint32_t* s_byte_init, const uint8_t* bytes, int len, uint32_t& sym, int lane) {
    const unsigned full = 0xFFFFFFFFu;
    sym = lane < len ? s_byte_init[bytes[lane]] : kNone;
    unsigned have = __ballot_sync(full, sym != kNone);
    unsigned want = len == 32 ? full : ((1u << len) - 1u);
    if (have != want) {
        int dst = __popc(have & ((1u << lane) - 1u));
        uint32_t out = kNone;
        for (int src = 0; src < 32; ++src) {
            uint32_t v = __shfl_sync(full, sym, src);
            int d = __shfl_sync(full, dst, src);
            if (((have >> src) & 1u) && d == lane) out = v;
        }
        sym = out;
    }
    return __popc(have);
}

}  // namespace ctk
#include "encode_long.cuh"
#include "encode_xlong.cuh"
namespace ctk {

#ifndef CTK_LB
#define CTK_LB 4
#endif
#ifndef CTK_OPAQUE_WARP
#define CTK_OPAQUE_WARP 0     // measured: 3.50 -> 3.44 ms without
#endif
#ifndef CTK_PIPELINE_ROUNDS
#define CTK_PIPELINE_ROUNDS 1
#endif
#ifndef CTK_COMPACT_UNROLL
#define CTK_COMPACT_UNROLL 6   // measured: 4 -> 6: 3.50 -> 3.47 ms
#endif
#ifndef CTK_BULK_CHUNK
#define CTK_BULK_CHUNK 0     // interior chunks by ONE 512-byte bulk async copy (TMA, cp.async.bulk + mbarrier) issued by one lane instead of 32 cp.async
#endif
#ifndef CTK_ASYNC_CHUNK
#define CTK_ASYNC_CHUNK 1    // the next slice's chunk travels global -> shared with cp.async while this one is processed (no registers)
#endif

template <int RW> struct RunId;
template <> struct RunId<2> { typedef uint16_t type; };
template <> struct RunId<4> { typedef uint32_t type; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// st.shared.u16 only where `bits` != 0: one ISETP + one predicated STS, no branch
__device__ __forceinline__ void sts16_if(uint32_t saddr, uint32_t v, uint32_t bits) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.u16 [%0], %1; }" :: "r"(saddr), "h"((unsigned short)v), "r"(bits) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t saddr, const void* g, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(saddr), "l"(g), "r"(src_bytes) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// ---- 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (the TMA engine moves the bytes)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{ .reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W; }" :: "r"(bar), "r"(parity) : "memory");
}

// RW = bytes per id in the slice runs: 2 when every id the tokenizer can emit is below 65 536 (half the scratch
// traffic of this kernel and of k_compact), else 4.  In a cache slot ids then take 16 bits each (p.id_bits == 16).
template <int RW>
__global__ void __launch_bounds__(FW * 32, CTK_LB) k_encode_slices(const FusedParams p) {
    typedef typename RunId<RW>::type run_t;
    const unsigned full = 0xFFFFFFFFu;
    __shared__ WarpSmem sm[FW];
    __shared__ uint32_t s_byte_init[256];
    __shared__ uint4 s_kmask[17];                       // byte masks: keep the first n bytes of 16
    const int tid = threadIdx.x, lane = tid & 31;
    int w = tid >> 5;
    if (FW >= 8) { if (tid < 256) s_byte_init[tid] = __ldg(p.t.byte_init + tid); }
    else for (int i = tid; i < 256; i += FW * 32) s_byte_init[i] = __ldg(p.t.byte_init + i);
    if (tid < 17) {
        uint32_t m[4];
        for (int j = 0; j < 4; ++j) { int b = tid - 4 * j; m[j] = b >= 4 ? 0xFFFFFFFFu : (b <= 0 ? 0u : ((1u << (8 * b)) - 1u)); }
        s_kmask[tid] = make_uint4(m[0], m[1], m[2], m[3]);
    }
#if CTK_OPAQUE_WARP
    asm volatile("" : "+r"(w));                          // keep the warp index in a register: ptxas otherwise rebuilds the warp's shared-memory base
                                                         // from the thread id (two S2R + four ALU) at half a dozen places of the slice loop
#endif
    WarpSmem& S = sm[w];
    if (lane < 4) {
        uint4* z = reinterpret_cast<uint4*>(lane & 1 ? S.buf[lane >> 1].pad1 : S.buf[lane >> 1].pad0);
        *z = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const CacheSlot* const cache = p.cache;
    const uint32_t cshift = p.cache_shift, idb = p.id_bits, idmask = p.id_bits >= 32 ? 0xFFFFFFFFu : (1u << p.id_bits) - 1u, ninl = p.n_inline;
    const uint8_t* const text = p.text;
    const uint32_t nb32 = (uint32_t)p.n_bytes;                           // a device call takes < 4 GiB: positions are 32-bit here
    const uint32_t n_slices = (uint32_t)p.n_slices;
    const uint32_t stride = gridDim.x * FW;
    uint32_t slice = blockIdx.x * FW + w;
    if (slice >= n_slices) return;

    uint32_t bulk_pending = 0, bulk_phase = 0;                             // per buffer: a bulk copy is in flight / the parity to wait for (warp-uniform)
    // this lane's 16 bytes of a slice's chunk -> shared memory, zero beyond either end of the text
    auto fetch = [&](uint32_t sl, int b) {
        const uint32_t q = sl * SLICE - LCTX + 16 * lane;                 // (wraps for lane 0 of slice 0, which reads nothing)
        uint32_t sz = 16;
        if (sl == 0 || sl * SLICE + (CHUNK - LCTX) > nb32) {              // warp-uniform: the chunk sticks out of the text
            const uint32_t left = nb32 - sl * SLICE + LCTX;               // bytes from the chunk's first byte to the end of the text
            const uint32_t mine = left > 16u * lane ? left - 16u * lane : 0u;
            sz = (sl == 0 && lane == 0) ? 0u : (mine < 16u ? mine : 16u);
        }
#if CTK_BULK_CHUNK
        if (!(sl == 0 || sl * SLICE + (CHUNK - LCTX) > nb32)) {           // interior chunk: 512 bytes, 16-byte aligned, whole
            if (lane == 0) bulk_load(smem_u32(S.buf[b].chunk), text + (sl * SLICE - LCTX), CHUNK, smem_u32(&S.bar[b]));
            bulk_pending |= 1u << b;
            return;
        }
#endif
#if CTK_ASYNC_CHUNK
        cp_async16_zfill(smem_u32(S.buf[b].chunk + 16 * lane), text + (sz ? q : 0), sz);
#else
        uint4 r = make_uint4(0, 0, 0, 0);
        if (sz) r = __ldg(reinterpret_cast<const uint4*>(text + q));
        if (sz && sz < 16) { const uint4 km = s_kmask[sz]; r.x &= km.x; r.y &= km.y; r.z &= km.z; r.w &= km.w; }
        *reinterpret_cast<uint4*>(S.buf[b].chunk + 16 * lane) = r;
#endif
    };
    int buf = 0;
#if CTK_BULK_CHUNK
    if (lane == 0) { mbar_init(smem_u32(&S.bar[0]), 1); mbar_init(smem_u32(&S.bar[1]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
#endif
#if CTK_ASYNC_CHUNK
    fetch(slice, 0);
#endif
    uint32_t fd_next = __ldg(p.first_doc + slice);
    for (; slice < n_slices; slice += stride, buf ^= 1) {
        const uint32_t lo = slice * SLICE;                                 // first owned byte; the chunk starts LCTX bytes before it
        run_t* const run = reinterpret_cast<run_t*>(p.runs) + (uint64_t)slice * STAGE;
        uint8_t* const chunk = S.buf[buf].chunk;
        uint32_t n_owned = 0, stage_cnt = 0, n_long = 0, first_k = 0, ownm = 0;
        // interior slice (warp-uniform): the whole chunk lies inside the text
        const bool edge = slice == 0 || lo + (CHUNK - LCTX) > nb32;
        // valid bytes from this lane's first byte on (only its sign and values below 16 matter)
        const int room = edge ? (int)min(nb32 - lo, 1024u) + LCTX - 16 * lane : 16;

        // ---- 1. the chunk: fetched one slice ahead, so its latency hides behind the previous slice's work
        const uint32_t fd = fd_next;
#if CTK_ASYNC_CHUNK
        cp_async_wait_all();
#if CTK_BULK_CHUNK
        if (bulk_pending & (1u << buf)) {
            mbar_wait(smem_u32(&S.bar[buf]), (bulk_phase >> buf) & 1u);
            bulk_phase ^= 1u << buf;
            bulk_pending &= ~(1u << buf);
        }
#endif
        __syncwarp();                                                      // every lane's 16 bytes have landed; the previous slice's readers are done
        if (slice + stride < n_slices) { fetch(slice + stride, buf ^ 1); fd_next = __ldg(p.first_doc + slice + stride); }
#else
        __syncwarp();
        fetch(slice, buf);
        if (slice + stride < n_slices) fd_next = __ldg(p.first_doc + slice + stride);
        __syncwarp();
#endif
        const uint4 v = *reinterpret_cast<const uint4*>(chunk + 16 * lane);
        const uint32_t d0 = fd & 0x7FFFFFFFu;
        // ---- document starts inside the chunk (position n_bytes = off[n_docs] counts as one)
        uint32_t ds16 = 0;
        const bool docs_here = (fd >> 31) != 0;                            // warp-uniform
        if (docs_here) {
            S.ds[lane] = 0;
            __syncwarp();
            for (uint64_t d = (uint64_t)d0 + lane;; d += 32) {            // off[d0] is the first offset at or after the chunk's first byte
                const uint32_t rel = d <= p.n_docs ? (uint32_t)__ldg(p.off + d) - lo + LCTX : 0xFFFFFFFFu;   // chunk-relative
                const bool in = rel < (uint32_t)CHUNK;
                if (in) atomicOr(&S.ds[rel >> 4], 1u << (rel & 15));
                if (!__all_sync(full, in)) break;
            }
            __syncwarp();
            ds16 = S.ds[lane];
        }
        // ---- 2. classes and boundaries
        uint32_t own16;
        {
            Masks16 m = classify16(chunk, 16 * lane, v.x, v.y, v.z, v.w, p.t.trie_index, p.t.trie_blocks);
            if (m.SUSP && p.check_nfc) atomicOr(p.err, ERRF_NFC_SUSPECT);
            const uint32_t pa = m.L | (m.N << 16), pb = m.W | (m.SP << 16), pc = m.AP | (m.CONT << 16);
            // neighbours need only the adjoining byte of each mask: four shuffles move all seven masks
            // (lane 0's "previous" and lane 31's "next" are its own values: those lanes own nothing, their result is not used)
            const uint32_t ua = __shfl_up_sync(full, __byte_perm(pa, pb, 0x7531), 1), ub = __shfl_up_sync(full, __byte_perm(pc, ds16, 0x7531), 1);
            const uint32_t na = __shfl_down_sync(full, __byte_perm(pa, pb, 0x6420), 1), nb = __shfl_down_sync(full, __byte_perm(pc, ds16, 0x6420), 1);
            // windows = [previous lane's high byte | own two bytes | next lane's low byte] of each 16-bit mask: two PRMT each
            #define WIN(u, o, nx, s1, s2) __byte_perm(__byte_perm(u, o, s1), nx, s2)
            const uint32_t S32 = start_window(WIN(ua, pa, na, 0x0540, 0x4210), WIN(ua, pa, na, 0x0761, 0x5210), WIN(ua, pb, na, 0x0542, 0x6210),
                                              WIN(ua, pb, na, 0x0763, 0x7210), WIN(ub, pc, nb, 0x0540, 0x4210), WIN(ub, pc, nb, 0x0761, 0x5210),
                                              WIN(ub, ds16, nb, 0x0542, 0x6210), chunk + 16 * lane - 8);
            #undef WIN
            own16 = (S32 >> 8) & 0xFFFFu;
        }
        {
            const uint32_t valid = room >= 16 ? 0xFFFFu : (room <= 0 ? 0u : ((1u << room) - 1u));
            ownm = (lane >= 1 && lane <= 28) ? (own16 & valid) : 0u;
            // ---- 3. compaction: list of owned pre-token starts, then the sentinel (first start after them)
            first_k = warp_excl_scan(__popc(ownm), n_owned, lane);
            uint32_t bits = ownm;
            const uint32_t lb = 16 * lane - 1;
            uint32_t la = smem_u32(S.list + first_k);
#pragma unroll
            for (int j = 0; j < CTK_COMPACT_UNROLL; ++j) {                 // predicated, no branches: a lane rarely has more starts
                sts16_if(la + 2 * j, lb + __ffs(bits), bits);
                bits &= bits - 1;                                          // 0 stays 0
            }
            if (bits) { uint16_t* lp = S.list + first_k + CTK_COMPACT_UNROLL; do { *lp++ = (uint16_t)(lb + __ffs(bits)); bits &= bits - 1; } while (bits); }
            // sentinel candidates: starts in the right context (lanes 29, 30) and the end of the text
            // (position n_bytes is a start thanks to its DS bit) wherever it falls
            uint32_t rc = 0;
            if (lane == 29 || lane == 30) rc = own16 & (room >= 16 ? 0xFFFFu : (room < 0 ? 0u : ((2u << room) - 1u)));
            else if (edge && lane >= 1 && lane <= 28 && room >= 0 && room < 16) rc = own16 & (1u << room);
            const unsigned bal = __ballot_sync(full, rc != 0);
            uint32_t sent = END_UNKNOWN;
            if (bal) { const int sl = __ffs(bal) - 1; const uint32_t r2 = __shfl_sync(full, rc, sl); sent = 16 * sl + (__ffs(r2) - 1); }
            // list[n_owned ..] = sentinel for a whole round and one more: the lanes beyond the last pre-token see length 0
            S.list[n_owned + lane] = (uint16_t)sent;
            if (lane < 2) S.list[n_owned + 32 + lane] = (uint16_t)sent;
        }
        __syncwarp();
        if (lane == 0) p.slice_first[slice] = n_owned ? S.list[0] : (uint16_t)0xFFFFu;

        if (p.ablate == 1) { if (lane == 0) { p.slice_cnt[slice] = n_owned; p.slice_info[slice] = 0; } continue; }
        // ---- 4. pre-tokens, 32 per round.  The round's key, hash and cache-slot load are ISSUED one round ahead (right after
        //      the previous round's compare, before its scan and stores), so the slot's L2 round trip runs behind that work.
        uint32_t pos, end, len, x0, x1, x2, x3, idx0, s0, s1, s2, s3, meta, t0, t1, t2;
#define CTK_ISSUE_ROUND(BK)                                                                                             \
        {                                                                                                               \
            const uint32_t k_ = (BK) + lane;                                                                            \
            pos = S.list[k_]; end = S.list[k_ + 1];                                                                     \
            len = end - pos;                       /* 0 beyond the last pre-token; huge when the end is unknown */      \
            const uint32_t pc = pos & (CHUNK - 1); /* (only a lane without a pre-token can hold END_UNKNOWN here) */     \
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(chunk + (pc & ~3u));                                 \
            const uint32_t a0 = wp[0], a1 = wp[1], a2 = wp[2], a3 = wp[3], a4 = wp[4];                                  \
            const uint32_t sh = (pc & 3u) * 8;                                                                          \
            const uint4 km = s_kmask[len < 16 ? len : 16];                                                              \
            x0 = __funnelshift_r(a0, a1, sh) & km.x; x1 = __funnelshift_r(a1, a2, sh) & km.y;                           \
            x2 = __funnelshift_r(a2, a3, sh) & km.z; x3 = __funnelshift_r(a3, a4, sh) & km.w;                           \
            idx0 = key_hash(x0, x1, x2, x3, len) >> cshift;                                                             \
            load_slot(cache + idx0, s0, s1, s2, s3, meta, t0, t1, t2);                                                  \
        }
        if (n_owned) CTK_ISSUE_ROUND(0u)
        for (uint32_t base_k = 0; base_k < n_owned; base_k += 32) {
            const uint32_t k = base_k + lane;
            const bool have = len != 0;
            // the table only holds pre-tokens of 1..16 bytes and an EMPTY / BUSY slot has 0xFF / 0xFE in its length byte: lanes
            // without a pre-token (len 0) or with a longer one can never compare equal
            if (p.ablate == 4) meta = META_BUSY;                           // measurement: cache off, every pre-token is merged where it stands
            bool match = (meta & 0xFFu) == len && (((s0 ^ x0) | (s1 ^ x1)) | ((s2 ^ x2) | (s3 ^ x3))) == 0;
            uint32_t ntok = __byte_perm(meta, 0, 0x4441);                  // (meta >> 8) & 0xFF
            bool fast = match && ntok <= ninl;
            if (!fast) ntok = 0;                                           // (also the lanes beyond the last pre-token, whose slot is whatever their zero key hashed to)
            // everything else: displaced keys walk on; then warp-cooperative, one pre-token after the other, in lane order
            unsigned slow = __ballot_sync(full, have && !fast);
            if (p.ablate == 2) slow = 0;
            if (slow) {
                uint32_t ins_slot = kNone;
                {
                    bool open = have && len <= 16 && !match;
                    if (open && meta == META_EMPTY) { ins_slot = idx0; open = false; }
                    if (p.ablate == 3 || p.ablate == 4) open = false;
                    for (int pr = 1; pr < PROBES && __any_sync(full, open); ++pr) {
                        if (open) {
                            uint32_t r0, r1, r2, r3, rm, v0, v1, v2;
                            const uint32_t idx = (idx0 + pr) & p.cache_mask;
                            load_slot(cache + idx, r0, r1, r2, r3, rm, v0, v1, v2);
                            if ((rm & 0xFFu) == len && (((r0 ^ x0) | (r1 ^ x1)) | ((r2 ^ x2) | (r3 ^ x3))) == 0) {
                                match = true; open = false; meta = rm; t0 = v0; t1 = v1; t2 = v2;
                            } else if (rm == META_EMPTY) { ins_slot = idx; open = false; }
                        }
                    }
                    ntok = (meta >> 8) & 0xFFu;
                    fast = match && ntok <= ninl;
                    if (!fast) ntok = 0;
                    slow = __ballot_sync(full, have && !fast);
                }
                uint32_t hit_total;
                const uint32_t E = warp_excl_scan(ntok, hit_total, lane);  // ids of fast lanes before each lane
                uint32_t extra = 0;                                        // ids of slow lanes handled so far
                unsigned lm = 0;                                           // lanes whose pre-token is long
                while (slow) {
                    const int src = __ffs(slow) - 1;
                    slow &= slow - 1;
                    const uint32_t spos = __shfl_sync(full, pos, src), slen = __shfl_sync(full, len, src);
                    if (slen > 32) { lm |= 1u << src; if (p.ablate == 9 && lane == 0) atomicAdd(p.err + 16, 1u); continue; }          // also END_UNKNOWN
                    const uint32_t o = stage_cnt + __shfl_sync(full, E, src) + extra;
                    int cnt = -1;
                    const uint32_t ins = __shfl_sync(full, ins_slot, src);
                    const uint32_t y0 = __shfl_sync(full, x0, src), y1 = __shfl_sync(full, x1, src), y2 = __shfl_sync(full, x2, src),
                                   y3 = __shfl_sync(full, x3, src);
                    if (__shfl_sync(full, (uint32_t)match, src)) {         // cached, but its ids live in the overflow pool
                        cnt = (int)((__shfl_sync(full, meta, src) >> 8) & 0xFFu);
                        const uint32_t rec = __shfl_sync(full, t0, src);
                        if (lane < cnt) run[o + lane] = (run_t)p.ovf_pool[(uint64_t)rec * 16 + lane];
                        if (p.ablate == 9 && lane == 0) atomicAdd(p.err + 17, 1u);
                    }
                    if (cnt < 0) {                                         // merge now
                        uint32_t sym;
                        if (p.ablate == 9 && lane == 0) atomicAdd(p.err + 19 + (slen > 16 ? 1 : 0) + (slen <= 16 && ins == kNone ? 2 : 0), 1u);
                        if (p.t.n_added == 0) {
                            int n = init_symbols32(s_byte_init, chunk + spos, (int)slen, sym, lane);
                            cnt = n ? bpe_warp32(p.t, sym, n) : 0;
                            if (lane < cnt) run[o + lane] = (run_t)sym;
                        } else {                                           // mod.rs:566-610: added tokens inside the word
                            int r = 0;
                            cnt = 0;
                            while (r < (int)slen) {
                                uint32_t aid;
                                const int pl = added_next_piece(p.t, chunk + spos + r, (int)slen - r, lane, &aid);
                                if (aid != kNone) { if (lane == 0) run[o + cnt] = (run_t)aid; cnt += 1; }
                                else {
                                    uint32_t ps;
                                    int n = init_symbols32(s_byte_init, chunk + spos + r, pl, ps, lane);
                                    int c = n ? bpe_warp32(p.t, ps, n) : 0;
                                    if (lane < c) run[o + cnt + lane] = (run_t)ps;
                                    cnt += c;
                                }
                                r += pl;
                            }
                            __syncwarp();
                            sym = lane < cnt ? (uint32_t)__ldcg(run + o + lane) : kNone;  // back into lanes for the cache entry
                        }
                        if (ins != kNone) {                                // publish in the batch cache
                            uint32_t w0 = 0, w1 = 0, w2 = 0;
                            bool ok = true;
                            if ((uint32_t)cnt <= ninl) {                   // pack ids, id_bits each, first id lowest
                                unsigned long long lo64 = 0, hi64 = 0;
                                for (int i = (int)ninl - 1; i >= 0; --i) {
                                    uint32_t ti = __shfl_sync(full, sym, i);
                                    hi64 = (hi64 << idb) | (lo64 >> (64 - idb));
                                    lo64 = (lo64 << idb) | (i < cnt ? ti : 0u);
                                }
                                w0 = (uint32_t)lo64; w1 = (uint32_t)(lo64 >> 32); w2 = (uint32_t)hi64;
                            } else {
                                uint32_t rec = 0;
                                if (lane == 0) rec = atomicAdd(p.ovf_cursor, 1u);
                                rec = __shfl_sync(full, rec, 0);
                                ok = rec < p.ovf_cap;
                                if (ok && lane < cnt) p.ovf_pool[(uint64_t)rec * 16 + lane] = sym;
                                w0 = rec;
                            }
                            if (ok && lane == 0) {
                                CacheSlot* sl = p.cache + ins;
                                if (atomicCAS(&sl->meta, META_EMPTY, META_BUSY) == META_EMPTY) {
                                    sl->k0 = y0 | ((uint64_t)y1 << 32); sl->k1 = y2 | ((uint64_t)y3 << 32);
                                    sl->tok[0] = w0; sl->tok[1] = w1; sl->tok[2] = w2;
                                    __threadfence();
                                    *reinterpret_cast<volatile uint32_t*>(&sl->meta) = slen | ((uint32_t)cnt << 8);
                                }
                            }
                        }
                    }
                    if (lane == src) ntok = (uint32_t)cnt;                 // so that the scan below covers slow lanes too
                    extra += (uint32_t)cnt;
                }
                while (lm) {                                               // long pre-tokens: remember them
                    const int src = __ffs(lm) - 1;
                    lm &= lm - 1;
                    const uint32_t lpv = __shfl_sync(full, pos, src), le = __shfl_sync(full, end, src);
                    if (lane == 0 && n_long < MAXLONG) {
                        S.l_k[n_long] = (uint16_t)(base_k + src); S.l_pos[n_long] = (uint16_t)lpv;
                        S.l_len[n_long] = (uint16_t)(le == END_UNKNOWN ? END_UNKNOWN : le - lpv);
                    }
                    ++n_long;
                }
            }
            // ids of all lower lanes (fast and slow): the slow lanes wrote theirs at exactly these offsets
            uint32_t e0 = t0, e1 = t1, e2 = t2;                            // this round's packed ids (the next round's load reuses t0..t2)
#if CTK_PIPELINE_ROUNDS
            if (base_k + 32 < n_owned) CTK_ISSUE_ROUND(base_k + 32)
#endif
            uint32_t round_total;
            const uint32_t o = stage_cnt + warp_excl_scan(ntok, round_total, lane);
            {
                const uint32_t en = fast ? ntok : 0u;                      // ids this lane unpacks from its slot (a slot may hold none: every byte dropped)
                const uint32_t mx = __reduce_max_sync(full, en);           // most pre-tokens are one or two ids: stop at the round's longest
                run_t* const dst = run + o;
                if (RW == 2) {                                             // 16 bits per id in the slot, 16-bit stores
                    if (en >= 1) dst[0] = (run_t)e0;
                    if (mx >= 2) {
                        if (en >= 2) dst[1] = (run_t)(e0 >> 16);
                        if (mx >= 3) {
                            if (en >= 3) dst[2] = (run_t)e1;
                            if (mx >= 4) {
                                if (en >= 4) dst[3] = (run_t)(e1 >> 16);
                                if (en >= 5) dst[4] = (run_t)e2;
                                if (en >= 6) dst[5] = (run_t)(e2 >> 16);
                            }
                        }
                    }
                } else {
                    if (en >= 1) dst[0] = (run_t)(e0 & idmask);
                    for (uint32_t i = 1; i < mx; ++i) {
                        e0 = __funnelshift_r(e0, e1, idb); e1 = __funnelshift_r(e1, e2, idb); e2 >>= idb;
                        if (i < en) dst[i] = (run_t)(e0 & idmask);
                    }
                }
            }
            // the list entry now becomes the pre-token's id offset inside the slice's run (ids_off, long ones)
            if (docs_here || n_long) { __syncwarp(); if (have) S.list[k] = (uint16_t)o; }
            stage_cnt += round_total;
#if !CTK_PIPELINE_ROUNDS
            if (base_k + 32 < n_owned) CTK_ISSUE_ROUND(base_k + 32)
#endif
        }
#undef CTK_ISSUE_ROUND
        __syncwarp();
        if (lane == 0) S.list[n_owned] = (uint16_t)stage_cnt;

        // ---- long pre-tokens (> 32 bytes): described here, merged by k_encode_long
        uint32_t desc0 = 0;
        if (n_long) {
            if (n_long > MAXLONG) { n_long = MAXLONG; if (lane == 0) atomicOr(p.err, ERRF_POOL); }
            if (lane == 0) desc0 = atomicAdd(p.desc_cursor, n_long);
            desc0 = __shfl_sync(full, desc0, 0);
            if (desc0 + n_long > p.desc_cap) { if (lane == 0) atomicOr(p.err, ERRF_POOL); n_long = 0; }
            __syncwarp();
            if (lane < (int)n_long) {
                LongDesc dd;
                dd.gstart = (uint64_t)(lo + S.l_pos[lane] - LCTX);
                dd.len = S.l_len[lane] == END_UNKNOWN ? 0xFFFFFFFFu : (uint32_t)S.l_len[lane];
                dd.slice = slice;
                dd.k_at = ((uint32_t)S.l_k[lane] << 16) | (uint32_t)S.list[S.l_k[lane]];
                dd.pool = 0; dd.cnt = 0; dd.chunk_end = (uint32_t)(CHUNK - 16) - (uint32_t)S.l_pos[lane];
                p.desc[desc0 + lane] = dd;
            }
        }
        __syncwarp();

        // ---- ids_off (relative to this slice) for the documents that start in the owned bytes, or at
        //      the very end of the text (owned by the last slice)
        if (docs_here) {
            const bool last_slice = slice + 1 == n_slices;
            for (uint64_t d = (uint64_t)d0 + lane;; d += 32) {
                const uint32_t crel = d <= p.n_docs ? (uint32_t)__ldg(p.off + d) - lo + LCTX : 0xFFFFFFFFu;
                const bool in = crel < (uint32_t)CHUNK;
                const bool own = in && crel >= (uint32_t)LCTX && (crel < (uint32_t)(LCTX + SLICE) || (last_slice && crel - LCTX == nb32 - lo));
                const uint32_t rel = own ? crel : 0u;
                uint32_t fk = __shfl_sync(full, first_k, rel >> 4), sb = __shfl_sync(full, ownm, rel >> 4);
                if (own) {
                    uint32_t k = fk + __popc(sb & ((1u << (rel & 15)) - 1u));
                    if (k > n_owned) k = n_owned;
                    p.ids_off_rel[d] = (unsigned long long)S.list[k] | ((unsigned long long)k << 32);   // + long ids: k_doc_fixup
                }
                if (!__all_sync(full, in)) break;
            }
        }
        if (lane == 0) {
            p.slice_cnt[slice] = stage_cnt;                                // k_encode_long adds its ids
            if (p.ablate == 9) atomicAdd(p.err + 24, n_owned);
            p.slice_info[slice] = stage_cnt | (n_long << 16);
            if (n_long) p.slice_desc[slice] = desc0;
        }
    }
}

// one warp per slice: the slice's run -> its final place (four loads in flight per lane).
// RW / OW = bytes per id in the runs / in the packed output (2 or 4 each)
template <int RW, int OW>
__global__ void __launch_bounds__(256) k_compact(const void* __restrict__ runs_v, const uint32_t* __restrict__ slice_base,
                                                 const uint32_t* __restrict__ slice_info, const uint32_t* __restrict__ slice_desc,
                                                 const LongDesc* __restrict__ desc, const uint32_t* __restrict__ long_pool,
                                                 uint64_t n_slices, void* __restrict__ out_v, uint64_t out_cap,
                                                 uint32_t* __restrict__ err) {
    typedef typename RunId<RW>::type run_t;
    typedef typename RunId<OW>::type out_t;
    const int lane = threadIdx.x & 31;
    const uint64_t s = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (s >= n_slices) return;
    const run_t* src = static_cast<const run_t*>(runs_v) + s * STAGE;
    // the first 128 ids of the run are fetched together with the slice's base and count, not after them: one dependent
    // round trip less per slice (a run has STAGE entries whatever its count: reading past the count is harmless)
    run_t v0 = __ldcs(src + lane), v1 = __ldcs(src + lane + 32), v2 = __ldcs(src + lane + 64), v3 = __ldcs(src + lane + 96);
    const uint32_t b = __ldg(slice_base + s), nxt = __ldg(slice_base + s + 1), inf = __ldg(slice_info + s);
    if ((uint64_t)nxt > out_cap) { if (lane == 0) atomicOr(err, ERRF_CAPACITY); return; }
    const uint32_t n_stage = inf & 0xFFFFu, n_long = inf >> 16;
    out_t* dst = static_cast<out_t*>(out_v) + b;
    if (n_long == 0) {
        for (uint32_t i0 = 0; i0 < n_stage; i0 += 128) {
            const uint32_t i = i0 + lane;
            if (i0) {
                if (i < n_stage) v0 = __ldcs(src + i);
                if (i + 32 < n_stage) v1 = __ldcs(src + i + 32);
                if (i + 64 < n_stage) v2 = __ldcs(src + i + 64);
                if (i + 96 < n_stage) v3 = __ldcs(src + i + 96);
            }
            if (i < n_stage) dst[i] = (out_t)v0;
            if (i + 32 < n_stage) dst[i + 32] = (out_t)v1;
            if (i + 64 < n_stage) dst[i + 64] = (out_t)v2;
            if (i + 96 < n_stage) dst[i + 96] = (out_t)v3;
        }
    } else {                                            // run ids interleaved with long pre-tokens' ids (rare)
        const LongDesc* dd = desc + slice_desc[s];
        for (uint32_t i = lane; i < n_stage; i += 32) {
            uint32_t add = 0;
            for (uint32_t q = 0; q < n_long; ++q) if ((dd[q].k_at & 0xFFFFu) <= i) add += dd[q].cnt;
            dst[i + add] = (out_t)src[i];
        }
        uint32_t before = 0;
        for (uint32_t q = 0; q < n_long; ++q) {
            const uint32_t* ls = long_pool + dd[q].pool;
            out_t* ld = dst + (dd[q].k_at & 0xFFFFu) + before;
            if (dd[q].pool != kNone)                       // kNone: a very long one, placed by k_xl_place
                for (uint32_t i = lane; i < dd[q].cnt; i += 32) ld[i] = (out_t)ls[i];
            before += dd[q].cnt;
        }
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return eng.cuda_fail(e_, #x); } while (0)

static constexpr size_t mid_smem(int n) { return (size_t)(2 * n * 32 + n * 4 + 256) * 4; }   // k_encode_mid<n>: symbols, keys, block minima, byte LUT

// Round-parallel merging of the very long pre-tokens k_encode_long set aside (encode_xlong.cuh).
// `cursor` = list entries << XL_IDX_SHIFT | symbols (incl. one separator per entry), read back by the caller.
struct XlState { const uint32_t* xs; uint32_t n, n_list; const uint32_t* sep_pos; uint32_t* region_dst; };

static int xlong_rounds(Engine& eng, const FusedParams& p, uint64_t cursor, uint32_t* ctrl, XlState& out, cudaStream_t st) {
    const uint64_t n_list64 = cursor >> XL_IDX_SHIFT, total = cursor & ((1ull << XL_IDX_SHIFT) - 1);
    if (n_list64 > p.desc_cap || total >= 0xFFFFFFF0ull) return eng.fail(CTK_ERR_CUDA, "internal scratch pool exhausted (very long pre-tokens)");
    const uint32_t n_list = (uint32_t)n_list64;
    Workspace& ws = eng.ws;
    eng.mark(nullptr, st);
    uint32_t *xa, *xb, *rank, *newid, *pos, *sep_pos;
    uint8_t* sv;
    void* tmp;
    CK(ws.get(32, (total + 2) * 4, (void**)&xa));
    CK(ws.get(33, (total + 2) * 4, (void**)&xb));
    CK(ws.get(34, total * 4, (void**)&rank));
    CK(ws.get(35, total * 4, (void**)&newid));
    CK(ws.get(36, total + 16, (void**)&sv));
    CK(ws.get(38, total * 4, (void**)&pos));
    CK(ws.get(39, (uint64_t)n_list * 4 + 16, (void**)&sep_pos));
    size_t t1 = 0, t2 = 0;
    cub::CountingInputIterator<uint32_t> idx_it(0);
    cub::TransformInputIterator<uint32_t, XlKeepFn, cub::CountingInputIterator<uint32_t>> keep_it(idx_it, XlKeepFn{xa, rank, sv});
    CK(cub::DeviceScan::InclusiveScan(nullptr, t1, sv, sv, XlSegOp(), (int)total, st));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, t2, keep_it, pos, (int)total, st));
    const size_t tmp_bytes = (t1 > t2 ? t1 : t2) + 16;
    CK(ws.get(40, tmp_bytes, &tmp));
    uint32_t* holes = ctrl + 8;
    uint32_t* n_out = ctrl + 9;
    {
        unsigned gy = n_list < 65535u ? n_list : 65535u;
        uint64_t per = total / n_list / 2048 + 1;
        unsigned gx = (unsigned)(per < 128 ? per : 128);
        k_xl_init<<<dim3(gx, gy), 256, 0, st>>>(p, p.xl_list, n_list, xa, holes);
        eng.launched(1);
    }
    CK(eng.publish({{holes, 1, 7}}, st));
    CK(cudaStreamSynchronize(st));
    bool drop_holes = eng.h_flags[7] != 0;                   // a byte without a vocab entry: one merge-free round removes them
    const uint32_t W = eng.model.max_token_span < 1 ? 1u : eng.model.max_token_span;
    const size_t smem = 2 * (size_t)(XL_TILE + 2 * W) * 4;
    uint32_t n = (uint32_t)total;
    int rounds = 0;
    for (;;) {
        const unsigned g256 = (n + 255) / 256;
        k_xl_rank<<<(n + XL_TILE - 1) / XL_TILE, XL_THREADS, smem, st>>>(p.t, xa, n, W, drop_holes ? 1 : 0, rank, newid, sv);
        size_t tb = tmp_bytes;
        CK(cub::DeviceScan::InclusiveScan(tmp, tb, sv, sv, XlSegOp(), (int)n, st));
        tb = tmp_bytes;
        cub::TransformInputIterator<uint32_t, XlKeepFn, cub::CountingInputIterator<uint32_t>> keep_now(idx_it, XlKeepFn{xa, rank, sv});
        CK(cub::DeviceScan::ExclusiveSum(tmp, tb, keep_now, pos, (int)n, st));
        k_xl_scatter<<<g256, 256, 0, st>>>(xa, rank, newid, sv, pos, n, xb, n_out);
        eng.launched(4);
        CK(eng.publish({{n_out, 1, 6}}, st));
        CK(cudaStreamSynchronize(st));
        const uint32_t n_new = eng.h_flags[6];
        uint32_t* t = xa; xa = xb; xb = t;
        ++rounds;
        if (!drop_holes && n_new == n) break;                  // nothing was selected: every region is final
        drop_holes = false;
        n = n_new;
    }
    eng.mark("xlong rounds", st);
    k_xl_seps<<<(n + 255) / 256, 256, 0, st>>>(xa, n, sep_pos);
    k_xl_count<<<(n_list + 255) / 256, 256, 0, st>>>(p, p.xl_list, n_list, sep_pos);
    eng.launched(2);
    eng.mark("k_xl_count", st);
    eng.xl_last_rounds = rounds;
    CK(cudaGetLastError());
    out = XlState{xa, n, n_list, sep_pos, xb};                 // xb is free now: region destinations
    return CTK_OK;
}

int encode_fused(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes,
                 uint32_t* d_ids_u32, uint64_t ids_cap, uint64_t* d_ids_off, uint64_t* n_ids_host, cudaStream_t st, bool check_nfc) {
    // eng.out_id_width == 2: the caller's buffer takes uint16 ids (ctk_encode_batch_device_ex; only offered when every id fits)
    void* const d_ids = d_ids_u32;
    const int out_width = eng.out_id_width == 2 && eng.run_width == 2 ? 2 : 4;
    if (n_bytes >= 0xFFFFF000ull) return eng.fail(CTK_ERR_ARG, "one device call handles less than 4 GiB of text");
    if (n_bytes == 0) {
        CK(cudaMemsetAsync(d_ids_off, 0, (n_docs + 1) * 8, st));
        if (n_ids_host) { CK(cudaStreamSynchronize(st)); *n_ids_host = 0; }
        return CTK_OK;
    }
    if (eng.fused_grid == 0) {
        int per_sm = 0, sms = 0;
        if (eng.run_width == 2) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_encode_slices<2>, FW * 32, 0));
        else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_encode_slices<4>, FW * 32, 0));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, eng.device));
        if (per_sm < 1) return eng.fail(CTK_ERR_CUDA, "encode kernel does not fit on an SM");
        eng.fused_grid = per_sm * sms;
        eng.long_grid = sms * 4;
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_encode_mid<MID_N0>, 32, mid_smem(MID_N0))); eng.mid_grid[0] = sms * (occ > 0 ? occ : 1);
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_encode_mid<MID_N1>, 32, mid_smem(MID_N1))); eng.mid_grid[1] = sms * (occ > 0 ? occ : 1);
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_encode_mid<MID_N2>, 32, mid_smem(MID_N2))); eng.mid_grid[2] = sms * (occ > 0 ? occ : 1);
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_encode_mid<MID_N3>, 32, mid_smem(MID_N3))); eng.mid_grid[3] = sms * (occ > 0 ? occ : 1);
    }
    Workspace& ws = eng.ws;
    FusedParams p{};
    p.t = eng.tables;
    p.text = d_text; p.n_bytes = n_bytes; p.off = d_off; p.n_docs = n_docs;
    p.n_slices = (n_bytes + SLICE - 1) / SLICE;
    p.n_tiles = (uint32_t)((p.n_slices + FW - 1) / FW);
    // small inputs: a small cache (the clear is part of every call) and the plain long path (fewer launches)
    // Slots never touched cost nothing but their share of the clear: a large table keeps hot keys in their first slot (a
    // displaced key sends its whole round through the probe loop) and leaves room for corpora with millions of distinct
    // pre-tokens.  2^22 slots = 128 MB, cleared in ~0.04 ms; small inputs use a prefix sized to them (the clear is part of every call).
    uint32_t cache_slots = 1u << CACHE_LOG2;
    if (const char* e = getenv("CTK_CACHE_LOG2")) { int v = atoi(e); if (v >= 10 && v <= CACHE_LOG2) cache_slots = 1u << v; }
    while (cache_slots > 1024 && (uint64_t)cache_slots * 4 > n_bytes) cache_slots >>= 1;   // ~one slot per 4-8 input bytes
    const uint32_t ovf_cap = 1u << 18;
    uint32_t *first_doc, *ctrl, *slice_base;
    CK(ws.get(0, (p.n_slices + 1) * 4, (void**)&first_doc));
    CK(ws.get(18, (uint64_t)(1u << CACHE_LOG2) * sizeof(CacheSlot), (void**)&p.cache));       // always the full table; a call uses a prefix
    CK(ws.get(19, (uint64_t)ovf_cap * 64, (void**)&p.ovf_pool));
    CK(ws.get(4, 256, (void**)&ctrl));
    CK(ws.get(1, (p.n_slices + 2) * 4, (void**)&p.slice_cnt));
    CK(ws.get(2, (p.n_slices + 2) * 4, (void**)&slice_base));
    CK(ws.get(3, (p.n_slices + 2) * 4, (void**)&p.slice_info));
    CK(ws.get(6, (p.n_slices + 2) * 4, (void**)&p.slice_desc));
    p.run_width = eng.run_width;
    CK(ws.get(8, p.n_slices * (uint64_t)STAGE * p.run_width, (void**)&p.runs));
    p.long_cap = n_bytes + 64;
    CK(ws.get(7, p.long_cap * 4, (void**)&p.long_pool));
    p.desc_cap = (uint32_t)(n_bytes / 33 + 16);
    CK(ws.get(9, (uint64_t)p.desc_cap * sizeof(LongDesc), (void**)&p.desc));
    CK(ws.get(41, (uint64_t)p.desc_cap * sizeof(XlEntry), (void**)&p.xl_list));
    CK(ws.get(44, (uint64_t)MID_LISTS * p.desc_cap * 4, (void**)&p.work_list));
    CK(ws.get(42, (p.n_slices + 2) * 2, (void**)&p.slice_first));
    CK(ws.get(43, (n_docs + 1) * 8, (void**)&p.ids_off_rel));
    p.xl_enabled = n_ids_host != nullptr && eng.model.merges_monotone && eng.model.max_token_span <= XL_MAX_WINDOW &&
                   eng.tables.n_added == 0 && !getenv("CTK_NO_XLONG");
    p.no_rounds = getenv("CTK_NO_ROUNDS") != nullptr;
    p.mid_enabled = n_bytes > (256u << 10) && eng.tables.n_added == 0 && (eng.model.pairs.empty() || eng.model.pairs.back().rank < (1u << 24)) && !getenv("CTK_NO_MID");   // rank << 8 | slot keys
    p.check_nfc = check_nfc ? 1 : 0;
    p.first_doc = first_doc; p.cache_mask = cache_slots - 1; p.ovf_cap = ovf_cap;
    p.cache_shift = 32;
    while ((1ull << (32 - p.cache_shift)) < cache_slots) --p.cache_shift;
    uint32_t max_id = eng.max_emit_id ? eng.max_emit_id : 1u;
    p.id_bits = 1;
    while ((1ull << p.id_bits) <= max_id) ++p.id_bits;
    if (p.run_width == 2) p.id_bits = 16;                                 // k_encode_slices<2> unpacks fixed 16-bit fields
    p.n_inline = 96 / p.id_bits;
    if (p.n_inline > MAXINLINE) p.n_inline = MAXINLINE;
    // ctrl words: [0] err flags, [2] desc cursor, [3] ovf cursor, [4..5] long cursor, [6..7] xlong cursor, [8] holes, [9] round size,
    //             [10..14] work-list lengths
    p.work_count = ctrl + 10;
    p.xl_cursor = reinterpret_cast<unsigned long long*>(ctrl + 6);
    p.err = ctrl; p.desc_cursor = ctrl + 2; p.ovf_cursor = ctrl + 3; p.long_cursor = reinterpret_cast<unsigned long long*>(ctrl + 4);
    p.ids_off = d_ids_off;
    { const char* a = getenv("CTK_ABLATE"); p.ablate = a ? atoi(a) : 0; }
    eng.mark(nullptr, st);
    if (!((eng.cache_persistent || eng.keep_cache_once) && eng.cache_valid)) {
        CK(cudaMemsetAsync(p.cache, 0xFF, (uint64_t)cache_slots * sizeof(CacheSlot), st));
        CK(cudaMemsetAsync(ctrl, 0, 256, st));
        eng.cache_valid = true;
        eng.cache_init_slots = cache_slots;
    } else {
        if (cache_slots > eng.cache_init_slots) {                      // kept cache, larger prefix than ever cleared: clear the new part
            CK(cudaMemsetAsync(p.cache + eng.cache_init_slots, 0xFF, (uint64_t)(cache_slots - eng.cache_init_slots) * sizeof(CacheSlot), st));
            eng.cache_init_slots = cache_slots;
        }
        CK(cudaMemsetAsync(ctrl, 0, 12, st));                          // keep the overflow cursor
        CK(cudaMemsetAsync(ctrl + 4, 0, 48, st));
    }
    eng.mark("memset(cache)", st);
    unsigned doc_grid = (unsigned)((n_docs + 1 + 255) / 256);
    k_first_doc<<<doc_grid, 256, 0, st>>>(d_off, n_docs, n_bytes, p.n_slices, first_doc, p.err);
    eng.launched(1); eng.mark("k_first_doc", st);
    unsigned grid = p.n_tiles < (uint32_t)eng.fused_grid ? p.n_tiles : (unsigned)eng.fused_grid;
    if (p.run_width == 2) k_encode_slices<2><<<grid, FW * 32, 0, st>>>(p);
    else k_encode_slices<4><<<grid, FW * 32, 0, st>>>(p);
    eng.launched(1); eng.mark("k_encode_slices", st);
    k_long_prep<<<eng.long_grid, 256, 0, st>>>(p);
    if (p.mid_enabled) {
        k_encode_mid<MID_N0><<<eng.mid_grid[0], 32, mid_smem(MID_N0), st>>>(p, 0);
        k_encode_mid<MID_N1><<<eng.mid_grid[1], 32, mid_smem(MID_N1), st>>>(p, 1);
        k_encode_mid<MID_N2><<<eng.mid_grid[2], 32, mid_smem(MID_N2), st>>>(p, 2);
        k_encode_mid<MID_N3><<<eng.mid_grid[3], 32, mid_smem(MID_N3), st>>>(p, 3);
        eng.launched(4);
    }
    k_encode_long<<<eng.long_grid, 256, 0, st>>>(p);
    eng.launched(2); eng.mark("k_long_prep+mid+long", st);
    size_t cub_bytes = 0;
    void* cub_tmp;
    CK(cudaMemsetAsync(p.slice_cnt + p.n_slices, 0, 4, st));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, p.slice_cnt, slice_base, p.n_slices + 1, st));
    CK(ws.get(5, cub_bytes + 16, &cub_tmp));
    XlState xl{};
    for (int pass = 0;; ++pass) {
        CK(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, p.slice_cnt, slice_base, p.n_slices + 1, st));
        eng.launched(1); eng.mark("scan(slice counts)", st);
        {
            const unsigned cg = (unsigned)((p.n_slices + 7) / 8);
#define CTK_COMPACT(RW_, OW_) k_compact<RW_, OW_><<<cg, 256, 0, st>>>(p.runs, slice_base, p.slice_info, p.slice_desc, p.desc, p.long_pool, \
                                                                     p.n_slices, d_ids, ids_cap, p.err)
            if (p.run_width == 2) { if (out_width == 2) CTK_COMPACT(2, 2); else CTK_COMPACT(2, 4); }
            else { if (out_width == 2) CTK_COMPACT(4, 2); else CTK_COMPACT(4, 4); }
#undef CTK_COMPACT
        }
        eng.launched(1); eng.mark("k_compact", st);
        if (pass == 1) {
            k_xl_dst<<<(xl.n_list + 255) / 256, 256, 0, st>>>(p, p.xl_list, xl.n_list, slice_base, xl.region_dst);
            if (out_width == 2) k_xl_place<uint16_t><<<(xl.n + 255) / 256, 256, 0, st>>>(xl.xs, xl.n, xl.n_list, xl.sep_pos, xl.region_dst, (uint16_t*)d_ids, ids_cap);
            else k_xl_place<uint32_t><<<(xl.n + 255) / 256, 256, 0, st>>>(xl.xs, xl.n, xl.n_list, xl.sep_pos, xl.region_dst, (uint32_t*)d_ids, ids_cap);
            eng.launched(2); eng.mark("k_xl_place", st);
        }
        k_doc_fixup<<<doc_grid, 256, 0, st>>>(d_off, n_docs, p.n_slices, slice_base, p.slice_info, p.slice_desc, p.desc, p.ids_off_rel, d_ids_off);
        eng.launched(1); eng.mark("k_doc_fixup", st);
        CK(cudaGetLastError());
        if (p.xl_enabled && pass == 0) CK(eng.publish({{ctrl + 6, 2, 4}}, st));
        int rc = eng.finish(p.err, d_ids_off, n_docs, n_ids_host, st);
        if (rc != CTK_OK || !p.xl_enabled || pass == 1) return rc;
        // very long pre-tokens were set aside (rare): merge them in rounds, then place every id again
        uint64_t cursor;
        memcpy(&cursor, eng.h_flags + 4, 8);
        if ((cursor >> XL_IDX_SHIFT) == 0) return rc;
        rc = xlong_rounds(eng, p, cursor, ctrl, xl, st);
        if (rc != CTK_OK) return rc;
    }
}

}  // namespace ctk
