// Fused single-pass encode kernel: text in, packed ids + per-document offsets out, one read of the
// text and one write of the ids (algorithmic traffic B + 4T + 16(D+1)).
//
// Work decomposition
//   CTA  = 8 warps = one tile of 8 x 448 bytes; tiles are handed out by an atomic ticket so that a
//          tile never waits on a tile that has not started (decoupled look-back, below).
//   warp = one 512-byte chunk: 16 bytes of left context + 448 OWNED bytes + 48 bytes of right
//          context.  Lane l holds bytes [16 l, 16 l + 16) of the chunk in registers (one coalesced
//          16-byte load per lane) and the chunk is mirrored in shared memory for unaligned access.
// Stages inside a warp (reference lines in brackets)
//   1 classify: SWAR class masks per lane, trie for non-ASCII   [pretokenizers.rs:13 \p{L} \p{N} \s]
//   2 boundaries: 32-bit window logic + warp shuffles           [pretokenizers.rs:13, :158-185]
//   3 compaction: ballot-free prefix sum of per-lane popcounts -> list of pre-token starts
//   4 per pre-token, 32 at a time:
//       <= 16 bytes: look up the batch's pre-token cache (one 32-byte L2 sector per probe); BPE of a
//                    pre-token is a pure function of its bytes, so each distinct pre-token of a batch is
//                    merged once and every other occurrence copies the ids  [bpe.rs:88-153 is pure]
//       miss or 17..32 bytes: warp-cooperative merge loop (bpe_warp32)        [bpe.rs:104-153]
//       > 32 bytes: deferred to the end of the warp's work, merged in global scratch
//   5 ids are staged in shared memory in pre-token order (mod.rs:562-612 `result.extend`)
// Then per CTA: sum of the 8 warp totals -> decoupled look-back over tile_state -> coalesced copy of
// the staged ids to their final place, and ids_off[d] for every document that starts in the tile.
#include "engine.hpp"
#include "start_window.cuh"

namespace ctk {

constexpr int FW = 8;            // warps per CTA
constexpr int SLICE = 448;       // bytes owned by a warp
constexpr int CHUNK = 512;       // bytes a warp looks at
constexpr int LCTX = 16;         // left context
constexpr int STAGE = 480;       // ids a warp can stage: owned pre-tokens of <= 32 bytes cover < 480 bytes
constexpr int MAXLONG = 14;      // pre-tokens longer than 32 bytes that can start in 448 bytes
constexpr uint32_t META_EMPTY = 0xFFFFFFFFu, META_BUSY = 0xFFFFFFFEu;
constexpr int PROBES = 4;
constexpr uint32_t END_UNKNOWN = 0xFFFFu;

struct FusedParams {
    DevTables t;
    const uint8_t* text; uint64_t n_bytes;
    const uint64_t* off; uint64_t n_docs;
    const uint32_t* first_doc; uint64_t n_slices;
    CacheSlot* cache; uint32_t cache_mask;
    uint32_t* ovf_pool; uint32_t ovf_cap; uint32_t* ovf_cursor;
    uint32_t* long_pool; unsigned long long long_cap; unsigned long long* long_cursor;
    unsigned long long* tile_state; uint32_t* ticket;
    uint32_t* out; uint64_t out_cap; uint64_t* ids_off; uint32_t* err;
};

struct __align__(16) WarpSmem {
    uint8_t pad0[16];
    uint8_t chunk[CHUNK];
    uint8_t pad1[16];
    uint32_t ds[32];
    uint32_t stage[STAGE];
    uint16_t list[SLICE + 8];
    uint16_t l_at[MAXLONG + 2], l_k[MAXLONG + 2], l_pos[MAXLONG + 2], l_len[MAXLONG + 2];
    uint32_t l_pool[MAXLONG + 2], l_cnt[MAXLONG + 2];
};

// first_doc[s] = smallest d with off[d] >= s*SLICE - LCTX ; also validates the offsets
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
    for (uint64_t s = s_lo; s <= s_hi; ++s) first_doc[s] = (uint32_t)d;
}

__device__ __forceinline__ void load_slot(const CacheSlot* p, uint64_t& k0, uint64_t& k1, uint32_t& meta, uint32_t& t0,
                                          uint32_t& t1, uint32_t& t2) {
    uint64_t c, d;
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(k0), "=l"(k1), "=l"(c), "=l"(d) : "l"(p));
    meta = (uint32_t)c; t0 = (uint32_t)(c >> 32); t1 = (uint32_t)d; t2 = (uint32_t)(d >> 32);
}

__device__ __forceinline__ uint32_t key_hash(uint64_t k0, uint64_t k1, uint32_t len) {
    uint64_t h = (k0 ^ (k1 * 0x9E3779B97F4A7C15ull) ^ len) * 0xD6E8FEB86659FD93ull;
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    return (uint32_t)(h >> 32);
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t& total) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(full, incl, o); if (lane >= o) incl += u; }
    total = __shfl_sync(full, incl, 31);
    return incl - v;
}

// initial ids of `len` bytes at text[...] into lanes (<= 32), unknown bytes dropped (bpe.rs:94-97); returns count
__device__ __forceinline__ int init_symbols32(const uint32_t* s_byte_init, const uint8_t* bytes, int len, uint32_t& sym) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
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

__global__ void __launch_bounds__(FW * 32, 5) k_encode_fused(const FusedParams p) {
    const unsigned full = 0xFFFFFFFFu;
    __shared__ WarpSmem sm[FW];
    __shared__ uint32_t s_byte_init[256];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_wtot[FW];
    __shared__ unsigned long long s_tile_base;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    s_byte_init[tid] = __ldg(p.t.byte_init + tid);
    if (tid == 0) s_tile = atomicAdd(p.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t slice = (uint64_t)tile * FW + w;
    const bool active = slice < p.n_slices;
    WarpSmem& S = sm[w];
    const long long lo = (long long)slice * SLICE, cb = lo - LCTX;     // chunk base (may be -16 for slice 0)
    const uint32_t d0 = active ? __ldg(p.first_doc + slice) : 0;
    uint32_t n_owned = 0, stage_cnt = 0, n_long = 0, first_k = 0, ownm = 0, long_total = 0;

    if (active) {
        // ---- 1. load the chunk: one 16-byte vector per lane, zero beyond the text
        const long long q = cb + 16 * lane;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (q >= 0 && (uint64_t)q < p.n_bytes) {
            v = __ldg(reinterpret_cast<const uint4*>(p.text + q));
            if ((uint64_t)q + 16 > p.n_bytes) {
                int keep = (int)(p.n_bytes - (uint64_t)q);             // 1..15 valid bytes
                uint32_t wv[4] = {v.x, v.y, v.z, v.w};
                for (int k = 0; k < 4; ++k) {
                    int b = keep - 4 * k;
                    wv[k] = b >= 4 ? wv[k] : (b <= 0 ? 0u : (wv[k] & ((1u << (8 * b)) - 1u)));
                }
                v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            }
        }
        *reinterpret_cast<uint4*>(S.chunk + 16 * lane) = v;
        if (lane == 0) { *reinterpret_cast<uint4*>(S.pad0) = make_uint4(0, 0, 0, 0); *reinterpret_cast<uint4*>(S.pad1) = make_uint4(0, 0, 0, 0); }
        S.ds[lane] = 0;
        __syncwarp();
        // ---- document starts inside the chunk (position n_bytes = off[n_docs] counts as one)
        for (uint64_t d = (uint64_t)d0 + lane;; d += 32) {
            uint64_t pos = d <= p.n_docs ? __ldg(p.off + d) : ~0ull;
            bool in = (long long)pos < cb + CHUNK && pos != ~0ull;
            if (in) { uint32_t rel = (uint32_t)((long long)pos - cb); atomicOr(&S.ds[rel >> 4], 1u << (rel & 15)); }
            if (!__all_sync(full, in)) break;
        }
        __syncwarp();
        // ---- 2. classes and boundaries
        Masks16 m = classify16(S.chunk, 16 * lane, v.x, v.y, v.z, v.w, p.t.trie_index, p.t.trie_blocks);
        uint32_t ds16 = S.ds[lane];
        uint32_t pa = m.L | (m.N << 16), pb = m.W | (m.SP << 16), pc = m.AP | (m.CONT << 16);
        uint32_t ua = __shfl_up_sync(full, pa, 1), ub = __shfl_up_sync(full, pb, 1), uc = __shfl_up_sync(full, pc, 1),
                 ud = __shfl_up_sync(full, ds16, 1);
        uint32_t na = __shfl_down_sync(full, pa, 1), nb = __shfl_down_sync(full, pb, 1), nc = __shfl_down_sync(full, pc, 1),
                 nd = __shfl_down_sync(full, ds16, 1);
        if (lane == 0) { ua = ub = uc = ud = 0; }
        if (lane == 31) { na = nb = nc = nd = 0; }
        uint32_t S32 = start_window(window(ua & 0xFFFF, m.L, na & 0xFFFF), window(ua >> 16, m.N, na >> 16),
                                    window(ub & 0xFFFF, m.W, nb & 0xFFFF), window(ub >> 16, m.SP, nb >> 16),
                                    window(uc & 0xFFFF, m.AP, nc & 0xFFFF), window(uc >> 16, m.CONT, nc >> 16),
                                    window(ud, ds16, nd), S.chunk + 16 * lane - 8);
        uint32_t own16 = (S32 >> 8) & 0xFFFFu;
        // positions at or beyond the end of the text are not pre-token starts of anything we own
        {
            long long room = (long long)p.n_bytes - q;               // valid positions in this group
            uint32_t valid = room >= 16 ? 0xFFFFu : (room <= 0 ? 0u : ((1u << room) - 1u));
            ownm = (lane >= 1 && lane <= 28) ? (own16 & valid) : 0u;
            // ---- 3. compaction: list of owned pre-token starts, then the sentinel (first start in the right context)
            uint32_t c = __popc(ownm);
            first_k = warp_excl_scan(c, n_owned);
            uint32_t bits = ownm, o = first_k;
            while (bits) { int b = __ffs(bits) - 1; bits &= bits - 1; S.list[o++] = (uint16_t)(16 * lane + b); }
            // sentinel candidates: starts in the right context (lanes 29, 30) and the end of the text
            // (position n_bytes is a start thanks to its DS bit) wherever it falls
            uint32_t rc = 0;
            if (lane == 29 || lane == 30) rc = own16 & (room >= 16 ? 0xFFFFu : (room < 0 ? 0u : ((2u << room) - 1u)));
            else if (lane >= 1 && lane <= 28 && room >= 0 && room < 16) rc = own16 & (1u << room);
            unsigned bal = __ballot_sync(full, rc != 0);
            uint32_t sent = END_UNKNOWN;
            if (bal) { int sl = __ffs(bal) - 1; uint32_t r2 = __shfl_sync(full, rc, sl); sent = 16 * sl + (__ffs(r2) - 1); }
            if (lane == 0) S.list[n_owned] = (uint16_t)sent;
        }
        __syncwarp();

        // ---- 4. pre-tokens, 32 per round
        for (uint32_t base_k = 0; base_k < n_owned; base_k += 32) {
            const uint32_t k = base_k + lane;
            const bool have = k < n_owned;
            uint32_t pos = have ? S.list[k] : 0, end = have ? S.list[k + 1] : 0;
            uint32_t len = end - pos;
            __syncwarp();
            // kind: 0 nothing, 1 cache hit (<= 3 ids inline), 2 cache hit (ids in the overflow pool), 3 needs merging, 4 long
            int kind = 0;
            uint32_t ntok = 0, t0 = 0, t1 = 0, t2 = 0, ins = kNone;
            uint64_t k0 = 0, k1 = 0;
            if (have) {
                if (end == END_UNKNOWN || len > 32) kind = 4;
                else if (len > 16) kind = 3;
                else {
                    const uint32_t* wp = reinterpret_cast<const uint32_t*>(S.chunk + (pos & ~3u));
                    uint32_t a0 = wp[0], a1 = wp[1], a2 = wp[2], a3 = wp[3], a4 = wp[4];
                    uint32_t sh = (pos & 3u) * 8;
                    uint32_t x0 = __funnelshift_r(a0, a1, sh), x1 = __funnelshift_r(a1, a2, sh), x2 = __funnelshift_r(a2, a3, sh),
                             x3 = __funnelshift_r(a3, a4, sh);
                    k0 = x0 | ((uint64_t)x1 << 32);
                    k1 = x2 | ((uint64_t)x3 << 32);
                    if (len < 8) { k0 &= (1ull << (8 * len)) - 1ull; k1 = 0; }
                    else if (len < 16) k1 &= (1ull << (8 * (len - 8))) - 1ull;
                    uint32_t h = key_hash(k0, k1, len);
                    kind = 3;
                    for (int pr = 0; pr < PROBES; ++pr) {
                        uint32_t idx = (h + pr) & p.cache_mask;
                        uint64_t s0, s1; uint32_t meta, u0, u1, u2;
                        load_slot(p.cache + idx, s0, s1, meta, u0, u1, u2);
                        if (meta == META_EMPTY) { ins = idx; break; }
                        if (meta < META_BUSY && (meta & 0xFFu) == len && s0 == k0 && s1 == k1) {
                            ntok = (meta >> 8) & 0xFFu; t0 = u0; t1 = u1; t2 = u2;
                            kind = ntok <= 3 ? 1 : 2;
                            break;
                        }
                    }
                }
            }
            uint32_t hit_total;
            uint32_t E = warp_excl_scan((kind == 1 || kind == 2) ? ntok : 0u, hit_total);
            // misses and 17..32-byte pre-tokens: merge cooperatively, one after the other, in lane order
            uint32_t extra = 0;
            unsigned mm = __ballot_sync(full, kind == 3);
            while (mm) {
                const int src = __ffs(mm) - 1;
                mm &= mm - 1;
                const uint32_t spos = __shfl_sync(full, pos, src), slen = __shfl_sync(full, len, src);
                uint32_t sym;
                int n = init_symbols32(s_byte_init, S.chunk + spos, (int)slen, sym);
                int cnt = n ? bpe_warp32(p.t, sym, n) : 0;
                const uint32_t o = stage_cnt + __shfl_sync(full, E, src) + extra;
                if (lane < cnt) S.stage[o + lane] = sym;
                const uint32_t sins = __shfl_sync(full, ins, src);
                if (sins != kNone) {                                   // publish in the batch cache
                    uint32_t a0 = __shfl_sync(full, sym, 0), a1 = __shfl_sync(full, sym, 1), a2 = __shfl_sync(full, sym, 2);
                    bool ok = true;
                    if (cnt > 3) {
                        uint32_t rec = 0;
                        if (lane == src) rec = atomicAdd(p.ovf_cursor, 1u);
                        rec = __shfl_sync(full, rec, src);
                        ok = rec < p.ovf_cap;
                        if (ok && lane < cnt) p.ovf_pool[(uint64_t)rec * 16 + lane] = sym;
                        a0 = rec;
                    }
                    if (ok && lane == src) {
                        CacheSlot* sl = p.cache + sins;
                        if (atomicCAS(&sl->meta, META_EMPTY, META_BUSY) == META_EMPTY) {
                            sl->k0 = k0; sl->k1 = k1; sl->tok[0] = a0; sl->tok[1] = a1; sl->tok[2] = a2;
                            __threadfence();
                            *reinterpret_cast<volatile uint32_t*>(&sl->meta) = slen | ((uint32_t)cnt << 8);
                        }
                    }
                }
                if (lane == src) { ntok = (uint32_t)cnt; kind = 5; }
                extra += (uint32_t)cnt;
            }
            uint32_t round_total;
            uint32_t F = warp_excl_scan(kind == 4 ? 0u : ntok, round_total);
            const uint32_t o = stage_cnt + F;
            if (kind == 1) {
                if (ntok > 0) S.stage[o] = t0;
                if (ntok > 1) S.stage[o + 1] = t1;
                if (ntok > 2) S.stage[o + 2] = t2;
            } else if (kind == 2) {
                for (uint32_t i = 0; i < ntok; ++i) S.stage[o + i] = p.ovf_pool[(uint64_t)t0 * 16 + i];
            }
            // the list entry now becomes the pre-token's id offset inside the warp's stage (for ids_off)
            if (have) S.list[k] = (uint16_t)o;
            unsigned lm = __ballot_sync(full, kind == 4);
            while (lm) {
                const int src = __ffs(lm) - 1;
                lm &= lm - 1;
                uint32_t at = __shfl_sync(full, o, src), lp = __shfl_sync(full, pos, src), le = __shfl_sync(full, end, src);
                if (lane == 0 && n_long < MAXLONG) {
                    S.l_at[n_long] = (uint16_t)at; S.l_k[n_long] = (uint16_t)(base_k + src); S.l_pos[n_long] = (uint16_t)lp;
                    S.l_len[n_long] = (uint16_t)(le == END_UNKNOWN ? END_UNKNOWN : le - lp);
                }
                ++n_long;
            }
            stage_cnt += round_total;
            __syncwarp();
        }
        if (lane == 0) S.list[n_owned] = (uint16_t)stage_cnt;

        // ---- long pre-tokens: merged in global scratch (symbols compacted in place)
        for (uint32_t j = 0; j < n_long; ++j) {
            const uint64_t gstart = (uint64_t)(cb + S.l_pos[j]);
            uint64_t len = S.l_len[j];
            if (len == END_UNKNOWN) {                                  // runs past the chunk: find its end
                uint64_t dl = 0, dh = p.n_docs;                        // last doc with off[d] <= gstart
                while (dl + 1 < dh) { uint64_t mid = (dl + dh) >> 1; if (__ldg(p.off + mid) <= gstart) dl = mid; else dh = mid; }
                TextView tv{p.text, p.n_bytes, nullptr, p.t.trie_index, p.t.trie_blocks, __ldg(p.off + dl), __ldg(p.off + dl + 1)};
                uint64_t e = 0;
                for (uint64_t i = (uint64_t)(cb + CHUNK - 16) + lane;; i += 32) {
                    bool s = i >= tv.dhi || tv.is_start(i);
                    unsigned b = __ballot_sync(full, s);
                    if (b) { e = i - lane + (__ffs(b) - 1); break; }
                }
                len = e - gstart;
            }
            unsigned long long po = 0;
            if (lane == 0) po = atomicAdd(p.long_cursor, (unsigned long long)len);
            po = __shfl_sync(full, po, 0);
            uint32_t cnt = 0;
            if (po + len <= p.long_cap) {
                uint32_t* sym = p.long_pool + po;
                uint32_t n = 0;
                for (uint64_t b0 = 0; b0 < len; b0 += 32) {
                    uint64_t i = b0 + lane;
                    uint32_t sv = i < len ? s_byte_init[__ldg(p.text + gstart + i)] : kNone;
                    unsigned hv = __ballot_sync(full, sv != kNone);
                    if (sv != kNone) sym[n + __popc(hv & ((1u << lane) - 1u))] = sv;
                    n += __popc(hv);
                }
                __syncwarp();
                cnt = (uint32_t)bpe_warp_long(p.t, sym, (int)n);
            } else if (lane == 0) atomicOr(p.err, ERRF_POOL);
            if (lane == 0) { S.l_pool[j] = (uint32_t)po; S.l_cnt[j] = cnt; }
            long_total += cnt;
        }
        __syncwarp();
    }

    // ---- per CTA: tile total, decoupled look-back, final positions
    if (lane == 0) s_wtot[w] = stage_cnt + long_total;
    __syncthreads();
    if (w == 0) {
        uint32_t tv = lane < FW ? s_wtot[lane] : 0u, tile_total;
        warp_excl_scan(tv, tile_total);
        volatile unsigned long long* st = p.tile_state;
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) st[0] = (2ull << 62) | tile_total;
        } else {
            if (lane == 0) st[tile] = (1ull << 62) | tile_total;
            long long look = (long long)tile - 1;
            for (;;) {
                long long idx = look - lane;
                unsigned long long sv;
                do { sv = idx >= 0 ? st[idx] : (2ull << 62); } while (__any_sync(full, (sv >> 62) == 0));
                unsigned pm = __ballot_sync(full, (sv >> 62) == 2);
                unsigned long long val = sv & ((1ull << 62) - 1);
                if (pm) {
                    int first = __ffs(pm) - 1;
                    if (lane > first) val = 0;
                }
                for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(full, val, o);
                excl += val;
                if (pm) break;
                look -= 32;
            }
            if (lane == 0) st[tile] = (2ull << 62) | (excl + tile_total);
        }
        if (lane == 0) s_tile_base = excl;
    }
    __syncthreads();
    if (!active) return;
    unsigned long long base = s_tile_base;
    for (int k = 0; k < w; ++k) base += s_wtot[k];
    const uint32_t wtotal = stage_cnt + long_total;
    if (base + wtotal > p.out_cap) { if (lane == 0) atomicOr(p.err, ERRF_CAPACITY); return; }
    uint32_t* out = p.out + base;
    if (n_long == 0) {
        for (uint32_t i = lane; i < stage_cnt; i += 32) out[i] = S.stage[i];
    } else {
        // staged ids interleaved with the long pre-tokens' ids (rare)
        uint32_t nl = n_long < MAXLONG ? n_long : MAXLONG;
        for (uint32_t i = lane; i < stage_cnt; i += 32) {
            uint32_t add = 0;
            for (uint32_t j = 0; j < nl; ++j) if (S.l_at[j] <= i) add += S.l_cnt[j];
            out[i + add] = S.stage[i];
        }
        uint32_t before = 0;
        for (uint32_t j = 0; j < nl; ++j) {
            const uint32_t* src = p.long_pool + S.l_pool[j];
            uint32_t* dst = out + S.l_at[j] + before;
            for (uint32_t i = lane; i < S.l_cnt[j]; i += 32) dst[i] = src[i];
            before += S.l_cnt[j];
        }
        if (n_long > MAXLONG && lane == 0) atomicOr(p.err, ERRF_POOL);
    }
    // ---- ids_off for the documents that start in the owned bytes (or at the very end of the text)
    const bool last_slice = slice + 1 == p.n_slices;
    for (uint64_t d = (uint64_t)d0 + lane;; d += 32) {
        uint64_t pos = d <= p.n_docs ? __ldg(p.off + d) : ~0ull;
        bool in = (long long)pos < cb + CHUNK && pos != ~0ull;
        bool own = in && (((long long)pos >= lo && (long long)pos < lo + SLICE) || (last_slice && pos == p.n_bytes && (long long)pos >= lo));
        uint32_t rel = own ? (uint32_t)((long long)pos - cb) : 0u;
        uint32_t fk = __shfl_sync(full, first_k, rel >> 4), sb = __shfl_sync(full, ownm, rel >> 4);
        if (own) {
            uint32_t k = fk + __popc(sb & ((1u << (rel & 15)) - 1u));
            if (k > n_owned) k = n_owned;
            unsigned long long tokoff = S.list[k];
            uint32_t nl = n_long < MAXLONG ? n_long : MAXLONG;
            for (uint32_t j = 0; j < nl; ++j) if (S.l_k[j] < k) tokoff += S.l_cnt[j];
            p.ids_off[d] = base + tokoff;
        }
        if (!__all_sync(full, in)) break;
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return eng.cuda_fail(e_, #x); } while (0)

int encode_fused(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes,
                 uint32_t* d_ids, uint64_t ids_cap, uint64_t* d_ids_off, uint64_t* n_ids_host, cudaStream_t st) {
    if (n_bytes >= 0xFFFFFFF0ull) return eng.fail(CTK_ERR_ARG, "one device call handles less than 4 GiB of text");
    if (n_bytes == 0) {
        CK(cudaMemsetAsync(d_ids_off, 0, (n_docs + 1) * 8, st));
        if (n_ids_host) { CK(cudaStreamSynchronize(st)); *n_ids_host = 0; }
        return CTK_OK;
    }
    Workspace& ws = eng.ws;
    FusedParams p{};
    p.t = eng.tables;
    p.text = d_text; p.n_bytes = n_bytes; p.off = d_off; p.n_docs = n_docs;
    p.n_slices = (n_bytes + SLICE - 1) / SLICE;
    uint32_t n_tiles = (uint32_t)((p.n_slices + FW - 1) / FW);
    const uint32_t cache_slots = 1u << 20, ovf_cap = 1u << 18;
    uint32_t *first_doc, *ctrl;
    CK(ws.get(0, (p.n_slices + 1) * 4, (void**)&first_doc));
    CK(ws.get(18, (uint64_t)cache_slots * sizeof(CacheSlot), (void**)&p.cache));
    CK(ws.get(19, (uint64_t)ovf_cap * 64, (void**)&p.ovf_pool));
    CK(ws.get(3, ((uint64_t)n_tiles + 1) * 8, (void**)&p.tile_state));
    CK(ws.get(4, 256, (void**)&ctrl));
    p.long_cap = n_bytes + 64;
    CK(ws.get(7, p.long_cap * 4, (void**)&p.long_pool));
    p.first_doc = first_doc; p.cache_mask = cache_slots - 1; p.ovf_cap = ovf_cap;
    // ctrl words: [0] err flags, [2] ticket, [3] ovf cursor, [4..5] long cursor
    p.err = ctrl; p.ticket = ctrl + 2; p.ovf_cursor = ctrl + 3; p.long_cursor = reinterpret_cast<unsigned long long*>(ctrl + 4);
    p.out = d_ids; p.out_cap = ids_cap; p.ids_off = d_ids_off;
    eng.mark(nullptr, st);
    if (!((eng.cache_persistent || eng.keep_cache_once) && eng.cache_valid)) {
        CK(cudaMemsetAsync(p.cache, 0xFF, (uint64_t)cache_slots * sizeof(CacheSlot), st));
        CK(cudaMemsetAsync(ctrl, 0, 256, st));
        eng.cache_valid = true;
    } else {
        CK(cudaMemsetAsync(ctrl, 0, 12, st));                          // keep the overflow cursor
        CK(cudaMemsetAsync(ctrl + 4, 0, 8, st));
    }
    CK(cudaMemsetAsync(p.tile_state, 0, ((uint64_t)n_tiles + 1) * 8, st));
    eng.mark("memset(cache,state)", st);
    k_first_doc<<<(unsigned)((n_docs + 1 + 255) / 256), 256, 0, st>>>(d_off, n_docs, n_bytes, p.n_slices, first_doc, p.err);
    eng.launched(1); eng.mark("k_first_doc", st);
    k_encode_fused<<<n_tiles, FW * 32, 0, st>>>(p);
    eng.launched(1); eng.mark("k_encode_fused", st);
    CK(cudaGetLastError());
    return eng.finish(p.err, d_ids_off, n_docs, n_ids_host, st);
}

}  // namespace ctk
