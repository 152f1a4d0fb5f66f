// BPE of the pre-tokens that are longer than 32 bytes (CJK runs, URLs, long digit or punctuation runs).
// Included by encode_fused.cu after FusedParams.
//   k_long_prep      length of those that run past their chunk, room in the long pool, work lists by length class
//   k_encode_mid<N>  33 .. 128 bytes, one LANE per pre-token (32 per warp): exact for every merge table
//   k_encode_long    the rest up to 256 bytes (and everything when k_encode_mid is off), one WARP per pre-token:
//                    round-parallel merging in shared memory (bpe_warp_rounds) when the table allows, else
//                    bpe_warp_regs / bpe_warp_long as described next; > 256 bytes are handed to encode_xlong.cuh
//
// Same order as the reference (bpe.rs:104-153): every iteration applies ONE merge, the lowest rank,
// leftmost on ties.  Up to 256 symbols live in registers, blocked K per lane, together with the cached
// (rank, new id) of the pair that starts at each symbol, so an iteration is: local min over K ranks ->
// redux.min -> ballot for the leftmost holder -> shift everything behind the consumed symbol left by
// one (register moves + one shuffle) -> re-probe only the two pairs that touch the new symbol.
// Longer ones fall back to bpe_warp_long (symbols in global scratch, every pair re-probed per merge).
#pragma once

namespace ctk {

constexpr uint32_t XL_MIN = 256;                 // pre-tokens longer than this take the round-parallel path (encode_xlong.cuh)

template <int K>
__device__ __forceinline__ int bpe_warp_regs(const DevTables& t, uint32_t (&s)[K], int n, int lane) {
    const unsigned full = 0xFFFFFFFFu;
    uint32_t r[K], v[K];
    {
        const uint32_t nx0 = __shfl_down_sync(full, s[0], 1);
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const uint32_t right = j + 1 < K ? s[j + 1 < K ? j + 1 : j] : nx0;
            r[j] = kNone; v[j] = 0;
            if (lane * K + j + 1 < n) { uint2 q = pair_lookup(t, s[j], right); r[j] = q.x; v[j] = q.y; }
        }
    }
    while (n > 1) {
        uint32_t m = r[0];
#pragma unroll
        for (int j = 1; j < K; ++j) m = r[j] < m ? r[j] : m;
        const uint32_t g = __reduce_min_sync(full, m);
        if (g == kNone) break;
        const int L = __ffs(__ballot_sync(full, m == g)) - 1;      // leftmost lane holding the lowest rank
        int myJ = K;
        uint32_t myV = 0;
#pragma unroll
        for (int j = K - 1; j >= 0; --j) if (r[j] == g) { myJ = j; myV = v[j]; }
        const int J = __shfl_sync(full, myJ, L);
        const uint32_t newid = __shfl_sync(full, myV, L);
        const int P = L * K + J;                                   // position of the merged pair's left symbol
        // close the gap behind the consumed symbol (position P + 1)
        {
            uint32_t ns = __shfl_down_sync(full, s[0], 1), nr = __shfl_down_sync(full, r[0], 1), nv = __shfl_down_sync(full, v[0], 1);
            if (lane == 31) { ns = kNone; nr = kNone; nv = 0; }        // nothing beyond the last lane: keep "no pair" there
            const int from = lane > L ? 0 : (lane == L ? J + 1 : K);
#pragma unroll
            for (int j = 0; j < K; ++j) {
                if (j >= from) {
                    s[j] = j + 1 < K ? s[j + 1 < K ? j + 1 : j] : ns;
                    r[j] = j + 1 < K ? r[j + 1 < K ? j + 1 : j] : nr;
                    v[j] = j + 1 < K ? v[j + 1 < K ? j + 1 : j] : nv;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < K; ++j) if (lane == L && j == J) s[j] = newid;
        --n;
        // only the pairs (P-1, P) and (P, P+1) changed: lane L re-probes the second, lane H the first
        const int Lp = J > 0 ? L : L - 1, Jp = J > 0 ? J - 1 : K - 1, H = (L + 1) & 31;
        uint32_t sprev = 0, right = 0;
        {
            uint32_t cand = 0;
#pragma unroll
            for (int j = 0; j < K; ++j) if (j == Jp) cand = s[j];
            sprev = __shfl_sync(full, cand, Lp < 0 ? 0 : Lp);
            const uint32_t nx = __shfl_down_sync(full, s[0], 1);
#pragma unroll
            for (int j = 0; j < K; ++j) if (j == J) right = j + 1 < K ? s[j + 1 < K ? j + 1 : j] : nx;
        }
        uint2 res = make_uint2(kNone, 0u);
        if (lane == L && P + 1 < n) res = pair_lookup(t, newid, right);
        if (lane == H && P >= 1) res = pair_lookup(t, sprev, newid);
        const uint32_t rr = __shfl_sync(full, res.x, H), rv = __shfl_sync(full, res.y, H);
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (lane == L && j == J) { r[j] = res.x; v[j] = res.y; }
        }
        if (P >= 1) {
#pragma unroll
            for (int j = 0; j < K; ++j) if (lane == Lp && j == Jp) { r[j] = rr; v[j] = rv; }
        }
    }
    return n;
}

template <int K>
__device__ __forceinline__ uint32_t long_in_regs(const FusedParams& p, const uint8_t* src, int len, uint32_t* out, int lane, bool& ok) {
    const unsigned full = 0xFFFFFFFFu;
    uint32_t s[K];
    bool unknown = false;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int i = lane * K + j;
        s[j] = i < len ? __ldg(p.t.byte_init + __ldg(src + i)) : kNone;
        unknown = unknown || (i < len && s[j] == kNone);
    }
    if (__any_sync(full, unknown)) { ok = false; return 0; }     // a byte without a vocab entry: general path compacts
    ok = true;
    const int m = bpe_warp_regs<K>(p.t, s, len, lane);
#pragma unroll
    for (int j = 0; j < K; ++j) if (lane * K + j < m) out[lane * K + j] = s[j];
    return (uint32_t)m;
}

// ------------------------------------------------------------------------------------------------
// Round-parallel merging inside one warp (33 .. RP_MAX symbols; symbols, positions and ranks in shared memory).
// Same result as one merge per iteration (bpe.rs:104-153) when DevTables::round_parallel holds; the rule and why
// it is exact are in encode_xlong.cuh, with one refinement: a pair (x, y) only has to look as far as a token
// around it can reach -- to the left as far as a token ENDING with x extends, to the right as far as a token
// STARTING with y extends (DevTables::reach, in initial symbols; loader.cpp: token_reach).  For scripts without
// spaces this is what makes the rounds short: the bytes of different characters merge at the same time.
constexpr int RP_MAX = 256;
struct __align__(16) RoundBuf {
    uint32_t sym[2][RP_MAX];
    uint32_t rk[RP_MAX], nv[RP_MAX];
    uint16_t pos[2][RP_MAX + 2];
};

// in: buf.sym[0][0..n), n >= 2.  out: final ids in buf.sym[which][0..m); returns m, sets `which`.
__device__ __forceinline__ int bpe_warp_rounds(const DevTables& t, RoundBuf& buf, int n, int lane, int& which, uint32_t* dbg) {
    const unsigned full = 0xFFFFFFFFu;
    int cur = 0;
    uint32_t d_rounds = 0, d_chunks = 0, d_trips = 0;
    for (int i = lane; i <= n; i += 32) buf.pos[0][i] = (uint16_t)i;
    __syncwarp();
    for (;;) {
        const uint32_t* sym = buf.sym[cur];
        const uint16_t* pos = buf.pos[cur];
        // 1. rank and product of every pair
        for (int j = lane; j < n; j += 32) {
            uint2 q = make_uint2(kNone, 0u);
            if (j + 1 < n) q = pair_lookup(t, sym[j], sym[j + 1]);
            buf.rk[j] = q.x; buf.nv[j] = q.y;
        }
        __syncwarp();
        // 2. selection and compaction, 32 pairs at a time; run state is carried across the chunks
        uint32_t carry_rank = kNone, carry_sel = 0, carry_bad = 0;
        int carry_start = 0, out = 0;
        uint32_t* nsym = buf.sym[cur ^ 1];
        uint16_t* npos = buf.pos[cur ^ 1];
        ++d_rounds;
        for (int c0 = 0; c0 < n; c0 += 32) {
            ++d_chunks;
            const int j = c0 + lane;
            const uint32_t r = j < n ? buf.rk[j] : kNone;
            bool blocked = false;
            uint32_t me = 0;
            if (j < n) me = sym[j];
            if (r != kNone) {
                const uint32_t wl = __ldg(t.reach + me) & 0xFFFFu, wr = __ldg(t.reach + sym[j + 1]) >> 16;
                const uint32_t p0 = pos[j], e1 = pos[j + 2];
                // both directions in one loop (its trip count is the longer reach, not the sum); most pairs are blocked
                // by an immediate neighbour and leave in the first trip
                int kl = j - 1, kr = j + 1;
                bool goL = kl >= 0 && p0 - pos[kl] <= wl, goR = kr + 1 < n && (uint32_t)pos[kr + 2] - e1 <= wr;
                while (goL || goR) {
                    const uint32_t a = goL ? buf.rk[kl] : kNone, b = goR ? buf.rk[kr] : kNone;
                    if (min(a, b) < r) { blocked = true; break; }
                    --kl; ++kr; ++d_trips;
                    goL = goL && kl >= 0 && p0 - pos[kl] <= wl;
                    goR = goR && kr + 1 < n && (uint32_t)pos[kr + 2] - e1 <= wr;
                }
            }
            uint32_t prev_r = __shfl_up_sync(full, r, 1);
            if (lane == 0) prev_r = carry_rank;
            const bool head = r == kNone || prev_r != r;
            const unsigned H = __ballot_sync(full, head), Bk = __ballot_sync(full, blocked);
            const unsigned upto = (2u << lane) - 1u;                  // lanes 0 .. lane (wraps to all ones for lane 31)
            const unsigned below = H & upto;
            bool bad; int start;
            if (below) {
                const int s = 31 - __clz(below);
                bad = (Bk & upto & ~((1u << s) - 1u)) != 0;
                start = c0 + s;
            } else {
                bad = carry_bad || (Bk & upto) != 0;
                start = carry_start;
            }
            const bool sel = r != kNone && !bad && (((j - start) & 1) == 0);
            const unsigned Sel = __ballot_sync(full, sel);
            const bool eaten = lane == 0 ? carry_sel != 0 : ((Sel >> (lane - 1)) & 1u) != 0;
            const bool keep = j < n && !eaten;
            const unsigned K = __ballot_sync(full, keep);
            if (keep) {
                const int o = out + __popc(K & ((1u << lane) - 1u));
                nsym[o] = sel ? buf.nv[j] : me;
                npos[o] = pos[j];
            }
            out += __popc(K);
            carry_rank = __shfl_sync(full, r, 31);
            carry_start = __shfl_sync(full, start, 31);
            carry_bad = __shfl_sync(full, (uint32_t)bad, 31);
            carry_sel = Sel >> 31;
        }
        if (lane == 0) npos[out] = pos[n];
        __syncwarp();
        if (out == n) break;                                          // nothing was selected: final
        n = out;
        cur ^= 1;
    }
    if (dbg) {
        d_trips = __reduce_max_sync(full, d_trips);
        if (lane == 0) { atomicAdd(dbg + 26, 1u); atomicAdd(dbg + 27, d_rounds); atomicAdd(dbg + 28, d_chunks); atomicAdd(dbg + 29, d_trips); }
    }
    which = cur;
    return n;
}

// ids of src[0..len) (no added tokens inside) -> out; returns their number
__device__ __forceinline__ uint32_t long_piece(const FusedParams& p, const uint8_t* src, uint64_t len, uint32_t* out, int lane,
                                               RoundBuf* rb) {
    const unsigned full = 0xFFFFFFFFu;
    bool ok = false;
    uint32_t cnt = 0;
    if (rb && len <= RP_MAX) {                                    // rounds in shared memory
        int n = 0;
        for (uint64_t b0 = 0; b0 < len; b0 += 32) {               // initial ids; bytes without a vocab entry are dropped
            const uint64_t q = b0 + lane;
            const uint32_t sv = q < len ? __ldg(p.t.byte_init + __ldg(src + q)) : kNone;
            const unsigned hv = __ballot_sync(full, sv != kNone);
            if (sv != kNone) rb->sym[0][n + __popc(hv & ((1u << lane) - 1u))] = sv;
            n += __popc(hv);
        }
        __syncwarp();
        int which = 0;
        if (n >= 2) n = bpe_warp_rounds(p.t, *rb, n, lane, which, p.ablate == 9 ? p.err : nullptr);
        for (int i = lane; i < n; i += 32) out[i] = rb->sym[which][i];
        __syncwarp();
        return (uint32_t)n;
    }
    if (len <= 64) cnt = long_in_regs<2>(p, src, (int)len, out, lane, ok);
    else if (len <= 128) cnt = long_in_regs<4>(p, src, (int)len, out, lane, ok);
    else if (len <= 256) cnt = long_in_regs<8>(p, src, (int)len, out, lane, ok);
    if (!ok) {                                                 // very long, or a byte that is dropped
        uint32_t n = 0;
        for (uint64_t b0 = 0; b0 < len; b0 += 32) {
            uint64_t q = b0 + lane;
            uint32_t sv = q < len ? __ldg(p.t.byte_init + __ldg(src + q)) : kNone;
            unsigned hv = __ballot_sync(full, sv != kNone);
            if (sv != kNone) out[n + __popc(hv & ((1u << lane) - 1u))] = sv;
            n += __popc(hv);
        }
        __syncwarp();
        cnt = (uint32_t)bpe_warp_long(p.t, out, (int)n);
    }
    return cnt;
}

// ------------------------------------------------------------------------------------------------
// k_long_prep: one thread per long pre-token.  Resolves the length of those that run past their chunk (the next
// owned start of a later slice, or the end of the text), reserves room in the long pool (one atomic per warp) and
// sorts the pre-token into a work list by length class (<= 48, 64, 96, 128 bytes: batches of similar pre-tokens, and
// more warps per SM for the short classes) or the rest.
constexpr int MID_N0 = 48, MID_N1 = 64, MID_N2 = 96, MID_N3 = 128;   // one k_encode_mid instantiation and one work list per class
constexpr int MID_LISTS = 5;                         // the four classes + the rest (k_encode_long)
__global__ void __launch_bounds__(256) k_long_prep(const FusedParams p) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    uint32_t n_desc = *p.desc_cursor;
    if (n_desc > p.desc_cap) n_desc = p.desc_cap;
    for (uint32_t i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < n_desc; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + lane;
    const bool have = i < n_desc;
    uint64_t len = 0;
    if (have) {
        const LongDesc dd = p.desc[i];
        len = dd.len;
        if (dd.len == 0xFFFFFFFFu) {
            uint64_t e = p.n_bytes;
            for (uint64_t s = (uint64_t)dd.slice + 1; s < p.n_slices; ++s) {
                const uint32_t f = p.slice_first[s];
                if (f != 0xFFFFu) { e = s * SLICE - LCTX + f; break; }
            }
            len = e - dd.gstart;
            p.desc[i].len = (uint32_t)len;
        }
    }
    const bool xl = have && p.xl_enabled && len > XL_MIN;          // placed by k_xl_place, no pool space
    // pool space: exclusive scan of the lengths inside the warp, one atomic for all
    unsigned long long need = have && !xl ? len : 0, incl = need;
    for (int o = 1; o < 32; o <<= 1) { unsigned long long u = __shfl_up_sync(full, incl, o); if (lane >= o) incl += u; }
    unsigned long long base = 0;
    if (lane == 31 && incl) base = atomicAdd(p.long_cursor, incl);
    base = __shfl_sync(full, base, 31);
    const unsigned long long po = base + incl - need;
    bool ok = have && (xl || po + len <= p.long_cap);
    if (have && !ok) atomicOr(p.err, ERRF_POOL);
    if (have) { p.desc[i].pool = xl ? kNone : (uint32_t)po; p.desc[i].cnt = 0; }
    int which = !ok ? MID_LISTS : MID_LISTS - 1;
    if (ok && !xl && p.mid_enabled) {
        which = len <= (uint64_t)MID_N0 ? 0 : len <= (uint64_t)MID_N1 ? 1 : len <= (uint64_t)MID_N2 ? 2 : len <= (uint64_t)MID_N3 ? 3 : which;
    }
#pragma unroll
    for (int w = 0; w < MID_LISTS; ++w) {
        const unsigned m = __ballot_sync(full, which == w);
        if (!m) continue;
        uint32_t b0 = 0;
        if (lane == __ffs(m) - 1) b0 = atomicAdd(p.work_count + w, (uint32_t)__popc(m));
        b0 = __shfl_sync(full, b0, __ffs(m) - 1);
        if (which == w) p.work_list[(uint64_t)w * p.desc_cap + b0 + __popc(m & ((1u << lane) - 1u))] = i;
    }
    }
}

// ------------------------------------------------------------------------------------------------
// k_encode_mid<N>: pre-tokens of 33 .. N bytes, ONE LANE EACH (32 per warp at a time).  The reference's loop
// (bpe.rs:104-153) verbatim and exact for every table: scan the pair ranks for the lowest, leftmost on ties,
// merge that one pair, refresh the two pairs next to it.  Symbols and pair ranks live in shared memory as
// [slot][lane] (conflict-free); a merged-away slot is marked DEAD and skipped, so nothing is ever shifted.
// One lane per pre-token needs ~20x fewer warp instructions than one warp per pre-token for CJK-like runs.
// (Tried and measured slower: the neighbours of a slot from a per-lane live bitmask in registers instead of
// walking DEAD slots -- 9.0 -> 14.6 ms on config 3.)
constexpr uint32_t MID_DEAD = 0xFFFFFFFEu;
template <int N>
__global__ void __launch_bounds__(32) k_encode_mid(const FusedParams p, int list_no) {
    extern __shared__ uint32_t smem_mid[];
    uint32_t* const sym = smem_mid;                                // [N][32]
    uint32_t* const rnk = smem_mid + N * 32;                       // [N][32]; (rank << 8 | slot) of the pair that STARTS at the slot, kNone if none
    uint32_t* const bm = smem_mid + 2 * N * 32;                    // [N/8][32]: minimum key of each block of 8 slots
    uint32_t* const s_init = bm + (N / 8) * 32;                    // [256]
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x;
    for (int k = lane; k < 256; k += 32) s_init[k] = __ldg(p.t.byte_init + k);
    __syncwarp();
    const uint32_t count = min(p.work_count[list_no], p.desc_cap);
    const uint32_t* list = p.work_list + (uint64_t)list_no * p.desc_cap;
    for (uint32_t base = blockIdx.x * 32u; base < count; base += gridDim.x * 32u) {
        const bool have = base + lane < count;
        uint32_t di = 0;
        LongDesc dd{};
        if (have) { di = list[base + lane]; dd = p.desc[di]; }
        const uint8_t* src = p.text + dd.gstart;
        // initial ids; bytes without a vocab entry are dropped (bpe.rs:94-97)
        int n = 0;
        const int len = have ? (int)dd.len : 0;
        for (int b = 0; b < len; b += 4) {                          // four byte loads in flight
            uint32_t c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = b + q < len ? __ldg(src + b + q) : 0u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t sv = s_init[c[q]];
                if (b + q < len && sv != kNone) { sym[n * 32 + lane] = sv; ++n; }
            }
        }
        for (int i = 0; i < n; i += 2) {                            // two independent lookups in flight
            uint2 r1 = make_uint2(kNone, 0u), r2 = r1;
            if (i + 2 < n) pair_lookup2(p.t, sym[i * 32 + lane], sym[(i + 1) * 32 + lane], sym[(i + 1) * 32 + lane], sym[(i + 2) * 32 + lane], r1, r2);
            else if (i + 1 < n) r1 = pair_lookup(p.t, sym[i * 32 + lane], sym[(i + 1) * 32 + lane]);
            rnk[i * 32 + lane] = r1.x == kNone ? kNone : (r1.x << 8) | (uint32_t)i;
            if (i + 1 < n) rnk[(i + 1) * 32 + lane] = r2.x == kNone ? kNone : (r2.x << 8) | (uint32_t)(i + 1);
        }
        const int nb = (__reduce_max_sync(full, n) + 7) >> 3;        // blocks of 8 slots every lane looks at
        for (int i = n; i < nb * 8; ++i) rnk[i * 32 + lane] = kNone;
        auto block_min = [&](int blk) {
            uint32_t v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = rnk[(blk * 8 + q) * 32 + lane];
            bm[blk * 32 + lane] = min(min(min(v[0], v[1]), min(v[2], v[3])), min(min(v[4], v[5]), min(v[6], v[7])));
        };
        for (int blk = 0; blk < nb; ++blk) block_min(blk);
        bool busy = n > 1;
        while (__any_sync(full, busy)) {
            // lowest rank, leftmost on ties = the minimum of (rank << 8 | slot): two levels, so that a merge costs a scan
            // of the block minima plus the refresh of the (at most three) blocks it touched, not a scan of every slot
            uint32_t m0 = kNone, m1 = kNone, m2 = kNone, m3 = kNone;
            int i = 0;
            for (; i + 4 <= nb; i += 4) {
                m0 = min(m0, bm[i * 32 + lane]); m1 = min(m1, bm[(i + 1) * 32 + lane]);
                m2 = min(m2, bm[(i + 2) * 32 + lane]); m3 = min(m3, bm[(i + 3) * 32 + lane]);
            }
            for (; i < nb; ++i) m0 = min(m0, bm[i * 32 + lane]);
            const uint32_t best = min(min(m0, m1), min(m2, m3));
            const int bi = (int)(best & 0xFFu);
            if (best == kNone) busy = false;
            if (busy) {
                int j = bi + 1;
                while (sym[j * 32 + lane] == MID_DEAD) ++j;         // right symbol of the pair
                const uint32_t a = sym[bi * 32 + lane], b = sym[j * 32 + lane];
                const uint32_t nid = pair_lookup(p.t, a, b).y;
                sym[bi * 32 + lane] = nid;
                sym[j * 32 + lane] = MID_DEAD;
                rnk[j * 32 + lane] = kNone;
                int k = j + 1;
                while (k < n && sym[k * 32 + lane] == MID_DEAD) ++k;
                int h = bi - 1;
                while (h >= 0 && sym[h * 32 + lane] == MID_DEAD) --h;
                uint2 rr = make_uint2(kNone, 0u), rl = rr;                // the two pairs next to the new symbol, together
                if (k < n && h >= 0) pair_lookup2(p.t, nid, sym[k * 32 + lane], sym[h * 32 + lane], nid, rr, rl);
                else if (k < n) rr = pair_lookup(p.t, nid, sym[k * 32 + lane]);
                else if (h >= 0) rl = pair_lookup(p.t, sym[h * 32 + lane], nid);
                rnk[bi * 32 + lane] = rr.x == kNone ? kNone : (rr.x << 8) | (uint32_t)bi;
                if (h >= 0) rnk[h * 32 + lane] = rl.x == kNone ? kNone : (rl.x << 8) | (uint32_t)h;
                block_min(bi >> 3);
                if ((j >> 3) != (bi >> 3)) block_min(j >> 3);
                if (h >= 0 && (h >> 3) != (bi >> 3)) block_min(h >> 3);
            }
        }
        if (have) {
            uint32_t* out = p.long_pool + dd.pool;
            uint32_t cnt = 0;
            for (int i = 0; i < n; ++i) { const uint32_t v = sym[i * 32 + lane]; if (v != MID_DEAD) out[cnt++] = v; }
            p.desc[di].cnt = cnt;
            atomicAdd(p.slice_cnt + dd.slice, cnt);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_encode_long(const FusedParams p) {
    const int lane = threadIdx.x & 31;
    __shared__ RoundBuf s_rb[8];
    RoundBuf* const rb = (p.t.round_parallel && !p.no_rounds) ? &s_rb[threadIdx.x >> 5] : nullptr;
    const uint32_t count = min(p.work_count[MID_LISTS - 1], p.desc_cap);
    const uint32_t* list = p.work_list + (uint64_t)(MID_LISTS - 1) * p.desc_cap;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < count; w += n_warps) {
        const uint32_t i = list[w];
        const LongDesc dd = p.desc[i];
        const uint64_t len = dd.len;
        uint32_t cnt = 0;
        if (dd.pool == kNone) {                                    // merged in rounds by encode_xlong.cuh, after this kernel
            if (lane == 0) {
                const unsigned long long t = atomicAdd(p.xl_cursor, (1ull << XL_IDX_SHIFT) | (unsigned long long)(len + 1));
                const unsigned long long idx = t >> XL_IDX_SHIFT;
                if (idx < p.desc_cap) p.xl_list[idx] = XlEntry{i, 0u, t & ((1ull << XL_IDX_SHIFT) - 1)};
            }
        } else {
            uint32_t* out = p.long_pool + dd.pool;
            const uint8_t* src = p.text + dd.gstart;
            if (p.t.n_added == 0) cnt = long_piece(p, src, len, out, lane, rb);
            else {                                                 // mod.rs:566-610: added tokens inside the word
                uint64_t r = 0;
                while (r < len) {
                    uint32_t aid;
                    const uint64_t left = len - r;
                    const int pl = added_next_piece(p.t, src + r, left > 0x7FFFFFFFull ? 0x7FFFFFFF : (int)left, lane, &aid);
                    if (aid != kNone) { if (lane == 0) out[cnt] = aid; cnt += 1; }
                    else cnt += long_piece(p, src + r, (uint64_t)pl, out + cnt, lane, rb);
                    r += (uint64_t)pl;
                    __syncwarp();
                }
            }
        }
        if (lane == 0) {
            p.desc[i].cnt = cnt;
            atomicAdd(p.slice_cnt + dd.slice, cnt);
        }
    }
}

}  // namespace ctk
