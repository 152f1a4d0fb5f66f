// k_encode_long: BPE of the pre-tokens that are longer than 32 bytes (CJK runs, URLs, long digit or
// punctuation runs), one warp per pre-token.  Included by encode_fused.cu after FusedParams.
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

__global__ void __launch_bounds__(256) k_encode_long(const FusedParams p) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    __shared__ RoundBuf s_rb[8];
    RoundBuf* const rb = (p.t.round_parallel && !p.no_rounds) ? &s_rb[threadIdx.x >> 5] : nullptr;
    uint32_t n_desc = *p.desc_cursor;
    if (n_desc > p.desc_cap) n_desc = p.desc_cap;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_desc; i += n_warps) {
        LongDesc dd = p.desc[i];
        uint64_t len = dd.len;
        if (dd.len == 0xFFFFFFFFu) {                               // runs past its chunk: it ends at the next start, which is
            uint64_t e = p.n_bytes;                                // the first owned start of a later slice (or the text's end)
            for (uint64_t s0 = (uint64_t)dd.slice + 1; s0 < p.n_slices; s0 += 32) {
                const uint64_t s = s0 + lane;
                const uint32_t f = s < p.n_slices ? p.slice_first[s] : 0xFFFFu;
                const unsigned b = __ballot_sync(full, f != 0xFFFFu);
                if (b) {
                    const int src = __ffs(b) - 1;
                    e = (s0 + src) * SLICE - LCTX + __shfl_sync(full, f, src);
                    break;
                }
            }
            len = e - dd.gstart;
            if (lane == 0) p.desc[i].len = (uint32_t)len;
        }
        unsigned long long po = 0;
        if (!(p.xl_enabled && len > XL_MIN)) {
            if (lane == 0) po = atomicAdd(p.long_cursor, (unsigned long long)len);
            po = __shfl_sync(full, po, 0);
        }
        uint32_t cnt = 0;
        if (p.xl_enabled && len > XL_MIN) {                        // merged in rounds by encode_xlong.cuh, after this kernel
            po = kNone;                                            // no room in the long pool: k_xl_place writes the output directly
            if (lane == 0) {
                const unsigned long long t = atomicAdd(p.xl_cursor, (1ull << XL_IDX_SHIFT) | (unsigned long long)(len + 1));
                const unsigned long long idx = t >> XL_IDX_SHIFT;
                if (idx < p.desc_cap) p.xl_list[idx] = XlEntry{i, 0u, t & ((1ull << XL_IDX_SHIFT) - 1)};
            }
        } else if (po + len <= p.long_cap) {
            uint32_t* out = p.long_pool + po;
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
        } else if (lane == 0) atomicOr(p.err, ERRF_POOL);
        if (lane == 0) {
            p.desc[i].pool = (uint32_t)po;
            p.desc[i].cnt = cnt;
            atomicAdd(p.slice_cnt + dd.slice, cnt);
        }
    }
}

}  // namespace ctk
