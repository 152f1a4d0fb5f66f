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

// ids of src[0..len) (no added tokens inside) -> out; returns their number
__device__ __forceinline__ uint32_t long_piece(const FusedParams& p, const uint8_t* src, uint64_t len, uint32_t* out, int lane) {
    const unsigned full = 0xFFFFFFFFu;
    bool ok = false;
    uint32_t cnt = 0;
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
            if (p.t.n_added == 0) cnt = long_piece(p, src, len, out, lane);
            else {                                                 // mod.rs:566-610: added tokens inside the word
                uint64_t r = 0;
                while (r < len) {
                    uint32_t aid;
                    const uint64_t left = len - r;
                    const int pl = added_next_piece(p.t, src + r, left > 0x7FFFFFFFull ? 0x7FFFFFFF : (int)left, lane, &aid);
                    if (aid != kNone) { if (lane == 0) out[cnt] = aid; cnt += 1; }
                    else cnt += long_piece(p, src + r, (uint64_t)pl, out + cnt, lane);
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
