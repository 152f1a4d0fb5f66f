// General (multi-kernel) encode pipeline.  Handles every input the hot path supports, including
// pre-tokens of any length; the fused single-pass kernel in encode_fused.cu is the fast path for
// pre-tokens up to 32 bytes and hands anything longer to the routines here.
//
//   k_docstart   scatter one bit per document start                      (mod.rs:694-696: docs are independent)
//   k_starts     pre-token boundaries, 1 bit per byte + per-block counts (pretokenizers.rs:13, :158-185)
//   k_list       compaction: bit map -> sorted list of pre-token starts
//   k_bpe        one warp per pre-token: byte -> initial id, merge loop   (bpe.rs:88-153)
//   k_emit       prefix-sum offsets -> packed ids + per-document offsets  (mod.rs:562-612 `result.extend`)
#include <cub/device/device_scan.cuh>

#include "device_common.cuh"
#include "engine.hpp"
#include "start_window.cuh"

namespace ctk {

__global__ void k_docstart(const uint64_t* __restrict__ off, uint64_t n_docs, uint64_t n_bytes,
                           uint32_t* __restrict__ ds, uint32_t* __restrict__ err) {
    uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    uint64_t p = off[d];
    if (d == 0 && p != 0) atomicOr(err, ERRF_OFFSETS);
    if (d == n_docs && p != n_bytes) atomicOr(err, ERRF_OFFSETS);
    if (d < n_docs && off[d + 1] < p) atomicOr(err, ERRF_OFFSETS);
    if (p < n_bytes) atomicOr(&ds[p >> 5], 1u << (p & 31));
}

// one thread per 32 byte positions -> one word of the start bitmap
__global__ void __launch_bounds__(256) k_starts(TextView tv, uint32_t* __restrict__ start_bits,
                                                uint32_t* __restrict__ block_counts) {
    uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t base = w * 32;
    uint32_t bits = 0;
    if (base < tv.n) {
        for (int k = 0; k < 32; ++k) {
            uint64_t i = base + k;
            if (i < tv.n && tv.is_start(i)) bits |= 1u << k;
        }
        start_bits[w] = bits;
    }
    int c = __popc(bits);
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    __shared__ int s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; ++k) t += s[k];
        block_counts[blockIdx.x] = (uint32_t)t;
    }
}

__global__ void __launch_bounds__(256) k_list(const uint32_t* __restrict__ start_bits, uint64_t n_words,
                                              const uint32_t* __restrict__ block_base, uint32_t* __restrict__ starts) {
    uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t bits = w < n_words ? start_bits[w] : 0u;
    int c = __popc(bits), lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
    __shared__ int s[8];
    if (lane == 31) s[wid] = incl;
    __syncthreads();
    int wbase = 0;
    for (int k = 0; k < wid; ++k) wbase += s[k];
    uint32_t o = block_base[blockIdx.x] + (uint32_t)(wbase + incl - c);
    while (bits) {
        int k = __ffs(bits) - 1;
        bits &= bits - 1;
        starts[o++] = (uint32_t)(w * 32 + k);
    }
}

// one warp per pre-token
__global__ void __launch_bounds__(256) k_bpe(DevTables t, const uint8_t* __restrict__ text, uint64_t n_bytes,
                                             const uint32_t* __restrict__ starts, uint32_t n_pre,
                                             uint32_t* __restrict__ tmp_ids, uint32_t* __restrict__ ntok) {
    const unsigned full = 0xFFFFFFFFu;
    uint64_t k = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (k >= n_pre) return;
    const int lane = threadIdx.x & 31;
    uint32_t s = starts[k];
    uint32_t e = (k + 1 < n_pre) ? starts[k + 1] : (uint32_t)n_bytes;
    int len = (int)(e - s);
    int m;
    if (len <= 32) {
        uint32_t sym = kNone;
        if (lane < len) sym = __ldg(t.byte_init + text[s + lane]);
        unsigned have = __ballot_sync(full, sym != kNone);       // bpe.rs:94-97: unknown chars are dropped
        int n = __popc(have);
        if (have != (len == 32 ? full : ((1u << len) - 1u))) {   // compact
            int dst = __popc(have & ((1u << lane) - 1u));
            uint32_t out = kNone;
            for (int src = 0; src < 32; ++src) {
                uint32_t v = __shfl_sync(full, sym, src);
                int d = __shfl_sync(full, dst, src);
                if (((have >> src) & 1u) && d == lane) out = v;
            }
            sym = out;
        }
        m = n ? bpe_warp32(t, sym, n) : 0;
        if (lane < m) tmp_ids[s + lane] = sym;
    } else {
        uint32_t* sym = tmp_ids + s;
        int n = 0;
        for (int base = 0; base < len; base += 32) {
            int i = base + lane;
            uint32_t v = i < len ? __ldg(t.byte_init + text[s + i]) : kNone;
            unsigned have = __ballot_sync(full, v != kNone);
            if (v != kNone) sym[n + __popc(have & ((1u << lane) - 1u))] = v;
            n += __popc(have);
        }
        __syncwarp();
        m = bpe_warp_long(t, sym, n);
    }
    if (lane == 0) ntok[k] = (uint32_t)m;
}

__global__ void __launch_bounds__(256) k_emit(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ tok_off,
                                              uint32_t n_pre, const uint32_t* __restrict__ tmp_ids,
                                              uint32_t* __restrict__ out, uint64_t out_cap, uint32_t* __restrict__ err) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k >= n_pre) return;
    uint32_t o = tok_off[k], c = tok_off[k + 1] - o, s = starts[k];
    if ((uint64_t)o + c > out_cap) { atomicOr(err, ERRF_CAPACITY); return; }
    for (uint32_t i = 0; i < c; ++i) out[o + i] = tmp_ids[s + i];
}

// ids_off[d] = number of ids produced by pre-tokens that start before text_off[d]
__global__ void k_doc_offsets(const uint64_t* __restrict__ text_off, uint64_t n_docs, const uint32_t* __restrict__ starts,
                              uint32_t n_pre, const uint32_t* __restrict__ tok_off, uint64_t* __restrict__ ids_off) {
    uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    uint64_t p = text_off[d];
    uint32_t lo = 0, hi = n_pre;                 // lower_bound(starts, p)
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (starts[mid] < p) lo = mid + 1; else hi = mid;
    }
    ids_off[d] = tok_off[lo];
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return eng.cuda_fail(e_, #x); } while (0)

// The same bitmap through the bit-parallel window logic of the fused kernel (start_window.cuh): one thread per 16-byte group,
// one CTA per 8 KiB (512 groups + one halo group on each side, masks exchanged through shared memory).  Output layout
// identical to k_starts: one word per 32 bytes + the number of starts per 8 KiB block.
constexpr int kFastGroups = 512;
__global__ void __launch_bounds__(kFastGroups + 32) k_starts_window(const uint8_t* __restrict__ text, uint64_t n, const uint32_t* __restrict__ ds,
                                                                    const uint8_t* __restrict__ trie_index, const uint8_t* __restrict__ trie_blocks,
                                                                    uint32_t* __restrict__ start_bits, uint32_t* __restrict__ block_counts) {
    __shared__ Masks16 sm[kFastGroups + 2];
    __shared__ uint32_t sds[kFastGroups + 2];
    __shared__ int scount;
    const int t = threadIdx.x;
    const int64_t g = (int64_t)blockIdx.x * kFastGroups + t - 1;               // t = 0 and t = 513 are the halo groups
    const int64_t n_groups = (int64_t)((n + 15) / 16);
    if (t == 0) scount = 0;
    if (t < kFastGroups + 2) {
        Masks16 m{};
        uint32_t d16 = 0;
        if (g >= 0 && g <= n_groups) {                                          // group n_groups: only zero padding, but position n is a document start
            const uint4 v = *reinterpret_cast<const uint4*>(text + g * 16);
            m = classify16(text + g * 16, 0, v.x, v.y, v.z, v.w, trie_index, trie_blocks);
            d16 = (ds[g >> 1] >> (16 * (g & 1))) & 0xFFFFu;
            const int64_t past = (int64_t)n - g * 16;                           // positions >= n count as document starts
            if (past < 16) d16 |= past <= 0 ? 0xFFFFu : (0xFFFFu << past) & 0xFFFFu;
        }
        sm[t] = m;
        sds[t] = d16;
    }
    __syncthreads();
    uint32_t s16 = 0;
    if (t >= 1 && t <= kFastGroups && g < n_groups) {
        const Masks16 &a = sm[t - 1], &b = sm[t], &c = sm[t + 1];
        const uint32_t S = start_window(window(a.L, b.L, c.L), window(a.N, b.N, c.N), window(a.W, b.W, c.W), window(a.SP, b.SP, c.SP),
                                        window(a.AP, b.AP, c.AP), window(a.CONT, b.CONT, c.CONT), window(sds[t - 1], sds[t], sds[t + 1]),
                                        text + g * 16 - 8);
        s16 = (S >> 8) & 0xFFFFu;
        const int64_t past = (int64_t)n - g * 16;
        if (past < 16) s16 &= (1u << past) - 1u;
    }
    // groups (t = 1, 2), (3, 4), ... share an output word: the odd t holds the low half
    __shared__ uint32_t sout[kFastGroups + 2];
    if (t < kFastGroups + 2) sout[t] = s16;
    __syncthreads();
    if (t >= 1 && t <= kFastGroups && (t & 1)) {
        const uint64_t w = ((uint64_t)blockIdx.x * kFastGroups + (t - 1)) >> 1;
        if (w * 32 < n) {
            const uint32_t bits = sout[t] | (sout[t + 1] << 16);
            start_bits[w] = bits;
            atomicAdd(&scount, __popc(bits));
        }
    }
    __syncthreads();
    if (t == 0) block_counts[blockIdx.x] = (uint32_t)scount;
}

// Pre-token start bitmap of a batch (one bit per byte) + the number of starts per block of 256 words (8 KiB);
// used by the rich `Encoding` path (encoding.cu), which needs the words themselves and not only their ids.
// ds: (n_words + 1) words of scratch for the document-start bits.
int starts_bitmap(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes, uint32_t* ds,
                  uint32_t* start_bits, uint32_t* block_counts, uint32_t* err, cudaStream_t st) {
    const uint64_t n_words = (n_bytes + 31) / 32;
    const uint32_t n_blocks = (uint32_t)((n_words + 255) / 256);
    CK(cudaMemsetAsync(ds, 0, (n_words + 1) * 4, st));
    k_docstart<<<(unsigned)((n_docs + 1 + 255) / 256), 256, 0, st>>>(d_off, n_docs, n_bytes, ds, err);
    if (getenv("CTK_SCALAR_STARTS") || (reinterpret_cast<uintptr_t>(d_text) & 15)) {   // debug / unaligned text: the scalar predicate
        TextView tv{d_text, n_bytes, ds, eng.tables.trie_index, eng.tables.trie_blocks};
        k_starts<<<n_blocks, 256, 0, st>>>(tv, start_bits, block_counts);
    } else {
        k_starts_window<<<n_blocks, kFastGroups + 32, 0, st>>>(d_text, n_bytes, ds, eng.tables.trie_index, eng.tables.trie_blocks,
                                                               start_bits, block_counts);
    }
    eng.launched(2);
    CK(cudaGetLastError());
    return CTK_OK;
}

int encode_general(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes,
                   uint32_t* d_ids, uint64_t ids_cap, uint64_t* d_ids_off, uint64_t* n_ids_host, cudaStream_t st) {
    if (n_bytes >= 0xFFFFFFF0ull) return eng.fail(CTK_ERR_ARG, "one device call handles less than 4 GiB of text");
    if (eng.tables.n_added) return eng.fail(CTK_ERR_UNSUPPORTED, "the debug general pipeline does not match added tokens inside words");
    if (n_bytes == 0) {
        CK(cudaMemsetAsync(d_ids_off, 0, (n_docs + 1) * 8, st));
        if (n_ids_host) { CK(cudaStreamSynchronize(st)); *n_ids_host = 0; }
        return CTK_OK;
    }
    uint64_t n_words = (n_bytes + 31) / 32;
    uint32_t n_blocks = (uint32_t)((n_words + 255) / 256);
    Workspace& ws = eng.ws;
    uint32_t *ds, *sb, *bc, *bb, *err;
    CK(ws.get(0, (n_words + 1) * 4, (void**)&ds));
    CK(ws.get(1, (n_words + 1) * 4, (void**)&sb));
    CK(ws.get(2, ((uint64_t)n_blocks + 1) * 4, (void**)&bc));
    CK(ws.get(3, ((uint64_t)n_blocks + 1) * 4, (void**)&bb));
    CK(ws.get(4, 256, (void**)&err));
    CK(cudaMemsetAsync(ds, 0, (n_words + 1) * 4, st));
    CK(cudaMemsetAsync(err, 0, 256, st));
    eng.mark(nullptr, st);
    k_docstart<<<(unsigned)((n_docs + 1 + 255) / 256), 256, 0, st>>>(d_off, n_docs, n_bytes, ds, err);
    eng.launched(1); eng.mark("k_docstart", st);
    TextView tv{d_text, n_bytes, ds, eng.tables.trie_index, eng.tables.trie_blocks};
    k_starts<<<n_blocks, 256, 0, st>>>(tv, sb, bc);
    eng.launched(1); eng.mark("k_starts", st);
    size_t cub_bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, bc, bb, n_blocks + 1, st));
    void* cub_tmp;
    CK(ws.get(5, cub_bytes + 16, &cub_tmp));
    CK(cudaMemsetAsync(bc + n_blocks, 0, 4, st));
    CK(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, bc, bb, n_blocks + 1, st));
    eng.launched(1); eng.mark("scan_blocks", st);
    uint32_t n_pre = 0;
    CK(cudaMemcpyAsync(&n_pre, bb + n_blocks, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    uint32_t *starts, *tmp_ids, *ntok, *tok_off;
    CK(ws.get(6, ((uint64_t)n_pre + 2) * 4, (void**)&starts));
    CK(ws.get(7, (n_bytes + 64) * 4, (void**)&tmp_ids));
    CK(ws.get(8, ((uint64_t)n_pre + 2) * 4, (void**)&ntok));
    CK(ws.get(9, ((uint64_t)n_pre + 2) * 4, (void**)&tok_off));
    eng.mark(nullptr, st);
    k_list<<<n_blocks, 256, 0, st>>>(sb, n_words, bb, starts);
    eng.launched(1); eng.mark("k_list", st);
    k_bpe<<<(unsigned)(((uint64_t)n_pre * 32 + 255) / 256), 256, 0, st>>>(eng.tables, d_text, n_bytes, starts, n_pre, tmp_ids, ntok);
    eng.launched(1); eng.mark("k_bpe", st);
    CK(cudaMemsetAsync(ntok + n_pre, 0, 4, st));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, ntok, tok_off, n_pre + 1, st));
    CK(ws.get(5, cub_bytes + 16, &cub_tmp));
    CK(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, ntok, tok_off, n_pre + 1, st));
    eng.launched(1); eng.mark("scan_ntok", st);
    k_emit<<<(n_pre + 255) / 256, 256, 0, st>>>(starts, tok_off, n_pre, tmp_ids, d_ids, ids_cap, err);
    k_doc_offsets<<<(unsigned)((n_docs + 1 + 255) / 256), 256, 0, st>>>(d_off, n_docs, starts, n_pre, tok_off, d_ids_off);
    eng.launched(2); eng.mark("k_emit+doc_offsets", st);
    CK(cudaGetLastError());
    return eng.finish(err, d_ids_off, n_docs, n_ids_host, st);
}

}  // namespace ctk
