// Metaspace pipelines on the device (second half).
//
// Reference: metaspace_pretokenize (src/pretokenizers.rs:188-200): text' = [replacement +] text with every U+0020 replaced by
// the replacement character; words = text' split on white space (other than the replacement), empty ones dropped.  Then, per
// word (src/huggingface/mod.rs:562-612): the in-word added-token scan, and BpeTokenizer::encode (src/bpe.rs:88-153) whose
// symbols are the word's CHARACTERS (unknown characters are dropped, :94-97) -- not bytes as in the ByteLevel pipeline.
//
//   k_meta_len / k_meta_copy   warp per text: length and bytes of text' (a text grows by the prefix and by
//                              (len(replacement) - 1) per space)
//   k_meta_bits                thread per 32 bytes of text': bitmaps "a word starts here" and "a word cannot continue here"
//                              (a white-space character or a text start)
//   k_meta_list                bitmap -> sorted list of word starts
//   k_meta_bpe                 warp per word: characters -> initial ids through a code-point hash table, compaction, added
//                              tokens inside the word (added_tokens.cuh, on raw bytes), bpe_warp32 / bpe_warp_long
//   k_meta_emit / k_meta_doc   prefix-sum offsets -> packed ids (2 or 4 bytes) + per-text id offsets
//
// This is the general (multi-kernel) shape of encode_general.cu, not the fused single-pass kernel: Metaspace is a widening row;
// the fused kernel's pre-token cache and slice geometry are tied to byte symbols.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "added_tokens.cuh"
#include "device_common.cuh"
#include "engine.hpp"

namespace ctk {
namespace {

__device__ __forceinline__ bool ws_cp(uint32_t c) {             // char::is_whitespace = White_Space
    return (c >= 9 && c <= 13) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x2028 ||
           c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}
// code point of the character whose lead byte is at i (valid UTF-8 assumed; truncated at hi)
__device__ __forceinline__ uint32_t cp_at(const uint8_t* t, uint64_t i, uint64_t hi, int& len) {
    const uint32_t c = t[i];
    if (c < 0xC0u) { len = 1; return c; }
    const int want = c < 0xE0u ? 2 : (c < 0xF0u ? 3 : 4);
    if (i + want > hi) { len = 1; return c; }
    len = want;
    if (want == 2) return ((c & 0x1Fu) << 6) | (t[i + 1] & 63u);
    if (want == 3) return ((c & 0x0Fu) << 12) | ((t[i + 1] & 63u) << 6) | (t[i + 2] & 63u);
    return ((c & 7u) << 18) | ((t[i + 1] & 63u) << 12) | ((t[i + 2] & 63u) << 6) | (t[i + 3] & 63u);
}
__device__ __forceinline__ int utf8_put(uint8_t* o, uint32_t cp) {
    if (cp < 0x80) { o[0] = (uint8_t)cp; return 1; }
    if (cp < 0x800) { o[0] = (uint8_t)(0xC0 | (cp >> 6)); o[1] = (uint8_t)(0x80 | (cp & 63)); return 2; }
    if (cp < 0x10000) { o[0] = (uint8_t)(0xE0 | (cp >> 12)); o[1] = (uint8_t)(0x80 | ((cp >> 6) & 63)); o[2] = (uint8_t)(0x80 | (cp & 63)); return 3; }
    o[0] = (uint8_t)(0xF0 | (cp >> 18)); o[1] = (uint8_t)(0x80 | ((cp >> 12) & 63)); o[2] = (uint8_t)(0x80 | ((cp >> 6) & 63)); o[3] = (uint8_t)(0x80 | (cp & 63));
    return 4;
}
__host__ __device__ inline int utf8_len(uint32_t cp) { return cp < 0x80 ? 1 : cp < 0x800 ? 2 : cp < 0x10000 ? 3 : 4; }

// warp per text: new_len[d] = prefix + len + (rep_len - 1) * spaces
__global__ void __launch_bounds__(256) k_meta_len(const uint8_t* __restrict__ text, const uint64_t* __restrict__ off, uint64_t n, int rep_len, int prefix,
                                                  uint64_t n_bytes, uint64_t* __restrict__ new_len, uint32_t* __restrict__ err) {
    const uint64_t d = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (d > n) return;
    if (d == n) { if (lane == 0) { new_len[d] = 0; if (off[d] != n_bytes) atomicOr(err, ERRF_OFFSETS); } return; }
    const uint64_t lo = off[d], hi = off[d + 1];
    if ((d == 0 && lo != 0) || hi < lo || hi > n_bytes) { if (lane == 0) { atomicOr(err, ERRF_OFFSETS); new_len[d] = 0; } return; }
    uint64_t spaces = 0;
    for (uint64_t i = lo + lane; i < hi; i += 32) spaces += text[i] == 0x20;
    for (int o = 16; o; o >>= 1) spaces += __shfl_xor_sync(0xFFFFFFFFu, spaces, o);
    if (lane == 0) new_len[d] = (hi - lo) + (uint64_t)(rep_len - 1) * spaces + (prefix ? rep_len : 0);
}
__global__ void __launch_bounds__(256) k_meta_copy(const uint8_t* __restrict__ text, const uint64_t* __restrict__ off, uint64_t n, uint32_t rep, int rep_len,
                                                   int prefix, const uint64_t* __restrict__ new_off, uint8_t* __restrict__ out) {
    const uint64_t d = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (d >= n) return;
    const uint64_t lo = off[d], hi = off[d + 1];
    uint64_t o = new_off[d];
    if (prefix) { if (lane == 0) utf8_put(out + o, rep); o += rep_len; }
    for (uint64_t base = lo; base < hi; base += 32) {
        const uint64_t i = base + lane;
        const bool sp = i < hi && text[i] == 0x20;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, sp);
        if (i < hi) {
            const uint64_t p = o + lane + (uint64_t)(rep_len - 1) * __popc(m & ((1u << lane) - 1u));
            if (sp) utf8_put(out + p, rep); else out[p] = text[i];
        }
        o += min((uint64_t)32, hi - base) + (uint64_t)(rep_len - 1) * __popc(m);
    }
}

// text starts as bits (a text start ends the previous word)
__global__ void k_meta_docbits(const uint64_t* __restrict__ off, uint64_t n, uint64_t n_bytes, uint32_t* __restrict__ ds) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n) return;
    const uint64_t p = off[d];
    if (p < n_bytes) atomicOr(&ds[p >> 5], 1u << (p & 31));
}
// thread per 32 bytes: S = a word starts here; X = a word cannot continue here (white space, or a text start)
__global__ void __launch_bounds__(256) k_meta_bits(const uint8_t* __restrict__ text, uint64_t n_bytes, const uint32_t* __restrict__ ds, uint64_t n_words,
                                                   uint32_t* __restrict__ S, uint32_t* __restrict__ X, uint32_t* __restrict__ cnt) {
    const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w > n_words) return;
    if (w == n_words) { cnt[w] = 0; return; }
    const uint32_t dsw = ds[w];
    uint32_t s = 0, x = 0;
    // white space of the byte before the group's first byte (whole characters: look back to its lead)
    auto ws_byte = [&](uint64_t i) -> bool {                  // byte i belongs to a white-space character
        uint64_t l = i;
        for (int k = 0; k < 3 && l > 0 && (text[l] & 0xC0u) == 0x80u; ++k) --l;
        int len;
        return ws_cp(cp_at(text, l, n_bytes, len));
    };
    bool prev_ws = w == 0 ? true : ws_byte(w * 32 - 1);
    for (int k = 0; k < 32; ++k) {
        const uint64_t i = w * 32 + k;
        if (i >= n_bytes) break;
        const bool docstart = (dsw >> k) & 1u;
        const bool lead = (text[i] & 0xC0u) != 0x80u;
        const bool ws = lead ? ws_byte(i) : prev_ws;          // continuation bytes inherit their character's class
        if (ws || docstart) x |= 1u << k;
        if (!ws && lead && (prev_ws || docstart)) s |= 1u << k;
        prev_ws = ws;
    }
    S[w] = s; X[w] = x;
    cnt[w] = __popc(s);
}
__global__ void k_meta_list(const uint32_t* __restrict__ S, const uint64_t* __restrict__ rank, uint64_t n_words, uint32_t* __restrict__ starts) {
    const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t b = S[w];
    uint64_t k = rank[w];
    while (b) { const int j = __ffs(b) - 1; b &= b - 1; starts[k++] = (uint32_t)(w * 32 + j); }
}

__device__ __forceinline__ uint32_t char_id(const uint2* __restrict__ tab, uint32_t mask, uint32_t cp) {
    uint32_t h = (cp * 0x9E3779B1u) >> 7;
    for (;;) {
        const uint2 e = __ldg(tab + (h & mask));
        if (e.x == cp) return e.y;
        if (e.x == 0xFFFFFFFFu) return kNone;
        ++h;
    }
}

// initial ids of the characters of bytes [lo, hi) -> sym[0..) (global scratch at the word's own byte offset: ids <= bytes);
// unknown characters are dropped (bpe.rs:94-97).  Returns the count (same on every lane).
__device__ __forceinline__ int meta_symbols(const uint8_t* __restrict__ text, uint64_t lo, uint64_t hi, const uint2* tab, uint32_t mask, uint32_t* sym, int lane) {
    int n = 0;
    for (uint64_t base = lo; base < hi; base += 32) {
        const uint64_t i = base + lane;
        uint32_t id = kNone;
        if (i < hi && (text[i] & 0xC0u) != 0x80u) { int len; id = char_id(tab, mask, cp_at(text, i, hi, len)); }
        const unsigned have = __ballot_sync(0xFFFFFFFFu, id != kNone);
        if (id != kNone) sym[n + __popc(have & ((1u << lane) - 1u))] = id;
        n += __popc(have);
    }
    __syncwarp();
    return n;
}

// warp per word
__global__ void __launch_bounds__(256) k_meta_bpe(DevTables t, const uint2* __restrict__ char_tab, uint32_t char_mask, const uint8_t* __restrict__ text,
                                                  uint64_t n_bytes, const uint32_t* __restrict__ X, const uint32_t* __restrict__ starts, uint32_t n_pre,
                                                  uint32_t* __restrict__ tmp_ids, uint32_t* __restrict__ ntok) {
    const unsigned full = 0xFFFFFFFFu;
    const uint64_t k = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (k >= n_pre) return;
    const int lane = threadIdx.x & 31;
    const uint64_t s = starts[k];
    // the word ends at the first stop bit after s (or at the end of the buffer)
    uint64_t e = n_bytes;
    {
        uint64_t w = (s + 1) >> 5;
        uint32_t bits = (s + 1) < n_bytes ? X[w] & (0xFFFFFFFFu << ((s + 1) & 31)) : 0u;
        const uint64_t n_words = (n_bytes + 31) / 32;
        while ((s + 1) < n_bytes) {
            if (bits) { e = w * 32 + (__ffs(bits) - 1); break; }
            if (++w >= n_words) break;
            bits = X[w];
        }
        if (e > n_bytes) e = n_bytes;
    }
    uint32_t* out = tmp_ids + s;                              // the word's ids: at most one per byte
    int cnt = 0;
    uint64_t r = s;
    while (r < e) {
        uint32_t aid = kNone;
        uint64_t pl = e - r;
        if (t.n_added) {
            const int rem_n = (int)min((uint64_t)0x7FFFFFFF, e - r);
            pl = (uint64_t)added_next_piece(t, text + r, rem_n, lane, &aid);      // mod.rs:566-610 on the word's own bytes
        }
        if (aid != kNone) { if (lane == 0) out[cnt] = aid; cnt += 1; }
        else {
            uint32_t* sym = out + cnt;
            const int n = meta_symbols(text, r, r + pl, char_tab, char_mask, sym, lane);
            int m = 0;
            if (n > 0 && n <= 32) {
                uint32_t v = lane < n ? sym[lane] : kNone;
                m = bpe_warp32(t, v, n);
                __syncwarp();
                if (lane < m) sym[lane] = v;
            } else if (n > 32) m = bpe_warp_long(t, sym, n);
            cnt += m;
        }
        __syncwarp();
        r += pl;
    }
    if (lane == 0) ntok[k] = (uint32_t)cnt;
}

template <class OutT>
__global__ void __launch_bounds__(256) k_meta_emit(const uint32_t* __restrict__ starts, const uint64_t* __restrict__ tok_off, uint32_t n_pre,
                                                   const uint32_t* __restrict__ tmp_ids, OutT* __restrict__ out, uint64_t out_cap, uint32_t* __restrict__ err) {
    const uint64_t k = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n_pre) return;
    const uint64_t o = tok_off[k], c = tok_off[k + 1] - o, s = starts[k];
    if (o + c > out_cap) { if (lane == 0) atomicOr(err, ERRF_CAPACITY); return; }
    for (uint64_t i = lane; i < c; i += 32) out[o + i] = (OutT)tmp_ids[s + i];
}
// ids_off[d] = ids of the words that start before the text's first byte
__global__ void k_meta_doc(const uint64_t* __restrict__ text_off, uint64_t n, const uint32_t* __restrict__ starts, uint32_t n_pre,
                           const uint64_t* __restrict__ tok_off, uint64_t* __restrict__ ids_off) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n) return;
    const uint64_t p = text_off[d];
    uint32_t lo = 0, hi = n_pre;
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (starts[mid] < p) lo = mid + 1; else hi = mid; }
    ids_off[d] = tok_off[lo];
}

struct U32to64 { __host__ __device__ uint64_t operator()(uint32_t v) const { return v; } };

// ---- Split stages in front of a Metaspace stage: an EMPTY text stays one (empty) piece in the reference (pretokenizers.rs:305-307)
// and the Metaspace stage turns it into the word "<replacement>"; the Split kernels know pieces by their bytes, so those ids
// are put in afterwards.  extra[d] = number of ids to add for text d (k if it is empty, else 0).
__global__ void k_meta_empty(const uint64_t* __restrict__ off, uint64_t n, uint32_t k, uint64_t* __restrict__ extra, uint32_t* __restrict__ any) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n) return;
    const bool empty = d < n && off[d + 1] == off[d];
    extra[d] = empty ? k : 0;
    if (empty) *any = 1u;
}
template <class T>
__global__ void __launch_bounds__(256) k_meta_insert(const T* __restrict__ src, const uint64_t* __restrict__ old_off, const uint64_t* __restrict__ add, uint64_t n,
                                                     const uint64_t* __restrict__ text_off, uint32_t id0, T* __restrict__ dst, uint64_t cap, uint64_t* __restrict__ new_off,
                                                     uint32_t* __restrict__ err) {
    const uint64_t d = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (d > n) return;
    const uint64_t o = old_off[d] + add[d];
    if (lane == 0) new_off[d] = o;
    if (d == n) return;
    const uint64_t len = old_off[d + 1] - old_off[d];
    const bool empty = text_off[d + 1] == text_off[d];
    if (o + len + (empty ? 1 : 0) > cap) { if (lane == 0) atomicOr(err, ERRF_CAPACITY); return; }
    if (empty && lane == 0) dst[o] = (T)id0;
    for (uint64_t i = lane; i < len; i += 32) dst[o + i] = src[old_off[d] + i];
}

}  // namespace

int metaspace_empty_texts(Engine& eng, const uint64_t* d_text_off, size_t n, uint32_t* d_ids_u32, uint64_t ids_cap, uint64_t* d_ids_off, uint64_t* n_ids_host,
                          cudaStream_t st) {
    const std::vector<uint32_t>& ids = eng.model.meta_empty_ids;
    if (ids.empty() || n == 0) return CTK_OK;
    Workspace& ws = eng.ws;
    uint64_t *extra, *add, *new_off; uint32_t* flag; void* cub_tmp; void* copy;
    const int width = eng.out_id_width == 2 && eng.run_width == 2 ? 2 : 4;
    if (cudaError_t e = ws.get(83, (n + 2) * 8, (void**)&extra)) return eng.cuda_fail(e, "workspace");
    if (cudaError_t e = ws.get(84, (n + 2) * 8, (void**)&add)) return eng.cuda_fail(e, "workspace");
    if (cudaError_t e = ws.get(85, (n + 2) * 8, (void**)&new_off)) return eng.cuda_fail(e, "workspace");
    if (cudaError_t e = ws.get(4, 256, (void**)&flag)) return eng.cuda_fail(e, "workspace");
    if (cudaError_t e = cudaMemsetAsync(flag, 0, 256, st)) return eng.cuda_fail(e, "memset");
    k_meta_empty<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(d_text_off, n, (uint32_t)ids.size(), extra, flag + 8);
    if (cudaError_t e = eng.publish({{flag + 8, 1, 8}, {d_ids_off + n, 2, 2}}, st)) return eng.cuda_fail(e, "publish");
    if (cudaError_t e = cudaStreamSynchronize(st)) return eng.cuda_fail(e, "sync");
    if (!eng.h_flags[8]) return CTK_OK;
    uint64_t total;
    memcpy(&total, eng.h_flags + 2, 8);
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, extra, add, n + 1, st);
    if (cudaError_t e = ws.get(5, cub_bytes + 16, &cub_tmp)) return eng.cuda_fail(e, "workspace");
    if (cudaError_t e = cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, extra, add, n + 1, st)) return eng.cuda_fail(e, "scan");
    if (cudaError_t e = ws.get(86, total * width + 64, &copy)) return eng.cuda_fail(e, "workspace");
    if (cudaError_t e = cudaMemcpyAsync(copy, d_ids_u32, total * width, cudaMemcpyDeviceToDevice, st)) return eng.cuda_fail(e, "copy");
    const unsigned g = (unsigned)(((n + 1) * 32 + 255) / 256);
    if (width == 2) k_meta_insert<uint16_t><<<g, 256, 0, st>>>(static_cast<const uint16_t*>(copy), d_ids_off, add, n, d_text_off, ids[0], reinterpret_cast<uint16_t*>(d_ids_u32), ids_cap, new_off, flag);
    else k_meta_insert<uint32_t><<<g, 256, 0, st>>>(static_cast<const uint32_t*>(copy), d_ids_off, add, n, d_text_off, ids[0], d_ids_u32, ids_cap, new_off, flag);
    if (cudaError_t e = cudaMemcpyAsync(d_ids_off, new_off, (n + 1) * 8, cudaMemcpyDeviceToDevice, st)) return eng.cuda_fail(e, "copy");
    eng.launched(3);
    uint64_t dummy;
    return eng.finish(flag, d_ids_off, n, n_ids_host ? n_ids_host : &dummy, st);
}


#define CKM(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return eng.cuda_fail(e_, #x); } while (0)

// code point -> id table of the single-character vocabulary entries (once, at creation)
int metaspace_upload(Engine& eng) {
    if (!eng.model.metaspace) return CTK_OK;
    uint32_t cap = 64;
    while (cap < eng.model.char_ids.size() * 2 + 2) cap <<= 1;
    std::vector<uint2> tab(cap, make_uint2(0xFFFFFFFFu, 0u));
    for (auto& kv : eng.model.char_ids) {
        uint32_t h = (kv.first * 0x9E3779B1u) >> 7;
        while (tab[h & (cap - 1)].x != 0xFFFFFFFFu) ++h;
        tab[h & (cap - 1)] = make_uint2(kv.first, kv.second);
    }
    void* p = nullptr;
    CKM(cudaMalloc(&p, (size_t)cap * sizeof(uint2)));
    eng.split_mem.push_back(p);
    CKM(cudaMemcpy(p, tab.data(), (size_t)cap * sizeof(uint2), cudaMemcpyHostToDevice));
    eng.meta_char_tab = p;
    eng.meta_char_mask = cap - 1;
    return CTK_OK;
}

// The Metaspace stage and everything after it, for texts that are already normalised (and split, if Split stages precede).
int encode_metaspace(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n, uint64_t n_bytes, uint32_t* d_ids_u32, uint64_t ids_cap,
                     uint64_t* d_ids_off, uint64_t* n_ids_host, cudaStream_t st) {
    Workspace& ws = eng.ws;
    const HostModel& m = eng.model;
    const int rep_len = utf8_len(m.meta_replacement), prefix = m.meta_prefix ? 1 : 0;
    const int out_width = eng.out_id_width == 2 && eng.run_width == 2 ? 2 : 4;
    uint32_t* err;
    CKM(ws.get(4, 256, (void**)&err));
    CKM(cudaMemsetAsync(err, 0, 256, st));
    eng.mark(nullptr, st);
    // ---- text' and its offsets
    uint64_t *new_len, *new_off;
    CKM(ws.get(73, (n + 2) * 8, (void**)&new_len));
    CKM(ws.get(74, (n + 2) * 8, (void**)&new_off));
    k_meta_len<<<(unsigned)(((n + 1) * 32 + 255) / 256), 256, 0, st>>>(d_text, d_off, n, rep_len, prefix, n_bytes, new_len, err);
    size_t cub_bytes = 0; void* cub_tmp;
    CKM(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, new_len, new_off, n + 1, st));
    CKM(ws.get(5, cub_bytes + 16, &cub_tmp));
    CKM(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, new_len, new_off, n + 1, st));
    CKM(eng.publish({{err, 1, 0}, {new_off + n, 2, 2}}, st));
    CKM(cudaStreamSynchronize(st));
    if (eng.h_flags[0] & ERRF_OFFSETS) return eng.fail(CTK_ERR_ARG, "offsets must start at 0, be non-decreasing and end at the buffer length");
    uint64_t B;
    memcpy(&B, eng.h_flags + 2, 8);
    if (B >= 0xFFFFF000ull) return eng.fail(CTK_ERR_ARG, "one device call handles less than 4 GiB of text");
    if (B == 0) {
        CKM(cudaMemsetAsync(d_ids_off, 0, (n + 1) * 8, st));
        if (n_ids_host) { CKM(cudaStreamSynchronize(st)); *n_ids_host = 0; }
        return CTK_OK;
    }
    uint8_t* text;
    CKM(ws.get(75, B + 128, (void**)&text));
    CKM(cudaMemsetAsync(text + B, 0, 64, st));
    k_meta_copy<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(d_text, d_off, n, m.meta_replacement, rep_len, prefix, new_off, text);
    // ---- words
    const uint64_t n_words = (B + 31) / 32;
    uint32_t *ds, *S, *X, *cnt, *starts, *tmp_ids, *ntok;
    uint64_t *rank, *tok_off;
    CKM(ws.get(76, (n_words + 2) * 4, (void**)&ds));
    CKM(ws.get(77, (n_words + 2) * 4, (void**)&S));
    CKM(ws.get(57, (n_words + 2) * 4, (void**)&X));
    CKM(ws.get(0, (n_words + 2) * 4, (void**)&cnt));
    CKM(ws.get(1, (n_words + 2) * 8, (void**)&rank));
    CKM(cudaMemsetAsync(ds, 0, (n_words + 2) * 4, st));
    k_meta_docbits<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(new_off, n, B, ds);
    k_meta_bits<<<(unsigned)((n_words + 1 + 255) / 256), 256, 0, st>>>(text, B, ds, n_words, S, X, cnt);
    cub::TransformInputIterator<uint64_t, U32to64, const uint32_t*> c_it(cnt, U32to64());
    CKM(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, c_it, rank, n_words + 1, st));
    CKM(ws.get(5, cub_bytes + 16, &cub_tmp));
    CKM(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, c_it, rank, n_words + 1, st));
    CKM(eng.publish({{rank + n_words, 2, 2}}, st));
    CKM(cudaStreamSynchronize(st));
    uint64_t n_pre64;
    memcpy(&n_pre64, eng.h_flags + 2, 8);
    const uint32_t n_pre = (uint32_t)n_pre64;
    CKM(ws.get(2, ((uint64_t)n_pre + 2) * 4, (void**)&starts));
    CKM(ws.get(3, ((uint64_t)n_pre + 2) * 4, (void**)&ntok));
    CKM(ws.get(6, ((uint64_t)n_pre + 2) * 8, (void**)&tok_off));
    CKM(ws.get(7, (B + 64) * 4, (void**)&tmp_ids));
    k_meta_list<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(S, rank, n_words, starts);
    if (n_pre) k_meta_bpe<<<(unsigned)(((uint64_t)n_pre * 32 + 255) / 256), 256, 0, st>>>(eng.tables, static_cast<const uint2*>(eng.meta_char_tab), eng.meta_char_mask,
                                                                                             text, B, X, starts, n_pre, tmp_ids, ntok);
    CKM(cudaMemsetAsync(ntok + n_pre, 0, 4, st));
    cub::TransformInputIterator<uint64_t, U32to64, const uint32_t*> t_it(ntok, U32to64());
    CKM(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, t_it, tok_off, n_pre + 1, st));
    CKM(ws.get(5, cub_bytes + 16, &cub_tmp));
    CKM(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, t_it, tok_off, n_pre + 1, st));
    if (n_pre) {
        const unsigned g = (unsigned)(((uint64_t)n_pre * 32 + 255) / 256);
        if (out_width == 2) k_meta_emit<uint16_t><<<g, 256, 0, st>>>(starts, tok_off, n_pre, tmp_ids, reinterpret_cast<uint16_t*>(d_ids_u32), ids_cap, err);
        else k_meta_emit<uint32_t><<<g, 256, 0, st>>>(starts, tok_off, n_pre, tmp_ids, d_ids_u32, ids_cap, err);
    }
    k_meta_doc<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(new_off, n, starts, n_pre, tok_off, d_ids_off);
    eng.launched(12);
    eng.mark("metaspace encode", st);
    CKM(cudaGetLastError());
    return eng.finish(err, d_ids_off, n, n_ids_host, st);
}

}  // namespace ctk
