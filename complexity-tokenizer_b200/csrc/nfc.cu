// Unicode NFC on the GPU (reference: Normalizer::NFC, src/normalizers.rs:45-47, the default
// normalizer per src/huggingface/parsing.rs:89).
//
// Almost all text is already NFC.  k_nfc_flag scans every byte once (vectorised) and marks the
// documents that contain an "NFC-suspect" code point (NFC_QC != Yes or ccc != 0; none exist below
// U+0300, i.e. below lead byte 0xCC).  If no document is marked the stage is a no-op.  Otherwise
// marked documents are normalised by a streaming decompose / canonical-reorder / compose pass
// (UAX #15) and every other document is copied, into a fresh packed buffer with new offsets.
#include <cub/device/device_scan.cuh>

#include <cstring>

#include "engine.hpp"

namespace ctk {

__device__ __forceinline__ int nfc_ccc(const NfcTables& t, uint32_t cp) {
    if (cp < 0x300) return 0;
    int lo = 0, hi = t.nq - 1;
    while (lo <= hi) { int mid = (lo + hi) >> 1; uint32_t k = t.qkey[mid]; if (cp < k) hi = mid - 1; else if (cp > k) lo = mid + 1; else return t.qval[mid]; }
    return 0;
}
__device__ __forceinline__ int nfc_decomp_idx(const NfcTables& t, uint32_t cp) {
    if (cp < 0xC0) return -1;
    int lo = 0, hi = t.nd - 1;
    while (lo <= hi) { int mid = (lo + hi) >> 1; uint32_t k = t.dkey[mid]; if (cp < k) hi = mid - 1; else if (cp > k) lo = mid + 1; else return mid; }
    return -1;
}
__device__ __forceinline__ uint32_t nfc_compose(const NfcTables& t, uint32_t a, uint32_t b) {
    if (a >= 0x1100 && a < 0x1113 && b >= 0x1161 && b < 0x1176) return 0xAC00 + ((a - 0x1100) * 21 + (b - 0x1161)) * 28;
    if (a >= 0xAC00 && a < 0xD7A4 && (a - 0xAC00) % 28 == 0 && b > 0x11A7 && b < 0x11C3) return a + (b - 0x11A7);
    uint64_t key = ((uint64_t)a << 21) | b;
    int lo = 0, hi = t.nc - 1;
    while (lo <= hi) { int mid = (lo + hi) >> 1; uint64_t k = t.ckey[mid]; if (key < k) hi = mid - 1; else if (key > k) lo = mid + 1; else return t.cval[mid]; }
    return 0;
}

constexpr int kSegMax = 48;      // code points buffered between two stable starters

struct NfcSink {                 // counts, and writes when out != nullptr
    uint8_t* out; uint64_t n;
    __device__ void put(uint32_t cp) {
        if (out) {
            if (cp < 0x80) out[n] = (uint8_t)cp;
            else if (cp < 0x800) { out[n] = 0xC0 | (cp >> 6); out[n + 1] = 0x80 | (cp & 63); }
            else if (cp < 0x10000) { out[n] = 0xE0 | (cp >> 12); out[n + 1] = 0x80 | ((cp >> 6) & 63); out[n + 2] = 0x80 | (cp & 63); }
            else { out[n] = 0xF0 | (cp >> 18); out[n + 1] = 0x80 | ((cp >> 12) & 63); out[n + 2] = 0x80 | ((cp >> 6) & 63); out[n + 3] = 0x80 | (cp & 63); }
        }
        n += cp < 0x80 ? 1 : cp < 0x800 ? 2 : cp < 0x10000 ? 3 : 4;
    }
};

struct NfcStream {
    const NfcTables& t;
    NfcSink& sink;
    uint32_t buf[kSegMax];
    uint8_t cc[kSegMax];
    int n = 0;
    bool overflow = false;
    __device__ NfcStream(const NfcTables& t_, NfcSink& s_) : t(t_), sink(s_) {}

    // compose buf[0..n) in place (it is canonically ordered), return new length
    __device__ void compose_buf() {
        if (n < 2) return;
        int w = 0, starter = -1, prev_cc = 0;
        for (int k = 0; k < n; ++k) {
            uint32_t c = buf[k]; int c_cc = cc[k];
            if (starter >= 0 && (prev_cc == 0 || prev_cc < c_cc)) {
                uint32_t comp = nfc_compose(t, buf[starter], c);
                if (comp) { buf[starter] = comp; continue; }
            }
            if (c_cc == 0) starter = w;
            prev_cc = c_cc;
            buf[w] = c; cc[w] = (uint8_t)c_cc; ++w;
        }
        n = w;
    }
    __device__ void flush() { compose_buf(); for (int k = 0; k < n; ++k) sink.put(buf[k]); n = 0; }
    // one fully decomposed code point
    __device__ void push_decomposed(uint32_t cp) {
        int c = nfc_ccc(t, cp);
        if (c == 0) {
            // a starter closes the segment unless it can still combine with a lone preceding starter
            compose_buf();
            if (n == 1 && cc[0] == 0) {
                uint32_t comp = nfc_compose(t, buf[0], cp);
                if (comp) { buf[0] = comp; return; }
            }
            for (int k = 0; k < n; ++k) sink.put(buf[k]);
            n = 0;
            buf[0] = cp; cc[0] = 0; n = 1;
            return;
        }
        if (n >= kSegMax) { overflow = true; flush(); }
        // canonical ordering: insert after the last mark with ccc <= c
        int j = n;
        while (j > 0 && cc[j - 1] > c) { buf[j] = buf[j - 1]; cc[j] = cc[j - 1]; --j; }
        buf[j] = cp; cc[j] = (uint8_t)c; ++n;
    }
    __device__ void push(uint32_t cp) {
        if (cp >= 0xAC00 && cp < 0xD7A4) {
            uint32_t s = cp - 0xAC00;
            push_decomposed(0x1100 + s / 588);
            push_decomposed(0x1161 + (s % 588) / 28);
            if (s % 28) push_decomposed(0x11A7 + s % 28);
            return;
        }
        // iterative full canonical decomposition (depth <= 4): expand the first element repeatedly
        uint32_t stack[8]; int sp = 0;
        stack[sp++] = cp;
        while (sp) {
            uint32_t c = stack[--sp];
            int di = nfc_decomp_idx(t, c);
            if (di < 0) { push_decomposed(c); continue; }
            uint32_t a = t.da[di], b = t.db[di];
            if (b && sp < 7) stack[sp++] = b;
            if (sp < 8) stack[sp++] = a;
        }
    }
};

__device__ __forceinline__ uint32_t dec_cp(const uint8_t* s, uint64_t n, uint64_t& i) {
    uint32_t c = s[i];
    if (c < 0x80) { ++i; return c; }
    if (c < 0xE0) { uint32_t r = ((c & 0x1F) << 6) | ((i + 1 < n ? s[i + 1] : 0) & 63); i += 2; return r; }
    if (c < 0xF0) { uint32_t r = ((c & 0x0F) << 12) | (((i + 1 < n ? s[i + 1] : 0) & 63) << 6) | ((i + 2 < n ? s[i + 2] : 0) & 63); i += 3; return r; }
    uint32_t r = ((c & 7) << 18) | (((i + 1 < n ? s[i + 1] : 0) & 63) << 12) | (((i + 2 < n ? s[i + 2] : 0) & 63) << 6) | ((i + 3 < n ? s[i + 3] : 0) & 63);
    i += 4;
    return r;
}

// ---- exact NFC of a segment with ANY number of combining marks per starter, in O(1) memory ----------------------
// The streaming pass above keeps a starter and its marks in a 48-entry buffer.  "Zalgo" text stacks more marks than
// that on one base; such a segment is redone here.  The fully decomposed stream is read through a restartable
// iterator; a block = [starter] + its run of marks.  Canonical order of the run = ascending combining class, stable
// inside a class, so it is produced by one pass over the run PER CLASS that occurs in it (<= 55 classes exist); the
// composition rule of UAX #15 (a mark composes with the starter unless a kept mark of an equal or higher class
// precedes it) only needs the current starter and the class of the last kept mark.  A dry pass finds the final
// starter, a second pass writes it and the marks that stay.  A starter whose marks all vanished is held back: the
// next starter may still compose with it (Hangul L+V+T, and a few others).
constexpr uint32_t kNoCp = 0xFFFFFFFFu;
struct NfcDIter {                       // iterator over the full canonical decomposition of text[i, end)
    const uint8_t* s; uint64_t i, end;
    uint32_t q[8]; int qn, qi;
    __device__ void init(const uint8_t* s_, uint64_t a, uint64_t b) { s = s_; i = a; end = b; qn = qi = 0; }
    __device__ bool next(const NfcTables& t, uint32_t& out) {
        if (qi < qn) { out = q[qi++]; return true; }
        if (i >= end) return false;
        const uint32_t cp = dec_cp(s, end, i);
        qn = 0; qi = 0;
        if (cp >= 0xAC00 && cp < 0xD7A4) {
            const uint32_t x = cp - 0xAC00;
            q[qn++] = 0x1100 + x / 588; q[qn++] = 0x1161 + (x % 588) / 28;
            if (x % 28) q[qn++] = 0x11A7 + x % 28;
        } else {
            uint32_t stack[8]; int sp = 0;
            stack[sp++] = cp;
            while (sp) {
                const uint32_t c = stack[--sp];
                const int di = nfc_decomp_idx(t, c);
                if (di < 0) { if (qn < 8) q[qn++] = c; continue; }
                const uint32_t a = t.da[di], b = t.db[di];
                if (b && sp < 7) stack[sp++] = b;
                if (sp < 8) stack[sp++] = a;
            }
        }
        out = q[qi++];
        return true;
    }
};

__device__ void nfc_segment_slow(const NfcTables& t, const uint8_t* text, uint64_t a, uint64_t b, NfcSink& sink) {
    NfcDIter it;
    it.init(text, a, b);
    uint32_t P = kNoCp;                  // starter held back (all of its marks composed into it)
    uint32_t c = 0;
    bool have = it.next(t, c);
    while (have) {
        uint32_t base = kNoCp;
        if (nfc_ccc(t, c) == 0) {
            base = c;
            if (P != kNoCp) { const uint32_t comp = nfc_compose(t, P, c); if (comp) base = comp; else sink.put(P); }
            P = kNoCp;
            have = it.next(t, c);
        }
        // the run of marks: `c` is its first element (if any), `run` re-reads what follows it
        const NfcDIter run = it;
        const uint32_t first = c;
        uint32_t present[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint64_t cnt = 0;
        while (have) {
            const int cc = nfc_ccc(t, c);
            if (cc == 0) break;
            present[cc >> 5] |= 1u << (cc & 31);
            ++cnt;
            have = it.next(t, c);
        }
        if (cnt == 0) { P = base; continue; }
        uint32_t sfin = base;
        uint64_t left = 0;
        for (int pass = 0; pass < 2; ++pass) {                       // 0: dry run, 1: write
            uint32_t st = base;
            int prev_cc = 0;
            if (pass == 1) {
                if (left == 0) break;
                if (base != kNoCp) sink.put(sfin);
            }
            for (int v = 1; v < 256; ++v) {
                if (!((present[v >> 5] >> (v & 31)) & 1u)) continue;
                NfcDIter r = run;
                uint32_t m = first;
                for (uint64_t k = 0; k < cnt; ++k) {
                    if (nfc_ccc(t, m) == v) {
                        bool composed = false;
                        if (st != kNoCp && (prev_cc == 0 || prev_cc < v)) {
                            const uint32_t x = nfc_compose(t, st, m);
                            if (x) { st = x; composed = true; }
                        }
                        if (!composed) { prev_cc = v; if (pass == 0) ++left; else sink.put(m); }
                    }
                    if (k + 1 < cnt) r.next(t, m);
                }
            }
            if (pass == 0) sfin = st;
        }
        if (left == 0) P = sfin;                                     // may still compose with the next starter
    }
    if (P != kNoCp) sink.put(P);
}

// one thread per 16 bytes: any suspect code point -> set its bit in `susp` and mark its document
__global__ void __launch_bounds__(256) k_nfc_flag(NfcTables t, const uint8_t* __restrict__ text, uint64_t n,
                                                  const uint64_t* __restrict__ off, uint64_t n_docs,
                                                  uint8_t* __restrict__ doc_flag, uint32_t* __restrict__ susp,
                                                  uint32_t* __restrict__ any) {
    uint64_t base = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 16;
    if (base >= n) return;
    uint4 v = *reinterpret_cast<const uint4*>(text + base);          // text buffers are 16-byte aligned and padded
    // quick reject: no byte >= 0xCC (SWAR: (x & 0x80) && ((x & 0x7F) + 0x34) & 0x80)
    auto hi = [](uint32_t x) { return (x & 0x80808080u) & (((x & 0x7F7F7F7Fu) + 0x34343434u)); };
    if (!(hi(v.x) | hi(v.y) | hi(v.z) | hi(v.w))) return;
    const uint8_t* p = text;
    for (int k = 0; k < 16; ++k) {
        uint64_t i = base + k;
        if (i >= n) break;
        uint32_t c = p[i];
        if (c < 0xCC) continue;
        uint64_t j = i;
        uint32_t cp = dec_cp(p, n, j);
        if (trie_nibble(t.trie_index, t.trie_blocks, cp) & 4u) {
            uint64_t lo = 0, hi2 = n_docs;                            // last doc with off[d] <= i
            while (lo + 1 < hi2) { uint64_t mid = (lo + hi2) >> 1; if (off[mid] <= i) lo = mid; else hi2 = mid; }
            doc_flag[lo] = 1;
            atomicOr(&susp[i >> 5], 1u << (i & 31));
            *any = 1;
        }
    }
}

// One warp per document.  Clean documents are copied.  In a flagged document only the SEGMENTS around
// suspect code points change: a segment is the code point before a run of suspect code points (a
// starter with NFC_QC=Yes, so nothing before it can interact) plus the run; lane 0 normalises it with
// the streaming UAX #15 pass, the whole warp copies the unchanged stretches in between.
// SLOW = false: the common kernel; a document in which a segment overflows the streaming buffer is only MARKED
// (doc_flag = 2) and left to the SLOW = true instantiation, which carries nfc_segment_slow (and its registers).
template <bool SLOW>
__global__ void __launch_bounds__(256) k_nfc_doc(NfcTables t, const uint8_t* __restrict__ text, const uint64_t* __restrict__ off,
                                                 uint64_t n_docs, uint8_t* __restrict__ doc_flag,
                                                 const uint32_t* __restrict__ susp, const uint64_t* __restrict__ new_off, uint8_t* out,
                                                 uint64_t* __restrict__ new_len, uint32_t* __restrict__ err) {
    const unsigned full = 0xFFFFFFFFu;
    uint64_t d = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (d >= n_docs) return;
    const int lane = threadIdx.x & 31;
    const uint64_t lo = off[d], hi = off[d + 1];
    uint8_t* o = out ? out + new_off[d] : nullptr;
    if (SLOW ? doc_flag[d] != 2 : doc_flag[d] == 2) return;
    if (!doc_flag[d]) {
        if (!out) { if (lane == 0) new_len[d] = hi - lo; return; }
        for (uint64_t i = lo + lane; i < hi; i += 32) o[i - lo] = text[i];
        return;
    }
    uint64_t cur_in = lo, cur_out = 0;
    bool overflow = false;
    for (uint64_t w0 = lo >> 5; w0 * 32 < hi; w0 += 32) {              // 32 bitmap words per step
        const uint64_t wi = w0 + lane, pb = wi * 32;
        uint32_t word = pb < hi ? susp[wi] : 0u;
        if (pb < lo) word &= (lo - pb >= 32) ? 0u : (~0u << (lo - pb));
        if (pb + 32 > hi) word &= (hi <= pb) ? 0u : ((hi - pb >= 32) ? ~0u : ((1u << (hi - pb)) - 1u));
        unsigned nz = __ballot_sync(full, word != 0);
        while (nz) {
            const int l = __ffs(nz) - 1;
            nz &= nz - 1;
            uint32_t wv = __shfl_sync(full, word, l);
            const uint64_t base = (w0 + l) * 32;
            while (wv) {
                const int b = __ffs(wv) - 1;
                wv &= wv - 1;
                const uint64_t p = base + b;                             // lead byte of a suspect code point
                if (p < cur_in) continue;                                // inside the segment just handled
                uint64_t s0 = p;                                         // segment start: the code point before, if any
                if (p > cur_in) { s0 = p - 1; while (s0 > cur_in && (text[s0] & 0xC0) == 0x80) --s0; }
                if (o) for (uint64_t i = cur_in + lane; i < s0; i += 32) o[cur_out + (i - cur_in)] = text[i];
                cur_out += s0 - cur_in;
                uint64_t seg_end = 0, seg_out = 0;
                if (lane == 0) {
                    NfcSink sink{o ? o + cur_out : nullptr, 0};
                    NfcStream st(t, sink);
                    uint64_t i = s0;
                    st.push(dec_cp(text, hi, i));                        // the base (or the first suspect one at a document start)
                    while (i < hi && ((susp[i >> 5] >> (i & 31)) & 1u)) st.push(dec_cp(text, hi, i));
                    st.flush();
                    if (SLOW) { if (st.overflow) { sink.n = 0; nfc_segment_slow(t, text, s0, i, sink); } }   // more marks on one base than the buffer holds
                    else overflow = overflow || st.overflow;
                    seg_end = i; seg_out = sink.n;
                }
                seg_end = __shfl_sync(full, seg_end, 0);
                seg_out = __shfl_sync(full, seg_out, 0);
                cur_in = seg_end;
                cur_out += seg_out;
            }
        }
    }
    if (o) for (uint64_t i = cur_in + lane; i < hi; i += 32) o[cur_out + (i - cur_in)] = text[i];
    cur_out += hi - cur_in;
    if (lane == 0) {
        if (overflow) { atomicOr(err, ERRF_NFC_LONG); doc_flag[d] = 2; }      // its length / bytes come from the SLOW kernel
        if (!out) new_len[d] = cur_out;
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return eng.cuda_fail(e_, #x); } while (0)

// In: d_text/d_off.  Out: *o_text/*o_off/*o_bytes = the normalised batch (the inputs themselves when
// nothing had to change).
int nfc_stage(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes,
              const uint8_t** o_text, const uint64_t** o_off, uint64_t* o_bytes, cudaStream_t st) {
    *o_text = d_text; *o_off = d_off; *o_bytes = n_bytes;
    eng.last_nfc_needed = false;
    if (!eng.model.nfc || n_bytes == 0 || n_docs == 0) return CTK_OK;
    Workspace& ws = eng.ws;
    uint8_t* flag; uint32_t* any;
    CK(ws.get(26, n_docs + 16, (void**)&flag));
    CK(ws.get(27, 64, (void**)&any));
    uint32_t* susp;
    CK(ws.get(31, (n_bytes / 32 + 2) * 4, (void**)&susp));
    eng.mark(nullptr, st);
    CK(cudaMemsetAsync(flag, 0, n_docs + 16, st));
    CK(cudaMemsetAsync(any, 0, 64, st));
    CK(cudaMemsetAsync(susp, 0, (n_bytes / 32 + 2) * 4, st));
    NfcTables t = eng.nfc;
    k_nfc_flag<<<(unsigned)(((n_bytes + 15) / 16 + 255) / 256), 256, 0, st>>>(t, d_text, n_bytes, d_off, n_docs, flag, susp, any);
    eng.launched(1); eng.mark("k_nfc_flag", st);
    CK(eng.publish({{any, 1, 8}}, st));
    CK(cudaStreamSynchronize(st));
    if (!eng.h_flags[8]) return CTK_OK;
    eng.last_nfc_needed = true;
    uint64_t *new_len, *new_off;
    CK(ws.get(28, (n_docs + 2) * 8, (void**)&new_len));
    CK(ws.get(29, (n_docs + 2) * 8, (void**)&new_off));
    unsigned grid = (unsigned)((n_docs * 32 + 255) / 256);
    eng.mark(nullptr, st);
    k_nfc_doc<false><<<grid, 256, 0, st>>>(t, d_text, d_off, n_docs, flag, susp, nullptr, nullptr, new_len, any + 1);
    eng.launched(1); eng.mark("k_nfc_doc(count)", st);
    CK(cudaMemsetAsync(new_len + n_docs, 0, 8, st));
    size_t cub_bytes = 0; void* cub_tmp;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, new_len, new_off, n_docs + 1, st));
    CK(ws.get(5, cub_bytes + 16, &cub_tmp));
    CK(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, new_len, new_off, n_docs + 1, st));
    eng.launched(1);
    uint64_t total = 0;
    CK(eng.publish({{new_off + n_docs, 2, 10}, {any + 1, 1, 8}}, st));
    CK(cudaStreamSynchronize(st));
    const bool long_runs = (eng.h_flags[8] & ERRF_NFC_LONG) != 0;    // "Zalgo" documents: counted (and later written) by the SLOW kernel
    if (long_runs) {
        k_nfc_doc<true><<<grid, 256, 0, st>>>(t, d_text, d_off, n_docs, flag, susp, nullptr, nullptr, new_len, any + 1);
        CK(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, new_len, new_off, n_docs + 1, st));
        eng.launched(2);
        CK(eng.publish({{new_off + n_docs, 2, 10}}, st));
        CK(cudaStreamSynchronize(st));
    }
    memcpy(&total, eng.h_flags + 10, 8);
    uint8_t* out;
    CK(ws.get(30, total + 64, (void**)&out));
    CK(cudaMemsetAsync(out + total, 0, 64, st));
    eng.mark(nullptr, st);
    k_nfc_doc<false><<<grid, 256, 0, st>>>(t, d_text, d_off, n_docs, flag, susp, new_off, out, nullptr, any + 1);
    if (long_runs) { k_nfc_doc<true><<<grid, 256, 0, st>>>(t, d_text, d_off, n_docs, flag, susp, new_off, out, nullptr, any + 1); eng.launched(1); }
    eng.launched(1); eng.mark("k_nfc_doc(write)", st);
    *o_text = out; *o_off = new_off; *o_bytes = total;
    return CTK_OK;
}

}  // namespace ctk

// ---------------------------------------------------------------------------------------------
// ByteLevel add_prefix_space (reference: byte_level_pretokenize, src/pretokenizers.rs:163-167):
// a document that is not empty and does not start with U+0020 gets one prepended (after normalisation).
namespace ctk {

__global__ void k_prefix_len(const uint8_t* __restrict__ text, const uint64_t* __restrict__ off, uint64_t n_docs,
                             uint64_t* __restrict__ new_len) {
    uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    if (d == n_docs) { new_len[d] = 0; return; }
    uint64_t lo = off[d], hi = off[d + 1];
    new_len[d] = (hi - lo) + ((hi > lo && text[lo] != ' ') ? 1 : 0);
}

__global__ void __launch_bounds__(256) k_prefix_copy(const uint8_t* __restrict__ text, const uint64_t* __restrict__ off,
                                                     uint64_t n_docs, const uint64_t* __restrict__ new_off, uint8_t* __restrict__ out) {
    uint64_t d = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (d >= n_docs) return;
    const int lane = threadIdx.x & 31;
    const uint64_t lo = off[d], hi = off[d + 1];
    uint8_t* o = out + new_off[d];
    const uint64_t shift = (new_off[d + 1] - new_off[d]) - (hi - lo);          // 0 or 1
    if (shift && lane == 0) o[0] = ' ';
    for (uint64_t i = lo + lane; i < hi; i += 32) o[shift + (i - lo)] = text[i];
}

#define CKP(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return eng.cuda_fail(e_, #x); } while (0)

int prefix_space_stage(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes,
                       const uint8_t** o_text, const uint64_t** o_off, uint64_t* o_bytes, cudaStream_t st) {
    *o_text = d_text; *o_off = d_off; *o_bytes = n_bytes;
    if (!eng.model.add_prefix_space || n_docs == 0 || n_bytes == 0) return CTK_OK;
    Workspace& ws = eng.ws;
    uint64_t *new_len, *new_off;
    uint8_t* out;
    CKP(ws.get(12, (n_docs + 2) * 8, (void**)&new_len));
    CKP(ws.get(13, (n_docs + 2) * 8, (void**)&new_off));
    CKP(ws.get(14, n_bytes + n_docs + 128, (void**)&out));
    eng.mark(nullptr, st);
    k_prefix_len<<<(unsigned)((n_docs + 1 + 255) / 256), 256, 0, st>>>(d_text, d_off, n_docs, new_len);
    size_t cub_bytes = 0; void* cub_tmp;
    CKP(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, new_len, new_off, n_docs + 1, st));
    CKP(ws.get(5, cub_bytes + 16, &cub_tmp));
    CKP(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, new_len, new_off, n_docs + 1, st));
    k_prefix_copy<<<(unsigned)((n_docs * 32 + 255) / 256), 256, 0, st>>>(d_text, d_off, n_docs, new_off, out);
    eng.launched(3); eng.mark("prefix_space", st);
    uint64_t total = 0;
    CKP(cudaMemcpyAsync(&total, new_off + n_docs, 8, cudaMemcpyDeviceToHost, st));
    CKP(cudaStreamSynchronize(st));
    CKP(cudaMemsetAsync(out + total, 0, 64, st));
    *o_text = out; *o_off = new_off; *o_bytes = total;
    return CTK_OK;
}

}  // namespace ctk
