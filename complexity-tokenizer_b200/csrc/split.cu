// Split pre-tokenizer stages on the device (reference src/pretokenizers.rs:298-433 applied through the
// Sequence arm, :114-124: every stage maps each piece of the previous stage to a list of pieces).
//
// A stage turns (text, offsets of n texts) into (text', offsets of n' >= 0 pieces) plus the index of every input text's
// first piece.  The pieces then go through the rest of the pipeline AS IF THEY WERE DOCUMENTS -- the ByteLevel stage restarts
// its own pattern at every piece (pretokenizers.rs:170 runs find_iter per piece), which is exactly what the fused encode
// kernel does at a document start -- and the per-piece id offsets are folded back to per-document offsets at the end.
//
//   k_split_mark     the only sequential part: find_iter is sequential by nature (a match starts where the previous one ended),
//                    but it is FRESH after every byte no match can contain (a "neutral" byte, known from the pattern's sets at
//                    load).  One thread per 512 bytes of the buffer walks the automaton from the first such safe start in its
//                    chunk to the first one of the next chunk and MARKS, in bitmaps with one bit per text byte, where a piece
//                    starts (and, for Removed, which bytes are kept).  The text streams through a 32-byte register window
//                    filled by 16-byte loads one block ahead; ASCII bytes index the transition table directly (trans_ascii,
//                    in shared memory), other characters go through the class tables.  A pattern without neutral bytes (one
//                    that covers every character, like the GPT-2 pattern itself) degrades to one thread per text.
//   everything else is data-parallel over the bitmaps (thread per 32 text bytes):
//   k_split_popc     per-word popcounts -> device scans -> rank of every word
//   k_split_offsets  bitmap -> sorted piece offsets (Removed: in the coordinates of the compacted text)
//   k_split_first    first piece of every input text = rank of its start
//   k_split_compact  Removed only: kept bytes -> new text
//   k_split_fold     composes "first piece of" maps of successive stages / folds id offsets back to documents
//
// Algorithmic bytes of a stage: B read + B/4 of bitmaps written and read + 8 P of piece offsets (P pieces).  The marking
// pass is latency-bound, not bandwidth-bound: one dependent shared-memory lookup per byte and thread, so its time is
// about (longest text) x (lookup latency) once the batch fills the machine.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "engine.hpp"
#include "split_walk.cuh"

namespace ctk {

namespace {

constexpr int SPLIT_THREADS = 128;
constexpr uint32_t SPLIT_SMEM_ASCII = 12 * 1024;        // uint16 entries of trans_ascii kept in shared memory (24 KB: 96 states)
constexpr uint32_t SPLIT_SMEM_TRANS = 4 * 1024;         // uint16 entries of trans kept in shared memory (8 KB)

// 32 bytes of text in registers: [base, base + 16) and the block after it, which is loaded while the first is consumed
struct WinReader {
    const uint8_t* t; uint32_t n;                       // the whole buffer (blocks are read whole only below n); below 4 GiB
    uint32_t base = 0x80000000u;                         // nothing loaded yet: every first access lands in a reload branch
    bool loaded = false;
    uint64_t c0 = 0, c1 = 0, n0 = 0, n1 = 0;
    __device__ __forceinline__ void load(uint32_t blk, uint64_t& lo, uint64_t& hi) const {
        if (blk + 16 <= n) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(t + blk));
            lo = v.x | ((uint64_t)v.y << 32); hi = v.z | ((uint64_t)v.w << 32);
        } else {
            lo = hi = 0;
            for (uint32_t k = blk; k < n && k - blk < 16u; ++k) {
                const uint64_t b = t[k];
                if (k - blk < 8) lo |= b << (8 * (k - blk)); else hi |= b << (8 * (k - blk - 8));
            }
        }
    }
    __device__ __forceinline__ uint32_t byte(uint32_t i) {
        uint32_t d = i - base;
        if (d >= 32u) {
            if (d < 48u && loaded) { c0 = n0; c1 = n1; base += 16; load(base + 16, n0, n1); }
            else { base = i & ~15u; load(base, c0, c1); load(base + 16, n0, n1); loaded = true; }
            d = i - base;
        }
        const uint64_t w = d < 16 ? (d < 8 ? c0 : c1) : (d < 24 ? n0 : n1);
        return (uint32_t)(w >> (8 * (d & 7))) & 0xFFu;
    }
};

// sets bits of a global bitmap through one cached word (positions come in non-decreasing order, so a word is flushed once)
struct BitWriter {
    uint32_t* bits; uint32_t word = ~0u; uint32_t acc = 0;
    __device__ __forceinline__ void flush() { if (acc) atomicOr(bits + word, acc); acc = 0; }
    __device__ __forceinline__ void set(uint32_t p) {
        const uint32_t w = p >> 5;
        if (w != word) { flush(); word = w; }
        acc |= 1u << (p & 31);
    }
    __device__ __forceinline__ void set_range(uint32_t a, uint32_t b) {      // [a, b)
        while (a < b) {
            const uint32_t w = a >> 5;
            if (w != word) { flush(); word = w; }
            const uint32_t lo = a & 31u, cnt = min(32u - lo, b - a);
            acc |= (cnt == 32 ? 0xFFFFFFFFu : ((1u << cnt) - 1u)) << lo;
            a += cnt;
        }
    }
};
struct MarkEmit {
    BitWriter starts, keep;
    uint32_t* doc_matched;
    __device__ __forceinline__ void boundary(uint32_t p) { starts.set(p); }
    __device__ __forceinline__ void span(uint32_t a, uint32_t b, bool starts_piece) { if (starts_piece) starts.set(a); keep.set_range(a, b); }
    __device__ __forceinline__ void matched() { if (doc_matched) *doc_matched = 1u; }
};

__device__ __forceinline__ bool neutral_byte(const uint4& m, uint32_t c) {
    if (c >= 128u) return false;
    const uint32_t w = c < 64u ? (c < 32u ? m.x : m.y) : (c < 96u ? m.z : m.w);
    return (w >> (c & 31u)) & 1u;
}

// The first safe start at or after x (x > 0), given the text d that contains position x - 1: the position after a neutral
// byte -- or, for the behaviours that do not look back (pairs != NULL), the position between two bytes no match can contain
// next to each other -- or the next text's start, whichever comes first.
__device__ __forceinline__ uint64_t safe_start(const uint8_t* __restrict__ text, const uint4& neutral, const uint32_t* __restrict__ pairs, uint64_t x,
                                               uint64_t next_text_start) {
    if (pairs) {
        uint32_t c1 = __ldg(text + x - 1);
        for (uint64_t p = x; p < next_text_start; ++p) {
            const uint32_t c2 = __ldg(text + p);
            if (c1 < 128u && c2 < 128u && ((__ldg(pairs + c1 * 4 + (c2 >> 5)) >> (c2 & 31u)) & 1u)) return p;
            c1 = c2;
        }
        return next_text_start;
    }
    for (uint64_t p = x - 1; p + 1 < next_text_start; ++p)
        if (neutral_byte(neutral, __ldg(text + p))) return p + 1;
    return next_text_start;
}

// One thread per SPLIT_CHUNK bytes of the whole buffer.  Thread j walks [S_j, S_j+1), S_j = the first safe start at or after
// j * SPLIT_CHUNK: a text's start or the position after a neutral byte (split_walk.cuh) -- so long texts are walked by many
// threads, and a pattern without neutral bytes degrades to one thread per text.
constexpr uint32_t SPLIT_CHUNK = 512;
__global__ void __launch_bounds__(SPLIT_THREADS) k_split_mark(SplitTables g, uint4 neutral, uint32_t n_trans, uint32_t n_trans_ascii, const uint8_t* __restrict__ text,
                                                              const uint64_t* __restrict__ off, uint64_t n, uint64_t n_bytes,
                                                              uint32_t* __restrict__ start_bits, uint32_t* __restrict__ keep_bits,
                                                              uint32_t* __restrict__ doc_matched) {
    __shared__ uint16_t s_ascii[SPLIT_SMEM_ASCII];
    __shared__ uint16_t s_trans[SPLIT_SMEM_TRANS];
    SplitTables s = g;
    if (g.trans_ascii && n_trans_ascii <= SPLIT_SMEM_ASCII) {
        for (uint32_t i = threadIdx.x; i < n_trans_ascii; i += blockDim.x) s_ascii[i] = g.trans_ascii[i];
        s.trans_ascii = s_ascii;
    }
    if (n_trans <= SPLIT_SMEM_TRANS) {
        for (uint32_t i = threadIdx.x; i < n_trans; i += blockDim.x) s_trans[i] = g.trans[i];
        s.trans = s_trans;
    }
    __syncthreads();
    const uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * SPLIT_CHUNK;
    if (x >= n_bytes) return;
    // d = the text that contains position max(x, 1) - 1: the largest d with off[d] <= that position
    const uint64_t probe = x ? x - 1 : 0;
    uint64_t lo_i = 0, hi_i = n;                             // invariant: off[lo_i] <= probe < off[hi_i]  (off[0] = 0, off[n] = n_bytes > probe)
    while (hi_i - lo_i > 1) {
        const uint64_t mid = (lo_i + hi_i) >> 1;
        if (__ldg(off + mid) <= probe) lo_i = mid; else hi_i = mid;
    }
    uint64_t d = lo_i;
    const uint32_t* pairs = (s.behavior == 1 || s.behavior == 3 || (s.behavior == 0 && !s.invert)) ? g.pair_impossible : nullptr;
    uint64_t seg_lo = x ? safe_start(text, neutral, pairs, x, __ldg(off + d + 1)) : 0;
    if (seg_lo >= n_bytes) return;
    // the end: the same rule at x + SPLIT_CHUNK (the next thread computes the same number)
    uint64_t seg_end = n_bytes;
    if (x + SPLIT_CHUNK < n_bytes) {
        uint64_t de = d;
        while (__ldg(off + de + 1) <= x + SPLIT_CHUNK - 1) ++de;
        seg_end = safe_start(text, neutral, pairs, x + SPLIT_CHUNK, __ldg(off + de + 1));
    }
    if (seg_end <= seg_lo) return;                           // (no safe start inside this chunk: an earlier thread walks through it)
    while (__ldg(off + d + 1) <= seg_lo) ++d;                // the text that contains seg_lo
    WinReader rd{text, (uint32_t)n_bytes};
    MarkEmit em;
    em.starts.bits = start_bits; em.keep.bits = keep_bits;
    while (seg_lo < seg_end) {
        const uint64_t lo = __ldg(off + d), hi = __ldg(off + d + 1);
        const uint64_t seg_hi = hi < seg_end ? hi : seg_end;
        if (seg_hi > seg_lo) {
            em.doc_matched = doc_matched ? doc_matched + d : nullptr;
            if (s.behavior != 0 && seg_lo == lo) em.starts.set((uint32_t)lo);          // the first piece starts where the text does
            split_walk<uint32_t>(s, rd, (uint32_t)lo, (uint32_t)hi, (uint32_t)seg_lo, (uint32_t)seg_hi, em);
        }
        seg_lo = seg_hi;
        ++d;
    }
    em.starts.flush();
    em.keep.flush();
}

// validates the offsets (the marking kernel relies on them)
__global__ void k_split_check(const uint64_t* __restrict__ off, uint64_t n, uint64_t n_bytes, uint32_t* __restrict__ err) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n) return;
    const uint64_t p = off[d];
    if ((d == 0 && p != 0) || (d == n && p != n_bytes) || (d < n && off[d + 1] < p) || p > n_bytes) atomicOr(err, ERRF_OFFSETS);
}

// Removed, not inverted: a text without any match is kept whole (pretokenizers.rs:305-307)
__global__ void k_split_nomatch(const uint64_t* __restrict__ off, uint64_t n, const uint32_t* __restrict__ doc_matched,
                                uint32_t* __restrict__ start_bits, uint32_t* __restrict__ keep_bits) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d >= n || doc_matched[d]) return;
    const uint64_t lo = off[d], hi = off[d + 1];
    if (hi <= lo) return;
    BitWriter sw{start_bits}, kw{keep_bits};
    sw.set((uint32_t)lo); sw.flush();
    kw.set_range((uint32_t)lo, (uint32_t)hi); kw.flush();
}

// thread per bitmap word: popcounts of both bitmaps (index n_words: 0, so that exclusive scans end with the totals)
__global__ void k_split_popc(const uint32_t* __restrict__ start_bits, const uint32_t* __restrict__ keep_bits, uint64_t n_words,
                             uint32_t* __restrict__ start_cnt, uint32_t* __restrict__ keep_cnt) {
    const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w > n_words) return;
    start_cnt[w] = w < n_words ? __popc(start_bits[w]) : 0;
    if (keep_bits) keep_cnt[w] = w < n_words ? __popc(keep_bits[w]) : 0;
}

// piece k starts at the k-th set bit; with a keep bitmap its offset is the number of kept bytes before it
__global__ void k_split_offsets(const uint32_t* __restrict__ start_bits, const uint64_t* __restrict__ start_rank, const uint32_t* __restrict__ keep_bits,
                                const uint64_t* __restrict__ keep_rank, uint64_t n_words, uint64_t n_bytes, uint64_t* __restrict__ piece_off) {
    const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w > n_words) return;
    if (w == n_words) { piece_off[start_rank[n_words]] = keep_bits ? keep_rank[n_words] : n_bytes; return; }
    uint32_t b = start_bits[w];
    uint64_t k = start_rank[w];
    const uint32_t kb = keep_bits ? keep_bits[w] : 0;
    while (b) {
        const int j = __ffs(b) - 1;
        b &= b - 1;
        piece_off[k++] = keep_bits ? keep_rank[w] + __popc(kb & ((1u << j) - 1u)) : w * 32 + j;
    }
}

// first piece of text d = number of piece starts before its first byte (d == n: all of them)
__global__ void k_split_first(const uint64_t* __restrict__ off, uint64_t n, const uint32_t* __restrict__ start_bits, const uint64_t* __restrict__ start_rank,
                              uint64_t n_words, uint64_t* __restrict__ first) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n) return;
    const uint64_t p = off[d], w = p >> 5;
    first[d] = w >= n_words ? start_rank[n_words] : start_rank[w] + __popc(start_bits[w] & ((1u << (p & 31)) - 1u));
}

__global__ void k_split_compact(const uint8_t* __restrict__ text, const uint32_t* __restrict__ keep_bits, const uint64_t* __restrict__ keep_rank,
                                uint64_t n_words, uint8_t* __restrict__ out) {
    const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t b = keep_bits[w];
    uint64_t o = keep_rank[w];
    while (b) {
        const int j = __ffs(b) - 1;
        b &= b - 1;
        out[o++] = text[w * 32 + j];
    }
}

// out[d] = inner[outer[d]] for d <= n: the first piece of text d after two stages / the id offset of document d
__global__ void k_split_fold(const uint64_t* __restrict__ outer, const uint64_t* __restrict__ inner, uint64_t n, uint64_t* __restrict__ out) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d <= n) out[d] = inner[outer[d]];
}

struct U32To64 { __host__ __device__ uint64_t operator()(uint32_t v) const { return v; } };

}  // namespace

#define CKS(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return eng.cuda_fail(e_, #x); } while (0)

// Uploads the automata of eng.model.split_stages (once, at creation).
int split_upload(Engine& eng) {
    for (const SplitStage& sg : eng.model.split_stages) {
        SplitTables t{};
        auto up = [&](const void* src, size_t bytes, const void** dst) -> cudaError_t {
            void* p = nullptr;
            cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
            if (e != cudaSuccess) return e;
            eng.split_mem.push_back(p);
            *dst = p;
            return bytes ? cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
        };
        CKS(up(sg.dfa.trans.data(), sg.dfa.trans.size() * 2, (const void**)&t.trans));
        CKS(up(sg.dfa.ascii_class.data(), sg.dfa.ascii_class.size(), (const void**)&t.ascii_class));
        CKS(up(sg.dfa.stage1.data(), sg.dfa.stage1.size() * 2, (const void**)&t.stage1));
        CKS(up(sg.dfa.blocks.data(), sg.dfa.blocks.size(), (const void**)&t.blocks));
        if (!sg.dfa.trans_ascii.empty()) CKS(up(sg.dfa.trans_ascii.data(), sg.dfa.trans_ascii.size() * 2, (const void**)&t.trans_ascii));
        if (!sg.dfa.pair_impossible.empty() && !getenv("CTK_SPLIT_NO_PAIRS")) CKS(up(sg.dfa.pair_impossible.data(), sg.dfa.pair_impossible.size() * 4, (const void**)&t.pair_impossible));
        t.n_classes = sg.dfa.n_classes; t.start = sg.dfa.start; t.behavior = sg.behavior; t.invert = sg.invert ? 1 : 0;
        eng.split_dev.push_back(t);
    }
    return CTK_OK;
}

// Runs every Split stage in order.  Out: the final pieces (text, offsets, count, bytes) and, when there is at least one
// stage, `first_piece` (n_docs + 1 entries: index of the first final piece of every document; the last entry = n pieces).
// Synchronises the stream (piece counts size the next stage's buffers).
int split_stages(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes,
                 const uint8_t** o_text, const uint64_t** o_off, size_t* o_n, uint64_t* o_bytes, const uint64_t** first_piece, cudaStream_t st) {
    *o_text = d_text; *o_off = d_off; *o_n = n_docs; *o_bytes = n_bytes; *first_piece = nullptr;
    if (eng.split_dev.empty()) return CTK_OK;
    if (n_bytes >= 0xFFFFF000ull) return eng.fail(CTK_ERR_ARG, "one device call handles less than 4 GiB of text");
    Workspace& ws = eng.ws;
    const uint8_t* text = d_text;
    const uint64_t* off = d_off;
    uint64_t n = n_docs, bytes = n_bytes;
    uint64_t* map = nullptr;                                   // first final piece of every document so far
    uint32_t* err;
    CKS(ws.get(58, 64, (void**)&err));
    CKS(cudaMemsetAsync(err, 0, 64, st));
    eng.mark(nullptr, st);
    for (size_t k = 0; k < eng.split_dev.size(); ++k) {
        const SplitTables& t = eng.split_dev[k];
        const SplitDfa& dfa = eng.model.split_stages[k].dfa;
        const int par = (int)(k & 1);                          // ping-pong: a stage reads the previous stage's outputs
        const bool removed = t.behavior == 0;
        const uint64_t n_words = (bytes + 31) / 32;
        uint32_t *start_bits, *keep_bits = nullptr, *start_cnt, *keep_cnt = nullptr;
        uint64_t *start_rank, *keep_rank = nullptr, *first, *piece_off;
        CKS(ws.get(50, (n_words + 2) * 4, (void**)&start_bits));
        CKS(ws.get(51, (n_words + 2) * 4, (void**)&start_cnt));
        CKS(ws.get(54, (n_words + 2) * 8, (void**)&start_rank));
        CKS(ws.get(52 + par, (n + 2) * 8, (void**)&first));
        CKS(cudaMemsetAsync(start_bits, 0, (n_words + 2) * 4, st));
        if (removed) {
            CKS(ws.get(47, (n_words + 2) * 4, (void**)&keep_bits));
            CKS(ws.get(48, (n_words + 2) * 4, (void**)&keep_cnt));
            CKS(ws.get(49, (n_words + 2) * 8, (void**)&keep_rank));
            CKS(cudaMemsetAsync(keep_bits, 0, (n_words + 2) * 4, st));
        }
        const unsigned dgrid = (unsigned)((n + 1 + 255) / 256), wgrid = (unsigned)((n_words + 1 + 255) / 256);
        const bool whole_if_no_match = removed && !t.invert;
        uint32_t* doc_matched = nullptr;
        if (whole_if_no_match) {
            CKS(ws.get(45, (n + 2) * 4, (void**)&doc_matched));
            CKS(cudaMemsetAsync(doc_matched, 0, (n + 2) * 4, st));
        }
        k_split_check<<<dgrid, 256, 0, st>>>(off, n, bytes, err);
        CKS(eng.publish({{err, 1, 0}}, st));                  // (the marking kernel walks the offsets: they are checked first)
        CKS(cudaStreamSynchronize(st));
        if (eng.h_flags[0] & ERRF_OFFSETS) return eng.fail(CTK_ERR_ARG, "offsets must start at 0, be non-decreasing and end at the buffer length");
        if (bytes) {
            const uint64_t n_chunks = (bytes + SPLIT_CHUNK - 1) / SPLIT_CHUNK;
            const uint32_t* nm = eng.model.split_stages[k].dfa.neutral;
            k_split_mark<<<(unsigned)((n_chunks + SPLIT_THREADS - 1) / SPLIT_THREADS), SPLIT_THREADS, 0, st>>>(
                t, make_uint4(nm[0], nm[1], nm[2], nm[3]), (uint32_t)dfa.trans.size(), (uint32_t)dfa.trans_ascii.size(), text, off, n, bytes, start_bits, keep_bits,
                doc_matched);
            if (whole_if_no_match) k_split_nomatch<<<dgrid, 256, 0, st>>>(off, n, doc_matched, start_bits, keep_bits);
        }
        eng.mark("split: k_split_mark", st);
        k_split_popc<<<wgrid, 256, 0, st>>>(start_bits, keep_bits, n_words, start_cnt, keep_cnt);
        cub::TransformInputIterator<uint64_t, U32To64, const uint32_t*> s_it(start_cnt, U32To64()), k_it(keep_cnt, U32To64());
        size_t cub_bytes = 0; void* cub_tmp;
        CKS(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, s_it, start_rank, n_words + 1, st));
        CKS(ws.get(5, cub_bytes + 16, &cub_tmp));
        CKS(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, s_it, start_rank, n_words + 1, st));
        if (removed) CKS(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, k_it, keep_rank, n_words + 1, st));
        CKS(eng.publish({{start_rank + n_words, 2, 2}, {removed ? keep_rank + n_words : start_rank + n_words, 2, 4}}, st));
        CKS(cudaStreamSynchronize(st));
        uint64_t n_new, new_bytes;
        memcpy(&n_new, eng.h_flags + 2, 8);
        memcpy(&new_bytes, eng.h_flags + 4, 8);
        if (!removed) new_bytes = bytes;
        CKS(ws.get(55 + par, (n_new + 2) * 8, (void**)&piece_off));
        uint8_t* out = nullptr;
        k_split_offsets<<<wgrid, 256, 0, st>>>(start_bits, start_rank, keep_bits, keep_rank, n_words, bytes, piece_off);
        k_split_first<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(off, n, start_bits, start_rank, n_words, first);
        if (removed) {
            CKS(ws.get(60 + par, new_bytes + 128, (void**)&out));
            CKS(cudaMemsetAsync(out + new_bytes, 0, 64, st));  // the encode kernels read whole 16-byte words
            k_split_compact<<<wgrid, 256, 0, st>>>(text, keep_bits, keep_rank, n_words, out);
        }
        eng.launched(removed ? 11 : 8);
        if (map) {                                             // documents -> pieces of this stage
            uint64_t* folded;
            CKS(ws.get(62 + par, (n_docs + 2) * 8, (void**)&folded));
            k_split_fold<<<(unsigned)((n_docs + 1 + 255) / 256), 256, 0, st>>>(map, first, n_docs, folded);
            eng.launched(1);
            map = folded;
        } else map = first;
        if (removed) { text = out; bytes = new_bytes; }
        off = piece_off;
        n = n_new;
    }
    eng.mark("split: ranks + offsets", st);
    CKS(cudaGetLastError());
    *o_text = text; *o_off = off; *o_n = (size_t)n; *o_bytes = bytes; *first_piece = map;
    return CTK_OK;
}

// ids_off[d] = piece_ids_off[first_piece[d]], d <= n_docs
int split_fold_ids(Engine& eng, const uint64_t* first_piece, const uint64_t* piece_ids_off, size_t n_docs, uint64_t* d_ids_off, cudaStream_t st) {
    k_split_fold<<<(unsigned)((n_docs + 1 + 255) / 256), 256, 0, st>>>(first_piece, piece_ids_off, n_docs, d_ids_off);
    eng.launched(1);
    CKS(cudaGetLastError());
    return CTK_OK;
}

}  // namespace ctk
