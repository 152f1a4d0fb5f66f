// Split pre-tokenizer stages on the device (reference src/pretokenizers.rs:298-433 applied through the
// Sequence arm, :114-124: every stage maps each piece of the previous stage to a list of pieces).
//
// A stage turns (text, offsets of n texts) into (text', offsets of n' >= 0 pieces) plus the index of every input text's
// first piece.  The pieces then go through the rest of the pipeline AS IF THEY WERE DOCUMENTS -- the ByteLevel stage restarts
// its own pattern at every piece (pretokenizers.rs:170 runs find_iter per piece), which is exactly what the fused encode
// kernel does at a document start -- and the per-piece id offsets are folded back to per-document offsets at the end.
//
//   k_split_count   one thread per text: walk the automaton, count pieces (and kept bytes for Removed)
//   (device scans)
//   k_split_write   the same walk, writing the piece offsets (Removed: also copies the kept spans into a new text)
//   k_split_fold    composes "first piece of" maps of successive stages / folds id offsets back to documents
//
// find_iter is sequential by nature (a match starts where the previous one ended), so a text is walked by ONE thread;
// the batch supplies the parallelism (190 726 documents in the headline workload).  The automaton's transition table sits
// in shared memory when it fits (48 KB), the class tables in L1/L2.
#include <cub/device/device_scan.cuh>

#include "engine.hpp"
#include "split_walk.cuh"

namespace ctk {

namespace {

struct CountEmit {
    uint64_t last = ~0ull, n = 0, bytes = 0;
    CTK_HD void boundary(uint64_t p) { if (p != last) { ++n; last = p; } }
    CTK_HD void span(uint64_t a, uint64_t b) { ++n; bytes += b - a; }
};
struct WriteEmit {
    uint64_t last = ~0ull;
    uint64_t* dst;                  // next piece-offset slot
    uint64_t out_pos;               // Removed: position in the new text
    const uint8_t* text; uint8_t* out;
    CTK_HD void boundary(uint64_t p) { if (p != last) { *dst++ = p; last = p; } }
    CTK_HD void span(uint64_t a, uint64_t b) {
        *dst++ = out_pos;
        for (uint64_t i = a; i < b; ++i) out[out_pos++] = text[i];
    }
};

constexpr int SPLIT_THREADS = 128;
constexpr uint32_t SPLIT_SMEM_ENTRIES = 24 * 1024;      // uint16 transitions kept in shared memory (48 KB)

__device__ __forceinline__ SplitTables stage_tables(const SplitTables& g, uint16_t* s_trans, uint32_t n_trans) {
    SplitTables s = g;
    if (n_trans <= SPLIT_SMEM_ENTRIES) {
        for (uint32_t i = threadIdx.x; i < n_trans; i += blockDim.x) s_trans[i] = g.trans[i];
        __syncthreads();
        s.trans = s_trans;
    }
    return s;
}

// n_pieces[d], kept[d] for d < n; both 0 at d == n (so that exclusive scans give totals)
__global__ void __launch_bounds__(SPLIT_THREADS) k_split_count(SplitTables g, uint32_t n_trans, const uint8_t* __restrict__ text,
                                                               const uint64_t* __restrict__ off, uint64_t n, uint64_t n_bytes,
                                                               uint64_t* __restrict__ n_pieces, uint64_t* __restrict__ kept,
                                                               uint32_t* __restrict__ err) {
    extern __shared__ uint16_t s_trans[];
    const SplitTables s = stage_tables(g, s_trans, n_trans);
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n) return;
    if (d == n) { n_pieces[d] = 0; kept[d] = 0; if (off[d] != n_bytes) atomicOr(err, ERRF_OFFSETS); return; }
    const uint64_t lo = off[d], hi = off[d + 1];
    if ((d == 0 && lo != 0) || hi < lo || hi > n_bytes) { atomicOr(err, ERRF_OFFSETS); n_pieces[d] = 0; kept[d] = 0; return; }
    CountEmit em;
    split_walk(s, text, lo, hi, em);
    if (s.behavior == 0) { n_pieces[d] = em.n; kept[d] = em.bytes; }
    else { n_pieces[d] = hi > lo ? em.n + 1 : 0; kept[d] = hi - lo; }
}

__global__ void __launch_bounds__(SPLIT_THREADS) k_split_write(SplitTables g, uint32_t n_trans, const uint8_t* __restrict__ text,
                                                               const uint64_t* __restrict__ off, uint64_t n,
                                                               const uint64_t* __restrict__ piece_base, const uint64_t* __restrict__ byte_base,
                                                               uint64_t* __restrict__ piece_off, uint8_t* __restrict__ out) {
    extern __shared__ uint16_t s_trans[];
    const SplitTables s = stage_tables(g, s_trans, n_trans);
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n) return;
    if (d == n) { piece_off[piece_base[n]] = s.behavior == 0 ? byte_base[n] : off[n]; return; }
    const uint64_t lo = off[d], hi = off[d + 1];
    if (piece_base[d + 1] == piece_base[d]) return;
    WriteEmit em;
    em.dst = piece_off + piece_base[d];
    em.out_pos = byte_base[d];
    em.text = text; em.out = out;
    if (s.behavior != 0) *em.dst++ = lo;                     // the first piece starts where the text does
    split_walk(s, text, lo, hi, em);
}

// out[d] = inner[outer[d]] for d <= n: the first piece of text d after two stages / the id offset of document d
__global__ void k_split_fold(const uint64_t* __restrict__ outer, const uint64_t* __restrict__ inner, uint64_t n, uint64_t* __restrict__ out) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d <= n) out[d] = inner[outer[d]];
}

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
        const uint32_t n_trans = (uint32_t)eng.model.split_stages[k].dfa.trans.size();
        const size_t smem = n_trans <= SPLIT_SMEM_ENTRIES ? (size_t)n_trans * 2 : 0;
        if (smem > 48 * 1024) return eng.fail(CTK_ERR_CUDA, "internal: split table does not fit");
        const int par = (int)(k & 1);                          // ping-pong: a stage reads the previous stage's outputs
        uint64_t *n_pieces, *kept, *piece_base, *byte_base, *piece_off;
        CKS(ws.get(50, (n + 2) * 8, (void**)&n_pieces));
        CKS(ws.get(51, (n + 2) * 8, (void**)&kept));
        CKS(ws.get(52 + par, (n + 2) * 8, (void**)&piece_base));
        CKS(ws.get(54, (n + 2) * 8, (void**)&byte_base));
        const unsigned grid = (unsigned)((n + 1 + SPLIT_THREADS - 1) / SPLIT_THREADS);
        k_split_count<<<grid, SPLIT_THREADS, smem, st>>>(t, n_trans, text, off, n, bytes, n_pieces, kept, err);
        size_t cub_bytes = 0; void* cub_tmp;
        CKS(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, n_pieces, piece_base, n + 1, st));
        CKS(ws.get(5, cub_bytes + 16, &cub_tmp));
        CKS(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, n_pieces, piece_base, n + 1, st));
        CKS(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, kept, byte_base, n + 1, st));
        CKS(eng.publish({{err, 1, 0}, {piece_base + n, 2, 2}, {byte_base + n, 2, 4}}, st));
        CKS(cudaStreamSynchronize(st));
        if (eng.h_flags[0] & ERRF_OFFSETS) return eng.fail(CTK_ERR_ARG, "offsets must start at 0, be non-decreasing and end at the buffer length");
        uint64_t n_new, new_bytes;
        memcpy(&n_new, eng.h_flags + 2, 8);
        memcpy(&new_bytes, eng.h_flags + 4, 8);
        CKS(ws.get(55 + par, (n_new + 2) * 8, (void**)&piece_off));
        uint8_t* out = nullptr;
        if (t.behavior == 0) {
            CKS(ws.get(60 + par, new_bytes + 128, (void**)&out));
            CKS(cudaMemsetAsync(out + new_bytes, 0, 64, st));  // the encode kernels read whole 16-byte words
        }
        k_split_write<<<grid, SPLIT_THREADS, smem, st>>>(t, n_trans, text, off, n, piece_base, byte_base, piece_off, out);
        eng.launched(5);
        if (map) {                                             // documents -> pieces of this stage
            uint64_t* folded;
            CKS(ws.get(62 + par, (n_docs + 2) * 8, (void**)&folded));
            k_split_fold<<<(unsigned)((n_docs + 1 + 255) / 256), 256, 0, st>>>(map, piece_base, n_docs, folded);
            eng.launched(1);
            map = folded;
        } else map = piece_base;
        if (t.behavior == 0) { text = out; bytes = new_bytes; }
        off = piece_off;
        n = n_new;
    }
    eng.mark("split stages", st);
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
