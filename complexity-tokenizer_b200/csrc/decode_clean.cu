// clean_up_tokenization_spaces (reference: src/huggingface/mod.rs:749-769), data-parallel and exact.
//
// The reference applies 15 ordered str::replace calls and then split_whitespace().join(" ").
// Every replace only DELETES U+0020 characters:
//   rules 1-6   " ." " ," " !" " ?" " :" " ;"      delete the space right before  . , ! ? : ;
//   rules 7-14  "\" " " \"" "' " " '" "( " " )" "[ " " ]"   delete the space right after " ' ( [  / right before " ' ) ]
//   rule 15     " - " -> "-"                        deletes the space on both sides of a hyphen, if both exist then
// so the outcome is decided per maximal RUN of U+0020 (length m, left neighbour X, right neighbour Y):
//   after rules 1-14:  rem = max(0, m - [X in "'(\[] - [Y in .,!?:;"')\]])     (one rule per side at most, and a
//   single str::replace pass never revisits a position), then rule 15, left to right and non-overlapping: a hyphen
//   with rem >= 1 on both sides takes one space from each; only a shared run with rem == 1 couples two hyphens.
// Finally every maximal White_Space REGION collapses to one ' ' if anything of it is left (another whitespace
// character, or a run with rem >= 1), to nothing if it is all gone or touches a document edge (trim).
// tests: the oracle applies the 15 replaces literally; tests/test_gpu_parity.py and test_gpu_properties.py compare.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "engine.hpp"

namespace ctk {

constexpr uint8_t F_WS = 1, F_SP = 2, F_DS = 4;

__device__ __forceinline__ int ws_len_at(const uint8_t* p, uint64_t i, uint64_t n) {   // White_Space char starting at i
    uint8_t c = p[i];
    if (c == 0x20 || (c >= 9 && c <= 13)) return 1;
    if (c == 0xC2 && i + 1 < n && (p[i + 1] == 0x85 || p[i + 1] == 0xA0)) return 2;
    if (i + 2 < n) {
        uint8_t d = p[i + 1], e = p[i + 2];
        if (c == 0xE1 && d == 0x9A && e == 0x80) return 3;
        if (c == 0xE2 && d == 0x80 && ((e >= 0x80 && e <= 0x8A) || e == 0xA8 || e == 0xA9 || e == 0xAF)) return 3;
        if (c == 0xE2 && d == 0x81 && e == 0x9F) return 3;
        if (c == 0xE3 && d == 0x80 && e == 0x80) return 3;
    }
    return 0;
}

// 16 bytes per thread (buffers are 16-byte aligned and padded); whitespace by SWAR for ASCII words, per byte otherwise
__global__ void __launch_bounds__(256) k_clean_flags(const uint8_t* __restrict__ R, uint64_t n, uint8_t* __restrict__ f,
                                                     uint8_t* __restrict__ emit) {
    const uint64_t base = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 16;
    if (base >= n) return;
    const uint4 v = *reinterpret_cast<const uint4*>(R + base);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t fo[4], eo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t x = w[q];
        if (!(x & 0x80808080u) && base + 4 * q + 4 <= n) {
            const uint32_t H = 0x80808080u;
            const uint32_t sp = ~((x ^ 0x20202020u) + 0x7F7F7F7Fu) & H;                 // byte == ' '
            const uint32_t ct = (x + 0x77777777u) & ~(x + 0x72727272u) & H;             // 9 <= byte <= 13
            const uint32_t ws = sp | ct;
            fo[q] = (ws >> 7) * F_WS | (sp >> 7) * F_SP;
            eo[q] = (~ws & H) >> 7;
        } else {
            uint32_t fw = 0, ew = 0;
            for (int k = 0; k < 4; ++k) {
                const uint64_t i = base + 4 * q + k;
                if (i >= n) break;
                const uint8_t c = (uint8_t)(x >> (8 * k));
                bool ws = false;
                if (c < 0x80) ws = c == 0x20 || (c >= 9 && c <= 13);
                else {
                    for (int back = 0; back <= 2 && (uint64_t)back <= i; ++back) {      // the lead is at most 2 bytes back
                        int L = ws_len_at(R, i - back, n);
                        if (L > back) { ws = true; break; }
                        if ((R[i - back] & 0xC0) != 0x80) break;                        // reached a lead that is not whitespace
                    }
                }
                fw |= (uint32_t)((ws ? F_WS : 0) | (c == 0x20 ? F_SP : 0)) << (8 * k);
                ew |= (uint32_t)(ws ? 0 : 1) << (8 * k);
            }
            fo[q] = fw; eo[q] = ew;
        }
    }
    *reinterpret_cast<uint4*>(f + base) = make_uint4(fo[0], fo[1], fo[2], fo[3]);
    *reinterpret_cast<uint4*>(emit + base) = make_uint4(eo[0], eo[1], eo[2], eo[3]);
}

__global__ void k_clean_docstarts(const uint64_t* __restrict__ raw_off, uint64_t n_docs, uint64_t n, uint8_t* __restrict__ f) {
    uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    uint64_t p = raw_off[d];
    if (p < n) f[p] |= F_DS;                        // several empty documents may share a start: same value written
}

// 16 positions per thread: bit k of the result = position base + k carries `bit` and starts a run of it
// (first of the document, or the byte before does not carry it)
__device__ __forceinline__ uint32_t run_starts16(const uint8_t* __restrict__ f, uint64_t base, uint64_t n, uint8_t bit) {
    const uint4 fv = *reinterpret_cast<const uint4*>(f + base);
    const uint32_t w[4] = {fv.x, fv.y, fv.z, fv.w};
    const uint32_t B = 0x01010101u * bit, D = 0x01010101u * F_DS;
    const int ds_shift = bit == F_WS ? 2 : 1;                                   // moves the F_DS bit onto `bit`
    if (!((w[0] | w[1] | w[2] | w[3]) & B)) return 0;
    uint32_t prev = base ? (uint32_t)f[base - 1] : 0u;                          // flags of the byte before the group
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t x = w[q];
        const uint32_t before = (x << 8) | prev;                                // flags of each byte's predecessor
        prev = x >> 24;
        const uint32_t t = x & B & (~before | ((x & D) >> ds_shift));           // has the bit, and (no predecessor bit or doc start)
#pragma unroll
        for (int k = 0; k < 4; ++k) if ((t >> (8 * k)) & bit) m |= 1u << (4 * q + k);
    }
    if (n - base < 16) m &= (1u << (n - base)) - 1u;
    return m;
}

// acts where a run of U+0020 starts
__device__ __forceinline__ void clean_run_at(const uint8_t* __restrict__ R, uint64_t n, const uint8_t* __restrict__ f,
                                             uint8_t* __restrict__ rr, uint64_t i) {
    uint8_t fi = f[i];
    uint64_t b = i + 1;
    while (b < n && (f[b] & (F_SP | F_DS)) == F_SP) ++b;
    uint64_t m = b - i;
    int dec = 0;
    if (!(fi & F_DS) && i > 0) { uint8_t X = R[i - 1]; dec += (X == '"' || X == '\'' || X == '(' || X == '['); }
    if (b < n && !(f[b] & F_DS)) {
        uint8_t Y = R[b];
        dec += (Y == '.' || Y == ',' || Y == '!' || Y == '?' || Y == ':' || Y == ';' || Y == '"' || Y == '\'' || Y == ')' || Y == ']');
    }
    uint64_t rem = m > (uint64_t)dec ? m - dec : 0;
    uint8_t v = (uint8_t)(rem > 3 ? 3 : rem);
    rr[i] = v;
    rr[b - 1] = v;
}

__global__ void __launch_bounds__(256) k_clean_runs(const uint8_t* __restrict__ R, uint64_t n, const uint8_t* __restrict__ f,
                                                    uint8_t* __restrict__ rr) {
    const uint64_t base = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 16;
    if (base >= n) return;
    uint32_t m = run_starts16(f, base, n, F_SP);
    while (m) { const int k = __ffs(m) - 1; m &= m - 1; clean_run_at(R, n, f, rr, base + k); }
}

// rule 15 along hyphen chains; one thread per byte, acts at chain heads
__device__ __forceinline__ void clean_hyphen_at(const uint8_t* __restrict__ R, uint64_t n, const uint8_t* __restrict__ f,
                                                const uint8_t* __restrict__ rr, uint8_t* __restrict__ fire, uint64_t h) {
    auto candidate = [&](uint64_t q) -> bool {      // '-' with a space on both sides inside one document
        return q > 0 && q + 1 < n && R[q] == '-' && !(f[q] & F_DS) && (f[q - 1] & F_SP) && !(f[q + 1] & F_DS) && (f[q + 1] & F_SP);
    };
    if (!candidate(h)) return;
    // dependent on the previous hyphen?  only through a shared left run with rem == 1 (then the run has <= 3 spaces)
    if (rr[h - 1] == 1) {
        uint64_t a = h - 1;
        while (a > 0 && !(f[a] & F_DS) && (f[a - 1] & F_SP)) --a;
        if (a > 0 && !(f[a] & F_DS) && candidate(a - 1)) return;               // the head of the chain handles us
    }
    uint64_t cur = h;
    int left = rr[cur - 1];
    for (;;) {
        int right = rr[cur + 1];
        int fr = left >= 1 && right >= 1;
        fire[cur] = (uint8_t)fr;
        if (right != 1) break;                                                  // a run with rem == 1 has <= 3 spaces: walk it
        uint64_t e = cur + 1;
        while (e < n && (f[e] & (F_SP | F_DS)) == F_SP) ++e;
        if (!(e < n && candidate(e))) break;
        left = 1 - fr;
        cur = e;
    }
}

__global__ void __launch_bounds__(256) k_clean_hyphens(const uint8_t* __restrict__ R, uint64_t n, const uint8_t* __restrict__ f,
                                                       const uint8_t* __restrict__ rr, uint8_t* __restrict__ fire) {
    const uint64_t base = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 16;
    if (base >= n) return;
    const uint4 v = *reinterpret_cast<const uint4*>(R + base);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t any = 0, z[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {                                              // bytes equal to '-' (exact zero-byte test)
        const uint32_t x = w[q] ^ 0x2D2D2D2Du;
        z[q] = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
        any |= z[q];
    }
    if (!any) return;
    for (int k = 0; k < 16; ++k)
        if ((z[k >> 2] >> (8 * (k & 3))) & 0x80u) { if (base + k < n) clean_hyphen_at(R, n, f, rr, fire, base + k); }
}

// acts where a whitespace region starts
__device__ __forceinline__ void clean_region_at(const uint8_t* __restrict__ R, uint64_t n, const uint8_t* __restrict__ f,
                                                const uint8_t* __restrict__ rr, const uint8_t* __restrict__ fire,
                                                uint8_t* __restrict__ emit, uint64_t i) {
    uint8_t fi = f[i];
    bool alive = false;
    int cur = -1;                                                               // rem of the run we are inside, -1 outside
    uint64_t e = i;
    for (;; ++e) {
        bool in = e < n && (f[e] & F_WS) && !(e > i && (f[e] & F_DS));
        bool sp = in && (f[e] & F_SP);
        if (cur >= 0 && !sp) {                                                  // the run ended at e
            if (e < n && !(f[e] & F_DS) && R[e] == '-') cur -= fire[e];
            alive = alive || cur >= 1;
            cur = -1;
        }
        if (!in) break;
        if (sp) {
            if (cur < 0) {                                                      // a run starts at e
                cur = rr[e];
                if (!(f[e] & F_DS) && e > 0 && R[e - 1] == '-') cur -= fire[e - 1];
            }
        } else alive = true;                                                    // another whitespace character: never deleted
    }
    bool interior = !(fi & F_DS) && i > 0 && e < n && !(f[e] & F_DS);
    if (alive && interior) emit[i] = 1;
}

__global__ void __launch_bounds__(256) k_clean_regions(const uint8_t* __restrict__ R, uint64_t n, const uint8_t* __restrict__ f,
                                                       const uint8_t* __restrict__ rr, const uint8_t* __restrict__ fire,
                                                       uint8_t* __restrict__ emit) {
    const uint64_t base = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 16;
    if (base >= n) return;
    uint32_t m = run_starts16(f, base, n, F_WS);
    while (m) { const int k = __ffs(m) - 1; m &= m - 1; clean_region_at(R, n, f, rr, fire, emit, base + k); }
}

// ---- survivors -> output, without a position per byte in global memory -----------------------------------------
// A tile = 4096 bytes = one CTA, 16 bytes per thread.  Pass 1 counts the tile's survivors; a scan of the tile counts
// places the tiles; pass 2 scans inside the CTA, lays the surviving bytes out in shared memory and writes them with
// aligned 16-byte stores, and gives the documents that start inside the tile their output offsets.
constexpr int CT = 4096, CTH = 256;

__global__ void __launch_bounds__(CTH) k_clean_tile_sums(const uint8_t* __restrict__ emit, uint64_t n, uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t s_part[CTH / 32];
    const uint64_t base = (uint64_t)blockIdx.x * CT + threadIdx.x * 16;
    uint32_t c = 0;
    if (base < n) {                                                  // emit[] is zero beyond n (64 bytes of padding)
        const uint4 e = *reinterpret_cast<const uint4*>(emit + base);
        c = __popc(e.x) + __popc(e.y) + __popc(e.z) + __popc(e.w);  // flags are 0 or 1
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < CTH / 32; ++w) tot += s_part[w];
        tile_sum[blockIdx.x] = tot;
    }
}

__global__ void __launch_bounds__(CTH) k_clean_write(const uint8_t* __restrict__ R, uint64_t n, const uint8_t* __restrict__ f,
                                                     const uint8_t* __restrict__ emit, const uint32_t* __restrict__ tile_base,
                                                     const uint64_t* __restrict__ raw_off, uint64_t n_docs,
                                                     uint8_t* __restrict__ out, uint64_t out_cap, uint64_t* __restrict__ out_off,
                                                     uint32_t* __restrict__ err) {
    __shared__ uint32_t s_thr[CTH + 1];                               // survivors before each thread's 16 bytes
    __shared__ uint32_t s_warp[CTH / 32];
    __shared__ __align__(16) uint8_t s_emit[CT];
    __shared__ __align__(16) uint8_t s_stage[CT + 16];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t t0 = (uint64_t)blockIdx.x * CT, gb = t0 + tid * 16;
    const uint32_t base = tile_base[blockIdx.x], total = tile_base[blockIdx.x + 1] - base;
    if ((uint64_t)base + total > out_cap) { if (tid == 0) atomicOr(err, ERRF_CAPACITY); return; }
    uint4 ev = make_uint4(0, 0, 0, 0), rv = ev, fv = ev;
    if (gb < n) {
        ev = *reinterpret_cast<const uint4*>(emit + gb);
        if (ev.x | ev.y | ev.z | ev.w) { rv = *reinterpret_cast<const uint4*>(R + gb); fv = *reinterpret_cast<const uint4*>(f + gb); }
    }
    *reinterpret_cast<uint4*>(s_emit + tid * 16) = ev;
    const uint32_t mine = __popc(ev.x) + __popc(ev.y) + __popc(ev.z) + __popc(ev.w);
    uint32_t incl = mine;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t off = incl - mine;
    for (int w = 0; w < wid; ++w) off += s_warp[w];
    s_thr[tid] = off;
    if (tid == CTH - 1) s_thr[CTH] = off + mine;
    const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(out) + base) & 15u);
    const uint32_t e[4] = {ev.x, ev.y, ev.z, ev.w}, r[4] = {rv.x, rv.y, rv.z, rv.w}, ff[4] = {fv.x, fv.y, fv.z, fv.w};
    if (mine) {
        uint8_t* d = s_stage + shift + off;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if ((e[k >> 2] >> (8 * (k & 3))) & 0xFFu)
                *d++ = ((ff[k >> 2] >> (8 * (k & 3))) & F_WS) ? (uint8_t)' ' : (uint8_t)(r[k >> 2] >> (8 * (k & 3)));
    }
    __syncthreads();
    if (total) {
        uint8_t* const g = out + base;
        const uint32_t head = min(total, (16u - shift) & 15u);
        if ((uint32_t)tid < head) g[tid] = s_stage[shift + tid];
        const uint32_t body = (total - head) >> 4;
        for (uint32_t v = tid; v < body; v += CTH)
            *reinterpret_cast<uint4*>(g + head + 16 * v) = *reinterpret_cast<const uint4*>(s_stage + shift + head + 16 * v);
        const uint32_t done = head + 16 * body;
        if ((uint32_t)tid < total - done) g[done + tid] = s_stage[shift + done + tid];
    }
    // output offsets of the documents that start in this tile (in the last tile also of those at the very end)
    const uint64_t t1 = t0 + CT < n ? t0 + CT : n;
    const bool last = t0 + CT >= n;
    uint64_t lo = 0, hi = n_docs + 1;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (raw_off[mid] >= t0) hi = mid; else lo = mid + 1; }
    for (uint64_t d = lo + tid; d <= n_docs; d += CTH) {
        const uint64_t pz = raw_off[d];
        if (!(pz < t1 || (last && pz == n))) break;
        const uint32_t q = (uint32_t)(pz - t0);                       // 0 .. CT
        uint32_t before = q < (uint32_t)CT ? s_thr[q >> 4] : s_thr[CTH];
        for (uint32_t k = q & ~15u; k < q && k < (uint32_t)CT; ++k) before += s_emit[k];
        out_off[d] = (uint64_t)base + before;
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return eng.cuda_fail(e_, #x); } while (0)

// raw: valid UTF-8 (so from_utf8_lossy is the identity), n bytes, raw_off[n_docs+1].
int clean_parallel(Engine& eng, const uint8_t* raw, const uint64_t* raw_off, size_t n_docs, uint64_t n, uint8_t* d_out,
                   uint64_t out_cap, uint64_t* d_out_off, uint32_t* err, cudaStream_t st) {
    if (n >= 0xFFFFFFF0ull) return eng.fail(CTK_ERR_ARG, "one decode call handles less than 4 GiB of text");
    Workspace& ws = eng.ws;
    uint8_t *f, *rr, *fire, *emit;
    CK(ws.get(14, n + 64, (void**)&f));
    CK(ws.get(15, n + 64, (void**)&rr));
    CK(ws.get(16, n + 64, (void**)&fire));
    CK(ws.get(17, n + 64, (void**)&emit));
    unsigned gb = (unsigned)((n + 255) / 256), gd = (unsigned)((n_docs + 1 + 255) / 256), g16 = (unsigned)(((n + 15) / 16 + 255) / 256);
    eng.mark(nullptr, st);
    CK(cudaMemsetAsync(fire, 0, n + 64, st));
    CK(cudaMemsetAsync(emit + n, 0, 64, st));
    if (n) {
        eng.mark("clean: memsets", st);
        k_clean_flags<<<g16, 256, 0, st>>>(raw, n, f, emit);
        k_clean_docstarts<<<gd, 256, 0, st>>>(raw_off, n_docs, n, f);
        eng.mark("clean: k_clean_flags", st);
        k_clean_runs<<<g16, 256, 0, st>>>(raw, n, f, rr);
        eng.mark("clean: k_clean_runs", st);
        k_clean_hyphens<<<g16, 256, 0, st>>>(raw, n, f, rr, fire);
        eng.mark("clean: k_clean_hyphens", st);
        k_clean_regions<<<g16, 256, 0, st>>>(raw, n, f, rr, fire, emit);
        eng.mark("clean: k_clean_regions", st);
        eng.launched(5);
    }
    const uint64_t n_tiles = (n + CT - 1) / CT;
    uint32_t *tile_sum, *tile_base;
    CK(ws.get(26, (n_tiles + 2) * 4, (void**)&tile_sum));
    CK(ws.get(46, (n_tiles + 2) * 4, (void**)&tile_base));
    if (n_tiles) k_clean_tile_sums<<<(unsigned)n_tiles, CTH, 0, st>>>(emit, n, tile_sum);
    CK(cudaMemsetAsync(tile_sum + n_tiles, 0, 4, st));
    size_t cub_bytes = 0; void* cub_tmp;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, tile_sum, tile_base, n_tiles + 1, st));
    CK(ws.get(5, cub_bytes + 16, &cub_tmp));
    CK(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, tile_sum, tile_base, n_tiles + 1, st));
    eng.mark("clean: tile sums + scan", st);
    if (n_tiles) k_clean_write<<<(unsigned)n_tiles, CTH, 0, st>>>(raw, n, f, emit, tile_base, raw_off, n_docs, d_out, out_cap, d_out_off, err);
    else CK(cudaMemsetAsync(d_out_off, 0, (n_docs + 1) * 8, st));
    eng.launched(3); eng.mark("clean: k_clean_write", st);
    return CTK_OK;
}

}  // namespace ctk
