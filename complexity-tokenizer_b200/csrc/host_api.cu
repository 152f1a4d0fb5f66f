// Host-buffer entry points of the C ABI: ctk_encode_batch / ctk_decode_batch.
//
// These are what the reference's PyO3 methods would call (bindings/tokenizer.rs:207-210, 226-238).
// encode_batch pipelines the batch in chunks of whole documents (4 .. 64 MiB) over three streams so that the PCIe
// copy in, the kernels and the copy out overlap:
//     st_h2d : text chunk c+1, c+2, ...      (pinned user memory is DMA'd directly)
//     st_comp: NFC check + fused encode of chunk c
//     st_d2h : ids of chunk c-1 into a pooled pinned result buffer
// The pre-token cache is cleared at the first chunk only: the chunks are one batch.
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <atomic>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <condition_variable>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include "engine.hpp"

namespace ctk {

// process-wide pool of pinned host buffers (pinning a GiB costs far more than encoding it)
struct PinnedPool {
    struct Buf { void* p; size_t cap; };
    std::mutex mu;
    std::vector<Buf> free_;
    cudaError_t get(size_t bytes, void** out, size_t* cap) {
        {
            std::lock_guard<std::mutex> lk(mu);
            int best = -1;
            for (int i = 0; i < (int)free_.size(); ++i)
                if (free_[i].cap >= bytes && (best < 0 || free_[i].cap < free_[best].cap)) best = i;
            if (best >= 0) { *out = free_[best].p; *cap = free_[best].cap; free_.erase(free_.begin() + best); return cudaSuccess; }
        }
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(out, want, cudaHostAllocDefault);
        if (e != cudaSuccess) { want = bytes + 64; e = cudaHostAlloc(out, want, cudaHostAllocDefault); }
        *cap = want;
        return e;
    }
    void put(void* p, size_t cap) {
        if (!p) return;
        std::lock_guard<std::mutex> lk(mu);
        if (free_.size() >= 8) {                         // drop the smallest
            int s = 0;
            for (int i = 1; i < (int)free_.size(); ++i) if (free_[i].cap < free_[s].cap) s = i;
            if (free_[s].cap < cap) { cudaFreeHost(free_[s].p); free_[s] = {p, cap}; } else cudaFreeHost(p);
            return;
        }
        free_.push_back({p, cap});
    }
};
static PinnedPool g_pinned;
cudaError_t pinned_get(size_t bytes, void** out, size_t* cap) { return g_pinned.get(bytes, out, cap); }
void pinned_put(void* p, size_t cap) { g_pinned.put(p, cap); }

struct Result {
    size_t n = 0;
    void *ids = nullptr, *off = nullptr, *bytes = nullptr;
    size_t ids_cap = 0, off_cap = 0, bytes_cap = 0;
};

// ---- narrow ids on the wire ------------------------------------------------------------------------------
// The result copy is as large as the input copy (4 bytes per ~4.4-byte token) and both share the PCIe link.
// Ids are therefore packed on the device to 2 bytes (every id < 65 536) or 3 bytes (< 2^24) before the copy out
// and widened to the uint32 the ABI promises by a few host threads while later chunks are still in flight.
// Measured on the B200 box (1 GiB, tools/diag_e2e_threads.sh): 25.3 ms plain, 24.1 ms with 8 threads and streaming
// stores, 26.9 ms with 4: host memory bandwidth, not the link, is what the narrower copy runs into.  Hence OPT-IN
// (CTK_WIDEN_THREADS=n); the default copies plain uint32.
__global__ void __launch_bounds__(256) k_pack_ids16(const uint32_t* __restrict__ ids, uint64_t n, uint32_t* __restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; 2 * i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t a = ids[2 * i], b = 2 * i + 1 < n ? ids[2 * i + 1] : 0u;
        out[i] = (a & 0xFFFFu) | (b << 16);
    }
}
__global__ void __launch_bounds__(256) k_pack_ids24(const uint32_t* __restrict__ ids, uint64_t n, uint32_t* __restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; 4 * i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = 4 * i + k < n ? ids[4 * i + k] & 0xFFFFFFu : 0u;
        out[3 * i] = v[0] | (v[1] << 24);
        out[3 * i + 1] = (v[1] >> 8) | (v[2] << 16);
        out[3 * i + 2] = (v[2] >> 16) | (v[3] << 8);
    }
}

static void widen_range(const uint8_t* src, uint32_t* dst, uint64_t lo, uint64_t hi, int width) {
    if (width == 2) {
        const uint16_t* s = reinterpret_cast<const uint16_t*>(src);
        uint64_t i = lo;
#if defined(__SSE2__)
        // streaming stores: the result is not read again by this thread, and a plain store would first read the line
        while (i < hi && (reinterpret_cast<uintptr_t>(dst + i) & 15)) { dst[i] = s[i]; ++i; }
        const __m128i z = _mm_setzero_si128();
        for (; i + 8 <= hi; i += 8) {
            const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_unpacklo_epi16(v, z));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 4), _mm_unpackhi_epi16(v, z));
        }
        _mm_sfence();
#endif
        for (; i < hi; ++i) dst[i] = s[i];
    } else {
        for (uint64_t i = lo; i < hi; ++i) { const uint8_t* q = src + 3 * i; dst[i] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16); }
    }
}

// A few persistent host threads.  One job = one ctk_encode_batch call; the caller publishes chunks as their copies are
// issued, every worker waits for a chunk's copy (cudaEventSynchronize) and widens its share of it.
struct WidenJob {
    struct Chunk { cudaEvent_t ev; const uint8_t* src; uint32_t* dst; uint64_t n; };
    std::vector<Chunk> chunks;                      // reserved up front: never reallocates while workers read it
    std::atomic<size_t> published{0};
    std::atomic<size_t> finished{0};                // chunk completions, counted once per worker
    std::atomic<bool> closed{false};
    int width = 4;
};
struct WidenPool {
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::vector<std::thread> threads;
    WidenJob* job = nullptr;
    uint64_t job_seq = 0;
    std::mutex use_mu;                              // one encode call at a time uses the pool; others copy plain uint32
    int active = 0;
    bool stop = false;
    int n_threads = 0;
    void start(int n) {
        std::lock_guard<std::mutex> lk(mu);
        if (!threads.empty() || n <= 0) return;
        n_threads = n;
        for (int t = 0; t < n; ++t) threads.emplace_back([this, t] { run(t); });
    }
    void run(int t) {
        uint64_t seen = 0;
        for (;;) {
            WidenJob* j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || (job && job_seq != seen); });
                if (stop) return;
                j = job; seen = job_seq; ++active;
            }
            size_t c = 0;
            for (;;) {
                while (c >= j->published.load(std::memory_order_acquire)) {
                    if (j->closed.load(std::memory_order_acquire) && c >= j->published.load(std::memory_order_acquire)) goto done;
                    std::this_thread::yield();
                }
                const WidenJob::Chunk& ch = j->chunks[c];
                cudaEventSynchronize(ch.ev);
                const uint64_t per = (ch.n + n_threads - 1) / n_threads;
                const uint64_t lo = std::min<uint64_t>(ch.n, per * t), hi = std::min<uint64_t>(ch.n, lo + per);
                widen_range(ch.src, ch.dst, lo, hi, j->width);
                j->finished.fetch_add(1, std::memory_order_release);
                ++c;
            }
        done:
            {
                std::lock_guard<std::mutex> lk(mu);
                --active;
            }
            cv_done.notify_all();
        }
    }
    void begin(WidenJob* j) {
        { std::lock_guard<std::mutex> lk(mu); job = j; ++job_seq; }
        cv.notify_all();
    }
    // all chunks published so far are widened by every worker
    void drain(WidenJob* j) {
        const size_t want = j->published.load() * (size_t)n_threads;
        while (j->finished.load(std::memory_order_acquire) < want) std::this_thread::yield();
    }
    void end(WidenJob* j) {
        j->closed.store(true, std::memory_order_release);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return active == 0 && j->finished.load() >= j->published.load() * (size_t)n_threads; });
        job = nullptr;
    }
    ~WidenPool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto& th : threads) if (th.joinable()) th.join();
    }
};
static WidenPool g_widen;

// ---- pageable input ------------------------------------------------------------------------------------------
// A caller's buffer is usually NOT page-locked (a Rust Vec, a NumPy array, an Arrow buffer, Python bytes).  The driver
// then stages it through its own bounce buffer on the calling thread: 9.2 GB/s for 1 GiB on the B200 box, against
// 41 GB/s from page-locked memory (tools/diag_pageable.py).  With the staging below: 25 GB/s at 4 threads, 33 at 8.  Here a helper thread copies chunk after chunk into pooled page-locked staging
// buffers with a few worker threads (the copy of chunk c+1 runs while chunk c is on the wire) and issues the DMA.
struct TaskPool {                                   // persistent workers; run(n, fn) = fn(0) .. fn(n-1), caller helps, returns when done
    struct Job { std::function<void(int)> fn; int n = 0; std::atomic<int> next{0}, done{0}; };
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::thread> threads;
    std::shared_ptr<Job> job;
    uint64_t gen = 0;
    bool stop = false;
    void start(int n) {
        std::lock_guard<std::mutex> lk(mu);
        while ((int)threads.size() < n) threads.emplace_back([this] { loop(); });
    }
    static void work(Job& j) {
        for (;;) {
            const int i = j.next.fetch_add(1, std::memory_order_relaxed);
            if (i >= j.n) return;
            j.fn(i);
            j.done.fetch_add(1, std::memory_order_release);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            std::shared_ptr<Job> j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen; j = job;
            }
            if (j) work(*j);
        }
    }
    void run(int n, std::function<void(int)> fn) {
        auto j = std::make_shared<Job>();
        j->fn = std::move(fn); j->n = n;
        { std::lock_guard<std::mutex> lk(mu); job = j; ++gen; }
        cv.notify_all();
        work(*j);
        while (j->done.load(std::memory_order_acquire) < n) std::this_thread::yield();
    }
    ~TaskPool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto& t : threads) if (t.joinable()) t.join();
    }
};
static TaskPool g_copy_pool;

static bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// DMA between the device and an UNALIGNED pinned host address runs ~10 % slower (tools/diag_pcie.py: 42.4 vs
// 46.6 GB/s per direction with both directions busy): copy the few bytes up to the next 4 KiB boundary of the
// host address separately, then the aligned rest.
static cudaError_t copy_host_aligned(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
    const uintptr_t host = kind == cudaMemcpyHostToDevice ? (uintptr_t)src : (uintptr_t)dst;
    const size_t head = (size_t)((0 - host) & 4095u);
    if (head && bytes > head + (64u << 10)) {
        cudaError_t e = cudaMemcpyAsync(dst, src, head, kind, st);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync((char*)dst + head, (const char*)src + head, bytes - head, kind, st);
    }
    return cudaMemcpyAsync(dst, src, bytes, kind, st);
}

static void free_result(Result* r) {
    if (!r) return;
    g_pinned.put(r->ids, r->ids_cap);
    g_pinned.put(r->off, r->off_cap);
    g_pinned.put(r->bytes, r->bytes_cap);
    delete r;
}

}  // namespace ctk

using namespace ctk;

#define CKE(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = eng->cuda_fail(e_, #x); goto done; } } while (0)

extern "C" {

int ctk_encode_batch(const ctk_tokenizer* tok, const uint8_t* text, const uint64_t* text_off, size_t n, ctk_result** res) {
    if (!tok || !text_off || !res) { set_last_error("NULL argument"); return CTK_ERR_ARG; }
    *res = nullptr;
    Engine* eng = const_cast<Engine*>(reinterpret_cast<const Engine*>(tok));
    if (text_off[0] != 0) { set_last_error("text_off[0] must be 0"); return CTK_ERR_ARG; }
    const uint64_t B = text_off[n];
    if (B && !text) { set_last_error("NULL text"); return CTK_ERR_ARG; }
    // Chunks of whole documents.  Steady state 64 MiB (fewer DMA operations measured faster than 16 or 32 MiB: tools/diag_e2e.py); the first chunks are small so that the copy out starts early
    // and the last ones shrink so that the drain (kernels + copy out of the last chunk) is short.  Each chunk is
    // copied to a 256-byte aligned place of its own on the device.
    uint64_t steady = 64ull << 20;
    if (const char* e = getenv("CTK_CHUNK_MB")) { long v = atol(e); if (v >= 1 && v <= 2048) steady = (uint64_t)v << 20; }
    const uint64_t small = std::min<uint64_t>(steady, 4ull << 20);
    struct Chunk { size_t d0, d1; uint64_t b0, b1, dev_text; size_t roff; };
    std::vector<Chunk> chunks;
    {
        size_t d = 0;
        uint64_t dev = 0;
        size_t ro = 0;
        while (d < n) {
            size_t e = d;
            uint64_t b0 = text_off[d];
            uint64_t target = std::min<uint64_t>(steady, small << std::min<size_t>(chunks.size(), 16));
            if (B - b0 < 2 * target) target = std::max<uint64_t>(small, (B - b0) / 2);
            while (e < n && (e == d || text_off[e + 1] - b0 <= target)) {
                if (text_off[e + 1] < text_off[e]) { set_last_error("text_off must be non-decreasing"); return CTK_ERR_ARG; }
                ++e;
            }
            uint64_t b1 = text_off[e];
            if (b1 - b0 >= 0xFFFFFFF0ull) { set_last_error("a single document of 4 GiB or more is not supported"); return CTK_ERR_ARG; }
            chunks.push_back({d, e, b0, b1, dev, ro});
            dev += ((b1 - b0 + 64 + 255) / 256) * 256;
            ro += (e - d) + 1;
            d = e;
        }
    }
    std::lock_guard<std::mutex> lk(eng->mu);
    int rc = CTK_OK;
    Result* r = new Result();
    r->n = n;
    uint8_t* d_text = nullptr; uint64_t *d_roff = nullptr, *d_ids_off = nullptr; uint32_t* d_ids = nullptr;
    uint64_t* h_roff = nullptr; size_t h_roff_cap = 0;
    uint64_t* h_ioff = nullptr; size_t h_ioff_cap = 0;
    uint64_t total = 0, dev_text_bytes = 0, ids_cap = 0;
    size_t n_roff = n + chunks.size() + 1;
    std::vector<cudaEvent_t> evs, evs_pk, evs_dh;
    std::vector<uint64_t> chunk_base;
    // narrow ids on the wire (see k_pack_ids16): staging buffers, the job the widening threads work on
    WidenJob job;
    std::unique_lock<std::mutex> widen_lock;
    bool packed = false;
    int width = 4;
    uint8_t *h_pack = nullptr, *d_pack = nullptr;
    size_t h_pack_cap = 0;
    uint64_t pk_off = 0, h2d_bytes = 0, d2h_bytes = 0;
    // pageable input: staging buffers, the helper thread that fills them and issues the copies, how far it got
    constexpr int NSTAGE = 3;
    void* stage_buf[NSTAGE] = {nullptr, nullptr, nullptr};
    size_t stage_cap[NSTAGE] = {0, 0, 0};
    cudaEvent_t stage_free[NSTAGE] = {nullptr, nullptr, nullptr};
    std::thread stager;
    std::atomic<size_t> h2d_issued{0};
    std::atomic<int> stager_err{0};
    bool staged = false;
    // CTK_TRACE=1: per-chunk timeline (H2D done, kernels done, D2H done; ms since the call started) on stderr
    const bool trace = getenv("CTK_TRACE") != nullptr;
    std::vector<cudaEvent_t> tr_h, tr_c, tr_d;
    std::vector<double> tr_host;
    cudaEvent_t tr0 = nullptr;
    auto now_ms = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double host0 = now_ms();
    CKE(cudaSetDevice(eng->device));
    dev_text_bytes = chunks.empty() ? 256 : chunks.back().dev_text + ((chunks.back().b1 - chunks.back().b0 + 64 + 255) / 256) * 256;
    ids_cap = B + B / 8 + n + 1024;
    CKE(eng->ws.get(20, dev_text_bytes, (void**)&d_text));
    CKE(eng->ws.get(21, n_roff * 8, (void**)&d_roff));
    CKE(eng->ws.get(22, n_roff * 8, (void**)&d_ids_off));
    CKE(eng->ws.get(23, ids_cap * 4, (void**)&d_ids));
    CKE(g_pinned.get(n_roff * 8, (void**)&h_roff, &h_roff_cap));
    CKE(g_pinned.get(n_roff * 8, (void**)&h_ioff, &h_ioff_cap));
    CKE(g_pinned.get((n + 1) * 8, &r->off, &r->off_cap));
    CKE(g_pinned.get((B / 3 + n + 1024) * 4, &r->ids, &r->ids_cap));
    {
        uint64_t max_emit = eng->model.id_present.empty() ? 0 : eng->model.id_present.size() - 1;
        for (const AddedTok& a : eng->model.added) if (a.may_match) max_emit = std::max<uint64_t>(max_emit, a.id);
        int T = 0;                                                   // opt-in: measured +5 % end to end for 8 busy host threads
        if (const char* e = getenv("CTK_WIDEN_THREADS")) T = atoi(e);
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0) T = std::min(T, std::max(1, hw - 2));
        if (T > 0 && B >= (16ull << 20) && max_emit < (1ull << 24)) {
            widen_lock = std::unique_lock<std::mutex>(g_widen.use_mu, std::try_to_lock);
            if (widen_lock.owns_lock()) { g_widen.start(T); packed = g_widen.n_threads > 0; }
        }
        if (packed) {
            width = max_emit < 65536 ? 2 : 3;
            CKE(eng->ws.get(45, (uint64_t)width * ids_cap + 256 * (chunks.size() + 2), (void**)&d_pack));
            CKE(g_pinned.get((uint64_t)width * (r->ids_cap / 4) + 256 * (chunks.size() + 2), (void**)&h_pack, &h_pack_cap));
            job.chunks.reserve(chunks.size());
            job.width = width;
            evs_pk.assign(chunks.size(), nullptr); evs_dh.assign(chunks.size(), nullptr);
            g_widen.begin(&job);
        }
    }
    // document offsets relative to their chunk
    for (const Chunk& c : chunks)
        for (size_t d = c.d0; d <= c.d1; ++d) h_roff[c.roff + (d - c.d0)] = text_off[d] - c.b0;
    CKE(cudaMemcpyAsync(d_roff, h_roff, n_roff * 8, cudaMemcpyHostToDevice, eng->st_h2d));
    evs.resize(chunks.size());
    if (trace) {
        tr_h.resize(chunks.size()); tr_c.resize(chunks.size()); tr_d.resize(chunks.size()); tr_host.resize(chunks.size());
        cudaEventCreate(&tr0);
        for (size_t c = 0; c < chunks.size(); ++c) { cudaEventCreate(&tr_h[c]); cudaEventCreate(&tr_c[c]); cudaEventCreate(&tr_d[c]); }
        cudaEventRecord(tr0, eng->st_h2d);
    }
    for (size_t c = 0; c < chunks.size(); ++c) {
        if (!eng->sync_ev_pool.empty()) { evs[c] = eng->sync_ev_pool.back(); eng->sync_ev_pool.pop_back(); }
        else CKE(cudaEventCreateWithFlags(&evs[c], cudaEventDisableTiming));
    }
    {
        const int hw = (int)std::thread::hardware_concurrency();
        int T = std::max(2, std::min(8, hw / 3));                     // measured (16 cores): 2 -> 16.6, 4 -> 25.1, 8 -> 32.8 GB/s
        if (const char* e = getenv("CTK_STAGE_THREADS")) T = atoi(e);
        staged = T > 0 && B >= (8ull << 20) && is_pageable(text);
        if (staged) {
            uint64_t biggest = 0;
            for (const Chunk& ch : chunks) biggest = std::max<uint64_t>(biggest, ch.b1 - ch.b0);
            for (int k = 0; k < NSTAGE; ++k) {
                CKE(g_pinned.get(biggest + 64, &stage_buf[k], &stage_cap[k]));
                CKE(cudaEventCreateWithFlags(&stage_free[k], cudaEventDisableTiming));
            }
            g_copy_pool.start(T - 1);
        }
    }
    if (staged) {
        stager = std::thread([&, T = (int)g_copy_pool.threads.size() + 1] {
            if (cudaSetDevice(eng->device) != cudaSuccess) { stager_err.store(1); h2d_issued.store(chunks.size()); return; }
            for (size_t c = 0; c < chunks.size(); ++c) {
                const Chunk& ch = chunks[c];
                const int k = (int)(c % NSTAGE);
                const uint64_t len = ch.b1 - ch.b0;
                cudaError_t e = cudaSuccess;
                if (c >= (size_t)NSTAGE) e = cudaEventSynchronize(stage_free[k]);       // the DMA that last read this buffer is done
                if (e == cudaSuccess && len) {
                    const int parts = (int)std::min<uint64_t>((uint64_t)T * 4, (len + (1u << 20) - 1) >> 20);
                    const uint64_t per = (len + parts - 1) / parts;
                    char* dst = (char*)stage_buf[k]; const char* src = (const char*)text + ch.b0;
                    g_copy_pool.run(parts, [=](int i) {
                        const uint64_t lo = per * (uint64_t)i, hi = std::min<uint64_t>(len, lo + per);
                        if (lo < hi) memcpy(dst + lo, src + lo, hi - lo);
                    });
                    e = cudaMemcpyAsync(d_text + ch.dev_text, stage_buf[k], len, cudaMemcpyHostToDevice, eng->st_h2d);
                }
                if (e == cudaSuccess) e = cudaEventRecord(stage_free[k], eng->st_h2d);
                if (e == cudaSuccess) e = cudaMemsetAsync(d_text + ch.dev_text + len, 0, 64, eng->st_h2d);
                if (e == cudaSuccess) e = cudaEventRecord(evs[c], eng->st_h2d);
                if (e == cudaSuccess && trace) cudaEventRecord(tr_h[c], eng->st_h2d);
                if (e != cudaSuccess) { stager_err.store((int)e); h2d_issued.store(chunks.size(), std::memory_order_release); return; }
                h2d_issued.store(c + 1, std::memory_order_release);
            }
        });
    } else {
        for (size_t c = 0; c < chunks.size(); ++c) {
            const Chunk& ch = chunks[c];
            if (ch.b1 > ch.b0) CKE(copy_host_aligned(d_text + ch.dev_text, text + ch.b0, ch.b1 - ch.b0, cudaMemcpyHostToDevice, eng->st_h2d));
            CKE(cudaMemsetAsync(d_text + ch.dev_text + (ch.b1 - ch.b0), 0, 64, eng->st_h2d));
            CKE(cudaEventRecord(evs[c], eng->st_h2d));
            if (trace) cudaEventRecord(tr_h[c], eng->st_h2d);
        }
        h2d_issued.store(chunks.size());
    }
    chunk_base.resize(chunks.size() + 1, 0);
    for (size_t c = 0; c < chunks.size(); ++c) {
        const Chunk& ch = chunks[c];
        while (h2d_issued.load(std::memory_order_acquire) <= c) std::this_thread::yield();   // the copy in of this chunk is enqueued
        if (stager_err.load()) { rc = eng->cuda_fail((cudaError_t)stager_err.load(), "staged copy in"); goto done; }
        CKE(cudaStreamWaitEvent(eng->st_comp, evs[c], 0));
        uint64_t cnt = 0;
        eng->keep_cache_once = c > 0;
        rc = encode_device(*eng, d_text + ch.dev_text, d_roff + ch.roff, ch.d1 - ch.d0, ch.b1 - ch.b0, d_ids + total, ids_cap - total,
                           d_ids_off + ch.roff, &cnt, eng->st_comp);
        eng->keep_cache_once = false;
        if (rc != CTK_OK) goto done;
        if (trace) { cudaEventRecord(tr_c[c], eng->st_comp); tr_host[c] = now_ms() - host0; }
        // copy out while the next chunk is being encoded
        if ((total + cnt + 1) * 4 > r->ids_cap) {                     // grow the pinned result (rare)
            if (packed) g_widen.drain(&job);                           // everything copied so far is widened into r->ids
            CKE(cudaStreamSynchronize(eng->st_d2h));
            void* nb; size_t ncap;
            uint64_t est = (uint64_t)((double)(total + cnt) * (double)B / (double)std::max<uint64_t>(ch.b1, 1) * 1.1) + 4096;
            CKE(g_pinned.get(std::max<uint64_t>(est, total + cnt + 1) * 4, &nb, &ncap));
            memcpy(nb, r->ids, total * 4);
            g_pinned.put(r->ids, r->ids_cap);
            r->ids = nb; r->ids_cap = ncap;
            if (packed) {                                              // the staging area follows; what it held is consumed
                g_pinned.put(h_pack, h_pack_cap);
                h_pack = nullptr;
                CKE(g_pinned.get((uint64_t)width * (r->ids_cap / 4) + 256 * (chunks.size() + 2), (void**)&h_pack, &h_pack_cap));
                pk_off = 0;
            }
        }
        if (cnt && packed) {
            const uint64_t pbytes = ((cnt * (uint64_t)width + 3) / 4) * 4;
            for (std::vector<cudaEvent_t>* v : {&evs_pk, &evs_dh}) {
                if (!eng->sync_ev_pool.empty()) { (*v)[c] = eng->sync_ev_pool.back(); eng->sync_ev_pool.pop_back(); }
                else CKE(cudaEventCreateWithFlags(&(*v)[c], cudaEventDisableTiming));
            }
            const uint64_t units = (cnt + (width == 2 ? 1 : 3)) / (width == 2 ? 2 : 4);
            const unsigned grid = (unsigned)std::min<uint64_t>((units + 255) / 256, 148 * 16);
            if (width == 2) k_pack_ids16<<<grid, 256, 0, eng->st_comp>>>(d_ids + total, cnt, reinterpret_cast<uint32_t*>(d_pack + pk_off));
            else k_pack_ids24<<<grid, 256, 0, eng->st_comp>>>(d_ids + total, cnt, reinterpret_cast<uint32_t*>(d_pack + pk_off));
            eng->launched(1);
            CKE(cudaEventRecord(evs_pk[c], eng->st_comp));
            CKE(cudaStreamWaitEvent(eng->st_d2h, evs_pk[c], 0));
            CKE(copy_host_aligned(h_pack + pk_off, d_pack + pk_off, pbytes, cudaMemcpyDeviceToHost, eng->st_d2h));
            CKE(cudaEventRecord(evs_dh[c], eng->st_d2h));
            job.chunks.push_back({evs_dh[c], h_pack + pk_off, (uint32_t*)r->ids + total, cnt});
            job.published.fetch_add(1, std::memory_order_release);
            pk_off += ((pbytes + 255) / 256) * 256;
            d2h_bytes += pbytes;
        } else if (cnt && !getenv("CTK_DIAG_NO_D2H")) {
            CKE(copy_host_aligned((uint32_t*)r->ids + total, d_ids + total, cnt * 4, cudaMemcpyDeviceToHost, eng->st_d2h));
            d2h_bytes += cnt * 4;
        }
        d2h_bytes += (ch.d1 - ch.d0 + 1) * 8;
        CKE(cudaMemcpyAsync(h_ioff + ch.roff, d_ids_off + ch.roff, (ch.d1 - ch.d0 + 1) * 8, cudaMemcpyDeviceToHost, eng->st_d2h));
        if (trace) cudaEventRecord(tr_d[c], eng->st_d2h);
        chunk_base[c] = total;
        total += cnt;
    }
    if (packed) { g_widen.end(&job); packed = false; }
    CKE(cudaStreamSynchronize(eng->st_d2h));
    h2d_bytes = B + n_roff * 8;
    eng->last_h2d_bytes = h2d_bytes; eng->last_d2h_bytes = d2h_bytes;
    if (trace) {
        fprintf(stderr, "[ctk trace] ids cross the link as %d bytes each\n", width);
        fprintf(stderr, "[ctk trace] %zu chunks, %.1f MiB in; host: pipeline issued+drained at %.3f ms\n", chunks.size(), B / 1048576.0, now_ms() - host0);
        for (size_t c = 0; c < chunks.size(); ++c) {
            float h = 0, k = 0, d = 0;
            cudaEventElapsedTime(&h, tr0, tr_h[c]); cudaEventElapsedTime(&k, tr0, tr_c[c]); cudaEventElapsedTime(&d, tr0, tr_d[c]);
            fprintf(stderr, "[ctk trace] chunk %2zu  h2d done %8.3f  kernels done %8.3f  d2h done %8.3f  (host saw kernels done at %8.3f)\n", c, h, k, d, tr_host[c]);
            cudaEventDestroy(tr_h[c]); cudaEventDestroy(tr_c[c]); cudaEventDestroy(tr_d[c]);
        }
        cudaEventDestroy(tr0);
    }
    {
        uint64_t* off = (uint64_t*)r->off;
        for (size_t c = 0; c < chunks.size(); ++c) {
            const Chunk& ch = chunks[c];
            for (size_t d = ch.d0; d < ch.d1; ++d) off[d] = chunk_base[c] + h_ioff[ch.roff + (d - ch.d0)];
        }
        off[n] = total;
    }
done:
    if (stager.joinable()) stager.join();                              // it only enqueues; nothing it waits on depends on this thread
    if (staged) {
        cudaStreamSynchronize(eng->st_h2d);                            // staging buffers go back to the pool only when no DMA reads them
        for (int k = 0; k < NSTAGE; ++k) { g_pinned.put(stage_buf[k], stage_cap[k]); if (stage_free[k]) cudaEventDestroy(stage_free[k]); }
    }
    if (packed) g_widen.end(&job);                                     // error path: the workers must let go of `job`
    for (cudaEvent_t ev : evs) if (ev) eng->sync_ev_pool.push_back(ev);
    for (cudaEvent_t ev : evs_pk) if (ev) eng->sync_ev_pool.push_back(ev);
    for (cudaEvent_t ev : evs_dh) if (ev) eng->sync_ev_pool.push_back(ev);
    g_pinned.put(h_pack, h_pack_cap);
    g_pinned.put(h_roff, h_roff_cap);
    g_pinned.put(h_ioff, h_ioff_cap);
    if (rc != CTK_OK) { cudaStreamSynchronize(eng->st_h2d); cudaStreamSynchronize(eng->st_d2h); free_result(r); return rc; }
    *res = reinterpret_cast<ctk_result*>(r);
    return CTK_OK;
}

int ctk_decode_batch(const ctk_tokenizer* tok, const uint32_t* ids, const uint64_t* ids_off, size_t n, int skip_special_tokens,
                     int clean_up_tokenization_spaces, ctk_result** res) {
    if (!tok || !ids_off || !res) { set_last_error("NULL argument"); return CTK_ERR_ARG; }
    *res = nullptr;
    Engine* eng = const_cast<Engine*>(reinterpret_cast<const Engine*>(tok));
    if (ids_off[0] != 0) { set_last_error("ids_off[0] must be 0"); return CTK_ERR_ARG; }
    for (size_t i = 0; i < n; ++i) if (ids_off[i + 1] < ids_off[i]) { set_last_error("ids_off must be non-decreasing"); return CTK_ERR_ARG; }
    const uint64_t T = ids_off[n];
    if (T && !ids) { set_last_error("NULL ids"); return CTK_ERR_ARG; }
    std::lock_guard<std::mutex> lk(eng->mu);
    int rc = CTK_OK;
    Result* r = new Result();
    r->n = n;
    uint32_t* d_ids; uint64_t *d_off, *d_out_off; uint8_t* d_out;
    uint64_t total = 0;
    cudaStream_t st = eng->st_comp;
    CKE(cudaSetDevice(eng->device));
    CKE(eng->ws.get(24, (T + 1) * 4, (void**)&d_ids));
    CKE(eng->ws.get(21, (n + 1) * 8, (void**)&d_off));
    CKE(eng->ws.get(22, (n + 1) * 8, (void**)&d_out_off));
    if (T) CKE(cudaMemcpyAsync(d_ids, ids, T * 4, cudaMemcpyHostToDevice, st));
    CKE(cudaMemcpyAsync(d_off, ids_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    rc = decode_device(*eng, d_ids, d_off, n, T, skip_special_tokens, clean_up_tokenization_spaces, nullptr, 0, d_out_off, &total, st);
    if (rc != CTK_OK) goto done;
    d_out = eng->last_decode_out;
    CKE(g_pinned.get((n + 1) * 8, &r->off, &r->off_cap));
    CKE(g_pinned.get(total + 1, &r->bytes, &r->bytes_cap));
    CKE(cudaMemcpyAsync(r->off, d_out_off, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (total) CKE(cudaMemcpyAsync(r->bytes, d_out, total, cudaMemcpyDeviceToHost, st));
    CKE(cudaStreamSynchronize(st));
done:
    if (rc != CTK_OK) { free_result(r); return rc; }
    *res = reinterpret_cast<ctk_result*>(r);
    return CTK_OK;
}

const uint32_t* ctk_result_ids(const ctk_result* res) { return (const uint32_t*)reinterpret_cast<const Result*>(res)->ids; }
const uint64_t* ctk_result_offsets(const ctk_result* res) { return (const uint64_t*)reinterpret_cast<const Result*>(res)->off; }
const uint8_t* ctk_result_bytes(const ctk_result* res) { return (const uint8_t*)reinterpret_cast<const Result*>(res)->bytes; }
// debug/test hook (host only): the worker pool that stages pageable input, as a parallel memcpy
int ctk_debug_parallel_copy(void* dst, const void* src, size_t bytes, int threads, int parts) {
    if (threads < 1 || parts < 1) return CTK_ERR_ARG;
    g_copy_pool.start(threads - 1);
    const size_t per = (bytes + parts - 1) / parts;
    g_copy_pool.run(parts, [=](int i) {
        const size_t lo = per * (size_t)i, hi = std::min(bytes, lo + per);
        if (lo < hi) memcpy((char*)dst + lo, (const char*)src + lo, hi - lo);
    });
    return CTK_OK;
}

void ctk_last_transfer_bytes(const ctk_tokenizer* tok, uint64_t* h2d, uint64_t* d2h) {
    const Engine* eng = reinterpret_cast<const Engine*>(tok);
    if (h2d) *h2d = eng->last_h2d_bytes;
    if (d2h) *d2h = eng->last_d2h_bytes;
}
size_t ctk_result_count(const ctk_result* res) { return reinterpret_cast<const Result*>(res)->n; }
void ctk_result_free(ctk_result* res) { free_result(reinterpret_cast<Result*>(res)); }

}  // extern "C"
