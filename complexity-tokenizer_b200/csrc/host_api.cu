// Host-buffer entry points of the C ABI: ctk_encode_batch / ctk_encode_batch_narrow / ctk_decode_batch.
//
// These are what the reference's PyO3 methods would call (bindings/tokenizer.rs:207-210, 226-238).
//
// One device: encode pipelines the batch in chunks of whole documents (4 .. 64 MiB) over three streams so that the PCIe
// copy in, the kernels and the copy out overlap:
//     st_h2d : text chunk c+1, c+2, ...      (pinned user memory is DMA'd directly)
//     st_comp: NFC check + fused encode of chunk c
//     st_d2h : ids of chunk c-1 into a pooled pinned result buffer
// The pre-token cache is cleared at the first chunk only: the chunks are one batch.
//
// Several devices (ctk_from_file_devices): the reference's encode_batch is ONE call that uses the whole machine
// (mod.rs:694-696, rayon over documents).  Here the documents are cut into contiguous ranges balanced by bytes, one per
// device; every device runs the single-device pipeline on its range from a host thread of its own, bound to the CPUs of
// the device's NUMA node, with page-locked buffers allocated on that node.  No data crosses between devices.  The result
// is one ctk_result made of per-device PARTS (zero-copy accessors below); ctk_result_ids() gathers them on demand.
//
// Narrow ids: when every id the tokenizer can emit is below 65 536 the device writes uint16 ids and the copy out is half
// as large (ctk_encode_batch_narrow).  Nothing is widened on the host: the binding's per-document copy into its own
// Vec<u32> / list[int] reads the 16-bit ids directly.  ctk_encode_batch keeps its uint32 contract.
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <thread>
#include <vector>

#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "engine.hpp"

namespace ctk {

// ---- NUMA placement --------------------------------------------------------------------------------------------
// No libnuma in the image: the node of a device comes from sysfs (PCI bus id -> numa_node), memory placement is the
// set_mempolicy system call around the page-locking allocation, thread placement is sched_setaffinity.
int device_numa_node(int device) {
    static std::mutex mu;
    static std::map<int, int> memo;
    std::lock_guard<std::mutex> lk(mu);
    auto it = memo.find(device);
    if (it != memo.end()) return it->second;
    int node = -1;
    char bus[64] = {0};
    if (!getenv("CTK_NO_NUMA") && cudaDeviceGetPCIBusId(bus, sizeof bus, device) == cudaSuccess) {
        for (char* c = bus; *c; ++c) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
        std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
        if (FILE* f = fopen(path.c_str(), "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    } else cudaGetLastError();
    memo[device] = node;
    return node;
}

static bool node_cpus(int node, cpu_set_t* set) {
    CPU_ZERO(set);
    if (node < 0) return false;
    char path[128];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    char buf[4096] = {0};
    const bool got = fgets(buf, sizeof buf, f) != nullptr;
    fclose(f);
    if (!got) return false;
    int any = 0;
    for (char* p = buf; *p;) {                               // "0-15,32-47"
        char* e;
        long a = strtol(p, &e, 10);
        if (e == p) break;
        long b = a;
        if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, set); ++any; }
        p = *e == ',' ? e + 1 : e;
        if (*e != ',') break;
    }
    return any > 0;
}

// binds the CALLING thread to the CPUs of `node` (only threads this library owns are ever bound)
static void bind_thread_to_node(int node) {
    cpu_set_t set;
    if (node_cpus(node, &set)) sched_setaffinity(0, sizeof set, &set);
}

struct ScopedMemPolicy {                                    // page-locked allocations inside the scope land on `node`
    bool on = false;
    explicit ScopedMemPolicy(int node) {
#ifdef SYS_set_mempolicy
        if (node >= 0 && node < 1024) {
            unsigned long mask[16] = {0};
            mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
            on = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, 1024ul) == 0;
        }
#endif
    }
    ~ScopedMemPolicy() {
#ifdef SYS_set_mempolicy
        if (on) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
#endif
    }
};

// ---- pools of pinned host buffers, one per NUMA node (pinning a GiB costs far more than encoding it) --------------
struct PinnedPool {
    struct Buf { void* p; size_t cap; };
    int node = -1;
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
        ScopedMemPolicy pol(node);
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(out, want, cudaHostAllocPortable);
        if (e != cudaSuccess) { cudaGetLastError(); want = bytes + 64; e = cudaHostAlloc(out, want, cudaHostAllocPortable); }
        *cap = want;
        return e;
    }
    void put(void* p, size_t cap) {
        if (!p) return;
        std::lock_guard<std::mutex> lk(mu);
        if (free_.size() >= 64) {                        // drop the smallest (8 was too few: a 4-device result holds 8-12 buffers, and
                                                         // re-pinning them every call cost 550 ms per 4 GiB batch, measured)
            int s = 0;
            for (int i = 1; i < (int)free_.size(); ++i) if (free_[i].cap < free_[s].cap) s = i;
            if (free_[s].cap < cap) { cudaFreeHost(free_[s].p); free_[s] = {p, cap}; } else cudaFreeHost(p);
            return;
        }
        free_.push_back({p, cap});
    }
};
static PinnedPool* pool_for_node(int node) {
    static std::mutex mu;
    static std::map<int, PinnedPool*> pools;               // never destroyed: buffers may outlive static destruction order
    std::lock_guard<std::mutex> lk(mu);
    PinnedPool*& p = pools[node];
    if (!p) { p = new PinnedPool(); p->node = node; }
    return p;
}
cudaError_t pinned_get(size_t bytes, void** out, size_t* cap) { return pool_for_node(-1)->get(bytes, out, cap); }
void pinned_put(void* p, size_t cap) { pool_for_node(-1)->put(p, cap); }

struct PBuf {                                               // a pooled pinned buffer that remembers where it came from
    void* p = nullptr; size_t cap = 0; PinnedPool* pool = nullptr;
    cudaError_t get(PinnedPool* from, size_t bytes) { release(); pool = from; return from->get(bytes, &p, &cap); }
    void release() { if (p && pool) pool->put(p, cap); p = nullptr; cap = 0; }
};

// One ctk_result.  Single device: ids/off/bytes.  Several devices: `parts` (one per device that got documents), the
// flat views are built on demand.
struct Result {
    size_t n = 0;
    int id_width = 4;
    PBuf ids, off, bytes;
    uint64_t total = 0;                                     // ids (encode) or bytes (decode)
    std::vector<Result*> parts;
    std::vector<size_t> part_first;                         // first document of each part
    std::mutex lazy_mu;
    std::vector<uint32_t> wide;                             // ctk_result_ids of a narrow / multi-part result
    std::vector<uint64_t> flat_off;
    std::vector<uint8_t> flat_bytes;
    bool have_wide = false, have_off = false, have_bytes = false;
};

static void free_result(Result* r) {
    if (!r) return;
    for (Result* q : r->parts) free_result(q);
    r->ids.release(); r->off.release(); r->bytes.release();
    delete r;
}

// ---- a few persistent host threads: staging of pageable input, gathering / widening of results -----------------------
struct TaskPool {                                   // run(n, fn) = fn(0) .. fn(n-1), caller helps, returns when done
    struct Job { std::function<void(int)> fn; int n = 0; std::atomic<int> next{0}, done{0}; };
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::thread> threads;
    std::shared_ptr<Job> job;
    uint64_t gen = 0;
    bool stop = false;
    int node = -1;
    void start(int n) {
        std::lock_guard<std::mutex> lk(mu);
        while ((int)threads.size() < n) threads.emplace_back([this] { bind_thread_to_node(node); loop(); });
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
static TaskPool* copy_pool_for_node(int node) {
    static std::mutex mu;
    static std::map<int, TaskPool*> pools;
    std::lock_guard<std::mutex> lk(mu);
    TaskPool*& p = pools[node];
    if (!p) { p = new TaskPool(); p->node = node; }
    return p;
}
static int default_copy_threads() {
    const int hw = (int)std::thread::hardware_concurrency();
    int T = std::max(2, std::min(8, hw / 3));                         // measured (16 cores): 2 -> 16.6, 4 -> 25.1, 8 -> 32.8 GB/s
    if (const char* e = getenv("CTK_STAGE_THREADS")) T = atoi(e);
    return T;
}
static void parallel_for_bytes(TaskPool* pool, int T, uint64_t total, std::function<void(uint64_t, uint64_t)> fn) {
    if (total == 0) return;
    const int parts = (int)std::min<uint64_t>((uint64_t)std::max(1, T) * 4, (total + (1u << 20) - 1) >> 20);
    if (parts <= 1 || T <= 1) { fn(0, total); return; }
    const uint64_t per = (total + parts - 1) / parts;
    pool->start(T - 1);
    pool->run(parts, [=](int i) { const uint64_t lo = per * (uint64_t)i, hi = std::min(total, lo + per); if (lo < hi) fn(lo, hi); });
}

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

#define CKE(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = eng->cuda_fail(e_, #x); goto done; } } while (0)

constexpr int RC_RETRY_CAPACITY = 101;                      // internal: the id buffer was too small (NFC grew the text); run again with the safe bound

// ---- encode, one device ---------------------------------------------------------------------------------------------
// text_off[0] may be > 0 (a shard of a larger batch): bytes text[text_off[0] .. text_off[n]).  Result offsets start at 0.
static int encode_single_try(Engine* eng, const uint8_t* text, const uint64_t* text_off, size_t n, bool narrow, bool big_cap, Result** res) {
    *res = nullptr;
    const uint64_t base0 = text_off[0];
    const uint64_t B = text_off[n] - base0;
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
            if (base0 + B - b0 < 2 * target) target = std::max<uint64_t>(small, (base0 + B - b0) / 2);
            while (e < n && (e == d || text_off[e + 1] - b0 <= target)) {
                if (text_off[e + 1] < text_off[e]) { set_last_error("text_off must be non-decreasing"); return CTK_ERR_ARG; }
                ++e;
            }
            uint64_t b1 = text_off[e];
            if (b1 - b0 >= 0xFFFFF000ull) { set_last_error("a single document of 4 GiB or more is not supported"); return CTK_ERR_ARG; }
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
    const int width = narrow && eng->run_width == 2 && !eng->use_general ? 2 : 4;   // (the debug multi-kernel pipeline only writes uint32)
    r->id_width = width;
    PinnedPool* const pool = pool_for_node(eng->numa_node);
    uint8_t* d_text = nullptr; uint64_t *d_roff = nullptr, *d_ids_off = nullptr; uint8_t* d_ids = nullptr;
    PBuf h_roff, h_ioff;
    uint64_t total = 0, dev_text_bytes = 0, ids_cap = 0;
    size_t n_roff = n + chunks.size() + 1;
    std::vector<cudaEvent_t> evs;
    std::vector<uint64_t> chunk_base;
    uint64_t d2h_bytes = 0;
    // pageable input: staging buffers, the helper thread that fills them and issues the copies, how far it got
    constexpr int NSTAGE = 3;
    PBuf stage_buf[NSTAGE];
    cudaEvent_t stage_free[NSTAGE] = {nullptr, nullptr, nullptr};
    std::thread stager;
    std::atomic<size_t> h2d_issued{0};
    std::atomic<int> stager_err{0};
    bool staged = false;
    TaskPool* const cpool = copy_pool_for_node(eng->numa_node);
    const int copyT = default_copy_threads();
    // CTK_TRACE=1: per-chunk timeline (H2D done, kernels done, D2H done; ms since the call started) on stderr
    const bool trace = getenv("CTK_TRACE") != nullptr;
    std::vector<cudaEvent_t> tr_h, tr_c, tr_d;
    std::vector<double> tr_host;
    cudaEvent_t tr0 = nullptr;
    auto now_ms = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double host0 = now_ms();
    CKE(cudaSetDevice(eng->device));
    dev_text_bytes = chunks.empty() ? 256 : chunks.back().dev_text + ((chunks.back().b1 - chunks.back().b0 + 64 + 255) / 256) * 256;
    // at most one id per byte of the NORMALISED text: NFC leaves almost all text as it is (first try) and never more than triples it
    ids_cap = (big_cap ? 3 * B : B + B / 8) + n + 1024;
    CKE(eng->ws.get(20, dev_text_bytes, (void**)&d_text));
    CKE(eng->ws.get(21, n_roff * 8, (void**)&d_roff));
    CKE(eng->ws.get(22, n_roff * 8, (void**)&d_ids_off));
    CKE(eng->ws.get(23, ids_cap * (uint64_t)width, (void**)&d_ids));
    CKE(h_roff.get(pool, n_roff * 8));
    CKE(h_ioff.get(pool, n_roff * 8));
    CKE(r->off.get(pool, (n + 1) * 8));
    CKE(r->ids.get(pool, (B / 3 + n + 1024) * (uint64_t)width));
    // document offsets relative to their chunk
    for (const Chunk& c : chunks)
        for (size_t d = c.d0; d <= c.d1; ++d) ((uint64_t*)h_roff.p)[c.roff + (d - c.d0)] = text_off[d] - c.b0;
    CKE(cudaMemcpyAsync(d_roff, h_roff.p, n_roff * 8, cudaMemcpyHostToDevice, eng->st_h2d));
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
    staged = copyT > 0 && B >= (8ull << 20) && is_pageable(text + base0);
    if (staged) {
        uint64_t biggest = 0;
        for (const Chunk& ch : chunks) biggest = std::max<uint64_t>(biggest, ch.b1 - ch.b0);
        for (int k = 0; k < NSTAGE; ++k) {
            CKE(stage_buf[k].get(pool, biggest + 64));
            CKE(cudaEventCreateWithFlags(&stage_free[k], cudaEventDisableTiming));
        }
        stager = std::thread([&] {
            bind_thread_to_node(eng->numa_node);
            if (cudaSetDevice(eng->device) != cudaSuccess) { stager_err.store(1); h2d_issued.store(chunks.size()); return; }
            for (size_t c = 0; c < chunks.size(); ++c) {
                const Chunk& ch = chunks[c];
                const int k = (int)(c % NSTAGE);
                const uint64_t len = ch.b1 - ch.b0;
                cudaError_t e = cudaSuccess;
                if (c >= (size_t)NSTAGE) e = cudaEventSynchronize(stage_free[k]);       // the DMA that last read this buffer is done
                if (e == cudaSuccess && len) {
                    char* dst = (char*)stage_buf[k].p; const char* src = (const char*)text + ch.b0;
                    parallel_for_bytes(cpool, copyT, len, [=](uint64_t lo, uint64_t hi) { memcpy(dst + lo, src + lo, hi - lo); });
                    e = cudaMemcpyAsync(d_text + ch.dev_text, stage_buf[k].p, len, cudaMemcpyHostToDevice, eng->st_h2d);
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
        eng->out_id_width = width;
        rc = encode_device(*eng, d_text + ch.dev_text, d_roff + ch.roff, ch.d1 - ch.d0, ch.b1 - ch.b0,
                           reinterpret_cast<uint32_t*>(d_ids + total * (uint64_t)width), ids_cap - total, d_ids_off + ch.roff, &cnt, eng->st_comp);
        eng->out_id_width = 4;
        eng->keep_cache_once = false;
        if (rc != CTK_OK) {
            if (!big_cap && (eng->h_flags[0] & ERRF_CAPACITY)) rc = RC_RETRY_CAPACITY;
            goto done;
        }
        if (trace) { cudaEventRecord(tr_c[c], eng->st_comp); tr_host[c] = now_ms() - host0; }
        // copy out while the next chunk is being encoded
        if ((total + cnt + 1) * (uint64_t)width > r->ids.cap) {           // grow the pinned result (rare)
            CKE(cudaStreamSynchronize(eng->st_d2h));
            PBuf nb;
            uint64_t est = (uint64_t)((double)(total + cnt) * (double)B / (double)std::max<uint64_t>(ch.b1 - base0, 1) * 1.1) + 4096;
            CKE(nb.get(pool, std::max<uint64_t>(est, total + cnt + 1) * (uint64_t)width));
            memcpy(nb.p, r->ids.p, total * (uint64_t)width);
            r->ids.release();
            r->ids = nb;
        }
        if (cnt && !getenv("CTK_DIAG_NO_D2H")) {
            CKE(copy_host_aligned((uint8_t*)r->ids.p + total * (uint64_t)width, d_ids + total * (uint64_t)width, cnt * (uint64_t)width,
                                  cudaMemcpyDeviceToHost, eng->st_d2h));
            d2h_bytes += cnt * (uint64_t)width;
        }
        d2h_bytes += (ch.d1 - ch.d0 + 1) * 8;
        CKE(cudaMemcpyAsync((uint64_t*)h_ioff.p + ch.roff, d_ids_off + ch.roff, (ch.d1 - ch.d0 + 1) * 8, cudaMemcpyDeviceToHost, eng->st_d2h));
        if (trace) cudaEventRecord(tr_d[c], eng->st_d2h);
        chunk_base[c] = total;
        total += cnt;
    }
    CKE(cudaStreamSynchronize(eng->st_d2h));
    eng->last_h2d_bytes = B + n_roff * 8; eng->last_d2h_bytes = d2h_bytes;
    if (trace) {
        fprintf(stderr, "[ctk trace] device %d (NUMA node %d): ids cross the link as %d bytes each\n", eng->device, eng->numa_node, width);
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
        uint64_t* off = (uint64_t*)r->off.p;
        const uint64_t* io = (const uint64_t*)h_ioff.p;
        for (size_t c = 0; c < chunks.size(); ++c) {
            const Chunk& ch = chunks[c];
            for (size_t d = ch.d0; d < ch.d1; ++d) off[d] = chunk_base[c] + io[ch.roff + (d - ch.d0)];
        }
        off[n] = total;
        r->total = total;
    }
done:
    if (stager.joinable()) stager.join();                              // it only enqueues; nothing it waits on depends on this thread
    if (staged) {
        cudaStreamSynchronize(eng->st_h2d);                            // staging buffers go back to the pool only when no DMA reads them
        for (int k = 0; k < NSTAGE; ++k) { stage_buf[k].release(); if (stage_free[k]) cudaEventDestroy(stage_free[k]); }
    }
    for (cudaEvent_t ev : evs) if (ev) eng->sync_ev_pool.push_back(ev);
    if (rc != CTK_OK) { cudaStreamSynchronize(eng->st_h2d); cudaStreamSynchronize(eng->st_comp); cudaStreamSynchronize(eng->st_d2h); }
    h_roff.release();
    h_ioff.release();
    if (rc != CTK_OK) { free_result(r); return rc; }
    *res = r;
    return CTK_OK;
}

static int encode_single(Engine* eng, const uint8_t* text, const uint64_t* text_off, size_t n, bool narrow, Result** res) {
    int rc = encode_single_try(eng, text, text_off, n, narrow, false, res);
    if (rc == RC_RETRY_CAPACITY) rc = encode_single_try(eng, text, text_off, n, narrow, true, res);
    return rc;
}

// ---- several devices: contiguous document ranges balanced by bytes ---------------------------------------------------
// cut[g] .. cut[g+1] = documents of device g.  `off` may be text or id offsets.
static std::vector<size_t> balanced_cuts(const uint64_t* off, size_t n, size_t G) {
    std::vector<size_t> cut(G + 1, n);
    cut[0] = 0;
    const uint64_t base = off[0], total = off[n] - base;
    for (size_t g = 1; g < G; ++g) {
        const uint64_t want = base + (uint64_t)((double)total * (double)g / (double)G);
        size_t d = (size_t)(std::lower_bound(off, off + n + 1, want) - off);
        if (d > n) d = n;
        cut[g] = std::max(d, cut[g - 1]);
    }
    cut[G] = n;
    return cut;
}

template <typename F>
static int run_on_peers(Engine* eng, const std::vector<size_t>& cut, Result* merged, F&& one) {
    const size_t G = eng->peers.size();
    std::vector<int> rcs(G, CTK_OK);
    std::vector<std::string> errs(G);
    std::vector<Result*> parts(G, nullptr);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; ++g) {
        if (cut[g + 1] == cut[g]) continue;
        th.emplace_back([&, g] {
            bind_thread_to_node(eng->peers[g]->numa_node);
            rcs[g] = one(eng->peers[g], cut[g], cut[g + 1] - cut[g], &parts[g]);
            if (rcs[g] != CTK_OK) errs[g] = ctk_last_error();
        });
    }
    for (auto& t : th) t.join();
    int rc = CTK_OK;
    for (size_t g = 0; g < G; ++g) if (rcs[g] != CTK_OK && rc == CTK_OK) { rc = rcs[g]; set_last_error(errs[g]); }
    for (size_t g = 0; g < G; ++g) {
        if (!parts[g]) continue;
        if (rc != CTK_OK) { free_result(parts[g]); continue; }
        merged->parts.push_back(parts[g]);
        merged->part_first.push_back(cut[g]);
        merged->total += parts[g]->total;
    }
    return rc;
}

static uint64_t multi_min_bytes() {
    if (const char* e = getenv("CTK_MULTI_MIN_MB")) return (uint64_t)atol(e) << 20;
    return 32ull << 20;                                     // below this one device is faster than starting threads on several
}

static int encode_host(const ctk_tokenizer* tok, const uint8_t* text, const uint64_t* text_off, size_t n, bool narrow, ctk_result** res) {
    if (!tok || !text_off || !res) { set_last_error("NULL argument"); return CTK_ERR_ARG; }
    *res = nullptr;
    Engine* eng = const_cast<Engine*>(reinterpret_cast<const Engine*>(tok));
    if (text_off[0] != 0) { set_last_error("text_off[0] must be 0"); return CTK_ERR_ARG; }
    const uint64_t B = text_off[n];
    if (B && !text) { set_last_error("NULL text"); return CTK_ERR_ARG; }
    Result* r = nullptr;
    int rc;
    if (eng->peers.size() > 1 && B >= multi_min_bytes()) {
        for (size_t i = 0; i < n; ++i) if (text_off[i + 1] < text_off[i]) { set_last_error("text_off must be non-decreasing"); return CTK_ERR_ARG; }
        r = new Result();
        r->n = n;
        r->id_width = narrow && eng->run_width == 2 && !eng->use_general ? 2 : 4;
        const std::vector<size_t> cut = balanced_cuts(text_off, n, eng->peers.size());
        rc = run_on_peers(eng, cut, r, [&](Engine* pe, size_t d0, size_t nd, Result** out) { return encode_single(pe, text, text_off + d0, nd, narrow, out); });
        if (rc != CTK_OK) { free_result(r); r = nullptr; }
    } else {
        rc = encode_single(eng, text, text_off, n, narrow, &r);
    }
    if (rc != CTK_OK) return rc;
    *res = reinterpret_cast<ctk_result*>(r);
    return CTK_OK;
}

// ---- decode, one device ---------------------------------------------------------------------------------------------
static int decode_single(Engine* eng, const uint32_t* ids, const uint64_t* ids_off, size_t n, int skip_special_tokens,
                         int clean_up_tokenization_spaces, Result** res) {
    *res = nullptr;
    const uint64_t base0 = ids_off[0];
    const uint64_t T = ids_off[n] - base0;
    std::lock_guard<std::mutex> lk(eng->mu);
    int rc = CTK_OK;
    Result* r = new Result();
    r->n = n;
    PinnedPool* const pool = pool_for_node(eng->numa_node);
    uint32_t* d_ids; uint64_t *d_off, *d_out_off; uint8_t* d_out;
    uint64_t total = 0;
    PBuf h_off;
    cudaStream_t st = eng->st_comp;
    CKE(cudaSetDevice(eng->device));
    CKE(eng->ws.get(24, (T + 1) * 4, (void**)&d_ids));
    CKE(eng->ws.get(21, (n + 1) * 8, (void**)&d_off));
    CKE(eng->ws.get(22, (n + 1) * 8, (void**)&d_out_off));
    if (T) CKE(cudaMemcpyAsync(d_ids, ids + base0, T * 4, cudaMemcpyHostToDevice, st));
    if (base0) {                                              // a shard: offsets relative to its first id
        CKE(h_off.get(pool, (n + 1) * 8));
        for (size_t i = 0; i <= n; ++i) ((uint64_t*)h_off.p)[i] = ids_off[i] - base0;
        CKE(cudaMemcpyAsync(d_off, h_off.p, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    } else {
        CKE(cudaMemcpyAsync(d_off, ids_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    }
    rc = decode_device(*eng, d_ids, d_off, n, T, skip_special_tokens, clean_up_tokenization_spaces, nullptr, 0, d_out_off, &total, st);
    if (rc != CTK_OK) goto done;
    d_out = eng->last_decode_out;
    CKE(r->off.get(pool, (n + 1) * 8));
    CKE(r->bytes.get(pool, total + 1));
    CKE(cudaMemcpyAsync(r->off.p, d_out_off, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (total) CKE(cudaMemcpyAsync(r->bytes.p, d_out, total, cudaMemcpyDeviceToHost, st));
    CKE(cudaStreamSynchronize(st));
    r->total = total;
done:
    if (rc != CTK_OK) cudaStreamSynchronize(st);
    h_off.release();
    if (rc != CTK_OK) { free_result(r); return rc; }
    *res = r;
    return CTK_OK;
}

// flat views of a multi-part / narrow result, built on first use with the copy threads
static void build_flat_offsets(Result* r) {
    std::lock_guard<std::mutex> lk(r->lazy_mu);
    if (r->have_off) return;
    r->flat_off.resize(r->n + 1);
    uint64_t base = 0;
    for (size_t q = 0; q < r->parts.size(); ++q) {
        const Result* p = r->parts[q];
        const uint64_t* po = (const uint64_t*)p->off.p;
        uint64_t* dst = r->flat_off.data() + r->part_first[q];
        for (size_t d = 0; d < p->n; ++d) dst[d] = base + po[d];
        base += p->total;
    }
    r->flat_off[r->n] = base;
    r->have_off = true;
}

}  // namespace ctk

using namespace ctk;

extern "C" {

int ctk_encode_batch(const ctk_tokenizer* tok, const uint8_t* text, const uint64_t* text_off, size_t n, ctk_result** res) {
    return encode_host(tok, text, text_off, n, false, res);
}

int ctk_encode_batch_narrow(const ctk_tokenizer* tok, const uint8_t* text, const uint64_t* text_off, size_t n, ctk_result** res) {
    return encode_host(tok, text, text_off, n, true, res);
}

int ctk_id_width(const ctk_tokenizer* tok) { return tok ? reinterpret_cast<const Engine*>(tok)->run_width : 4; }

int ctk_decode_batch(const ctk_tokenizer* tok, const uint32_t* ids, const uint64_t* ids_off, size_t n, int skip_special_tokens,
                     int clean_up_tokenization_spaces, ctk_result** res) {
    if (!tok || !ids_off || !res) { set_last_error("NULL argument"); return CTK_ERR_ARG; }
    *res = nullptr;
    Engine* eng = const_cast<Engine*>(reinterpret_cast<const Engine*>(tok));
    if (ids_off[0] != 0) { set_last_error("ids_off[0] must be 0"); return CTK_ERR_ARG; }
    for (size_t i = 0; i < n; ++i) if (ids_off[i + 1] < ids_off[i]) { set_last_error("ids_off must be non-decreasing"); return CTK_ERR_ARG; }
    const uint64_t T = ids_off[n];
    if (T && !ids) { set_last_error("NULL ids"); return CTK_ERR_ARG; }
    Result* r = nullptr;
    int rc;
    if (eng->peers.size() > 1 && T * 4 >= multi_min_bytes()) {
        r = new Result();
        r->n = n;
        const std::vector<size_t> cut = balanced_cuts(ids_off, n, eng->peers.size());
        rc = run_on_peers(eng, cut, r, [&](Engine* pe, size_t d0, size_t nd, Result** out) {
            return decode_single(pe, ids, ids_off + d0, nd, skip_special_tokens, clean_up_tokenization_spaces, out);
        });
        if (rc != CTK_OK) { free_result(r); r = nullptr; }
    } else {
        rc = decode_single(eng, ids, ids_off, n, skip_special_tokens, clean_up_tokenization_spaces, &r);
    }
    if (rc != CTK_OK) return rc;
    *res = reinterpret_cast<ctk_result*>(r);
    return CTK_OK;
}

int ctk_result_id_width(const ctk_result* res) { return reinterpret_cast<const Result*>(res)->id_width; }
size_t ctk_result_parts(const ctk_result* res) {
    const Result* r = reinterpret_cast<const Result*>(res);
    return r->parts.empty() ? 1 : r->parts.size();
}
int ctk_result_part(const ctk_result* res, size_t i, size_t* first_item, size_t* n_items, const void** data, const uint64_t** offsets) {
    const Result* r = reinterpret_cast<const Result*>(res);
    const Result* p = r;
    size_t first = 0;
    if (!r->parts.empty()) { if (i >= r->parts.size()) return CTK_ERR_ARG; p = r->parts[i]; first = r->part_first[i]; }
    else if (i != 0) return CTK_ERR_ARG;
    if (first_item) *first_item = first;
    if (n_items) *n_items = p->n;
    if (data) *data = p->ids.p ? p->ids.p : p->bytes.p;
    if (offsets) *offsets = (const uint64_t*)p->off.p;
    return CTK_OK;
}

const void* ctk_result_ids_raw(const ctk_result* res) {
    const Result* r = reinterpret_cast<const Result*>(res);
    return r->parts.empty() ? r->ids.p : nullptr;           // several parts: use ctk_result_part (or ctk_result_ids, which gathers)
}

const uint32_t* ctk_result_ids(const ctk_result* res) {
    Result* r = const_cast<Result*>(reinterpret_cast<const Result*>(res));
    if (r->parts.empty() && r->id_width == 4) return (const uint32_t*)r->ids.p;
    std::lock_guard<std::mutex> lk(r->lazy_mu);
    if (!r->have_wide) {                                     // gather the parts and / or widen, with the copy threads
        r->wide.resize(r->total + 1);
        TaskPool* pool = copy_pool_for_node(-1);
        const int T = default_copy_threads();
        std::vector<const Result*> src;
        if (r->parts.empty()) src.push_back(r); else for (const Result* p : r->parts) src.push_back(p);
        uint64_t base = 0;
        for (const Result* p : src) {
            uint32_t* dst = r->wide.data() + base;
            if (p->id_width == 2) {
                const uint16_t* s = (const uint16_t*)p->ids.p;
                parallel_for_bytes(pool, T, p->total * 2, [=](uint64_t lo, uint64_t hi) { for (uint64_t i = lo / 2; i < hi / 2; ++i) dst[i] = s[i]; });
            } else {
                const uint8_t* s = (const uint8_t*)p->ids.p;
                parallel_for_bytes(pool, T, p->total * 4, [=](uint64_t lo, uint64_t hi) { memcpy((uint8_t*)dst + lo, s + lo, hi - lo); });
            }
            base += p->total;
        }
        r->have_wide = true;
    }
    return r->wide.data();
}

const uint64_t* ctk_result_offsets(const ctk_result* res) {
    Result* r = const_cast<Result*>(reinterpret_cast<const Result*>(res));
    if (r->parts.empty()) return (const uint64_t*)r->off.p;
    build_flat_offsets(r);
    return r->flat_off.data();
}

const uint8_t* ctk_result_bytes(const ctk_result* res) {
    Result* r = const_cast<Result*>(reinterpret_cast<const Result*>(res));
    if (r->parts.empty()) return (const uint8_t*)r->bytes.p;
    std::lock_guard<std::mutex> lk(r->lazy_mu);
    if (!r->have_bytes) {
        r->flat_bytes.resize(r->total + 1);
        uint64_t base = 0;
        for (const Result* p : r->parts) { if (p->total) memcpy(r->flat_bytes.data() + base, p->bytes.p, p->total); base += p->total; }
        r->have_bytes = true;
    }
    return r->flat_bytes.data();
}

// debug/test hook (host only): the worker pool that stages pageable input, as a parallel memcpy
int ctk_debug_parallel_copy(void* dst, const void* src, size_t bytes, int threads, int parts) {
    if (threads < 1 || parts < 1) return CTK_ERR_ARG;
    TaskPool* pool = copy_pool_for_node(-1);
    pool->start(threads - 1);
    const size_t per = (bytes + parts - 1) / parts;
    pool->run(parts, [=](int i) {
        const size_t lo = per * (size_t)i, hi = std::min(bytes, lo + per);
        if (lo < hi) memcpy((char*)dst + lo, (const char*)src + lo, hi - lo);
    });
    return CTK_OK;
}

void ctk_last_transfer_bytes(const ctk_tokenizer* tok, uint64_t* h2d, uint64_t* d2h) {
    const Engine* eng = reinterpret_cast<const Engine*>(tok);
    uint64_t a = 0, b = 0;
    if (eng->peers.size() > 1) for (const Engine* p : eng->peers) { a += p->last_h2d_bytes; b += p->last_d2h_bytes; }
    else { a = eng->last_h2d_bytes; b = eng->last_d2h_bytes; }
    if (h2d) *h2d = a;
    if (d2h) *d2h = b;
}
size_t ctk_result_count(const ctk_result* res) { return reinterpret_cast<const Result*>(res)->n; }
void ctk_result_free(ctk_result* res) { free_result(reinterpret_cast<Result*>(res)); }

}  // extern "C"
