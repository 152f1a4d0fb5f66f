// Engine: one loaded tokenizer bound to one GPU (device tables, scratch workspace, stream-ordered calls).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <initializer_list>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ctk.h"
#include "device_common.cuh"
#include "model.hpp"
#include "split_walk.cuh"

namespace ctk {

// device-side error flags (OR-ed into a word the host reads back)
enum : uint32_t { ERRF_OFFSETS = 1, ERRF_CAPACITY = 2, ERRF_UTF8 = 4, ERRF_POOL = 8, ERRF_NFC_LONG = 16, ERRF_NFC_SUSPECT = 32,
                  ERRF_ALIGN = 64, ERRF_PANIC = 128 };
constexpr int CTK_RETRY_NFC = 100;     // internal: the optimistic encode met a code point that needs NFC; run the normaliser and encode again

void set_last_error(const std::string& s);
extern std::atomic<uint64_t> g_kernel_launches;

// Grow-only device scratch, one buffer per slot.  Growing frees the old buffer (cudaFree
// synchronises the device, so nothing in flight can still be using it).
struct Workspace {
    static constexpr int kSlots = 96;
    void* p[kSlots] = {};
    size_t cap[kSlots] = {};
    cudaError_t get(int slot, size_t bytes, void** out) {
        if (bytes > cap[slot]) {
            if (p[slot]) cudaFree(p[slot]);
            p[slot] = nullptr; cap[slot] = 0;
            size_t want = bytes + bytes / 8 + 256;
            cudaError_t e = cudaMalloc(&p[slot], want);
            if (e != cudaSuccess) return e;
            cap[slot] = want;
        }
        *out = p[slot];
        return cudaSuccess;
    }
    void release() { for (int i = 0; i < kSlots; ++i) { if (p[i]) cudaFree(p[i]); p[i] = nullptr; cap[i] = 0; } }
};

struct DecodeTables {
    const uint8_t* blob = nullptr;
    const uint32_t* off = nullptr;      // n_ids + 1
    const uint8_t* special = nullptr;   // n_ids
    // compact copies for the gather: length per id (255 = look at off[]), the same with special tokens zeroed,
    // and one 16-byte record per id {first 12 bytes, length} so that a token costs ONE table access per pass
    const uint8_t* len8 = nullptr;
    const uint8_t* len8_skip = nullptr;
    const uint4* rec = nullptr;
    uint32_t n_ids = 0;
};

struct NfcTables {                      // canonical decomposition / composition / ccc, sorted for binary search
    const uint32_t *dkey, *da, *db; int nd;
    const uint64_t* ckey; const uint32_t* cval; int nc;
    const uint32_t* qkey; const uint8_t* qval; int nq;
    const uint8_t *trie_index, *trie_blocks;
};

struct RichTables {
    const uint32_t* tok_info = nullptr;     // [n ids] decoded byte length | vocabulary-string byte length << 16
    const uint32_t* special_bits = nullptr; // bit per id: id is a value of the special_tokens map (encoding.rs:77-83)
    uint32_t n_special_words = 0;
    const uint16_t* byte_map2 = nullptr;    // [256] UTF-8 of the byte-mapped char: low byte first, high byte 0 for one-byte chars
};

struct Engine {
    HostModel model;
    int device = 0;
    int numa_node = -1;                 // host memory node next to the device (sysfs), -1 unknown: page-locked buffers are placed there
    // ctk_from_file_devices: one engine per device, peers[0] == this (the handle); empty for a single-device tokenizer.
    // The host-buffer entry points split a batch over the peers by contiguous document ranges (host_api.cu).
    std::vector<Engine*> peers;
    DevTables tables{};
    DecodeTables dec{};
    NfcTables nfc{};
    RichTables rich{};
    void* d_table_mem[32] = {};
    std::vector<SplitTables> split_dev;  // device tables of model.split_stages (split.cu)
    std::vector<void*> split_mem;
    const void* meta_char_tab = nullptr; // Metaspace pipelines: code point -> id of the single-character vocabulary entries (metaspace.cu)
    uint32_t meta_char_mask = 0;
    Workspace ws;
    std::mutex mu;                      // serialises device work issued through this tokenizer
    bool cache_persistent = false;
    bool cache_valid = false;           // the pre-token cache holds entries of earlier calls
    uint32_t cache_init_slots = 0;      // slots [0, this) hold entries or EMPTY; the rest of the table was never cleared since the last full clear
    bool use_general = false;           // debug: run the multi-kernel pipeline instead of the fused kernel
    uint32_t max_emit_id = 0;           // largest id encode can emit (vocabulary and in-word added tokens)
    int run_width = 4;                  // bytes per id in the encode kernels' scratch runs: 2 when max_emit_id < 65 536
    int out_id_width = 4;               // bytes per id of the CURRENT device call's output buffer (4, or 2 via the _ex / narrow entry points)
    int long_grid = 0;
    int dec_write_grid = 0;             // persistent CTAs of k_dec_write
    uint32_t dec_hot = 0;               // records k_dec_write keeps in shared memory (ids 0 .. dec_hot - 1)
    int dec_sums_grid = 0;              // persistent CTAs of k_dec_tile_sums_smem (0: not computed yet, -1: the length table does not fit in shared memory)
    int mid_grid[4] = {};               // co-resident single-warp CTAs of the four k_encode_mid instantiations
    uint64_t last_h2d_bytes = 0, last_d2h_bytes = 0;   // bytes the last host-buffer encode call moved over PCIe
    int xl_last_rounds = 0;             // rounds the last very-long-pre-token pass took (diagnostics)
    int fused_grid = 0;                 // co-resident CTAs of k_encode_fused (SMs x occupancy), computed once
    // pinned staging for error flags / counters
    uint32_t* h_flags = nullptr;        // 64 words of MAPPED pinned memory; h_flags_dev is the same memory as the device sees it
    uint32_t* h_flags_dev = nullptr;
    // Small device->host readbacks (flags, totals) are stored by a one-thread kernel straight into mapped host
    // memory.  A 4-byte cudaMemcpyAsync would queue on the D2H copy engine behind the host pipeline's 64 MiB result
    // copies and hold the next chunk's kernels back for a millisecond (measured: 31 -> 4x GB/s end to end).
    struct Pub { const void* src; int words; int dst_word; };
    cudaError_t publish(std::initializer_list<Pub> items, cudaStream_t st);
    uint8_t* last_decode_out = nullptr;
    cudaStream_t st_h2d = nullptr, st_comp = nullptr, st_d2h = nullptr;   // host-buffer pipeline: copy in / compute / copy out
    // Almost all text is already NFC.  With `nfc_optimistic` the encode kernels run on the raw text and only REPORT an
    // NFC-suspect code point (same trie bit the normaliser's own scan uses); the call then runs again through the
    // normaliser, and later calls scan first, until a call comes back clean.  Saves one pass over the text.
    bool nfc_optimistic = true;
    bool last_nfc_needed = false;
    bool keep_cache_once = false;       // next encode call continues the current batch (chunked host-buffer path)   // decode_device(d_out = NULL) leaves its output here

    // optional per-kernel timing (CUDA events on the launching stream), for bench.py's roofline line
    bool profile = false;
    struct Mark { const char* name; cudaEvent_t ev; };
    std::vector<Mark> marks;
    std::vector<cudaEvent_t> ev_pool;                             // timing events (profile marks)
    std::vector<cudaEvent_t> sync_ev_pool;                        // cudaEventDisableTiming events (host pipeline)
    std::map<std::string, std::pair<double, uint64_t>> prof;      // name -> (total ms, launches)
    void mark(const char* name, cudaStream_t st) {                // call once before the first kernel (name = nullptr) and after each kernel
        if (!profile) return;
        cudaEvent_t ev;
        if (!ev_pool.empty()) { ev = ev_pool.back(); ev_pool.pop_back(); }
        else if (cudaEventCreate(&ev) != cudaSuccess) return;
        cudaEventRecord(ev, st);
        marks.push_back({name, ev});
    }
    void collect_marks() {                                        // after a synchronise
        for (size_t i = 0; i < marks.size(); ++i) {
            if (i > 0 && marks[i].name) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, marks[i - 1].ev, marks[i].ev) == cudaSuccess) {
                    auto& e = prof[marks[i].name]; e.first += ms; e.second += 1;
                }
            }
        }
        for (auto& m : marks) ev_pool.push_back(m.ev);
        marks.clear();
    }

    int fail(int code, const std::string& msg) { set_last_error(msg); return code; }
    int cuda_fail(cudaError_t e, const char* what) {
        set_last_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what);
        return CTK_ERR_CUDA;
    }
    void launched(int n) { g_kernel_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
    // common tail of an encode/decode device call: optionally synchronise, check the error flags,
    // return the total (last entry of the n+1 offsets array)
    int finish(const uint32_t* d_err, const uint64_t* d_off_out, size_t n, uint64_t* total_host, cudaStream_t st);
};

int nfc_stage(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes,
              const uint8_t** o_text, const uint64_t** o_off, uint64_t* o_bytes, cudaStream_t st);
int prefix_space_stage(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes,
                       const uint8_t** o_text, const uint64_t** o_off, uint64_t* o_bytes, cudaStream_t st);
int metaspace_empty_texts(Engine& eng, const uint64_t* d_text_off, size_t n, uint32_t* d_ids, uint64_t ids_cap, uint64_t* d_ids_off, uint64_t* n_ids_host,
                          cudaStream_t st);
int split_upload(Engine& eng);
int metaspace_upload(Engine& eng);
int encode_metaspace(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n, uint64_t n_bytes, uint32_t* d_ids, uint64_t ids_cap,
                     uint64_t* d_ids_off, uint64_t* n_ids_host, cudaStream_t st);
int split_stages(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes,
                 const uint8_t** o_text, const uint64_t** o_off, size_t* o_n, uint64_t* o_bytes, const uint64_t** first_piece, cudaStream_t st);
int split_fold_ids(Engine& eng, const uint64_t* first_piece, const uint64_t* piece_ids_off, size_t n_docs, uint64_t* d_ids_off, cudaStream_t st);
int encode_device(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n, uint64_t n_bytes, uint32_t* d_ids,
                  uint64_t ids_cap, uint64_t* d_ids_off, uint64_t* n_ids_host, cudaStream_t st);
int encode_fused(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes,
                 uint32_t* d_ids, uint64_t ids_cap, uint64_t* d_ids_off, uint64_t* n_ids_host, cudaStream_t st, bool check_nfc = false);
int encode_general(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes,
                   uint32_t* d_ids, uint64_t ids_cap, uint64_t* d_ids_off, uint64_t* n_ids_host, cudaStream_t st);
int starts_bitmap(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_docs, uint64_t n_bytes, uint32_t* ds,
                  uint32_t* start_bits, uint32_t* block_counts, uint32_t* err, cudaStream_t st);
cudaError_t pinned_get(size_t bytes, void** out, size_t* cap);      // process-wide pool of page-locked buffers (host_api.cu)
void pinned_put(void* p, size_t cap);
int device_numa_node(int device);

// ---- rich `Encoding` outputs (encoding.cu) ------------------------------------------------------
struct RichOut {                            // device pointers into the engine's workspace, valid until its next call
    uint64_t n_rows = 0, total = 0, n_tokens = 0, max_row = 0;
    const uint64_t* row_off = nullptr;      // n_rows + 1
    const uint64_t* row_full = nullptr;     // n_rows: length before truncation and padding
    const uint32_t* ids = nullptr;          // total
    const uint8_t *attention = nullptr, *type_ids = nullptr, *special = nullptr;   // total each
    const uint64_t* tok_off = nullptr;      // n_texts + 1: tokens of each text before post-processing
    const uint32_t* raw_ids = nullptr;      // n_tokens
    const uint2* offsets = nullptr;         // n_tokens (byte start, byte end) in the original text; NULL if not asked for
    const uint32_t* word_ids = nullptr;     // n_tokens
};
int encode_rich_device(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n_texts, uint64_t n_bytes,
                       const ctk_encoding_options& opt, RichOut* out, cudaStream_t st);

int clean_parallel(Engine& eng, const uint8_t* raw, const uint64_t* raw_off, size_t n_docs, uint64_t n, uint8_t* d_out,
                   uint64_t out_cap, uint64_t* d_out_off, uint32_t* err, cudaStream_t st);
int decode_device(Engine& eng, const uint32_t* d_ids, const uint64_t* d_ids_off, size_t n, uint64_t total_ids,
                  int skip_special, int cleanup, uint8_t* d_out, uint64_t out_cap, uint64_t* d_out_off,
                  uint64_t* n_bytes_host, cudaStream_t st);

}  // namespace ctk
