// C ABI (include/ctk.h): engine life cycle, table upload, host-buffer entry points.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <algorithm>
#include <new>
#include <vector>

#include "engine.hpp"
#include "start_window.cuh"
#include "unicode_trie_gen.h"

namespace ctk {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
std::atomic<uint64_t> g_kernel_launches{0};

struct PubArgs { const uint32_t* src[4]; int words[4]; int dst[4]; int n; };
__global__ void k_publish(PubArgs a, uint32_t* __restrict__ host) {
    for (int k = 0; k < a.n; ++k)
        for (int w = threadIdx.x; w < a.words[k]; w += blockDim.x) host[a.dst[k] + w] = a.src[k][w];
    __threadfence_system();
}

cudaError_t Engine::publish(std::initializer_list<Pub> items, cudaStream_t st) {
    PubArgs a{};
    for (const Pub& it : items) {
        if (a.n >= 4) return cudaErrorInvalidValue;
        a.src[a.n] = static_cast<const uint32_t*>(it.src); a.words[a.n] = it.words; a.dst[a.n] = it.dst_word; ++a.n;
    }
    k_publish<<<1, 32, 0, st>>>(a, h_flags_dev);
    launched(1);
    return cudaGetLastError();
}

int Engine::finish(const uint32_t* d_err, const uint64_t* d_off_out, size_t n, uint64_t* total_host, cudaStream_t st) {
    if (!total_host) return CTK_OK;                       // asynchronous use: flags are checked by the next synchronous call
    cudaError_t e = publish({{d_err, 2, 0}, {d_off_out + n, 2, 2}}, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "finish");
    collect_marks();
    uint32_t f = h_flags[0];
    if (f & ERRF_NFC_SUSPECT) return CTK_RETRY_NFC;         // nothing else about this attempt counts
    if (f & ERRF_OFFSETS) return fail(CTK_ERR_ARG, "offsets must start at 0, be non-decreasing and end at the buffer length");
    if (f & ERRF_UTF8) return fail(CTK_ERR_INVALID_DATA, "input text is not valid UTF-8");
    if (f & ERRF_CAPACITY) return fail(CTK_ERR_ARG, "output capacity too small");
    if (f & ERRF_POOL) return fail(CTK_ERR_CUDA, "internal scratch pool exhausted");
    memcpy(total_host, h_flags + 2, 8);
    return CTK_OK;
}

static int upload(Engine& eng) {
    cudaError_t e;
#define UP(slot, dst, src, bytes)                                                                     \
    do {                                                                                              \
        size_t b_ = (bytes);                                                                          \
        e = cudaMalloc(&eng.d_table_mem[slot], b_ ? b_ : 16);                                         \
        if (e != cudaSuccess) return eng.cuda_fail(e, "cudaMalloc(tables)");                          \
        if (b_) { e = cudaMemcpy(eng.d_table_mem[slot], (src), b_, cudaMemcpyHostToDevice);           \
                  if (e != cudaSuccess) return eng.cuda_fail(e, "cudaMemcpy(tables)"); }              \
        dst = reinterpret_cast<decltype(dst)>(eng.d_table_mem[slot]);                                 \
    } while (0)
    const HostModel& m = eng.model;
    // pair table: linear probing, power-of-two capacity >= 4 * entries
    uint64_t cap = 64;
    while (cap < m.pairs.size() * 4) cap <<= 1;
    std::vector<PairSlot> slots(cap, PairSlot{kNone, kNone, kNone, 0});
    for (const PairEntry& p : m.pairs) {
        uint32_t h = pair_hash(p.a, p.b) & (uint32_t)(cap - 1);
        while (slots[h].a != kNone) h = (h + 1) & (uint32_t)(cap - 1);
        slots[h] = PairSlot{p.a, p.b, p.rank, p.new_id};
    }
    eng.tables.pair_mask = (uint32_t)(cap - 1);
    UP(0, eng.tables.pairs, slots.data(), cap * sizeof(PairSlot));
    UP(1, eng.tables.byte_init, m.byte_init_id, 256 * 4);
    UP(2, eng.tables.trie_index, CTK_TRIE_INDEX, sizeof(CTK_TRIE_INDEX));
    UP(3, eng.tables.trie_blocks, CTK_TRIE_BLOCKS, sizeof(CTK_TRIE_BLOCKS));
    eng.dec.n_ids = (uint32_t)m.id_present.size();
    UP(4, eng.dec.blob, m.dec_blob.data(), m.dec_blob.size());
    UP(5, eng.dec.off, m.dec_off.data(), m.dec_off.size() * 4);
    UP(6, eng.dec.special, m.dec_special.data(), m.dec_special.size());
    {
        const size_t nid = m.dec_special.size();
        std::vector<uint8_t> l8((nid + 1 + 31) / 16 * 16, 0), l8s((nid + 1 + 31) / 16 * 16, 0);   // padded: k_dec_tile_sums_smem copies them 16 bytes at a time
        std::vector<uint4> rec(nid + 1, make_uint4(0, 0, 0, 0));

        for (size_t id = 0; id < nid; ++id) {
            const uint32_t b = m.dec_off[id], L = m.dec_off[id + 1] - b;
            l8[id] = (uint8_t)(L < 255 ? L : 255);
            l8s[id] = m.dec_special[id] ? 0 : l8[id];
            uint32_t w[3] = {0, 0, 0};
            for (uint32_t k = 0; k < L && k < 12; ++k) w[k >> 2] |= (uint32_t)m.dec_blob[b + k] << (8 * (k & 3));
            rec[id] = make_uint4(w[0], w[1], w[2], L);
        }
        UP(18, eng.dec.len8, l8.data(), l8.size());
        UP(19, eng.dec.len8_skip, l8s.data(), l8s.size());
        UP(20, eng.dec.rec, rec.data(), rec.size() * sizeof(uint4));
    }
    eng.nfc.nd = CTK_N_DECOMP; eng.nfc.nc = CTK_N_COMP; eng.nfc.nq = CTK_N_CCC;
    UP(7, eng.nfc.dkey, CTK_DECOMP_KEY, sizeof(CTK_DECOMP_KEY));
    UP(8, eng.nfc.da, CTK_DECOMP_A, sizeof(CTK_DECOMP_A));
    UP(9, eng.nfc.db, CTK_DECOMP_B, sizeof(CTK_DECOMP_B));
    UP(10, eng.nfc.ckey, CTK_COMP_KEY, sizeof(CTK_COMP_KEY));
    UP(11, eng.nfc.cval, CTK_COMP_VAL, sizeof(CTK_COMP_VAL));
    UP(12, eng.nfc.qkey, CTK_CCC_KEY, sizeof(CTK_CCC_KEY));
    UP(13, eng.nfc.qval, CTK_CCC_VAL, sizeof(CTK_CCC_VAL));
    eng.nfc.trie_index = eng.tables.trie_index; eng.nfc.trie_blocks = eng.tables.trie_blocks;
    {   // added tokens that can occur inside one pre-token (loader.cpp: analyse_added)
        std::vector<uint8_t> blob; std::vector<uint4> meta;
        for (const AddedTok& a : m.added) {
            if (!a.may_match) continue;
            meta.push_back(make_uint4((uint32_t)blob.size(), (uint32_t)a.bytes.size(), a.id,
                                      (a.single_word ? 1u : 0u) | (a.lstrip ? 2u : 0u) | (a.rstrip ? 4u : 0u)));
            blob.insert(blob.end(), a.bytes.begin(), a.bytes.end());
        }
        eng.tables.n_added = (uint32_t)meta.size();
        UP(14, eng.tables.added_blob, blob.data(), blob.size());
        UP(15, eng.tables.added_meta, meta.data(), meta.size() * sizeof(uint4));
        UP(16, eng.tables.mapped_alnum, CTK_MAPPED_ALNUM, sizeof(CTK_MAPPED_ALNUM));
    }
    {   // rich Encoding outputs (encoding.cu)
        const size_t nid = m.dec_special.size();
        std::vector<uint32_t> info(nid + 1, 0);
        for (size_t id = 0; id < nid; ++id) {
            const uint32_t L = m.dec_off[id + 1] - m.dec_off[id], S = id < m.token_str_len.size() ? m.token_str_len[id] : 0u;
            info[id] = (L < 0xFFFFu ? L : 0xFFFFu) | ((S < 0xFFFFu ? S : 0xFFFFu) << 16);
        }
        UP(21, eng.rich.tok_info, info.data(), info.size() * 4);
        uint32_t max_sp = 0;
        for (auto& kv : m.specials) if (kv.second < (1u << 27) && kv.second > max_sp) max_sp = kv.second;
        std::vector<uint32_t> bits(max_sp / 32 + 1, 0);
        for (auto& kv : m.specials) if (kv.second < (1u << 27)) bits[kv.second >> 5] |= 1u << (kv.second & 31);
        eng.rich.n_special_words = m.specials.empty() ? 0u : (uint32_t)bits.size();
        UP(22, eng.rich.special_bits, bits.data(), bits.size() * 4);
        uint32_t b2c[256];
        byte_map(b2c);
        uint16_t map2[256];
        for (int b = 0; b < 256; ++b) {
            const uint32_t cp = b2c[b];
            map2[b] = cp < 0x80u ? (uint16_t)cp : (uint16_t)((0xC0u | (cp >> 6)) | ((0x80u | (cp & 63u)) << 8));
        }
        UP(23, eng.rich.byte_map2, map2, sizeof map2);
    }
    {
        uint64_t mx = m.id_present.empty() ? 0 : m.id_present.size() - 1;
        for (const AddedTok& a : m.added) if (a.may_match) mx = std::max<uint64_t>(mx, a.id);
        eng.max_emit_id = (uint32_t)std::min<uint64_t>(mx, 0xFFFFFFFFull);
        eng.run_width = mx < 65536 && !getenv("CTK_WIDE_RUNS") ? 2 : 4;
    }
    eng.tables.round_parallel = m.round_parallel ? 1u : 0u;
    UP(17, eng.tables.reach, m.reach.data(), m.reach.size() * 4);
#undef UP
    e = cudaHostAlloc((void**)&eng.h_flags, 256, cudaHostAllocMapped);
    if (e != cudaSuccess) return eng.cuda_fail(e, "cudaHostAlloc");
    memset(eng.h_flags, 0, 256);
    e = cudaHostGetDevicePointer((void**)&eng.h_flags_dev, eng.h_flags, 0);
    if (e != cudaSuccess) return eng.cuda_fail(e, "cudaHostGetDevicePointer");
    for (cudaStream_t* sp : {&eng.st_h2d, &eng.st_comp, &eng.st_d2h}) {
        e = cudaStreamCreateWithFlags(sp, cudaStreamNonBlocking);
        if (e != cudaSuccess) return eng.cuda_fail(e, "cudaStreamCreate");
    }
    return CTK_OK;
}

// [Split stages ->] ByteLevel prefix space -> fused encode, on text that is already normalised.  With Split stages the pieces
// travel as documents and the per-piece id offsets are folded back to per-document offsets (split.cu).
static int encode_normalised(Engine& eng, const uint8_t* t, const uint64_t* o, size_t n, uint64_t b, uint32_t* d_ids, uint64_t ids_cap,
                             uint64_t* d_ids_off, uint64_t* n_ids_host, cudaStream_t st, bool check_nfc) {
    const uint64_t* first_piece = nullptr;
    const uint64_t* const text_off = o;                      // offsets of the normalised texts, before any Split stage
    size_t n_pieces = n;
    int rc = split_stages(eng, t, o, n, b, &t, &o, &n_pieces, &b, &first_piece, st);
    if (rc != CTK_OK) return rc;
    if (eng.model.metaspace) {                               // pretokenizers.rs:188-200: the last stage is Metaspace, symbols are characters
        if (!first_piece) return encode_metaspace(eng, t, o, n, b, d_ids, ids_cap, d_ids_off, n_ids_host, st);
        uint64_t* piece_ids_off;
        cudaError_t e = eng.ws.get(59, (n_pieces + 2) * 8, (void**)&piece_ids_off);
        if (e != cudaSuccess) return eng.cuda_fail(e, "workspace");
        uint64_t total = 0;
        rc = encode_metaspace(eng, t, o, n_pieces, b, d_ids, ids_cap, piece_ids_off, &total, st);
        if (rc != CTK_OK) return rc;
        rc = split_fold_ids(eng, first_piece, piece_ids_off, n, d_ids_off, st);
        if (rc != CTK_OK) return rc;
        if (n_ids_host) *n_ids_host = total;
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return eng.cuda_fail(e, "split fold");
        return metaspace_empty_texts(eng, text_off, n, d_ids, ids_cap, d_ids_off, n_ids_host, st);
    }
    rc = prefix_space_stage(eng, t, o, n_pieces, b, &t, &o, &b, st);
    if (rc != CTK_OK) return rc;
    if (!first_piece) {
        if (eng.use_general && !check_nfc) return encode_general(eng, t, o, n, b, d_ids, ids_cap, d_ids_off, n_ids_host, st);
        return encode_fused(eng, t, o, n, b, d_ids, ids_cap, d_ids_off, n_ids_host, st, check_nfc);
    }
    uint64_t* piece_ids_off;
    cudaError_t e = eng.ws.get(59, (n_pieces + 2) * 8, (void**)&piece_ids_off);
    if (e != cudaSuccess) return eng.cuda_fail(e, "workspace");
    uint64_t total = 0;
    rc = encode_fused(eng, t, o, n_pieces, b, d_ids, ids_cap, piece_ids_off, &total, st, check_nfc);
    if (rc != CTK_OK) return rc;
    rc = split_fold_ids(eng, first_piece, piece_ids_off, n, d_ids_off, st);
    if (rc != CTK_OK) return rc;
    if (n_ids_host) {
        *n_ids_host = total;
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return eng.cuda_fail(e, "split fold");
    }
    return CTK_OK;
}

// normaliser -> pre-tokenise -> BPE -> emit, all on the device (mod.rs:551-613 for every document)
int encode_device(Engine& eng, const uint8_t* d_text, const uint64_t* d_off, size_t n, uint64_t n_bytes, uint32_t* d_ids,
                         uint64_t ids_cap, uint64_t* d_ids_off, uint64_t* n_ids_host, cudaStream_t st) {
    if (n_bytes && (reinterpret_cast<uintptr_t>(d_text) & 15)) return eng.fail(CTK_ERR_ARG, "device text buffer must be 16-byte aligned");
    const uint8_t* t2; const uint64_t* o2; uint64_t b2;
    int rc;
    const bool sync_call = n_ids_host != nullptr || !eng.split_dev.empty();     // (Split stages synchronise anyway)
    if (eng.model.nfc && eng.nfc_optimistic && sync_call && !eng.use_general && !eng.model.metaspace && !getenv("CTK_NO_NFC_OPTIMISM")) {
        const bool keep = eng.keep_cache_once;
        uint64_t dummy;
        rc = encode_normalised(eng, d_text, d_off, n, n_bytes, d_ids, ids_cap, d_ids_off, n_ids_host ? n_ids_host : &dummy, st, true);
        if (rc != CTK_RETRY_NFC) return rc;
        eng.nfc_optimistic = false;                          // this text needs the normaliser: scan first from now on
        eng.keep_cache_once = keep;
    }
    rc = nfc_stage(eng, d_text, d_off, n, n_bytes, &t2, &o2, &b2, st);
    if (rc != CTK_OK) return rc;
    if (n_bytes) eng.nfc_optimistic = !eng.last_nfc_needed;  // a clean call switches the optimistic path back on
    return encode_normalised(eng, t2, o2, n, b2, d_ids, ids_cap, d_ids_off, n_ids_host, st, false);
}

static int create(const uint8_t* json, size_t len, int device, ctk_tokenizer** out) {
    if (!out) { set_last_error("out is NULL"); return CTK_ERR_ARG; }
    *out = nullptr;
    Engine* eng = new (std::nothrow) Engine();
    if (!eng) { set_last_error("out of memory"); return CTK_ERR_CUDA; }
    std::string err;
    int rc = load_model(json, len, eng->model, err);
    if (rc != CTK_OK) { set_last_error(err); delete eng; return rc; }
    if (eng->model.pairs.size() >= (1u << 26)) { set_last_error("too many merges"); delete eng; return CTK_ERR_UNSUPPORTED; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_last_error(std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
        delete eng;
        return CTK_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_last_error("bad device index"); delete eng; return CTK_ERR_ARG; }
    eng->device = device;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { rc = eng->cuda_fail(e, "cudaSetDevice"); delete eng; return rc; }
    eng->numa_node = device_numa_node(device);
    rc = upload(*eng);
    if (rc == CTK_OK) rc = split_upload(*eng);
    if (rc == CTK_OK) rc = metaspace_upload(*eng);
    if (rc != CTK_OK) { delete eng; return rc; }
    *out = reinterpret_cast<ctk_tokenizer*>(eng);
    return CTK_OK;
}

}  // namespace ctk

using namespace ctk;

extern "C" {

int ctk_from_json(const uint8_t* json, size_t len, int device, ctk_tokenizer** out) {
    if (!json) { set_last_error("json is NULL"); return CTK_ERR_ARG; }
    return create(json, len, device, out);
}

int ctk_from_file(const char* path, int device, ctk_tokenizer** out) {
    if (!path) { set_last_error("path is NULL"); return CTK_ERR_ARG; }
    std::ifstream f(path, std::ios::binary);
    if (!f) { set_last_error(std::string("cannot open ") + path + ": No such file or directory (os error 2)"); return CTK_ERR_IO; }
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (f.bad()) { set_last_error(std::string("read error on ") + path); return CTK_ERR_IO; }
    return create((const uint8_t*)data.data(), data.size(), device, out);
}

static void free_engine(Engine* eng);

void ctk_free(ctk_tokenizer* tok) {
    if (!tok) return;
    Engine* eng = reinterpret_cast<Engine*>(tok);
    std::vector<Engine*> peers = eng->peers;
    for (Engine* p : peers) if (p != eng) free_engine(p);
    free_engine(eng);
}

// Replaces HuggingFaceTokenizer::from_file for a tokenizer that drives several GPUs from one process: the reference's
// encode_batch is one call that uses the whole machine (mod.rs:694-696).  Tables are replicated on every device.
int ctk_from_json_devices(const uint8_t* json, size_t len, int n_devices, const int* device_ids, ctk_tokenizer** out) {
    if (!json || !out) { set_last_error("NULL argument"); return CTK_ERR_ARG; }
    *out = nullptr;
    std::vector<int> ids;
    if (n_devices <= 0) {                                                  // all visible devices
        int nd = 0;
        cudaError_t e = cudaGetDeviceCount(&nd);
        if (e != cudaSuccess || nd == 0) {
            set_last_error(std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
            return CTK_ERR_CUDA;
        }
        for (int i = 0; i < nd; ++i) ids.push_back(i);
    } else {
        if (!device_ids) { set_last_error("device_ids is NULL"); return CTK_ERR_ARG; }
        ids.assign(device_ids, device_ids + n_devices);
        for (size_t i = 0; i < ids.size(); ++i)
            for (size_t j = 0; j < i; ++j) if (ids[i] == ids[j]) { set_last_error("a device is listed twice"); return CTK_ERR_ARG; }
    }
    std::vector<Engine*> engines;
    for (int d : ids) {
        ctk_tokenizer* t = nullptr;
        const int rc = create(json, len, d, &t);
        if (rc != CTK_OK) { for (Engine* e : engines) free_engine(e); return rc; }
        engines.push_back(reinterpret_cast<Engine*>(t));
    }
    if (engines.size() > 1) engines[0]->peers = engines;
    *out = reinterpret_cast<ctk_tokenizer*>(engines[0]);
    return CTK_OK;
}

int ctk_from_file_devices(const char* path, int n_devices, const int* device_ids, ctk_tokenizer** out) {
    if (!path) { set_last_error("path is NULL"); return CTK_ERR_ARG; }
    std::ifstream f(path, std::ios::binary);
    if (!f) { set_last_error(std::string("cannot open ") + path + ": No such file or directory (os error 2)"); return CTK_ERR_IO; }
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (f.bad()) { set_last_error(std::string("read error on ") + path); return CTK_ERR_IO; }
    return ctk_from_json_devices((const uint8_t*)data.data(), data.size(), n_devices, device_ids, out);
}

size_t ctk_n_devices(const ctk_tokenizer* tok) {
    const Engine* eng = reinterpret_cast<const Engine*>(tok);
    return eng->peers.empty() ? 1 : eng->peers.size();
}
int ctk_device_at(const ctk_tokenizer* tok, size_t i) {
    const Engine* eng = reinterpret_cast<const Engine*>(tok);
    if (eng->peers.empty()) return i == 0 ? eng->device : -1;
    return i < eng->peers.size() ? eng->peers[i]->device : -1;
}
int ctk_numa_node(const ctk_tokenizer* tok, size_t i) {
    const Engine* eng = reinterpret_cast<const Engine*>(tok);
    if (eng->peers.empty()) return i == 0 ? eng->numa_node : -1;
    return i < eng->peers.size() ? eng->peers[i]->numa_node : -1;
}

static void free_engine(Engine* eng) {
    cudaSetDevice(eng->device);
    cudaDeviceSynchronize();
    eng->ws.release();
    for (void* p : eng->d_table_mem) if (p) cudaFree(p);
    for (void* p : eng->split_mem) if (p) cudaFree(p);
    if (eng->h_flags) cudaFreeHost(eng->h_flags);
    for (cudaStream_t sp : {eng->st_h2d, eng->st_comp, eng->st_d2h}) if (sp) cudaStreamDestroy(sp);
    for (cudaEvent_t ev : eng->ev_pool) cudaEventDestroy(ev);
    for (cudaEvent_t ev : eng->sync_ev_pool) cudaEventDestroy(ev);
    delete eng;
}

size_t ctk_vocab_size(const ctk_tokenizer* tok) { return reinterpret_cast<const Engine*>(tok)->model.vocab.size(); }

int ctk_token_to_id(const ctk_tokenizer* tok, const uint8_t* token, size_t len, uint32_t* id) {
    const HostModel& m = reinterpret_cast<const Engine*>(tok)->model;
    auto it = m.vocab.find(std::string((const char*)token, len));
    if (it == m.vocab.end()) return 0;
    if (id) *id = it->second;
    return 1;
}

const uint8_t* ctk_id_to_token(const ctk_tokenizer* tok, uint32_t id, size_t* len) {
    const HostModel& m = reinterpret_cast<const Engine*>(tok)->model;
    if (id >= m.id_present.size() || !m.id_present[id]) return nullptr;
    if (len) *len = m.id_to_token[id].size();
    return (const uint8_t*)m.id_to_token[id].data();
}

size_t ctk_n_special_tokens(const ctk_tokenizer* tok) { return reinterpret_cast<const Engine*>(tok)->model.specials.size(); }

const uint8_t* ctk_special_token(const ctk_tokenizer* tok, size_t i, size_t* len, uint32_t* id) {
    const HostModel& m = reinterpret_cast<const Engine*>(tok)->model;
    if (i >= m.specials.size()) return nullptr;
    if (len) *len = m.specials[i].first.size();
    if (id) *id = m.specials[i].second;
    return (const uint8_t*)m.specials[i].first.data();
}

int ctk_device_numa_node(int device) { return device_numa_node(device); }
int ctk_device(const ctk_tokenizer* tok) { return reinterpret_cast<const Engine*>(tok)->device; }
size_t ctk_decode_max_bytes(const ctk_tokenizer* tok) { return reinterpret_cast<const Engine*>(tok)->model.dec_max_bytes; }
const char* ctk_last_error(void) { return g_last_error.c_str(); }
uint64_t ctk_kernel_launches(void) { return g_kernel_launches.load(); }
void ctk_profile_enable(ctk_tokenizer* tok, int on) {
    Engine* eng = reinterpret_cast<Engine*>(tok);
    std::lock_guard<std::mutex> lk(eng->mu);
    eng->profile = on != 0;
    eng->prof.clear();
}
size_t ctk_profile_report(ctk_tokenizer* tok, char* buf, size_t cap) {
    Engine* eng = reinterpret_cast<Engine*>(tok);
    std::lock_guard<std::mutex> lk(eng->mu);
    std::string s;
    for (auto& kv : eng->prof) {
        char line[256];
        snprintf(line, sizeof line, "%s\t%.6f\t%llu\n", kv.first.c_str(), kv.second.first, (unsigned long long)kv.second.second);
        s += line;
    }
    if (buf && cap) { size_t n = s.size() < cap - 1 ? s.size() : cap - 1; memcpy(buf, s.data(), n); buf[n] = 0; }
    return s.size();
}
void ctk_set_cache_persistent(ctk_tokenizer* tok, int persistent) {
    Engine* eng = reinterpret_cast<Engine*>(tok);
    eng->cache_persistent = persistent != 0;
    for (Engine* p : eng->peers) p->cache_persistent = persistent != 0;
}

int ctk_encode_batch_device_ex(const ctk_tokenizer* tok, const uint8_t* d_text, const uint64_t* d_text_off, size_t n,
                               uint64_t total_bytes, void* d_ids, uint64_t ids_cap, int id_width, uint64_t* d_ids_off,
                               uint64_t* n_ids_host, void* stream) {
    if (!tok || !d_text_off || !d_ids_off || (total_bytes && (!d_text || !d_ids))) { set_last_error("NULL argument"); return CTK_ERR_ARG; }
    Engine* eng = const_cast<Engine*>(reinterpret_cast<const Engine*>(tok));
    if (id_width != 4 && !(id_width == 2 && eng->run_width == 2 && !eng->use_general)) {
        set_last_error("id_width must be 4, or 2 when ctk_id_width(tok) == 2");
        return CTK_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(eng->mu);
    cudaError_t e = cudaSetDevice(eng->device);
    if (e != cudaSuccess) return eng->cuda_fail(e, "cudaSetDevice");
    eng->out_id_width = id_width;
    const int rc = encode_device(*eng, d_text, d_text_off, n, total_bytes, static_cast<uint32_t*>(d_ids), ids_cap, d_ids_off, n_ids_host, (cudaStream_t)stream);
    eng->out_id_width = 4;
    return rc;
}

int ctk_encode_batch_device(const ctk_tokenizer* tok, const uint8_t* d_text, const uint64_t* d_text_off, size_t n,
                            uint64_t total_bytes, uint32_t* d_ids, uint64_t ids_cap, uint64_t* d_ids_off,
                            uint64_t* n_ids_host, void* stream) {
    if (!tok || !d_text_off || !d_ids_off || (total_bytes && (!d_text || !d_ids))) { set_last_error("NULL argument"); return CTK_ERR_ARG; }
    Engine* eng = const_cast<Engine*>(reinterpret_cast<const Engine*>(tok));
    std::lock_guard<std::mutex> lk(eng->mu);
    cudaError_t e = cudaSetDevice(eng->device);
    if (e != cudaSuccess) return eng->cuda_fail(e, "cudaSetDevice");
    return encode_device(*eng, d_text, d_text_off, n, total_bytes, d_ids, ids_cap, d_ids_off, n_ids_host, (cudaStream_t)stream);
}

int ctk_decode_batch_device(const ctk_tokenizer* tok, const uint32_t* d_ids, const uint64_t* d_ids_off, size_t n,
                            uint64_t total_ids, int skip_special_tokens, int clean_up_tokenization_spaces,
                            uint8_t* d_text_out, uint64_t text_cap, uint64_t* d_text_off_out, uint64_t* n_bytes_host,
                            void* stream) {
    if (!tok || !d_ids_off || !d_text_off_out || !d_text_out || (total_ids && !d_ids)) { set_last_error("NULL argument"); return CTK_ERR_ARG; }
    Engine* eng = const_cast<Engine*>(reinterpret_cast<const Engine*>(tok));
    std::lock_guard<std::mutex> lk(eng->mu);
    cudaError_t e = cudaSetDevice(eng->device);
    if (e != cudaSuccess) return eng->cuda_fail(e, "cudaSetDevice");
    return decode_device(*eng, d_ids, d_ids_off, n, total_ids, skip_special_tokens, clean_up_tokenization_spaces, d_text_out,
                         text_cap, d_text_off_out, n_bytes_host, (cudaStream_t)stream);
}

// debug: select the multi-kernel general pipeline (1) or the fused kernel (0, default)
void ctk_debug_use_general(ctk_tokenizer* tok, int on) { reinterpret_cast<Engine*>(tok)->use_general = on != 0; }

// ---- debug/test hooks (host only, no GPU needed): the scalar start predicate on host memory ----------
// out_bits: one bit per byte position, set where a pre-token starts.  Used by the CPU test-suite to
// check the device predicate's logic against the oracle's regex restatement.
int ctk_debug_starts_host(const uint8_t* text, uint64_t n, const uint64_t* off, size_t n_docs, uint32_t* out_bits) {
    uint64_t n_words = (n + 31) / 32 + 1;
    std::vector<uint32_t> ds(n_words, 0);
    for (size_t d = 0; d <= n_docs; ++d) if (off[d] < n) ds[off[d] >> 5] |= 1u << (off[d] & 31);
    TextView tv{text, n, ds.data(), CTK_TRIE_INDEX, CTK_TRIE_BLOCKS};
    for (uint64_t w = 0; w < n_words; ++w) out_bits[w] = 0;
    for (uint64_t i = 0; i < n; ++i) if (tv.is_start(i)) out_bits[i >> 5] |= 1u << (i & 31);
    return CTK_OK;
}

// Same, through the bit-parallel window logic the fused kernel uses (classify16 + start_window), with the
// 16-byte groups / neighbour windows emulated on the host.
int ctk_debug_starts_window_host(const uint8_t* text, uint64_t n, const uint64_t* off, size_t n_docs, uint32_t* out_bits) {
    uint64_t n_groups = (n + 15) / 16 + 2;                       // one group of padding on each side
    std::vector<uint8_t> buf((n_groups + 2) * 16, 0);
    uint8_t* chunk = buf.data() + 16;                            // chunk[0] = first padding group
    if (n) memcpy(chunk + 16, text, n);
    std::vector<uint32_t> ds(n_groups, 0);
    for (size_t d = 0; d <= n_docs; ++d) { uint64_t p = off[d] + 16; if (off[d] <= n) ds[p >> 4] |= 1u << (p & 15); }
    std::vector<Masks16> m(n_groups);
    for (uint64_t g = 0; g < n_groups; ++g) {
        uint32_t w[4];
        memcpy(w, chunk + g * 16, 16);
        m[g] = classify16(chunk, (int)(g * 16), w[0], w[1], w[2], w[3], CTK_TRIE_INDEX, CTK_TRIE_BLOCKS);
    }
    uint64_t n_words = (n + 31) / 32 + 1;
    for (uint64_t w = 0; w < n_words; ++w) out_bits[w] = 0;
    Masks16 z{};
    for (uint64_t g = 1; g + 1 < n_groups; ++g) {
        const Masks16 &a = m[g - 1], &b = m[g], &c = (g + 1 < n_groups) ? m[g + 1] : z;
        uint32_t S = start_window(window(a.L, b.L, c.L), window(a.N, b.N, c.N), window(a.W, b.W, c.W), window(a.SP, b.SP, c.SP),
                                  window(a.AP, b.AP, c.AP), window(a.CONT, b.CONT, c.CONT), window(ds[g - 1], ds[g], ds[g + 1]),
                                  chunk + g * 16 - 8);
        for (int k = 0; k < 16; ++k) {
            uint64_t pos = (g - 1) * 16 + k;
            if (pos < n && ((S >> (8 + k)) & 1u)) out_bits[pos >> 5] |= 1u << (pos & 31);
        }
    }
    return CTK_OK;
}

// the product's code point table (class in bits 0-1: 0 Other, 1 L, 2 N, 3 White_Space; bit 2: NFC-suspect), for the CPU
// test that checks every code point against Python's `regex` / `unicodedata`
// TEST HOOK (no device involved, never on a product path): compiles `pattern` like the loader does and walks ONE text through the
// automaton on the host with the same split_walk() the device kernels run.  pieces = (start, end) pairs.
// Returns 0, CTK_ERR_UNSUPPORTED (pattern outside the subset), or -1 when the `regex` crate would reject the pattern.
namespace { struct HostPieces {
    uint64_t *cuts, *spans; size_t nc = 0, ns = 0; bool any = false;
    CTK_HD void boundary(uint64_t p) { if (nc == 0 || cuts[nc - 1] != p) cuts[nc++] = p; }
    CTK_HD void span(uint64_t a, uint64_t b, bool starts_piece) { if (starts_piece || ns == 0) { spans[ns++] = a; spans[ns++] = b; } else spans[ns - 1] = b; }
    CTK_HD void matched() { any = true; }
}; }
int ctk_debug_split_pieces(const char* pattern, int behavior, int invert, const uint8_t* text, uint64_t n, uint64_t* pieces, size_t cap_pairs, size_t* n_pairs,
                           uint32_t* n_states, uint32_t* n_classes, int segment) {
    std::string json = std::string("{\"model\":{\"vocab\":{},\"merges\":[]},\"pre_tokenizer\":{\"type\":\"Sequence\",\"pretokenizers\":[{\"type\":\"Split\",\"pattern\":{\"Regex\":");
    json += '"';
    for (const char* c = pattern; *c; ++c) {
        if (*c == '"' || *c == '\\') { json += '\\'; json += *c; }
        else if ((unsigned char)*c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", (unsigned)(unsigned char)*c); json += b; }
        else json += *c;
    }
    json += "\"},\"behavior\":\"Isolated\"},{\"type\":\"ByteLevel\"}]}}";
    HostModel m;
    std::string err;
    int rc = load_model(reinterpret_cast<const uint8_t*>(json.data()), json.size(), m, err);
    if (rc != CTK_OK) { set_last_error(err); return rc; }
    if (m.split_stages.empty()) return -1;
    const SplitDfa& d = m.split_stages[0].dfa;
    if (n_states) *n_states = d.n_states;
    if (n_classes) *n_classes = d.n_classes;
    SplitTables t{d.trans.data(), d.ascii_class.data(), d.stage1.data(), d.blocks.data(), d.n_classes, d.start, behavior, invert,
                  d.trans_ascii.empty() ? nullptr : d.trans_ascii.data(), d.pair_impossible.data()};
    PtrReader rd{text};
    std::vector<uint64_t> cuts(n + 2), spans(2 * n + 4);
    HostPieces hp;
    hp.cuts = cuts.data(); hp.spans = spans.data();
    // the text is walked in the segments the device kernel would use when `segment` > 0 (safe starts after neutral bytes)
    {
        const uint32_t* nm = d.neutral;
        uint64_t seg_lo = 0;
        while (seg_lo < n) {
            uint64_t seg_hi = n;
            if (segment > 0) {
                const uint64_t x = (seg_lo / (uint64_t)segment + 1) * (uint64_t)segment;
                const bool pairs = behavior == 1 || behavior == 3 || (behavior == 0 && !invert);       // no look-back: cut between impossible pairs
                for (uint64_t p = x ? x - 1 : 0; p + 1 < n && x < n; ++p) {
                    if (text[p] < 128 && ((nm[text[p] >> 5] >> (text[p] & 31)) & 1u)) { seg_hi = p + 1; break; }
                    if (pairs && p + 1 >= x && text[p] < 128 && text[p + 1] < 128 &&
                        ((d.pair_impossible[text[p] * 4 + (text[p + 1] >> 5)] >> (text[p + 1] & 31)) & 1u)) { seg_hi = p + 1; break; }
                }
            }
            split_walk<uint64_t>(t, rd, (uint64_t)0, n, seg_lo, seg_hi, hp);
            seg_lo = seg_hi;
        }
        if (behavior == 0 && !invert && !hp.any && n) hp.span(0, n, true);       // pretokenizers.rs:305-307
    }
    std::vector<uint64_t> out;
    if (behavior == 0) out.assign(spans.begin(), spans.begin() + hp.ns);
    else if (n) {
        uint64_t prev = 0;
        for (size_t k = 0; k < hp.nc; ++k) { out.push_back(prev); out.push_back(cuts[k]); prev = cuts[k]; }
        out.push_back(prev); out.push_back(n);
    }
    *n_pairs = out.size() / 2;
    if (out.size() / 2 > cap_pairs) { set_last_error("capacity"); return CTK_ERR_ARG; }
    for (size_t i = 0; i < out.size(); ++i) pieces[i] = out[i];
    return 0;
}

int ctk_debug_cp_classes(uint32_t first, uint32_t count, uint8_t* out) {
    for (uint32_t i = 0; i < count; ++i) out[i] = (uint8_t)trie_nibble(CTK_TRIE_INDEX, CTK_TRIE_BLOCKS, first + i);
    return CTK_OK;
}

// model-only load (no device): returns the CTK_* code the loader gives, for CPU tests of the load rules
int ctk_debug_load_only(const uint8_t* json, size_t len, uint64_t* n_pairs, uint64_t* vocab_size, int* nfc, int* may_match) {
    HostModel m;
    std::string err;
    int rc = load_model(json, len, m, err);
    if (rc != CTK_OK) { set_last_error(err); return rc; }
    if (n_pairs) *n_pairs = m.pairs.size();
    if (vocab_size) *vocab_size = m.vocab.size();
    if (nfc) *nfc = m.nfc;
    if (may_match) *may_match = m.any_added_may_match;
    return CTK_OK;
}

// model-only load: the post-processor reduced to items (-1 = the ids, else a literal id), for CPU tests of the load rules
int ctk_debug_post_processor(const uint8_t* json, size_t len, int64_t* items, size_t cap, size_t* n) {
    HostModel m;
    std::string err;
    int rc = load_model(json, len, m, err);
    if (rc != CTK_OK) { set_last_error(err); return rc; }
    for (size_t i = 0; i < m.pp_items.size() && i < cap; ++i) items[i] = m.pp_items[i];
    if (n) *n = m.pp_items.size();
    return CTK_OK;
}

// model-only load: is the merge table monotone (round-parallel path for very long pre-tokens allowed), and the
// longest merged token in initial symbols
int ctk_debug_merge_props(const uint8_t* json, size_t len, int* monotone, uint32_t* max_span) {
    HostModel m;
    std::string err;
    int rc = load_model(json, len, m, err);
    if (rc != CTK_OK) { set_last_error(err); return rc; }
    if (monotone) *monotone = m.merges_monotone | (m.round_parallel ? 2 : 0);
    if (max_span) *max_span = m.max_token_span;
    return CTK_OK;
}

}  // extern "C"

// rounds of the last very-long-pre-token pass (diagnostics)
extern "C" int ctk_debug_xlong_rounds(ctk_tokenizer* tok) { return reinterpret_cast<ctk::Engine*>(tok)->xl_last_rounds; }

// debug: 16 counters the encode kernel keeps when CTK_ABLATE=9 (slow-path categories), see encode_fused.cu
extern "C" int ctk_debug_counters(ctk_tokenizer* tok, uint32_t* out16) {
    ctk::Engine* eng = reinterpret_cast<ctk::Engine*>(tok);
    if (!eng->ws.p[4]) return CTK_ERR_ARG;
    cudaSetDevice(eng->device);
    return cudaMemcpy(out16, (const uint32_t*)eng->ws.p[4] + 16, 64, cudaMemcpyDeviceToHost) == cudaSuccess ? CTK_OK : CTK_ERR_CUDA;
}
