/* ctk.h -- C ABI of the B200-native batched encode/decode path for complexity-tokenizer.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: its
 * boundary is the PyO3 class `Tokenizer` (src/bindings/tokenizer.rs:11-14) whose hot-path methods
 * forward to `HuggingFaceTokenizer` (src/huggingface/mod.rs).  Each entry point below names the
 * reference function it replaces; INTEGRATION.md shows the Rust `extern "C"` block and the
 * three-line bodies a maintainer would put in src/bindings/tokenizer.rs.
 *
 * Conventions
 *   - Batches are PACKED: one contiguous byte (or id) buffer plus n+1 uint64 offsets; item i is
 *     [off[i], off[i+1]).  No per-string pointers cross the boundary.
 *   - Text is UTF-8 and must be valid (the reference takes &str, which guarantees it).
 *   - A ctk_tokenizer is immutable after creation and may be used from several host threads.
 *   - All compute runs on the GPU.  There is no CPU fallback: if no usable device or the CUDA
 *     library is missing the call fails with CTK_ERR_CUDA.
 *   - Return value: 0 ok, else one of CTK_ERR_*; ctk_last_error() gives a thread-local message.
 */
#ifndef CTK_H
#define CTK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTK_OK 0
#define CTK_ERR_IO 1            /* file could not be read            (io::Error -> PyIOError)          */
#define CTK_ERR_INVALID_DATA 2  /* not a tokenizer.json / bad UTF-8  (io::ErrorKind::InvalidData)      */
#define CTK_ERR_UNSUPPORTED 3   /* pipeline outside the hot path (non-ByteLevel, NFKC, ...)            */
#define CTK_ERR_CUDA 4          /* CUDA failure, no device, out of memory                              */
#define CTK_ERR_ARG 5           /* bad argument (null pointer, unaligned device buffer, ...)           */
#define CTK_ERR_PANIC 6         /* the reference panics on this input (pyo3 PanicException); message says where */

typedef struct ctk_tokenizer ctk_tokenizer;   /* opaque */
typedef struct ctk_result ctk_result;         /* opaque, owns host (pinned) result buffers */

/* ---- load ---------------------------------------------------------------------------------
 * Replaces HuggingFaceTokenizer::from_file (src/huggingface/mod.rs:159-166) and ::from_str /
 * ::from_buffer (:169-180): parses tokenizer.json with the reference's rules (mod.rs:31-116,
 * :247-334; bpe.rs:52-79; parsing.rs defaults) and uploads the tables to `device`. */
int ctk_from_file(const char* path, int device, ctk_tokenizer** out);
int ctk_from_json(const uint8_t* json, size_t len, int device, ctk_tokenizer** out);
/* One tokenizer that drives SEVERAL devices from one process: the reference's encode_batch / decode_batch are single calls
 * that use the whole machine (rayon over documents, src/huggingface/mod.rs:694-696, :771-785).  Tables are replicated on
 * every listed device (n_devices <= 0: every visible device); the host-buffer entry points below cut a batch into
 * contiguous document ranges balanced by bytes, one per device, each served by a host thread bound to the device's NUMA
 * node.  No data crosses between devices.  The device-resident entry points use the first device. */
int ctk_from_file_devices(const char* path, int n_devices, const int* device_ids, ctk_tokenizer** out);
int ctk_from_json_devices(const uint8_t* json, size_t len, int n_devices, const int* device_ids, ctk_tokenizer** out);
size_t ctk_n_devices(const ctk_tokenizer* tok);
int ctk_device_at(const ctk_tokenizer* tok, size_t i);      /* i-th device, -1 if out of range */
int ctk_numa_node(const ctk_tokenizer* tok, size_t i);      /* host NUMA node next to the i-th device, -1 if unknown */
int ctk_device_numa_node(int device);                       /* the same for any visible device (sysfs: PCI bus id -> numa_node) */
void ctk_free(ctk_tokenizer* tok);

/* ---- cheap getters (src/bindings/tokenizer.rs:271-289) ------------------------------------- */
size_t ctk_vocab_size(const ctk_tokenizer* tok);                               /* mod.rs:856-858 */
/* mod.rs:860-862: returns 1 and sets *id if present, 0 if absent */
int ctk_token_to_id(const ctk_tokenizer* tok, const uint8_t* token, size_t len, uint32_t* id);
/* mod.rs:864-866: returns pointer to the token's UTF-8 bytes (owned by tok) or NULL */
const uint8_t* ctk_id_to_token(const ctk_tokenizer* tok, uint32_t id, size_t* len);
/* mod.rs:868-870: number of special tokens; i-th content/id */
size_t ctk_n_special_tokens(const ctk_tokenizer* tok);
const uint8_t* ctk_special_token(const ctk_tokenizer* tok, size_t i, size_t* len, uint32_t* id);
int ctk_device(const ctk_tokenizer* tok);

/* ---- encode_batch, host buffers ------------------------------------------------------------
 * Replaces HuggingFaceTokenizer::encode_batch (src/huggingface/mod.rs:694-696) and, with n = 1,
 * ::encode (:551-613).  text/text_off are host memory (pinned memory is copied by DMA directly).
 * On success *res owns: ids (uint32, packed) and ids_off (uint64[n+1]).  Bit-exact with the
 * reference's Vec<Vec<u32>>. */
int ctk_encode_batch(const ctk_tokenizer* tok, const uint8_t* text, const uint64_t* text_off, size_t n,
                     ctk_result** res);
const uint32_t* ctk_result_ids(const ctk_result* res);
const uint64_t* ctk_result_offsets(const ctk_result* res);   /* n+1 entries */
size_t ctk_result_count(const ctk_result* res);              /* n */

/* ---- narrow ids and per-device parts (what a binding should use for large batches) ----------------
 * The copy of the ids back to the host is as large as the copy of the text to the device.  When every id the tokenizer
 * can emit is below 65 536 (ctk_id_width(tok) == 2) ctk_encode_batch_narrow writes and returns uint16 ids: half the
 * device-to-host bytes, and nothing is widened on the host -- the binding's per-document copy into its own Vec<u32>
 * (src/bindings: Vec<Vec<u32>> -> list[list[int]]) reads the 16-bit ids directly.  Otherwise it equals ctk_encode_batch.
 *   ctk_result_id_width  2 or 4: element size of the ids of this result
 *   ctk_result_parts     1, or the number of devices that received documents (ctk_from_file_devices)
 *   ctk_result_part      part i: its first item, item count, ids (id_width each; bytes for a decode result) and
 *                        offsets[n_items + 1] RELATIVE to the part.  Parts are in item order.  Zero-copy.
 * ctk_result_ids / ctk_result_offsets / ctk_result_bytes still work on any result: they gather the parts (and widen
 * narrow ids) into one buffer on first use. */
int ctk_id_width(const ctk_tokenizer* tok);
int ctk_encode_batch_narrow(const ctk_tokenizer* tok, const uint8_t* text, const uint64_t* text_off, size_t n,
                            ctk_result** res);
int ctk_result_id_width(const ctk_result* res);
const void* ctk_result_ids_raw(const ctk_result* res);       /* single-part results; NULL when there are several parts */
size_t ctk_result_parts(const ctk_result* res);
int ctk_result_part(const ctk_result* res, size_t i, size_t* first_item, size_t* n_items, const void** data,
                    const uint64_t** offsets);

/* ---- decode_batch, host buffers ------------------------------------------------------------
 * Replaces HuggingFaceTokenizer::decode_batch_with_options (src/huggingface/mod.rs:775-785);
 * decode_batch (:771-773) is skip_special_tokens=0, clean_up_tokenization_spaces=1; decode /
 * decode_with_options (:698-709) are n = 1.  Result: UTF-8 bytes (packed) + uint64[n+1] offsets. */
int ctk_decode_batch(const ctk_tokenizer* tok, const uint32_t* ids, const uint64_t* ids_off, size_t n,
                     int skip_special_tokens, int clean_up_tokenization_spaces, ctk_result** res);
const uint8_t* ctk_result_bytes(const ctk_result* res);

void ctk_result_free(ctk_result* res);

/* ---- device-resident variants (inputs already in HBM, outputs left in HBM) -------------------
 * Same semantics; every pointer is a device pointer on ctk_device(tok).  Used for the roofline
 * measurement and by callers that keep corpora on the GPU.
 *   d_text        16-byte aligned, total_bytes = text_off[n]; d_text_off uint64[n+1]
 *   d_ids         capacity ids_cap uint32 (total_bytes + n always suffices: at most 1 id per byte)
 *   d_ids_off     uint64[n+1]; d_ids_off[n] = total ids
 * `stream` is a cudaStream_t (NULL = legacy default stream).  The call enqueues work and returns;
 * *n_ids_host, if not NULL, makes the call synchronise and receive the total id count.
 * Scratch memory is owned by the tokenizer and grown on demand (one concurrent device call per
 * tokenizer; the host-buffer entry points above serialise internally). */
int ctk_encode_batch_device(const ctk_tokenizer* tok, const uint8_t* d_text, const uint64_t* d_text_off,
                            size_t n, uint64_t total_bytes, uint32_t* d_ids, uint64_t ids_cap,
                            uint64_t* d_ids_off, uint64_t* n_ids_host, void* stream);
/*   d_text_out    capacity text_cap bytes; d_text_off_out uint64[n+1].  Required capacity is at
 *                 most ctk_decode_max_bytes(tok) * total_ids (3x that if ids decode to invalid
 *                 UTF-8 that must be replaced by U+FFFD). */
int ctk_decode_batch_device(const ctk_tokenizer* tok, const uint32_t* d_ids, const uint64_t* d_ids_off,
                            size_t n, uint64_t total_ids, int skip_special_tokens,
                            int clean_up_tokenization_spaces, uint8_t* d_text_out, uint64_t text_cap,
                            uint64_t* d_text_off_out, uint64_t* n_bytes_host, void* stream);
size_t ctk_decode_max_bytes(const ctk_tokenizer* tok);   /* longest decoded token, in bytes */
/* ctk_encode_batch_device with the element size of d_ids chosen by the caller: id_width 4, or 2 when ctk_id_width(tok) == 2 */
int ctk_encode_batch_device_ex(const ctk_tokenizer* tok, const uint8_t* d_text, const uint64_t* d_text_off,
                               size_t n, uint64_t total_bytes, void* d_ids, uint64_t ids_cap, int id_width,
                               uint64_t* d_ids_off, uint64_t* n_ids_host, void* stream);

/* ---- rich `Encoding` outputs ----------------------------------------
 * Replaces, for a whole batch and on the GPU,
 *   HuggingFaceTokenizer::encode_to_encoding / encode_pair_to_encoding / encode_batch_to_encoding /
 *   encode_batch_pairs_to_encoding                       (src/huggingface/mod.rs:340-395, :481-488)
 *   encode_batch_with_padding / encode_batch_pairs_with_padding            (mod.rs:490-545)
 *   PyTokenizer::__call__ (truncate + pad of a batch)     (src/bindings/tokenizer.rs:33-201)
 * with the arithmetic of Encoding::{from_ids, merge, mark_special_tokens, truncate, pad} (src/encoding.rs:44-267)
 * and PostProcessor::process(ids, None) (src/postprocessors.rs:34-188).
 *
 * A ROW is one Encoding: one text, or (pair != 0) texts 2r and 2r+1 merged.  All rows come back packed: `row_off`
 * (n_rows + 1) indexes input_ids / attention_mask / token_type_ids / special_tokens_mask; with padding every row has
 * the same length and the arrays are a dense row-major matrix.  Offsets and word ids exist per token BEFORE
 * post-processing (the reference neither shifts nor pads them, mod.rs:372-383, encoding.rs:86-129) and are indexed by
 * `token_off` (n_texts + 1).  Offsets are BYTE positions in the original text, found the way
 * pre_tokenize_with_offsets does (mod.rs:447-478: str::find of the byte-mapped word from a running position). */
typedef struct ctk_encoding_options {
    int add_special_tokens;   /* 1: encode_to_encoding (no added-token scan, post-processor, special mask);
                                 0: encode + Encoding::from_ids (bindings/tokenizer.rs:88-96), no offsets */
    int pair;                 /* 1: texts 2r, 2r+1 are (text, text_pair) of row r (Encoding::merge, type id 1) */
    int truncation;           /* 1: rows longer than max_length keep their first max_length entries (encoding.rs:132-232) */
    uint64_t max_length;
    int padding;              /* 0 none; 1 to the longest row of the batch; 2 to pad_to (encoding.rs:86-129) */
    uint64_t pad_to;
    int pad_left;
    int want_offsets;         /* 1: also compute offsets + word ids (only with add_special_tokens = 1) */
} ctk_encoding_options;
typedef struct ctk_encodings ctk_encodings;   /* opaque, owns host (pinned) result buffers */
int ctk_encode_batch_to_encoding(const ctk_tokenizer* tok, const uint8_t* text, const uint64_t* text_off, size_t n_texts,
                                 const ctk_encoding_options* opt, ctk_encodings** res);
size_t ctk_encodings_rows(const ctk_encodings* res);
const uint64_t* ctk_encodings_row_offsets(const ctk_encodings* res);       /* n_rows + 1 */
const uint64_t* ctk_encodings_row_full_lengths(const ctk_encodings* res);  /* n_rows: before truncation / padding */
const uint32_t* ctk_encodings_input_ids(const ctk_encodings* res);
const uint8_t* ctk_encodings_attention_mask(const ctk_encodings* res);
const uint8_t* ctk_encodings_type_ids(const ctk_encodings* res);
const uint8_t* ctk_encodings_special_tokens_mask(const ctk_encodings* res);
const uint64_t* ctk_encodings_token_offsets(const ctk_encodings* res);     /* n_texts + 1 */
const uint32_t* ctk_encodings_token_ids(const ctk_encodings* res);         /* ids before post-processing */
const uint32_t* ctk_encodings_offsets(const ctk_encodings* res);           /* 2 per token: start, end; NULL if not computed */
const uint32_t* ctk_encodings_word_ids(const ctk_encodings* res);          /* NULL if not computed */
void ctk_encodings_free(ctk_encodings* res);
/* What the loaded post-processor does to one sequence: writes up to cap items (-1 = the ids, else a literal id), returns the count. */
size_t ctk_post_processor_items(const ctk_tokenizer* tok, int64_t* items, size_t cap);
uint32_t ctk_pad_token(const ctk_tokenizer* tok, const uint8_t** token, size_t* len);   /* mod.rs:504-509 */

/* ---- BPE training ----------------------------------------------------------
 * Replaces BpeTrainer::train (src/bpe_trainer.rs:100-228; Python: BpeTrainer.train, src/bindings/trainers.rs:261-269).
 * texts are packed like everywhere else; words are split on Unicode White_Space (:248).  The word histogram and the
 * whole merge loop run on `device`.  Where the reference decides in hash-iteration order (:152-155 best pair,
 * :305-306 characters of equal frequency) this library decides by (count, left symbol index, right symbol index) and
 * by code point -- every output is one the reference can produce; oracle/py_trainer.py states the rule.
 * Fields mirror BpeTrainerConfig (:13-31). */
typedef struct ctk_bpe_trainer_config {
    uint64_t vocab_size;                       /* :16 */
    uint32_t min_frequency;                    /* :18 */
    const uint8_t* special_tokens;             /* :20, packed UTF-8 */
    const uint64_t* special_off;               /* n_special + 1 */
    size_t n_special;
    const uint32_t* initial_alphabet;          /* :24, code points; NULL = None */
    size_t n_alphabet;
    int64_t limit_alphabet;                    /* :26; -1 = None */
    const uint8_t* continuing_subword_prefix;  /* :28; NULL = None */
    size_t prefix_len;
    const uint8_t* end_of_word_suffix;         /* :30; NULL = None */
    size_t suffix_len;
} ctk_bpe_trainer_config;
typedef struct ctk_trained ctk_trained;        /* opaque: the (vocab, merges) pair train() returns */
typedef struct ctk_train_stats {
    uint64_t n_bytes, n_words, n_unique_words, n_symbols, n_merges, kernel_launches;
    uint32_t table_rebuilds;                   /* times the pair table grew (full recount each time) */
    uint32_t cluster_size;                     /* CTAs of the cluster the merge loop ran in; 0 = three kernels per merge */
    uint32_t stop_reason;                      /* 0 none, 1 no pairs left (:147), 2 below min_frequency (:162), 3 vocabulary full (:141) */
    double ms_words, ms_merges;                /* stream time of the word histogram (includes waiting for device allocations) / of the merge loop (CUDA events) */
    double ms_words_kernels;                   /* word histogram: kernels only */
} ctk_train_stats;
int ctk_train_bpe(const ctk_bpe_trainer_config* cfg, int device, const uint8_t* text, const uint64_t* text_off, size_t n_texts,
                  ctk_trained** out);
/* Replaces the training half of HuggingFaceTokenizer::train_new_from_iterator (src/huggingface/mod.rs:1231-1275): the texts go
 * through THIS tokenizer's normaliser and pre-tokenizer on the device (mod.rs:1257-1270) and the trainer runs on the pre-tokens
 * (each byte-mapped pre-token is one word, bpe_trainer.rs:248) -- no pre-tokenisation on the CPU.  The caller passes the
 * trainer configuration mod.rs:1244-1252 builds (vocab_size, min_frequency 2, special_tokens = all_special_tokens()) and
 * assembles the new tokenizer from the result (mod.rs:1276-1322).  Texts are host memory, packed as everywhere else. */
int ctk_train_new_from_texts(const ctk_tokenizer* tok, const ctk_bpe_trainer_config* cfg, const uint8_t* text, const uint64_t* text_off,
                             size_t n_texts, ctk_trained** out);
/* Every symbol string the training knows, by symbol index, and its id in the returned vocabulary map (-1: not in it).  Returns the count. */
size_t ctk_trained_symbols(const ctk_trained* t, const uint8_t** bytes, const uint64_t** off, const int64_t** vocab_id);
/* Merges in order, two symbol indices each.  Returns the count. */
size_t ctk_trained_merges(const ctk_trained* t, const uint32_t** pairs);
/* Count of each merge's pair at the moment it was chosen (bpe_trainer.rs:161); non-increasing along the list. */
const uint32_t* ctk_trained_merge_counts(const ctk_trained* t);
void ctk_trained_stats(const ctk_trained* t, ctk_train_stats* stats);
void ctk_trained_free(ctk_trained* t);

/* ---- diagnostics --------------------------------------------------------------------------- */
const char* ctk_last_error(void);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t ctk_kernel_launches(void);
/* Per-kernel timing with CUDA events on the launching stream (off by default).  The report is text,
 * one line per kernel: name <TAB> total milliseconds <TAB> launches; returns the full length. */
void ctk_profile_enable(ctk_tokenizer* tok, int on);
size_t ctk_profile_report(ctk_tokenizer* tok, char* buf, size_t cap);
/* Bytes the last ctk_encode_batch / ctk_encode_batch_narrow call moved host->device and device->host (all devices). */
void ctk_last_transfer_bytes(const ctk_tokenizer* tok, uint64_t* h2d, uint64_t* d2h);
/* Reset the per-batch pre-token cache policy: 0 = clear at the start of every encode call
 * (default; every call does all of its own work), 1 = keep entries across calls. */
void ctk_set_cache_persistent(ctk_tokenizer* tok, int persistent);

#ifdef __cplusplus
}
#endif
#endif /* CTK_H */
