/* ORACLE (test infrastructure, not product code) -- C core.
 *
 * CPU restatement of the reference's encode_batch / decode_batch hot loops
 * (Complexity-ML/complexity-tokenizer v0.3.3), with the reference's data-structure choices
 * (hash-map rank probes, full pair rescan per merge, one string per pre-token, per-document NFC)
 * so that it can also serve as the all-core CPU baseline ("port") in bench.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The tokenizer.json load rules live in oracle/py_oracle.py (the twin), which
 * hands this core already-built tables; tests check core == twin.
 *
 * PARITY PINNING: see oracle/py_oracle.py header (reference cannot be built here; pinned on the
 * reference's own known answers + HF tokenizers cross-check).
 *
 * Restated reference lines (relative to /root/reference):
 *   normalizers.rs:47 (NFC)                        -> nfc_normalize()
 *   pretokenizers.rs:13 (pattern), :158-185        -> find_iter_next(), encode_doc()
 *   huggingface/mod.rs:551-613, :616-675           -> encode_doc(), find_added()
 *   bpe.rs:88-153                                  -> bpe_word()
 *   huggingface/mod.rs:694-696, :771-785 (par_iter)-> orc_encode_batch(), orc_decode_batch()
 *   huggingface/mod.rs:711-769, decoders.rs:94-119 -> decode_doc(), utf8_lossy(), clean_up()
 *
 * Build: gcc -O2 -std=c11 -shared -fPIC -pthread oracle/oracle_core.c -o oracle/_build/liboracle.so
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "unicode_ranges_gen.h"

/* ------------------------------------------------------------------ small vectors */
typedef struct { uint32_t* p; size_t n, cap; } vec32;
typedef struct { uint8_t* p; size_t n, cap; } vec8;
static void v32_push(vec32* v, uint32_t x) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 64; v->p = (uint32_t*)realloc(v->p, v->cap * 4); }
    v->p[v->n++] = x;
}
static void v8_reserve(vec8* v, size_t extra) {
    if (v->n + extra > v->cap) { while (v->n + extra > v->cap) v->cap = v->cap ? v->cap * 2 : 256; v->p = (uint8_t*)realloc(v->p, v->cap); }
}
static void v8_push(vec8* v, uint8_t x) { v8_reserve(v, 1); v->p[v->n++] = x; }
static void v8_append(vec8* v, const uint8_t* s, size_t n) { v8_reserve(v, n); memcpy(v->p + v->n, s, n); v->n += n; }
static void v8_put_utf8(vec8* v, uint32_t cp) {
    if (cp < 0x80) v8_push(v, (uint8_t)cp);
    else if (cp < 0x800) { v8_push(v, 0xC0 | (cp >> 6)); v8_push(v, 0x80 | (cp & 63)); }
    else if (cp < 0x10000) { v8_push(v, 0xE0 | (cp >> 12)); v8_push(v, 0x80 | ((cp >> 6) & 63)); v8_push(v, 0x80 | (cp & 63)); }
    else { v8_push(v, 0xF0 | (cp >> 18)); v8_push(v, 0x80 | ((cp >> 12) & 63)); v8_push(v, 0x80 | ((cp >> 6) & 63)); v8_push(v, 0x80 | (cp & 63)); }
}

/* ------------------------------------------------------------------ Unicode lookups */
static int range_lookup(const uint32_t (*tab)[3], int n, uint32_t cp) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if (cp < tab[mid][0]) hi = mid - 1;
        else if (cp > tab[mid][1]) lo = mid + 1;
        else return (int)tab[mid][2];
    }
    return 0;
}
static int cp_class(uint32_t cp) { return range_lookup(ORC_CLASS_RANGES, ORC_N_CLASS_RANGES, cp); }
static int cp_ccc(uint32_t cp) { return cp < 0x300 ? 0 : range_lookup(ORC_CCC_RANGES, ORC_N_CCC_RANGES, cp); }
static const uint32_t* cp_decomp(uint32_t cp) {
    int lo = 0, hi = ORC_N_DECOMP - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if (cp < ORC_DECOMP[mid][0]) hi = mid - 1;
        else if (cp > ORC_DECOMP[mid][0]) lo = mid + 1;
        else return ORC_DECOMP[mid];
    }
    return NULL;
}
static uint32_t cp_compose(uint32_t a, uint32_t b) {
    /* Hangul */
    if (a >= 0x1100 && a < 0x1113 && b >= 0x1161 && b < 0x1176) return 0xAC00 + ((a - 0x1100) * 21 + (b - 0x1161)) * 28;
    if (a >= 0xAC00 && a < 0xD7A4 && (a - 0xAC00) % 28 == 0 && b > 0x11A7 && b < 0x11C3) return a + (b - 0x11A7);
    int lo = 0, hi = ORC_N_COMP - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        const uint32_t* e = ORC_COMP[mid];
        if (a < e[0] || (a == e[0] && b < e[1])) hi = mid - 1;
        else if (a > e[0] || (a == e[0] && b > e[1])) lo = mid + 1;
        else return e[2];
    }
    return 0;
}

static uint32_t utf8_next(const uint8_t* s, size_t n, size_t* i) {   /* input is valid UTF-8 (&str) */
    uint8_t c = s[*i];
    if (c < 0x80) { (*i)++; return c; }
    if (c < 0xE0 && *i + 1 < n + 0) { uint32_t r = ((c & 0x1F) << 6) | (s[*i + 1] & 63); *i += 2; return r; }
    if (c < 0xF0) { uint32_t r = ((c & 0x0F) << 12) | ((s[*i + 1] & 63) << 6) | (s[*i + 2] & 63); *i += 3; return r; }
    uint32_t r = ((c & 7) << 18) | ((s[*i + 1] & 63) << 12) | ((s[*i + 2] & 63) << 6) | (s[*i + 3] & 63);
    *i += 4;
    return r;
}

/* ------------------------------------------------------------------ NFC (UAX #15) */
static void decompose_push(vec32* out, uint32_t cp) {
    if (cp >= 0xAC00 && cp < 0xD7A4) {
        uint32_t s = cp - 0xAC00;
        v32_push(out, 0x1100 + s / 588);
        v32_push(out, 0x1161 + (s % 588) / 28);
        if (s % 28) v32_push(out, 0x11A7 + s % 28);
        return;
    }
    const uint32_t* d = cp < 0xC0 ? NULL : cp_decomp(cp);
    if (!d) { v32_push(out, cp); return; }
    decompose_push(out, d[1]);
    if (d[2]) decompose_push(out, d[2]);
}
/* returns 1 and fills `out` if normalisation changed anything; 0 if text is already NFC */
static int nfc_normalize(const uint8_t* s, size_t n, vec8* out, vec32* tmp) {
    size_t i;
    for (i = 0; i < n; ++i) if (s[i] >= 0xCC) break;     /* all code points < U+0300: NFC-stable */
    if (i == n) return 0;
    tmp->n = 0;
    for (i = 0; i < n;) decompose_push(tmp, utf8_next(s, n, &i));
    uint32_t* a = tmp->p;
    size_t m = tmp->n;
    /* canonical ordering */
    for (size_t k = 1; k < m; ++k) {
        int c = cp_ccc(a[k]);
        if (!c) continue;
        size_t j = k;
        uint32_t v = a[k];
        while (j > 0) { int cj = cp_ccc(a[j - 1]); if (cj <= c) break; a[j] = a[j - 1]; --j; }   /* cj==0 stops (0<=c) */
        a[j] = v;
    }
    /* canonical composition */
    size_t w = 0;
    long starter = -1;      /* index in a[0..w) of the last starter, -1 if none yet */
    int prev_cc = 0;        /* ccc of the last kept char; 0 <=> that char is the starter itself */
    for (size_t k = 0; k < m; ++k) {
        uint32_t c = a[k];
        int cc = cp_ccc(c);
        if (starter >= 0 && (prev_cc == 0 || prev_cc < cc)) {      /* not blocked (UAX #15 D115) */
            uint32_t comp = cp_compose(a[starter], c);
            if (comp) { a[starter] = comp; continue; }
        }
        if (cc == 0) starter = (long)w;
        prev_cc = cc;
        a[w++] = c;
    }
    out->n = 0;
    for (size_t k = 0; k < w; ++k) v8_put_utf8(out, a[k]);
    return !(out->n == n && memcmp(out->p, s, n) == 0);
}

/* ------------------------------------------------------------------ tokenizer tables */
typedef struct { uint64_t key; uint32_t rank; uint32_t used; } pair_slot;
typedef struct {
    /* bpe.rs: merge_ranks (pair -> rank) as an open-addressing hash, merges[rank].new_id compacted */
    pair_slot* pairs; uint64_t pair_mask;
    uint32_t* merge_new_id; size_t n_ops;
    /* single-char vocab lookups: mapped code point (< 0x180) -> id or -1 */
    int64_t char_id[0x180];
    /* byte -> mapped code point (pretokenizers.rs:130-153) */
    uint32_t byte_cp[256];
    /* added tokens (content as UTF-8) */
    size_t n_added; uint8_t** added; size_t* added_len; uint32_t* added_id; uint8_t* added_flags; /* 1 single_word 2 lstrip 4 rstrip */
    int nfc, add_prefix_space;
    /* decode: id -> token string (UTF-8), special flag */
    uint32_t max_id; uint8_t** tok; uint32_t* tok_len; uint8_t* tok_special;
} orc_tok;

static uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

static int pair_get(const orc_tok* t, uint32_t a, uint32_t b, uint32_t* rank) {
    uint64_t key = ((uint64_t)a << 32) | b, h = mix64(key) & t->pair_mask;
    for (;;) {
        const pair_slot* s = &t->pairs[h];
        if (!s->used) return 0;
        if (s->key == key) { *rank = s->rank; return 1; }
        h = (h + 1) & t->pair_mask;
    }
}

orc_tok* orc_new(const uint32_t* pair_a, const uint32_t* pair_b, const uint32_t* pair_rank, size_t n_pairs,
                 const uint32_t* merge_new_id, size_t n_ops,
                 const int64_t* char_id /* 0x180 */,
                 const uint8_t* added_blob, const uint64_t* added_off, const uint32_t* added_id,
                 const uint8_t* added_flags, size_t n_added,
                 const uint8_t* tok_blob, const uint64_t* tok_off, const uint32_t* tok_ids, const uint8_t* tok_special,
                 size_t n_tok, int nfc, int add_prefix_space) {
    orc_tok* t = (orc_tok*)calloc(1, sizeof(orc_tok));
    uint64_t cap = 16;
    while (cap < n_pairs * 2 + 2) cap <<= 1;
    t->pairs = (pair_slot*)calloc(cap, sizeof(pair_slot));
    t->pair_mask = cap - 1;
    for (size_t i = 0; i < n_pairs; ++i) {
        uint64_t key = ((uint64_t)pair_a[i] << 32) | pair_b[i], h = mix64(key) & t->pair_mask;
        while (t->pairs[h].used && t->pairs[h].key != key) h = (h + 1) & t->pair_mask;
        t->pairs[h].key = key; t->pairs[h].rank = pair_rank[i]; t->pairs[h].used = 1;
    }
    t->merge_new_id = (uint32_t*)malloc((n_ops + 1) * 4);
    memcpy(t->merge_new_id, merge_new_id, n_ops * 4);
    t->n_ops = n_ops;
    memcpy(t->char_id, char_id, sizeof(t->char_id));
    {   /* bytes_to_unicode */
        int n = 0;
        for (int b = 0; b < 256; ++b) {
            int keep = (b >= '!' && b <= '~') || (b >= 0xA1 && b <= 0xAC) || (b >= 0xAE);
            t->byte_cp[b] = keep ? (uint32_t)b : (uint32_t)(256 + n++);
        }
    }
    t->n_added = n_added;
    t->added = (uint8_t**)calloc(n_added + 1, sizeof(uint8_t*));
    t->added_len = (size_t*)calloc(n_added + 1, sizeof(size_t));
    t->added_id = (uint32_t*)calloc(n_added + 1, 4);
    t->added_flags = (uint8_t*)calloc(n_added + 1, 1);
    for (size_t i = 0; i < n_added; ++i) {
        size_t L = (size_t)(added_off[i + 1] - added_off[i]);
        t->added[i] = (uint8_t*)malloc(L + 1);
        memcpy(t->added[i], added_blob + added_off[i], L);
        t->added_len[i] = L; t->added_id[i] = added_id[i]; t->added_flags[i] = added_flags[i];
    }
    t->nfc = nfc; t->add_prefix_space = add_prefix_space;
    uint32_t mx = 0;
    for (size_t i = 0; i < n_tok; ++i) if (tok_ids[i] > mx) mx = tok_ids[i];
    t->max_id = n_tok ? mx : 0;
    t->tok = (uint8_t**)calloc((size_t)mx + 2, sizeof(uint8_t*));
    t->tok_len = (uint32_t*)calloc((size_t)mx + 2, 4);
    t->tok_special = (uint8_t*)calloc((size_t)mx + 2, 1);
    for (size_t i = 0; i < n_tok; ++i) {
        size_t L = (size_t)(tok_off[i + 1] - tok_off[i]);
        uint32_t id = tok_ids[i];
        free(t->tok[id]);
        t->tok[id] = (uint8_t*)malloc(L + 1);
        memcpy(t->tok[id], tok_blob + tok_off[i], L);
        t->tok_len[id] = (uint32_t)L; t->tok_special[id] = tok_special[i];
    }
    return t;
}

void orc_free(orc_tok* t) {
    if (!t) return;
    free(t->pairs); free(t->merge_new_id);
    for (size_t i = 0; i < t->n_added; ++i) free(t->added[i]);
    free(t->added); free(t->added_len); free(t->added_id); free(t->added_flags);
    if (t->tok) for (uint32_t i = 0; i <= t->max_id; ++i) free(t->tok[i]);
    free(t->tok); free(t->tok_len); free(t->tok_special);
    free(t);
}

/* ------------------------------------------------------------------ pre-token pattern
 * 's|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+   leftmost-first, at byte i. */
static int class_at(const uint8_t* s, size_t n, size_t i, size_t* next) {
    size_t j = i;
    uint32_t cp = utf8_next(s, n, &j);
    *next = j;
    if (cp < 0x80) {
        if ((cp | 0x20) >= 'a' && (cp | 0x20) <= 'z') return 1;
        if (cp >= '0' && cp <= '9') return 2;
        if (cp == ' ' || (cp >= 9 && cp <= 13)) return 3;
        return 0;
    }
    return cp_class(cp);
}
static size_t run_of(const uint8_t* s, size_t n, size_t i, int want) {
    while (i < n) { size_t nx; if (class_at(s, n, i, &nx) != want) break; i = nx; }
    return i;
}
static size_t find_iter_next(const uint8_t* s, size_t n, size_t i) {
    if (s[i] == '\'' && i + 1 < n) {
        uint8_t c1 = s[i + 1], c2 = i + 2 < n ? s[i + 2] : 0;
        if (c1 == 's') return i + 2;
        if (c1 == 't') return i + 2;
        if (c1 == 'r' && c2 == 'e') return i + 3;
        if (c1 == 'v' && c2 == 'e') return i + 3;
        if (c1 == 'm') return i + 2;
        if (c1 == 'l' && c2 == 'l') return i + 3;
        if (c1 == 'd') return i + 2;
    }
    static const int order[3] = {1, 2, 0};
    for (int a = 0; a < 3; ++a) {
        if (s[i] == ' ') { size_t e = run_of(s, n, i + 1, order[a]); if (e > i + 1) return e; }
        size_t e = run_of(s, n, i, order[a]);
        if (e > i) return e;
    }
    size_t e = run_of(s, n, i, 3);
    return e > i ? e : i + 1;
}

/* ------------------------------------------------------------------ added tokens (mod.rs:637-675) */
static int cp_is_alnum_mapped(uint32_t cp) { int c = cp_class(cp); return c == 1 || c == 2; }
static int cp_is_ws(uint32_t cp) { return cp_class(cp) == 3; }
static uint32_t last_cp_before(const uint8_t* s, size_t pos) {
    size_t k = pos - 1;
    while (k > 0 && (s[k] & 0xC0) == 0x80) --k;
    size_t j = k;
    return utf8_next(s, pos, &j);
}
/* returns position or -1 */
static long find_added(const orc_tok* t, size_t ti, const uint8_t* text, size_t n) {
    size_t L = t->added_len[ti];
    if (L > n) return -1;
    const uint8_t* f = (const uint8_t*)memmem(text, n, t->added[ti], L);
    if (!f) return -1;
    size_t pos = (size_t)(f - text), end = pos + L;
    uint8_t fl = t->added_flags[ti];
    if (fl & 1) {
        int before_ok = pos == 0 || !cp_is_alnum_mapped(last_cp_before(text, pos));
        size_t j = end;
        int after_ok = end >= n || !cp_is_alnum_mapped(utf8_next(text, n, &j));
        if (!before_ok || !after_ok) return -1;
    }
    if ((fl & 2) && pos > 0 && !cp_is_ws(last_cp_before(text, pos))) return -1;
    if ((fl & 4) && end < n) { size_t j = end; if (!cp_is_ws(utf8_next(text, n, &j))) return -1; }
    return (long)pos;
}

/* ------------------------------------------------------------------ BPE (bpe.rs:88-153) */
static void bpe_word(const orc_tok* t, const uint8_t* w, size_t n, vec32* out, vec32* tmp) {
    tmp->n = 0;
    for (size_t i = 0; i < n;) {
        uint32_t cp = utf8_next(w, n, &i);
        if (cp < 0x180 && t->char_id[cp] >= 0) v32_push(tmp, (uint32_t)t->char_id[cp]);
    }
    uint32_t* a = tmp->p;
    size_t m = tmp->n;
    for (;;) {
        long best = -1; uint32_t best_rank = 0;
        for (size_t i = 0; i + 1 < m; ++i) {
            uint32_t r;
            if (pair_get(t, a[i], a[i + 1], &r)) if (best < 0 || r < best_rank) { best = (long)i; best_rank = r; }
        }
        if (best < 0) break;
        if (best_rank >= t->n_ops) abort();            /* the reference panics here */
        a[best] = t->merge_new_id[best_rank];
        memmove(a + best + 1, a + best + 2, (m - (size_t)best - 2) * 4);
        --m;
    }
    for (size_t i = 0; i < m; ++i) v32_push(out, a[i]);
}

typedef struct { vec8 norm, word, pre; vec32 tmp, tmp2; } scratch;

static void encode_doc(const orc_tok* t, const uint8_t* text, size_t n, vec32* out, scratch* sc) {
    const uint8_t* s = text;
    if (t->nfc && nfc_normalize(text, n, &sc->norm, &sc->tmp2)) { s = sc->norm.p; n = sc->norm.n; }
    if (t->add_prefix_space && n && s[0] != ' ') {
        sc->pre.n = 0; v8_push(&sc->pre, ' '); v8_append(&sc->pre, s, n);
        s = sc->pre.p; n = sc->pre.n;
    }
    for (size_t i = 0; i < n;) {
        size_t e = find_iter_next(s, n, i);
        sc->word.n = 0;
        for (size_t k = i; k < e; ++k) v8_put_utf8(&sc->word, t->byte_cp[s[k]]);
        i = e;
        const uint8_t* rem = sc->word.p;
        size_t rn = sc->word.n;
        while (rn) {
            long best = -1;
            for (size_t a = 0; a < t->n_added; ++a)
                if (find_added(t, a, rem, rn) == 0 && (best < 0 || t->added_len[a] > t->added_len[best])) best = (long)a;
            if (best >= 0) { v32_push(out, t->added_id[best]); rem += t->added_len[best]; rn -= t->added_len[best]; continue; }
            size_t nxt = rn;
            for (size_t a = 0; a < t->n_added; ++a) { long p = find_added(t, a, rem, rn); if (p > 0 && (size_t)p < nxt) nxt = (size_t)p; }
            bpe_word(t, rem, nxt, out, &sc->tmp);
            rem += nxt; rn -= nxt;
        }
    }
}

/* ------------------------------------------------------------------ decode */
static void utf8_lossy(const uint8_t* b, size_t n, vec8* out) {   /* String::from_utf8_lossy: maximal subparts -> U+FFFD */
    size_t i = 0;
    while (i < n) {
        uint8_t c = b[i];
        if (c < 0x80) { v8_push(out, c); ++i; continue; }
        size_t need = 0; uint8_t lo = 0x80, hi = 0xBF;
        if (c >= 0xC2 && c <= 0xDF) need = 1;
        else if (c == 0xE0) { need = 2; lo = 0xA0; }
        else if (c >= 0xE1 && c <= 0xEC) need = 2;
        else if (c == 0xED) { need = 2; hi = 0x9F; }
        else if (c >= 0xEE && c <= 0xEF) need = 2;
        else if (c == 0xF0) { need = 3; lo = 0x90; }
        else if (c >= 0xF1 && c <= 0xF3) need = 3;
        else if (c == 0xF4) { need = 3; hi = 0x8F; }
        if (!need) { v8_put_utf8(out, 0xFFFD); ++i; continue; }
        size_t k = 1; int ok = 1;
        for (; k <= need; ++k) {
            if (i + k >= n) { ok = 0; break; }
            uint8_t d = b[i + k];
            uint8_t l = k == 1 ? lo : 0x80, h = k == 1 ? hi : 0xBF;
            if (d < l || d > h) { ok = 0; break; }
        }
        if (ok) { v8_append(out, b + i, need + 1); i += need + 1; }
        else { v8_put_utf8(out, 0xFFFD); i += k; }
    }
}
static void replace_all(vec8* src, vec8* dst, const char* pat, const char* rep) {
    size_t pl = strlen(pat), rl = strlen(rep);
    dst->n = 0;
    size_t i = 0;
    while (i < src->n) {
        if (i + pl <= src->n && memcmp(src->p + i, pat, pl) == 0) { v8_append(dst, (const uint8_t*)rep, rl); i += pl; }
        else v8_push(dst, src->p[i++]);
    }
    vec8 t = *src; *src = *dst; *dst = t;
}
static void clean_up(vec8* s, vec8* tmp) {     /* mod.rs:749-769 */
    static const char* R[15][2] = {{" .", "."}, {" ,", ","}, {" !", "!"}, {" ?", "?"}, {" :", ":"}, {" ;", ";"},
        {"\" ", "\""}, {" \"", "\""}, {"' ", "'"}, {" '", "'"}, {"( ", "("}, {" )", ")"}, {"[ ", "["}, {" ]", "]"}, {" - ", "-"}};
    for (int k = 0; k < 15; ++k) replace_all(s, tmp, R[k][0], R[k][1]);
    tmp->n = 0;
    int pending = 0, any = 0;
    for (size_t i = 0; i < s->n;) {
        size_t j = i;
        uint32_t cp = utf8_next(s->p, s->n, &j);
        if (cp_class(cp) == 3) { pending = 1; }
        else { if (pending && any) v8_push(tmp, ' '); pending = 0; any = 1; v8_append(tmp, s->p + i, j - i); }
        i = j;
    }
    vec8 t = *s; *s = *tmp; *tmp = t;
}
static void decode_doc(const orc_tok* t, const uint32_t* ids, size_t n, int skip_special, int cleanup, vec8* out, vec8* raw, vec8* tmp) {
    raw->n = 0;
    for (size_t i = 0; i < n; ++i) {
        uint32_t id = ids[i];
        if (id > t->max_id || !t->tok[id]) continue;
        if (skip_special && t->tok_special[id]) continue;
        const uint8_t* s = t->tok[id];
        size_t L = t->tok_len[id];
        for (size_t k = 0; k < L;) {                     /* decoders.rs:100-116 */
            uint32_t cp = utf8_next(s, L, &k);
            if (cp == 0x120) { v8_push(raw, ' '); continue; }
            int hit = 0;
            if ((cp >= '!' && cp <= '~') || (cp >= 0xA1 && cp <= 0xAC) || (cp >= 0xAE && cp <= 0xFF)) { v8_push(raw, (uint8_t)cp); hit = 1; }
            else if (cp >= 0x100 && cp < 0x100 + 68) {
                for (int b = 0; b < 256; ++b) if (t->byte_cp[b] == cp) { v8_push(raw, (uint8_t)b); hit = 1; break; }
            }
            if (!hit && cp < 0x80) v8_push(raw, (uint8_t)cp);
        }
    }
    out->n = 0;
    utf8_lossy(raw->p, raw->n, out);
    if (cleanup) clean_up(out, tmp);
}

/* ------------------------------------------------------------------ batch drivers (par_iter) */
typedef struct {
    const orc_tok* t; const uint8_t* text; const uint64_t* off; size_t n; vec32* out; atomic_size_t* next;
    const uint32_t* ids; int skip_special, cleanup; vec8* outb;
} job;
static void* enc_worker(void* p) {
    job* j = (job*)p;
    scratch sc; memset(&sc, 0, sizeof(sc));
    for (;;) {
        size_t d = atomic_fetch_add(j->next, 16);
        if (d >= j->n) break;
        for (size_t k = d; k < d + 16 && k < j->n; ++k)
            encode_doc(j->t, j->text + j->off[k], (size_t)(j->off[k + 1] - j->off[k]), &j->out[k], &sc);
    }
    free(sc.norm.p); free(sc.word.p); free(sc.pre.p); free(sc.tmp.p); free(sc.tmp2.p);
    return NULL;
}
static void* dec_worker(void* p) {
    job* j = (job*)p;
    vec8 raw = {0}, tmp = {0};
    for (;;) {
        size_t d = atomic_fetch_add(j->next, 16);
        if (d >= j->n) break;
        for (size_t k = d; k < d + 16 && k < j->n; ++k)
            decode_doc(j->t, j->ids + j->off[k], (size_t)(j->off[k + 1] - j->off[k]), j->skip_special, j->cleanup, &j->outb[k], &raw, &tmp);
    }
    free(raw.p); free(tmp.p);
    return NULL;
}
static void run_threads(void* (*fn)(void*), job* j, int n_threads) {
    if (n_threads <= 1) { fn(j); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, fn, j);
    for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    free(th);
}

typedef struct { uint32_t* ids; uint64_t* off; uint8_t* bytes; } orc_result;

/* texts: packed UTF-8 + n+1 offsets.  Result: packed ids + n+1 offsets (free with orc_result_free). */
orc_result* orc_encode_batch(const orc_tok* t, const uint8_t* text, const uint64_t* off, size_t n, int n_threads) {
    vec32* outs = (vec32*)calloc(n + 1, sizeof(vec32));
    atomic_size_t next = 0;
    job j = {t, text, off, n, outs, &next, NULL, 0, 0, NULL};
    run_threads(enc_worker, &j, n_threads);
    orc_result* r = (orc_result*)calloc(1, sizeof(orc_result));
    r->off = (uint64_t*)malloc((n + 1) * 8);
    uint64_t tot = 0;
    for (size_t i = 0; i < n; ++i) { r->off[i] = tot; tot += outs[i].n; }
    r->off[n] = tot;
    r->ids = (uint32_t*)malloc((tot + 1) * 4);
    for (size_t i = 0; i < n; ++i) { memcpy(r->ids + r->off[i], outs[i].p, outs[i].n * 4); free(outs[i].p); }
    free(outs);
    return r;
}
orc_result* orc_decode_batch(const orc_tok* t, const uint32_t* ids, const uint64_t* off, size_t n, int skip_special,
                             int cleanup, int n_threads) {
    vec8* outs = (vec8*)calloc(n + 1, sizeof(vec8));
    atomic_size_t next = 0;
    job j = {t, NULL, off, n, NULL, &next, ids, skip_special, cleanup, outs};
    run_threads(dec_worker, &j, n_threads);
    orc_result* r = (orc_result*)calloc(1, sizeof(orc_result));
    r->off = (uint64_t*)malloc((n + 1) * 8);
    uint64_t tot = 0;
    for (size_t i = 0; i < n; ++i) { r->off[i] = tot; tot += outs[i].n; }
    r->off[n] = tot;
    r->bytes = (uint8_t*)malloc(tot + 1);
    for (size_t i = 0; i < n; ++i) { memcpy(r->bytes + r->off[i], outs[i].p, outs[i].n); free(outs[i].p); }
    free(outs);
    return r;
}
const uint32_t* orc_result_ids(const orc_result* r) { return r->ids; }
const uint8_t* orc_result_bytes(const orc_result* r) { return r->bytes; }
const uint64_t* orc_result_off(const orc_result* r) { return r->off; }
void orc_result_free(orc_result* r) { if (r) { free(r->ids); free(r->off); free(r->bytes); free(r); } }

/* standalone NFC for tests: returns malloc'ed buffer (caller frees with orc_free_buf) */
uint8_t* orc_nfc(const uint8_t* s, size_t n, size_t* out_n) {
    vec8 o = {0}; vec32 tmp = {0};
    uint8_t* r;
    if (nfc_normalize(s, n, &o, &tmp)) { r = (uint8_t*)malloc(o.n + 1); memcpy(r, o.p, o.n); *out_n = o.n; }
    else { r = (uint8_t*)malloc(n + 1); memcpy(r, s, n); *out_n = n; }
    free(o.p); free(tmp.p);
    return r;
}
void orc_free_buf(uint8_t* p) { free(p); }
