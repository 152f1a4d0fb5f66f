// The walk of one text through a compiled Split pattern (regex_dfa.hpp): find_iter + the five behaviours of
// regex_split_with_behavior (reference src/pretokenizers.rs:298-433), written once for the device kernels (split.cu) and for
// the host-side test hook ctk_debug_split_pieces (api.cu), which lets the CPU test-suite exercise the compiler + walk
// without a GPU.  The product path is the device one.
//
// Every behaviour but Removed cuts the text into adjacent pieces, so it is described by BOUNDARIES:
//   Isolated            at every match start and every match end                                       (:332-347)
//   MergedWithPrevious  at a match end, unless the next match starts right there                       (:348-375)
//   MergedWithNext      at every match start                                                           (:376-404)
//   Contiguous          at a match start with a gap before it, at a match end with a gap after it      (:405-428)
// Removed keeps SPANS: the matches, or with `invert` the gaps between them                             (:313-331)
// A text without any match stays one piece whatever the behaviour (:305-307).
#pragma once
#include <cstdint>

#include "device_common.cuh"

namespace ctk {

struct SplitTables {
    const uint16_t* trans;          // [n_states * n_classes]: next state | 0x8000 if it accepts; state 0 is dead
    const uint8_t* ascii_class;     // [128]
    const uint16_t* stage1;         // [0x1100]
    const uint8_t* blocks;          // [n_blocks * 256]
    uint32_t n_classes, start;
    int behavior, invert;
};

// code point at t[i] (i < end) and its length; ill-formed input (cannot come from a Rust &str) degrades to single bytes
CTK_HD uint32_t split_decode(const uint8_t* t, uint64_t i, uint64_t end, uint32_t& len) {
    const uint32_t c = t[i];
    if (c < 0xC0u) { len = 1; return c; }
    const uint32_t want = c < 0xE0u ? 2u : (c < 0xF0u ? 3u : 4u);
    if (i + want > end) { len = 1; return c; }
    len = want;
    if (want == 2) return ((c & 0x1Fu) << 6) | (t[i + 1] & 63u);
    if (want == 3) return ((c & 0x0Fu) << 12) | ((t[i + 1] & 63u) << 6) | (t[i + 2] & 63u);
    return ((c & 7u) << 18) | ((t[i + 1] & 63u) << 12) | ((t[i + 2] & 63u) << 6) | (t[i + 3] & 63u);
}

CTK_HD uint32_t split_class(const SplitTables& s, uint32_t cp) {
    if (cp < 128u) return s.ascii_class[cp];
    if (cp > 0x10FFFFu) cp = 0x10FFFFu;
    return s.blocks[(uint32_t)s.stage1[cp >> 8] * 256u + (cp & 255u)];
}

// Emit must provide: void boundary(uint64_t pos)  (strictly inside the text; repeats allowed)
//                    void span(uint64_t a, uint64_t b)  (Removed only; a < b)
template <class Emit>
CTK_HD void split_walk(const SplitTables& s, const uint8_t* t, uint64_t lo, uint64_t hi, Emit& em) {
    uint64_t pos = lo, last_end = lo;
    bool any = false;
    while (pos < hi) {
        // the leftmost match at or after pos: try every character position in turn (the automaton is anchored)
        uint64_t a = pos, b = 0;
        for (; a < hi;) {
            uint32_t st = s.start, len;
            uint64_t q = a;
            uint32_t first_len = 1;
            while (q < hi) {
                const uint32_t cp = split_decode(t, q, hi, len);
                if (q == a) first_len = len;
                const uint32_t nx = s.trans[st * s.n_classes + split_class(s, cp)];
                st = nx & 0x7FFFu;
                if (st == 0) break;
                q += len;
                if (nx & 0x8000u) b = q;
            }
            if (b) break;
            a += first_len;
        }
        if (!b) break;
        // ---- one match [a, b)
        switch (s.behavior) {
            case 0:                                              // Removed
                if (s.invert) { if (a > last_end) em.span(last_end, a); }
                else em.span(a, b);
                break;
            case 1:                                              // Isolated
                if (a > lo) em.boundary(a);
                if (b < hi) em.boundary(b);
                break;
            case 2:                                              // MergedWithPrevious
                if (any && a > last_end) em.boundary(last_end);
                break;
            case 3:                                              // MergedWithNext
                if (a > lo) em.boundary(a);
                break;
            default:                                             // Contiguous
                if (a > last_end) { if (any) em.boundary(last_end); em.boundary(a); }
                break;
        }
        any = true;
        last_end = b;
        pos = b;
    }
    if (s.behavior == 0) {
        if (!any) { if (hi > lo) em.span(lo, hi); }
        else if (s.invert && last_end < hi) em.span(last_end, hi);
    } else if ((s.behavior == 2 || s.behavior == 4) && any && last_end < hi) em.boundary(last_end);
}

}  // namespace ctk
