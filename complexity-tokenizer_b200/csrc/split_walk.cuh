// The walk of one text through a compiled Split pattern (regex_dfa.hpp): find_iter + the five behaviours of
// regex_split_with_behavior (reference src/pretokenizers.rs:298-433), written once for the device kernel (split.cu) and for
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
    const uint16_t* trans_ascii;    // [n_states * 128] the same transitions indexed by an ASCII byte directly; may be NULL
    const uint32_t* pair_impossible;// [512] bit (c1 * 128 + c2): no match contains c1 c2 adjacent (regex_dfa.hpp)
};

// plain memory reader (host hook); the device kernel reads through a register window (split.cu)
struct PtrReader {
    const uint8_t* t;
    CTK_HD uint32_t byte(uint64_t i) { return t[i]; }
};

// code point at position i (i < end) and its length; ill-formed input (cannot come from a Rust &str) degrades to single bytes
template <class Pos, class Reader>
CTK_HD uint32_t split_decode(Reader& r, uint32_t c, Pos i, Pos end, uint32_t& len) {
    if (c < 0xC0u) { len = 1; return c; }
    const uint32_t want = c < 0xE0u ? 2u : (c < 0xF0u ? 3u : 4u);
    if (i + want > end) { len = 1; return c; }
    len = want;
    if (want == 2) return ((c & 0x1Fu) << 6) | (r.byte(i + 1) & 63u);
    if (want == 3) return ((c & 0x0Fu) << 12) | ((r.byte(i + 1) & 63u) << 6) | (r.byte(i + 2) & 63u);
    return ((c & 7u) << 18) | ((r.byte(i + 1) & 63u) << 12) | ((r.byte(i + 2) & 63u) << 6) | (r.byte(i + 3) & 63u);
}

CTK_HD uint32_t split_class(const SplitTables& s, uint32_t cp) {
    if (cp < 128u) return s.ascii_class[cp];
    if (cp > 0x10FFFFu) cp = 0x10FFFFu;
    return s.blocks[(uint32_t)s.stage1[cp >> 8] * 256u + (cp & 255u)];
}

// Emit must provide: void boundary(uint64_t pos)  (strictly inside the text; repeats allowed)
//                    void span(uint64_t a, uint64_t b, bool starts_piece)  (Removed only; a < b)
//                    void matched()  (the text has at least one match: see the no-match rule above)
//
// A text may be walked in SEGMENTS by different threads: [seg_lo, seg_hi) inside the text [lo, hi), where every segment
// border inside the text is a SAFE START -- the position right after a byte that no match can contain (a "neutral" byte:
// one that belongs to none of the pattern's sets), so that no match spans the border and the scan is fresh there.  The
// byte before such a border is a gap byte: with it as the "previous end" the behaviours that look back (MergedWithPrevious,
// Contiguous, inverted Removed) take the same decisions a single walk of the whole text takes.  The behaviours that do NOT
// look back (Isolated, MergedWithNext, Removed keeping the matches) may also be cut between two bytes that no match can
// contain next to each other (pair_impossible): every attempt that starts before such a border dies at it, so the scan
// arrives there fresh -- this is what lets a pattern that covers every character (the GPT-2 pattern itself) be walked in
// parallel.
// Pos: the integer type of text positions (the device walks buffers below 4 GiB with 32-bit positions: half the instructions).
template <class Pos, class Reader, class Emit>
CTK_HD void split_walk(const SplitTables& s, Reader& rd, Pos lo, Pos hi, Pos seg_lo, Pos seg_hi, Emit& em) {
    const bool continuing = seg_lo > lo;                 // the byte at seg_lo - 1 is a gap byte of this text
    Pos last_end = continuing ? seg_lo - 1 : lo;
    bool any = false;
    // ONE flat loop, one automaton step per iteration: find_iter tries an anchored match at every character position in
    // turn (`a`), `q` runs ahead of it while the automaton lives, `b` is the last accepting position seen.  (Nested loops --
    // positions, then steps -- split a warp into lane groups that never meet again: 5 of 32 lanes active, measured.)
    Pos a = seg_lo, q = seg_lo, b = 0;
    uint32_t st = s.start, first_len = 1;
    const bool ascii_table = s.trans_ascii != nullptr;
    while (a < seg_hi) {
        uint32_t nx = 0, len = 1;
        if (q < seg_hi) {
            const uint32_t c = rd.byte(q);
            if (c < 128u && ascii_table) nx = s.trans_ascii[st * 128u + c];
            else {
                const uint32_t cp = split_decode(rd, c, q, seg_hi, len);
                nx = s.trans[st * s.n_classes + split_class(s, cp)];
            }
            if (q == a) first_len = len;
        }
        if (nx & 0x7FFFu) {                                      // the automaton lives: one character consumed
            st = nx & 0x7FFFu;
            q += len;
            if (nx & 0x8000u) b = q;
            continue;
        }
        if (!b) { a += first_len; q = a; st = s.start; continue; }      // no match at a: the next character position
        // ---- one match [a, b)
        switch (s.behavior) {
            case 0:                                              // Removed
                if (s.invert) {
                    const Pos from = any || !continuing ? last_end : seg_lo;
                    if (a > from) em.span(from, a, any || !continuing);       // (the first gap of a later segment continues the previous segment's)
                } else em.span(a, b, true);
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
        if (!any) em.matched();
        any = true;
        last_end = b;
        a = q = b;
        b = 0;
        st = s.start;
    }
    if (s.behavior == 0) {
        if (s.invert) {
            const Pos from = any || !continuing ? last_end : seg_lo;
            if (from < seg_hi) em.span(from, seg_hi, any || !continuing);
        }
        // (not inverted and no match anywhere in the text: the whole text is kept -- decided per text by the caller)
    } else if ((s.behavior == 2 || s.behavior == 4) && any && last_end < seg_hi) em.boundary(last_end);
}

}  // namespace ctk
