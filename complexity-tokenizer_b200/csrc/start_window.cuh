// Bit-parallel pre-token boundary detection, 16 bytes per lane.
//
// Same rule as TextView::is_start in device_common.cuh (the reference's pattern,
// src/pretokenizers.rs:13), evaluated on 32-bit windows of per-byte property masks:
//   window bit k  <->  byte (g*16 - 8 + k) of the chunk, g = this lane's 16-byte group
//   bits 8..23 are the lane's own bytes, 0..7 the previous lane's last 8, 24..31 the next lane's first 8.
// Host+device so that tests/test_host.py can check it against the scalar predicate without a GPU.
#pragma once
#include "device_common.cuh"

namespace ctk {

struct Masks16 {            // one bit per byte of a 16-byte group
    uint32_t L, N, W;       // class of the code point containing the byte (Other = none of them)
    uint32_t SP, AP;        // byte == 0x20, byte == '\''
    uint32_t CONT;          // UTF-8 continuation byte
    uint32_t SUSP;          // != 0: the group touches an NFC-suspect code point (NFC_QC != Yes or ccc != 0; trie bit 2)
};

CTK_HD uint32_t movemask4(uint32_t hi) {             // bit 7 of each byte -> 4 bits
    return (((hi >> 7) & 0x01010101u) * 0x01020408u) >> 24;
}

// SWAR classification of 4 ASCII bytes (every byte < 0x80): per-byte flags in bit 7
CTK_HD void classify_word_ascii(uint32_t x, uint32_t& l, uint32_t& n, uint32_t& w, uint32_t& sp, uint32_t& ap) {
    const uint32_t H = 0x80808080u;
    uint32_t y = x | 0x20202020u;
    l = (y + 0x1F1F1F1Fu) & ~(y + 0x05050505u) & H;                   // 'a' <= y <= 'z'
    n = (x + 0x50505050u) & ~(x + 0x46464646u) & H;                   // '0' <= x <= '9'
    sp = ~((x ^ 0x20202020u) + 0x7F7F7F7Fu) & H;                      // x == ' '
    ap = ~((x ^ 0x27272727u) + 0x7F7F7F7Fu) & H;                      // x == '\''
    w = sp | ((x + 0x77777777u) & ~(x + 0x72727272u) & H);            // 9 <= x <= 13
}

// Classify the 16-byte group that starts at chunk[pos]; `chunk` must be readable from pos-3 to pos+18
// (continuation bytes look back for their lead, lead bytes look ahead for their tail).
CTK_HD Masks16 classify16(const uint8_t* chunk, int pos, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3,
                          const uint8_t* trie_index, const uint8_t* trie_blocks) {
    Masks16 m;
    uint32_t l, n, w, sp, ap;
    m.CONT = 0;
    m.SUSP = 0;
    uint32_t words[4] = {w0, w1, w2, w3};
    // bit 7 of byte k times 0x00204081 lands on bit 28 + k (no carries: all partial products are distinct bits);
    // four words are funnelled into the top 16 bits of an accumulator, 3 instructions per word and mask
    uint32_t aL = 0, aN = 0, aW = 0, aS = 0, aA = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        uint32_t x = words[k] & 0x7F7F7F7Fu;                           // non-ASCII bytes are fixed up below
        classify_word_ascii(x, l, n, w, sp, ap);
        uint32_t ascii = ~words[k] & 0x80808080u;
        aL = (aL >> 4) | (((l & ascii) * 0x00204081u) & 0xF0000000u);
        aN = (aN >> 4) | (((n & ascii) * 0x00204081u) & 0xF0000000u);
        aW = (aW >> 4) | (((w & ascii) * 0x00204081u) & 0xF0000000u);
        aS = (aS >> 4) | (((sp & ascii) * 0x00204081u) & 0xF0000000u);
        aA = (aA >> 4) | (((ap & ascii) * 0x00204081u) & 0xF0000000u);
    }
    m.L = aL >> 16; m.N = aN >> 16; m.W = aW >> 16; m.SP = aS >> 16; m.AP = aA >> 16;
    uint32_t non_ascii = movemask4(w0 & 0x80808080u) | (movemask4(w1 & 0x80808080u) << 4) |
                         (movemask4(w2 & 0x80808080u) << 8) | (movemask4(w3 & 0x80808080u) << 12);
    if (non_ascii) {                                                   // rare path: multi-byte code points, ONE lookup per character
        // continuation bytes 10xxxxxx: bit 7 set, bit 6 clear
        uint32_t cont = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; ++k) cont |= movemask4(words[k] & 0x80808080u & ~(words[k] << 1)) << (4 * k);
        m.CONT = cont;
        uint32_t leads = non_ascii & ~cont;
        // continuation bytes at the start of the group belong to a character that began in the previous group
        const uint32_t run0 = cont & ~(cont + 1u);
        if (run0) {
            int lead = pos - 1;
            for (int k = 0; k < 2 && (chunk[lead] & 0xC0u) == 0x80u; ++k) --lead;
            const uint32_t c = chunk[lead];
            uint32_t cp;
            if (c < 0xE0u) cp = ((c & 0x1Fu) << 6) | (chunk[lead + 1] & 63u);
            else if (c < 0xF0u) cp = ((c & 0x0Fu) << 12) | ((chunk[lead + 1] & 63u) << 6) | (chunk[lead + 2] & 63u);
            else cp = ((c & 7u) << 18) | ((chunk[lead + 1] & 63u) << 12) | ((chunk[lead + 2] & 63u) << 6) | (chunk[lead + 3] & 63u);
            const uint32_t nib = trie_nibble(trie_index, trie_blocks, cp);
            m.SUSP |= nib & 4u;
            const uint32_t cls = nib & 3u;
            if (cls == CLS_L) m.L |= run0;
            else if (cls == CLS_N) m.N |= run0;
            else if (cls == CLS_W) m.W |= run0;
        }
        while (leads) {
#if defined(__CUDA_ARCH__)
            const int bb = __ffs(leads) - 1;
#else
            const int bb = __builtin_ctz(leads);
#endif
            leads &= leads - 1;
            const int i = pos + bb;
            const uint32_t c = chunk[i];
            uint32_t cp, len;
            if (c < 0xE0u) { cp = ((c & 0x1Fu) << 6) | (chunk[i + 1] & 63u); len = 2; }
            else if (c < 0xF0u) { cp = ((c & 0x0Fu) << 12) | ((chunk[i + 1] & 63u) << 6) | (chunk[i + 2] & 63u); len = 3; }
            else { cp = ((c & 7u) << 18) | ((chunk[i + 1] & 63u) << 12) | ((chunk[i + 2] & 63u) << 6) | (chunk[i + 3] & 63u); len = 4; }
            const uint32_t nib = trie_nibble(trie_index, trie_blocks, cp);
            m.SUSP |= nib & 4u;
            const uint32_t cls = nib & 3u;
            // the lead and the continuation bytes that follow it inside this group (never more than len - 1)
            const uint32_t span = (((1u << len) - 1u) << bb) & 0xFFFFu & (cont | (1u << bb));
            if (cls == CLS_L) m.L |= span;
            else if (cls == CLS_N) m.N |= span;
            else if (cls == CLS_W) m.W |= span;
        }
    }
    return m;
}

// 16-bit masks of the previous / own / next group -> 32-bit window
CTK_HD uint32_t window(uint32_t prev16, uint32_t own16, uint32_t next16) {
    return ((prev16 >> 8) & 0xFFu) | ((own16 & 0xFFFFu) << 8) | ((next16 & 0xFFu) << 24);
}

// Pre-token starts for window bits 8..23 (the lane's own 16 bytes).  DS = document-start bits (the
// position one past the end of the text counts as a document start).  `wbytes` points at the byte of
// window bit 0 (readable for 35 bytes); it is only dereferenced next to an apostrophe.
CTK_HD uint32_t start_window(uint32_t L, uint32_t N, uint32_t W, uint32_t SP, uint32_t AP, uint32_t CONT, uint32_t DS,
                             const uint8_t* wbytes) {
    const uint32_t nb = ~DS, nbn = ~(DS >> 1);
#define P1(X) (((X) << 1) & nb)          /* property of the previous byte, same document */
#define N1(X) (((X) >> 1) & nbn)         /* property of the next byte, same document */
    const uint32_t notW = ~W;
    const uint32_t O = ~(L | N | W);
    const uint32_t G = SP & ~P1(W) & N1(notW);                        // " ?" : a single space glued to what follows
    uint32_t C1 = 0, C2 = 0;                                          // contraction starts of length 2 / 3
    uint32_t cand = AP & (DS | P1(L | N) | P1(W & ~G)) & 0x1FFFFFF8u; // apostrophe where the regex tries a new match
    while (cand) {
#if defined(__CUDA_ARCH__)
        int j = __ffs(cand) - 1;
#else
        int j = __builtin_ctz(cand);
#endif
        cand &= cand - 1;
        if ((DS >> (j + 1)) & 1u) continue;
        uint32_t c1 = wbytes[j + 1];
        if (c1 == 's' || c1 == 't' || c1 == 'm' || c1 == 'd') { C1 |= 1u << j; continue; }
        if ((DS >> (j + 2)) & 1u) continue;
        uint32_t c2 = wbytes[j + 2];
        if ((c1 == 'r' && c2 == 'e') || (c1 == 'v' && c2 == 'e') || (c1 == 'l' && c2 == 'l')) C2 |= 1u << j;
    }
    const uint32_t CA = C1 | C2;
    const uint32_t IN = P1(CA) | P1(P1(C2));                          // letters inside a contraction
    const uint32_t AFT = P1(P1(C1)) | P1(P1(P1(C2)));                 // first byte after a contraction
    const uint32_t SAME = (P1(L) & L) | (P1(N) & N) | (P1(O) & O);
    uint32_t S = (W & ~P1(W)) | (notW & (DS | AFT | (~IN & ~P1(G) & ~SAME)));
#undef P1
#undef N1
    return (S & ~CONT) | DS;
}

}  // namespace ctk
