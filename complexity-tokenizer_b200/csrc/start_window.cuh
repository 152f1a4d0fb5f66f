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

// byte select (PRMT): result byte i = byte (sel >> 4i) & 7 of {b:a}
CTK_HD uint32_t ctk_prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
#endif
}

// Per-byte flags (bit 7 of every byte) of the four TRANSPOSED words -> one word whose byte j holds, in bits 0..3,
// the flags of positions 4j .. 4j+3; nib_compress packs those four nibbles into 16 bits (bit k = position k).
CTK_HD uint32_t flag_combine(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
    return (f0 >> 7) | (f1 >> 6) | (f2 >> 5) | (f3 >> 4);
}
CTK_HD uint32_t nib_compress(uint32_t F) {
    const uint32_t G = (F | (F >> 4)) & 0x00FF00FFu;
    return (G | (G >> 8)) & 0xFFFFu;
}

// Classify the 16-byte group that starts at chunk[pos]; `chunk` must be readable from pos-3 to pos+18
// (continuation bytes look back for their lead, lead bytes look ahead for their tail).
//
// ASCII bytes are classified with SWAR range tests on words that are first TRANSPOSED (word k = bytes k, 4+k, 8+k,
// 12+k): the per-byte flags of the four words then combine with four shifts into nibbles that are already in
// position order, and one 16-bit mask costs ten instructions instead of a multiply-and-funnel per word and mask.
CTK_HD Masks16 classify16(const uint8_t* chunk, int pos, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3,
                          const uint8_t* trie_index, const uint8_t* trie_blocks) {
    Masks16 m;
    m.CONT = 0;
    m.SUSP = 0;
    const uint32_t H = 0x80808080u;
    uint32_t T[4];
    {
        const uint32_t a = ctk_prmt(w0, w1, 0x5140u), b = ctk_prmt(w2, w3, 0x5140u);
        const uint32_t c = ctk_prmt(w0, w1, 0x7362u), d = ctk_prmt(w2, w3, 0x7362u);
        T[0] = ctk_prmt(a, b, 0x5410u); T[1] = ctk_prmt(a, b, 0x7632u);
        T[2] = ctk_prmt(c, d, 0x5410u); T[3] = ctk_prmt(c, d, 0x7632u);
    }
    uint32_t l[4], n[4], sp[4], ap[4], ct[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        const uint32_t x = T[k] & 0x7F7F7F7Fu;                         // non-ASCII bytes are masked out below
        const uint32_t y = x | 0x20202020u;
        l[k] = (y + 0x1F1F1F1Fu) & ~(y + 0x05050505u) & H;             // 'a' <= y <= 'z'
        n[k] = (x + 0x50505050u) & ~(x + 0x46464646u) & H;             // '0' <= x <= '9'
        sp[k] = ~((x ^ 0x20202020u) + 0x7F7F7F7Fu) & H;                // x == ' '
        ap[k] = ~((x ^ 0x27272727u) + 0x7F7F7F7Fu) & H;                // x == '\''
        ct[k] = (x + 0x77777777u) & ~(x + 0x72727272u) & H;            // 9 <= x <= 13
    }
    const uint32_t NA = flag_combine(T[0] & H, T[1] & H, T[2] & H, T[3] & H);   // non-ASCII bytes, nibble form
    m.L = nib_compress(flag_combine(l[0], l[1], l[2], l[3]) & ~NA);
    m.SP = nib_compress(flag_combine(sp[0], sp[1], sp[2], sp[3]) & ~NA);
    m.N = 0; m.AP = 0; m.W = m.SP;
    if ((n[0] | n[1] | n[2] | n[3]) | (ap[0] | ap[1] | ap[2] | ap[3]) | (ct[0] | ct[1] | ct[2] | ct[3])) {   // digits, apostrophes, tabs / newlines: rare
        m.N = nib_compress(flag_combine(n[0], n[1], n[2], n[3]) & ~NA);
        m.AP = nib_compress(flag_combine(ap[0], ap[1], ap[2], ap[3]) & ~NA);
        m.W |= nib_compress(flag_combine(ct[0], ct[1], ct[2], ct[3]) & ~NA);
    }
    const uint32_t non_ascii = NA;
    if (non_ascii) {                                                   // rare path: multi-byte code points, ONE lookup per character
        // continuation bytes 10xxxxxx: bit 7 set, bit 6 clear
        const uint32_t cont = nib_compress(flag_combine(T[0] & H & ~(T[0] << 1), T[1] & H & ~(T[1] << 1), T[2] & H & ~(T[2] << 1),
                                                        T[3] & H & ~(T[3] << 1)));
        m.CONT = cont;
        uint32_t leads = nib_compress(NA) & ~cont;
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
