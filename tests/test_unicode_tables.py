"""The Unicode arithmetic the reference leaves to unpinned crates: `regex` 1.10 classes \\p{L} \\p{N} \\s and
`unicode-normalization` 0.1 NFC.  The product's trie (csrc/unicode_trie_gen.h) and the oracle's ranges are generated from
Python's unicodedata (Unicode 15.0).  Here EVERY code point is checked against two independent sources available in this
image: the `regex` module (its own, newer Unicode tables) and unicodedata.  General_Category and White_Space of an assigned
code point are stable across Unicode versions, so the sources must agree on everything assigned in 15.0."""
import sys
import unicodedata as ud

import numpy as np
import pytest

WHITE_SPACE = {0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000} | set(range(0x2000, 0x200B))


def _product_classes(lib):
    out = np.zeros(0x110000, dtype=np.uint8)
    assert lib.ctk_debug_cp_classes(0, 0x110000, out.ctypes.data) == 0
    return out


def test_every_code_point_class_against_unicodedata_and_regex(built_lib):
    import regex
    import py_oracle
    got = _product_classes(built_lib)
    rx = {1: regex.compile(r'\p{L}'), 2: regex.compile(r'\p{N}'), 3: regex.compile(r'\s')}
    bad_ud, bad_rx, bad_twin, newer = [], [], [], 0
    for cp in range(0x110000):
        if 0xD800 <= cp < 0xE000:
            continue                                  # surrogates never occur in a &str
        ch = chr(cp)
        cat = ud.category(ch)
        want = 3 if cp in WHITE_SPACE else 1 if cat[0] == 'L' else 2 if cat[0] == 'N' else 0
        cls = int(got[cp]) & 3
        if cls != want:
            bad_ud.append(cp)
        if py_oracle.char_class(ch) != want:
            bad_twin.append(cp)
        r = 3 if rx[3].match(ch) else 1 if rx[1].match(ch) else 2 if rx[2].match(ch) else 0
        if r != want:
            if cat == 'Cn':
                newer += 1                            # assigned after Unicode 15.0: outside what the tables (and the corpora) cover
            else:
                bad_rx.append(cp)
    assert not bad_ud, 'trie disagrees with unicodedata at %s' % [hex(c) for c in bad_ud[:10]]
    assert not bad_twin
    assert not bad_rx, 'regex module disagrees with the tables on code points assigned in 15.0: %s' % [hex(c) for c in bad_rx[:10]]
    assert sorted(cp for cp in range(0x110000) if regex.match(r'\s', chr(cp)) and not (0xD800 <= cp < 0xE000)) == sorted(WHITE_SPACE), \
        'White_Space is the 25 code points SURVEY.md 3.2(i) lists'
    print('code points the newer regex tables classify differently only because they were assigned after 15.0:', newer, file=sys.stderr)


def test_nfc_suspect_bit_covers_everything_nfc_can_change(built_lib):
    """bit 2 of the trie nibble must be set for every code point with NFC_QC != Yes or ccc != 0: text without such a code
    point is its own NFC (UAX #15 quick check), which is what lets the encode kernel skip the normaliser"""
    got = _product_classes(built_lib)
    missing = []
    for cp in range(0x110000):
        if 0xD800 <= cp < 0xE000:
            continue
        ch = chr(cp)
        needs = ud.combining(ch) != 0 or not ud.is_normalized('NFC', ch) or ud.normalize('NFC', ch) != ch
        if needs and not (int(got[cp]) & 4):
            missing.append(cp)
    # NFC_QC=Maybe code points are exactly those that can compose with a PRECEDING character: check through pairs of a few bases
    for cp in range(0x300, 0x3100):
        ch = chr(cp)
        if int(got[cp]) & 4:
            continue
        for base in 'aeEकᄀ가カ':
            if ud.normalize('NFC', base + ch) != base + ch:
                missing.append(cp)
                break
    assert not missing, [hex(c) for c in missing[:10]]


def test_nfc_of_the_c_oracle_against_unicodedata():
    """the oracle's C NFC (streaming UAX #15) on every assigned code point alone and after a base letter"""
    import c_oracle
    cps = [cp for cp in range(0x80, 0x30000) if not (0xD800 <= cp < 0xE000) and ud.category(chr(cp)) != 'Cn']
    step = 97
    texts = [chr(cp) for cp in cps[::1]][:20000] + ['e' + chr(cp) for cp in cps[::step]] + ['ᄀ' + chr(cp) for cp in range(0x1161, 0x1176)] + \
            ['가' + chr(cp) for cp in range(0x11A8, 0x11C3)]
    for i in range(0, len(texts), 5000):
        chunk = texts[i:i + 5000]
        joined = '\x00'.join(chunk)
        assert c_oracle.nfc(joined.encode('utf-8')).decode('utf-8') == ud.normalize('NFC', joined)
