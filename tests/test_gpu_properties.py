"""More GPU parity: randomised fuzzing against the oracle, slice/chunk boundary stress, long pre-tokens,
the two device pipelines against each other, and size-independent properties at larger sizes
(round trip, batch-splitting invariance, idempotence).  Bit-exact everywhere.  Run with -m gpu."""
import unicodedata

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle(path):
    import c_oracle
    return c_oracle.COracle.from_file(path)


def _tok(path):
    import complexity_tokenizer as ct
    return ct.Tokenizer.from_file(path)


ALPHABET = ["a", "b", "e", "s", "t", "r", "v", "l", "m", "d", "'", " ", " ", " ", "\n", "\t", "1", "9", ".", ",", "!", "-", "é", "ü",
            "ñ", "中", "文", "あ", "　", " ", "\U0001F600", "\U0001F44D", "́", "̧", "x", "'s", "'ll", "  ", "Ⅷ", "²", "_", "٣",
            "ß", "'re", "'ve", "'d", "'m", "'t", "Å", "豈", "각", "ᄀ", "ᅡ", "ᆨ", "the", " the", " of", "ing", "tion", "\r\n", "\x00", "=-"]


def _random_docs(rng, n_docs, max_len):
    docs = []
    for _ in range(n_docs):
        k = int(rng.integers(0, max_len))
        docs.append(''.join(ALPHABET[int(i)] for i in rng.integers(0, len(ALPHABET), size=k)))
    return docs


@pytest.mark.parametrize('cfg', ['config1', 'config3'])
def test_fuzz_many_small_documents(built_lib, tok_paths, cfg):
    """8 000 random documents of 0-60 pieces in one batch: document, slice and chunk boundaries everywhere."""
    rng = np.random.default_rng(17)
    docs = _random_docs(rng, 8000, 60)
    tok, orc = _tok(tok_paths[cfg]), _oracle(tok_paths[cfg])
    got, want = tok.encode_batch(docs), orc.encode_batch(docs)
    bad = [i for i, (g, w) in enumerate(zip(got, want)) if g != w]
    assert not bad, (bad[:5], docs[bad[0]] if bad else None)
    back = tok.decode_batch_with_options(got, False, False)
    assert back == [unicodedata.normalize('NFC', d) for d in docs]
    assert tok.decode_batch(got) == orc.decode_batch(got)


def test_slice_boundary_stress(built_lib, tok_paths):
    """Documents whose lengths straddle the 448-byte slice / 512-byte chunk geometry, and pre-tokens of
    1..700 bytes placed across slice boundaries (short, 17..32, > 32, > 256, unknown-end paths)."""
    tok, orc = _tok(tok_paths['config2']), _oracle(tok_paths['config2'])
    docs = []
    for n in list(range(430, 470)) + list(range(880, 912)) + [447, 448, 449, 463, 464, 465, 479, 480, 481, 495, 496, 497, 511, 512, 513]:
        docs.append(('ab ' * n)[:n])
        docs.append('x' * n)
        docs.append(' ' * n)
    for plen in [1, 15, 16, 17, 31, 32, 33, 34, 63, 64, 65, 127, 128, 129, 255, 256, 257, 300, 511, 700]:
        for lead in [0, 1, 430, 440, 447, 448, 460, 478, 479, 480, 481]:
            docs.append('.' * lead + ' ' + 'q' * plen + ' tail words here')
            docs.append('é' * (lead // 2) + ' ' + '中' * (plen // 3 + 1) + '!')
    got, want = tok.encode_batch(docs), orc.encode_batch(docs)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, (i, len(docs[i]), docs[i][:40])


def test_long_pretokens_match_oracle(built_lib, tok_paths):
    """config-4 shape at sizes the O(n^2) oracle finishes quickly: letter runs, space runs, punctuation runs."""
    import synth
    tok, orc = _tok(tok_paths['config2']), _oracle(tok_paths['config2'])
    docs = [d.decode() for d in synth.gen_long_docs(doc_bytes=3000, n_docs=12)]
    assert tok.encode_batch(docs) == orc.encode_batch(docs)


def test_fused_and_general_pipelines_agree(built_lib, tok_paths):
    """The multi-kernel general pipeline and the fused kernels are two implementations of one function."""
    import complexity_tokenizer as ct
    import synth
    for cfg, kind in (('config2', 'ascii'), ('config3', 'mixed')):
        tok = _tok(tok_paths[cfg])
        text, offs = synth.gen_corpus(kind, 321, 6 << 20, doc_median=2000)
        a_ids, a_off = tok.encode_packed(text, offs)
        ct._lib().ctk_debug_use_general(tok._h, 1)
        b_ids, b_off = tok.encode_packed(text, offs)
        ct._lib().ctk_debug_use_general(tok._h, 0)
        assert np.array_equal(a_off, b_off) and np.array_equal(a_ids, b_ids)


def test_properties_at_larger_size(built_lib, tok_paths):
    """256 MiB of the bench corpus (the oracle would need minutes): properties that need no oracle.
    (1) decode(encode(x)) == x byte-exact through decode_batch_with_options(False, False);
    (2) idempotence: a second encode gives identical ids;
    (3) batch-splitting invariance: encoding the two halves separately gives the same ids (documents are independent);
    (4) a 1 % sample of documents equals the oracle."""
    import synth
    tok, orc = _tok(tok_paths['config2']), _oracle(tok_paths['config2'])
    text, offs = synth.gen_corpus('ascii', 5000, 256 << 20, doc_median=4096, doc_min=256, doc_max=65536)
    ids, ioff = tok.encode_packed(text, offs)
    assert int(ioff[-1]) == ids.size and ids.size > 0
    import os
    os.environ['CTK_WIDEN_THREADS'] = '4'                         # opt-in narrow copy (2 bytes per id here): same result
    try:
        ids_n, ioff_n = tok.encode_packed(text, offs)
    finally:
        del os.environ['CTK_WIDEN_THREADS']
    assert np.array_equal(ids, ids_n) and np.array_equal(ioff, ioff_n)
    b, boff = tok.decode_packed(ids, ioff, False, False)
    assert np.array_equal(boff, offs) and np.array_equal(b, text)
    ids2, ioff2 = tok.encode_packed(text, offs)
    assert np.array_equal(ids, ids2) and np.array_equal(ioff, ioff2)
    h = (len(offs) - 1) // 2
    cut = int(offs[h])
    ia, oa = tok.encode_packed(text[:cut], offs[:h + 1])
    ib, ob = tok.encode_packed(text[cut:], offs[h:] - offs[h])
    assert np.array_equal(np.concatenate([ia, ib]), ids)
    assert np.array_equal(np.concatenate([oa[:-1], ob + oa[-1]]), ioff)
    n = len(offs) - 1
    for d in range(0, n, 100):
        w, _ = orc.encode_packed(text[int(offs[d]):int(offs[d + 1])], np.array([0, offs[d + 1] - offs[d]], dtype=np.uint64), threads=1)
        assert np.array_equal(ids[int(ioff[d]):int(ioff[d + 1])], w), d


def test_unsupported_and_errors(built_lib, tok_paths, small_tok_json):
    import json
    import complexity_tokenizer as ct
    with pytest.raises(IOError):
        ct.Tokenizer.from_file('/nonexistent/tokenizer.json')
    with pytest.raises(IOError):
        ct.Tokenizer.from_str('{"model": ')
    tj = json.loads(small_tok_json)
    tj['pre_tokenizer'] = {'type': 'Whitespace'}
    with pytest.raises(ct.UnsupportedTokenizerError):
        ct.Tokenizer.from_str(json.dumps(tj))
    tok = ct.Tokenizer.from_str(small_tok_json)
    with pytest.raises(TypeError):
        tok.encode_batch('not a list')
    with pytest.raises(UnicodeEncodeError):
        tok.encode('\ud800')
    assert tok.encode('') == [] and tok.decode([]) == ''


def test_add_prefix_space(built_lib, small_tok_json):
    """ByteLevel{add_prefix_space: true} (pretokenizers.rs:163-167): a leading space is added to non-empty
    documents that do not start with one -- after NFC, before the pattern."""
    import json
    import c_oracle
    import complexity_tokenizer as ct
    tj = json.loads(small_tok_json)
    tj['pre_tokenizer'] = {'type': 'ByteLevel', 'add_prefix_space': True, 'trim_offsets': True}
    js = json.dumps(tj, ensure_ascii=False)
    tok, orc = ct.Tokenizer.from_str(js), c_oracle.COracle.from_str(js)
    rng = np.random.default_rng(23)
    docs = ['', ' ', 'a', ' a', "'s", " 's", '\n x', 'hello world', 'é', '中文', '12 34', '  two'] + _random_docs(rng, 3000, 40)
    got, want = tok.encode_batch(docs), orc.encode_batch(docs)
    assert want == orc.twin.encode_batch(docs)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, (i, docs[i])


def _byte_tokenizer_json():
    """vocab = the 256 byte symbols, id == byte value: decode(ids) reproduces arbitrary byte strings"""
    import json
    import py_oracle
    vocab = {c: b for b, c in py_oracle.BYTE_ENCODER.items()}
    return json.dumps({'model': {'type': 'BPE', 'vocab': vocab, 'merges': []}}, ensure_ascii=False)


def test_cleanup_fuzz_against_sequential_replaces(built_lib):
    """clean_up_tokenization_spaces (mod.rs:749-769): the data-parallel formulation on the GPU against the oracle,
    which applies the 15 str::replace calls literally.  Space/punctuation/hyphen-heavy random strings."""
    import c_oracle
    import complexity_tokenizer as ct
    js = _byte_tokenizer_json()
    tok, orc = ct.Tokenizer.from_str(js), c_oracle.COracle.from_str(js)
    rng = np.random.default_rng(41)
    pieces = [' ', ' ', ' ', ' ', '-', '-', '.', ',', '"', "'", '(', ')', '[', ']', '!', '?', ':', ';', 'a', 'b', 'xyz', '\n', '\t',
              ' ', '　', 'é', ' - ', ' -', '- ', '  ', '   ', ' .', '" ', " '", '( ', ' )', '--', ' - - ', '\r\n']
    batch = []
    for _ in range(30000):
        k = int(rng.integers(0, 24))
        s = ''.join(pieces[int(i)] for i in rng.integers(0, len(pieces), size=k))
        batch.append(list(s.encode('utf-8')))
    batch += [list(b' - - - '), list(b'say " hi " now'), list(b'  .  ,'), list(b'( a ) [ b ]'), list(b'a -  - b'), list(b'(-  -)'),
              list(b' -  -  - '), list(b'x - - - - - - y'), list(b'- - -'), list(b' '), [], list(b'-'), list(b' - ')]
    got = tok.decode_batch(batch)
    want = orc.decode_batch(batch)
    bad = [i for i, (g, w) in enumerate(zip(got, want)) if g != w]
    assert not bad, (len(bad), bytes(batch[bad[0]]), got[bad[0]], want[bad[0]])
    # invalid UTF-8 mixed in: from_utf8_lossy + clean-up through the per-document path
    weird = [list(b'a \xe2\x82 . b'), list(b'\xff - \xc3'), list(b'ok . fine'), list(b'\xf0\x9f\x98 ,')]
    for opts in ((False, True), (False, False)):
        assert tok.decode_batch_with_options(weird, *opts) == orc.decode_batch(weird, *opts)


def test_added_tokens_inside_words(built_lib, small_tok_json):
    """mod.rs:566-675: added tokens are searched inside each byte-mapped pre-token (longest at position 0,
    first occurrence only, single_word / lstrip / rstrip flags).  Tokens such as <s> can never match; all-letter,
    all-digit or all-punctuation ones can."""
    import json
    import c_oracle
    import complexity_tokenizer as ct
    tj = json.loads(small_tok_json)
    nid = max(tj['model']['vocab'].values()) + 1
    extra = [('hello', {}), ('he', {}), ('hell', {}), ('Ġworld', {}), ('ing', {'single_word': True}), ('tion', {'rstrip': True}),
             ('pre', {'lstrip': True}), ('...', {}), ('42', {}), ('zz', {'single_word': True, 'special': True}), ('Ġthe', {'rstrip': True}),
             ('é', {}), ('<mask>', {'special': True})]
    for k, (content, flags) in enumerate(extra):
        t = {'id': nid + k, 'content': content, 'special': False, 'single_word': False, 'lstrip': False, 'rstrip': False, 'normalized': False}
        t.update(flags)
        tj['added_tokens'].append(t)
    js = json.dumps(tj, ensure_ascii=False)
    tok, orc = ct.Tokenizer.from_str(js), c_oracle.COracle.from_str(js)
    rng = np.random.default_rng(5)
    words = ['hello', 'hell', 'he', 'help', 'shell', 'hellohello', ' world', 'world', ' worldly', 'ing', 'sing', 'inging', 'ing!', 'tion',
             'nation', 'nations', 'pre', 'prefix', 'unpre', '...', '....', '.....', '42', '1423', '4242', 'zz', 'zzz', 'azz', ' the', ' then',
             'the', 'é', 'été', 'cafés', '<mask>', 'x', ' ', '\n', ',', "'s", 'hello' * 9, 'ab' * 30 + 'hello' + 'cd' * 40, '4' * 40 + '42' * 30]
    docs = [''.join(words[int(i)] if rng.random() < 0.7 else ' ' + words[int(i)] for i in rng.integers(0, len(words), size=int(rng.integers(0, 12))))
            for _ in range(4000)] + words
    got, want = tok.encode_batch(docs), orc.encode_batch(docs)
    assert orc.encode_batch(docs[:300]) == orc.twin.encode_batch(docs[:300])
    bad = [i for i, (g, w) in enumerate(zip(got, want)) if g != w]
    assert not bad, (len(bad), docs[bad[0]], got[bad[0]], want[bad[0]])
    assert any(nid <= t < nid + len(extra) for ids in want for t in ids)          # the added tokens really fire


# ---- very long pre-tokens: the round-parallel path (csrc/encode_xlong.cuh) ---------------------------------

def _xlong_docs():
    import synth
    docs = [d.decode() for d in synth.gen_long_docs(doc_bytes=9000, n_docs=8)]
    docs += ['a' * 5000, 'ab' * 3000, ' ' * 4097 + 'word', 'x' + ' ' * 3000, '=' * 257, '-' * 256, 'q' * 258 + ' ' + 'z' * 300,
             'The start of a normal sentence, then ' + 'lorem' * 400 + ' and an ordinary tail. ' * 30,
             '\n'.join('w' * n for n in (255, 256, 257, 258, 300, 447, 448, 449, 600, 1000))]
    return docs


def test_xlong_rounds_match_oracle(built_lib, tok_paths):
    """Pre-tokens of 257 .. 9000 bytes (letter runs, runs of one symbol, spaces, punctuation) merged in parallel
    rounds must equal the reference's one-merge-at-a-time order (bpe.rs:104-153) as restated by the oracle."""
    import complexity_tokenizer as ct
    for cfg in ('config2', 'config1'):
        tok, orc = _tok(tok_paths[cfg]), _oracle(tok_paths[cfg])
        docs = _xlong_docs()
        got, want = tok.encode_batch(docs), orc.encode_batch(docs)
        for i, (g, w) in enumerate(zip(got, want)):
            assert g == w, (cfg, i, len(docs[i]), docs[i][:30])
        assert ct._lib().ctk_debug_xlong_rounds(tok._h) > 0          # the rounds did run


def test_xlong_rounds_with_the_uniform_window(built_lib, tok_paths):
    """A monotone table without per-token reach windows (forced here with CTK_NO_REACH; in the field: a table whose
    merge products are not the concatenation of their parts) uses the uniform window W = longest token in the grid-wide
    rounds and the register path in k_encode_long.  Same ids as the oracle."""
    import os
    import complexity_tokenizer as ct
    os.environ['CTK_NO_REACH'] = '1'
    try:
        tok = _tok(tok_paths['config2'])
    finally:
        del os.environ['CTK_NO_REACH']
    orc = _oracle(tok_paths['config2'])
    docs = _xlong_docs()
    assert tok.encode_batch(docs) == orc.encode_batch(docs)
    assert ct._lib().ctk_debug_xlong_rounds(tok._h) > 0


def test_xlong_rounds_equal_sequential_device_path(built_lib, tok_paths):
    """Same inputs through the sequential warp path (CTK_NO_XLONG) and the rounds: identical ids."""
    import os
    tok = _tok(tok_paths['config2'])
    docs = _xlong_docs()
    a = tok.encode_batch(docs)
    os.environ['CTK_NO_XLONG'] = '1'
    try:
        b = tok.encode_batch(docs)
    finally:
        del os.environ['CTK_NO_XLONG']
    assert a == b


def test_xlong_with_dropped_bytes_and_non_monotone_table(built_lib, small_tok_json):
    """(1) a byte whose mapped char is not in the vocab is dropped before merging (bpe.rs:94-97), also inside a very long
    pre-token; (2) a non-monotone merge table keeps the sequential path and still matches the oracle."""
    import json
    import c_oracle
    import complexity_tokenizer as ct
    tj = json.loads(small_tok_json)
    v = tj['model']['vocab']
    del v['q']
    tj['model']['merges'] = [m for m in tj['model']['merges'] if 'q' not in (m if isinstance(m, str) else ''.join(m))]
    for k in [k for k in v if 'q' in k]:
        del v[k]
    docs = ['qa' * 700 + 'q', 'the' * 300 + 'q' * 10 + 'and' * 200, 'q' * 600, ' ' * 500 + 'q']
    data = json.dumps(tj, ensure_ascii=False)
    tok, orc = ct.Tokenizer.from_str(data), c_oracle.COracle.from_str(data)
    assert tok.encode_batch(docs) == orc.encode_batch(docs)
    assert ct._lib().ctk_debug_xlong_rounds(tok._h) > 0
    tj2 = json.loads(small_tok_json)
    ms = tj2['model']['merges']
    tj2['model']['merges'] = ms[300:] + ms[:300]                  # pairs now outrank the merges that make their parts
    data2 = json.dumps(tj2, ensure_ascii=False)
    tok2, orc2 = ct.Tokenizer.from_str(data2), c_oracle.COracle.from_str(data2)
    docs2 = [d[:1500] for d in _xlong_docs()]
    assert tok2.encode_batch(docs2) == orc2.encode_batch(docs2)
    assert ct._lib().ctk_debug_xlong_rounds(tok2._h) == 0


def test_config4_full_size_properties(built_lib, tok_paths):
    """BASELINE config 4 at full size (64 documents of 1 MiB; single pre-tokens of up to 2^20 symbols; the O(n^2)
    reference order would need hours on the CPU): size-independent properties.
    (1) decode(encode(x)) == x byte-exact; (2) idempotence; (3) a document encodes the same alone as in the batch;
    (4) a 1 MiB run of spaces is ceil-log-many tokens of the longest space tokens: every id decodes to spaces only."""
    import synth
    tok = _tok(tok_paths['config2'])
    docs = synth.gen_long_docs()
    text, offs = synth.pack(docs)
    ids, ioff = tok.encode_packed(text, offs)
    b, boff = tok.decode_packed(ids, ioff, False, False)
    assert np.array_equal(boff, offs) and np.array_equal(b, text)
    ids2, ioff2 = tok.encode_packed(text, offs)
    assert np.array_equal(ids, ids2) and np.array_equal(ioff, ioff2)
    for d in (0, 1, 2, 3, 63):
        one, _ = tok.encode_packed(text[int(offs[d]):int(offs[d + 1])], np.array([0, int(offs[d + 1] - offs[d])], dtype=np.uint64))
        assert np.array_equal(one, ids[int(ioff[d]):int(ioff[d + 1])]), d
    sp = ids[int(ioff[2]):int(ioff[3])]
    assert sp.size <= (1 << 20) // 2
    assert set(tok.decode([int(x) for x in np.unique(sp)]).replace(' ', '')) == set()


def test_warp_rounds_equal_sequential_merging_on_cjk(built_lib, tok_paths):
    """Pre-tokens of 33..256 bytes (CJK runs in config 3) are merged in parallel rounds inside one warp
    (encode_long.cuh: bpe_warp_rounds, per-symbol reach windows).  Same ids as the one-merge-per-iteration device
    path (CTK_NO_ROUNDS), which test_corpus_bit_exact pins on the oracle for a prefix of the same corpus."""
    import os
    import synth
    tok = _tok(tok_paths['config3'])
    text, offs = synth.gen_corpus('mixed', 3003, 24 << 20, doc_median=4096)
    os.environ['CTK_WIDEN_THREADS'] = '3'                         # opt-in: ids cross PCIe packed to 3 bytes (100K vocab) ...
    try:
        a_ids, a_off = tok.encode_packed(text, offs)
    finally:
        del os.environ['CTK_WIDEN_THREADS']
    os.environ['CTK_NO_ROUNDS'] = '1'
    os.environ['CTK_NO_MID'] = '1'                                # ... and here as plain uint32, merged sequentially
    try:
        b_ids, b_off = tok.encode_packed(text, offs)
    finally:
        del os.environ['CTK_NO_ROUNDS'], os.environ['CTK_NO_MID']
    assert np.array_equal(a_off, b_off) and np.array_equal(a_ids, b_ids)


def test_optimistic_nfc_switches_back_and_forth(built_lib, tok_paths):
    """The encode kernels run on the raw text and only report NFC-suspect code points; such a call is repeated through
    the normaliser and later calls scan first until one comes back clean (engine.hpp: nfc_optimistic).  Whatever the
    order of clean and dirty batches, the ids are the oracle's (normalizers.rs:45-47 then the rest of the pipeline)."""
    tok, orc = _tok(tok_paths['config1']), _oracle(tok_paths['config1'])
    clean = ['plain ascii text, nothing to normalise', 'déjà vu: precomposed é stays', '中文 and emoji \U0001F600']
    dirty = ['café with a combining acute', 'Å ring, Å angstrom sign, 각 jamo', 'x' * 500 + 'é' + 'y' * 500]
    for batch in (clean, dirty, dirty, clean, clean, dirty + clean, clean):
        assert tok.encode_batch(batch) == orc.encode_batch(batch)
        for t in batch[:2]:
            assert tok.encode(t) == orc.twin.encode(t)


def test_nfc_with_long_combining_sequences(built_lib, tok_paths):
    """"Zalgo" text: far more combining marks on one base than the normaliser's streaming buffer holds (48).  Such
    segments are redone by the O(1)-memory multi-pass routine (nfc.cu: nfc_segment_slow): canonical reordering over the
    whole run, composition with the base, Hangul chains.  Ids equal the oracle's; the raw decode equals
    unicodedata.normalize('NFC', text) (normalizers.rs:45-47)."""
    rng = np.random.default_rng(99)
    marks = [chr(c) for c in list(range(0x300, 0x34F)) + list(range(0x591, 0x5BE)) + list(range(0x64B, 0x653)) + [0x0F71, 0x0F72, 0x0F74, 0x1DC0, 0x20D0, 0x302A, 0x3099]]
    bases = ['a', 'e', 'o', 'A', 'x', 'ᄀ', '가', 'क', 'α', ' ', '中']
    docs = []
    for n in (40, 47, 48, 49, 50, 60, 100, 257, 1000):
        for _ in range(3):
            s = 'start '
            for _ in range(int(rng.integers(1, 4))):
                s += bases[int(rng.integers(0, len(bases)))] + ''.join(marks[int(i)] for i in rng.integers(0, len(marks), size=n))
                s += ['', ' ', 'ᅡᆨ', 'b'][int(rng.integers(0, 4))]
            docs.append(s + ' end')
    docs.append('́' * 300)                                         # marks with no base at the start of a document
    docs.append('a' + '̧́' * 200 + 'z')                              # two classes alternating: stable order inside each class
    for cfg in ('config1', 'config3'):
        tok, orc = _tok(tok_paths[cfg]), _oracle(tok_paths[cfg])
        got, want = tok.encode_batch(docs), orc.encode_batch(docs)
        bad = [i for i, (g, w) in enumerate(zip(got, want)) if g != w]
        assert not bad, bad[:5]
        assert tok.decode_batch_with_options(got, False, False) == [unicodedata.normalize('NFC', d) for d in docs]
