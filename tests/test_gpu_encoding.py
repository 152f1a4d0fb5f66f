"""Parity of the rich `Encoding` outputs: the CUDA path, through the C ABI and the shim, against the
oracle's restatement of mod.rs:340-545 / encoding.rs / postprocessors.rs / bindings/tokenizer.rs:33-201.
Bit-exact: ids, masks, type ids, tokens, byte offsets, word ids, sequence ids, overflowing windows.  Needs a GPU."""
import json

import numpy as np
import pytest

from test_encoding_oracle import TEMPLATE

pytestmark = pytest.mark.gpu

TEXTS = [
    "", " ", "a", "Hello, world!", "Hello world hello world", "don't stop", "a  b", "a b c d e f g", "x ", "  lead",
    "it's 12:30 o'clock", "tabs\tand\nnewlines\r\n\r\n  indented and more words after the break and more",
    "one two three\n\nfour five six seven four five", "the the the the\nthe the the", "été À la carte", "naïve café\nrésumé",
    "<s>literal specials</s><pad>", "x" * 40 + " " + "y" * 300, " " * 40, "=-_*" * 30 + " end", "a\n", "\n", "\n\n\nabc abc",
    "Ġ literal G-dot Ġword ĠĠ", "Ċ and Ċ again\nĊ", "word " * 50, "ab" * 20 + "\n" + "ab" * 20 + " ab ab", "abc\tabc abc",
    "emoji \U0001F600 end", "中文 abc", "café café café", "Ã© mojibake then é real",
    "é decomposed", "ohm Ω sign", "I'll've 's't", "a" * 1000, "b c " * 200,
]
PAIRS = [("Hello world", "second text here"), ("", "only b"), ("only a", ""), ("a b c", "a b c"), ("x\ny", "p q\n\nr")]


def _tj(small_tok_json, pp):
    tj = json.loads(small_tok_json)
    if pp is not None:
        tj['post_processor'] = pp
    return tj


def _both(tj):
    import complexity_tokenizer as ct
    import py_encoding as pe
    import py_oracle
    s = json.dumps(tj, ensure_ascii=False)
    return ct.Tokenizer.from_str(s), pe.RichOracle(py_oracle.OracleTokenizer(tj), tj)


def _same(got, want, what=''):
    assert got.ids == want.ids, what
    assert got.type_ids == want.type_ids, what
    assert got.attention_mask == want.attention_mask, what
    assert got.special_tokens_mask == want.special_tokens_mask, what
    assert got.tokens == want.tokens, what
    assert got.offsets == [tuple(o) for o in want.offsets], what
    assert got.word_ids == want.word_ids, what
    assert got.sequence_ids == want.sequence_ids, what
    assert got.n_overflowing == len(want.overflowing), what
    for a, b in zip(got.overflowing, want.overflowing):
        _same(a, b, what + ' (overflow)')


def _split_panics(orc, texts):
    import py_oracle
    ok, bad = [], []
    for t in texts:
        try:
            orc.encode_to_encoding(t)
            ok.append(t)
        except py_oracle.ReferencePanic:
            bad.append(t)
    return ok, bad


PPS = [None, TEMPLATE, {'type': 'RobertaProcessing'}, {'type': 'BertProcessing'}]


@pytest.mark.parametrize('pp', PPS, ids=['none', 'template', 'roberta', 'bert'])
def test_encode_batch_to_encoding_bit_exact(built_lib, small_tok_json, pp):
    import complexity_tokenizer as ct
    tok, orc = _both(_tj(small_tok_json, pp))
    ok, bad = _split_panics(orc, TEXTS)
    assert len(ok) > 25 and bad                               # both kinds are exercised
    got = tok.encode_batch_to_encoding(ok)
    for t, g in zip(ok, got):
        _same(g, orc.encode_to_encoding(t), repr(t))
    for t in ok[:8]:
        _same(tok.encode_to_encoding(t), orc.encode_to_encoding(t), repr(t))
    for t in bad:                                             # the reference panics (mod.rs:461): so does the call
        with pytest.raises(ct.PanicException):
            tok.encode_batch_to_encoding(ok[:3] + [t])
    got = tok.encode_batch_pairs_to_encoding(PAIRS)
    for (a, b), g in zip(PAIRS, got):
        _same(g, orc.encode_to_encoding(a, b), repr((a, b)))


@pytest.mark.parametrize('cfg', ['config1', 'config3'])
def test_real_tokenizers(built_lib, tok_paths, cfg):
    """The 32K (no specials in added_tokens ... ) and 100K (INL-shape, specials as added tokens, NFC) tokenizers."""
    import complexity_tokenizer as ct
    import py_encoding as pe
    import py_oracle
    tj = json.load(open(tok_paths[cfg], encoding='utf-8'))
    tok, orc = ct.Tokenizer.from_file(tok_paths[cfg]), pe.RichOracle(py_oracle.OracleTokenizer(tj), tj)
    ok, _ = _split_panics(orc, TEXTS)
    for t, g in zip(ok, tok.encode_batch_to_encoding(ok)):
        _same(g, orc.encode_to_encoding(t), repr(t))


@pytest.mark.parametrize('kw', [
    dict(), dict(add_special_tokens=False), dict(padding=True), dict(padding='max_length', max_length=20),
    dict(padding='left'), dict(truncation=True, max_length=7), dict(truncation=True, max_length=7, stride=3),
    dict(truncation=True, max_length=9, padding='max_length'), dict(truncation=True, max_length=5, padding=True, add_special_tokens=False),
    dict(padding='longest', truncation=True, max_length=600), dict(truncation=True, max_length=0),
    dict(add_special_tokens=False, truncation=True, max_length=4, stride=1, padding='left'),
])
def test_call_matches_the_reference_semantics(built_lib, small_tok_json, kw):
    tok, orc = _both(_tj(small_tok_json, TEMPLATE))
    ok, _ = _split_panics(orc, TEXTS)
    batch = tok(ok, **kw)
    want = orc.call(ok, **kw)
    assert len(batch) == len(want)
    assert batch.input_ids == [e.ids for e in want]
    assert batch.attention_mask == [e.attention_mask for e in want]
    assert batch.token_type_ids == [e.type_ids for e in want]
    for g, w in zip(batch.encodings(), want):
        _same(g, w, repr(kw))
    # pairs, and the single-text form
    a, b = [p[0] for p in PAIRS], [p[1] for p in PAIRS]
    for g, w in zip(tok(a, b, **kw).encodings(), orc.call(a, b, **kw)):
        _same(g, w, 'pairs ' + repr(kw))
    _same(tok(ok[4], **kw)[0], orc.call(ok[4], **kw)[0], 'single ' + repr(kw))
    _same(tok(a[0], b[0], **kw)[0], orc.call(a[0], b[0], **kw)[0], 'single pair ' + repr(kw))
    if kw.get('padding') and not (kw.get('padding') == 'max_length' and not kw.get('truncation')):
        m = tok(ok, **kw).as_numpy('input_ids')
        assert m.shape[0] == len(ok) and m.tolist() == batch.input_ids


def test_padding_variants_and_truncation_method(built_lib, small_tok_json):
    tok, orc = _both(_tj(small_tok_json, TEMPLATE))
    ok, _ = _split_panics(orc, TEXTS)
    for args in ((None, False), (None, True), (16, False), (16, True), (0, False)):
        for g, w in zip(tok.encode_batch_with_padding(ok, *args), orc.encode_batch_with_padding(ok, *args)):
            _same(g, w, repr(args))
    for g, w in zip(tok.encode_batch_pairs_with_padding(PAIRS, 12, True), orc.encode_batch_with_padding(PAIRS, 12, True, pairs=True)):
        _same(g, w, 'pairs padded')
    long_text = ok[-1]
    for ml, st in ((8, 0), (8, 3), (512, 0), (3, 2)):
        _same(tok.encode_with_truncation(long_text, None, ml, st), orc.encode_to_encoding(long_text, None, ml, st), repr((ml, st)))
        _same(tok.encode_with_truncation(long_text, "pair text", ml, st), orc.encode_to_encoding(long_text, "pair text", ml, st), repr((ml, st)))
    # stride >= max_length only loops forever (encoding.rs:190-193) for a row that is longer than max_length: a short one passes
    _same(tok.encode_with_truncation("a", None, 512, 600), orc.encode_to_encoding("a"), 'short row, degenerate stride')
    assert tok("a", truncation=True, max_length=64, stride=64)[0].ids == tok("a")[0].ids
    with pytest.raises(ValueError):
        tok.encode_with_truncation(long_text, None, 4, 4)
    with pytest.raises(ValueError):
        tok(long_text, truncation=True, max_length=4, stride=9)


def test_corpus_bit_exact(built_lib, tok_paths):
    """Documents of the config-1 corpus (newlines, double spaces: the find chain leaves the rails, as it does in the reference)."""
    import complexity_tokenizer as ct
    import py_encoding as pe
    import py_oracle
    import synth
    tj = json.load(open(tok_paths['config1'], encoding='utf-8'))
    tok, orc = ct.Tokenizer.from_file(tok_paths['config1']), pe.RichOracle(py_oracle.OracleTokenizer(tj), tj)
    text, offs = synth.gen_corpus('english', 1001, 256 << 10)
    raw = text.tobytes()
    docs = [raw[int(offs[i]):int(offs[i + 1])].decode('utf-8') for i in range(len(offs) - 1)]
    got = tok.encode_batch_to_encoding(docs)
    for d, g in zip(docs, got):
        w = orc.encode_to_encoding(d)
        assert g.ids == w.ids and g.offsets == [tuple(o) for o in w.offsets] and g.word_ids == w.word_ids
        assert g.special_tokens_mask == w.special_tokens_mask
    # the ids of the Encoding path are encode_batch's ids when no added token can occur inside a word
    assert [g.ids for g in got] == tok.encode_batch(docs)


def test_ids_ignore_added_tokens_on_this_path(built_lib, small_tok_json):
    """mod.rs:407 calls bpe.encode on whole words: an added token that `encode` would match inside a word is NOT matched here."""
    tj = _tj(small_tok_json, None)
    tj['added_tokens'].append({'id': 70000, 'content': 'zzq', 'special': False, 'single_word': False, 'lstrip': False,
                               'rstrip': False, 'normalized': False})
    tok, orc = _both(tj)
    t = 'azzqb zzq'
    assert 70000 in tok.encode(t) and tok.encode(t) == orc.tok.encode(t)
    _same(tok.encode_to_encoding(t), orc.encode_to_encoding(t))
    assert 70000 not in tok.encode_to_encoding(t).ids
    assert 70000 in tok(t, add_special_tokens=False)[0].ids
    assert tok.encode(t) == orc.tok.encode(t)                 # the switch does not leak into later encode calls


def test_bit_parallel_start_bitmap_equals_the_scalar_predicate(built_lib, tok_paths):
    """k_starts_window (16 bytes per thread, the fused kernel's window logic) == k_starts (scalar, per byte) == the host's scalar
    predicate, on mixed French / CJK / emoji text with document boundaries at arbitrary bytes and ragged sizes."""
    import ctypes
    import os
    import complexity_tokenizer as ct
    import synth
    lib = built_lib
    lib.ctk_debug_starts_device.restype = ctypes.c_int
    lib.ctk_debug_starts_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    tok = ct.Tokenizer.from_file(tok_paths['config3'])
    for kind, seed, size in (('mixed', 3003, 3 << 20), ('english', 1001, (1 << 20) + 12345), ('mixed', 7, 8191), ('mixed', 8, 16), ('ascii', 9, 1)):
        text, offs = synth.gen_corpus(kind, seed, size, doc_median=700)
        text = np.ascontiguousarray(text[:int(offs[-1])])
        n, nd = int(text.size), len(offs) - 1
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        nw = (n + 31) // 32
        got = {}
        for mode in ('window', 'scalar'):
            if mode == 'scalar':
                os.environ['CTK_SCALAR_STARTS'] = '1'
            try:
                out = np.zeros(nw + 1, dtype=np.uint32)
                assert lib.ctk_debug_starts_device(tok._h, text.ctypes.data, n, offs.ctypes.data, nd, out.ctypes.data) == 0
                got[mode] = out[:nw].copy()
            finally:
                os.environ.pop('CTK_SCALAR_STARTS', None)
        host = np.zeros(nw + 2, dtype=np.uint32)
        assert lib.ctk_debug_starts_host(text.ctypes.data, n, offs.ctypes.data, nd, host.ctypes.data) == 0
        assert np.array_equal(got['scalar'], host[:nw]), (kind, size)
        assert np.array_equal(got['window'], host[:nw]), (kind, size)


def test_edges_errors_and_threads(built_lib, small_tok_json):
    """Empty batches, empty texts, argument errors of the C entry point, and concurrent use of one tokenizer from host threads."""
    import ctypes
    import threading
    import complexity_tokenizer as ct
    tok, orc = _both(_tj(small_tok_json, TEMPLATE))
    assert len(tok([])) == 0 and tok([]).input_ids == [] and tok.encode_batch_to_encoding([]) == []
    for g, w in zip(tok(['', '', ''], padding=True).encodings(), orc.call(['', '', ''], padding=True)):
        _same(g, w)
    for g, w in zip(tok(['', ''], add_special_tokens=False, padding='max_length', max_length=3).encodings(),
                    orc.call(['', ''], add_special_tokens=False, padding='max_length', max_length=3)):
        _same(g, w)
    with pytest.raises(TypeError):
        tok(123)
    lib = built_lib
    opt = ct._EncodingOptions(1, 1, 0, 0, 0, 0, 0, 0)            # pair mode with an odd number of texts
    off = np.array([0, 1, 2, 3], dtype=np.uint64)
    buf = np.frombuffer(b'abc', dtype=np.uint8)
    res = ctypes.c_void_p()
    assert lib.ctk_encode_batch_to_encoding(tok._h, buf.ctypes.data, off.ctypes.data, 3, ctypes.byref(opt), ctypes.byref(res)) == ct.CTK_ERR_ARG
    assert lib.ctk_encode_batch_to_encoding(tok._h, buf.ctypes.data, off.ctypes.data, 3, None, ctypes.byref(res)) == ct.CTK_ERR_ARG
    bad = np.array([1, 2, 3, 3], dtype=np.uint64)                # offsets must start at 0
    opt.pair = 0
    assert lib.ctk_encode_batch_to_encoding(tok._h, buf.ctypes.data, bad.ctypes.data, 3, ctypes.byref(opt), ctypes.byref(res)) == ct.CTK_ERR_ARG
    # a template without $A: the reference underflows at mod.rs:377
    tj = _tj(small_tok_json, {'type': 'TemplateProcessing', 'single': [{'SpecialToken': {'id': '<s>', 'type_id': 0}}]})
    t2, _ = _both(tj)
    with pytest.raises(ct.PanicException):
        t2.encode_to_encoding('abc')
    assert t2('abc', add_special_tokens=False)[0].ids == t2.encode('abc')
    # threads: the engine serialises calls; results must not mix
    ok, _ = _split_panics(orc, TEXTS)
    want_call = [e.ids for e in orc.call(ok, padding=True)]
    want_ids = orc.tok.encode_batch(ok)
    want_off = [e.offsets for e in (orc.encode_to_encoding(t) for t in ok)]
    errs = []

    def work(kind):
        try:
            for _ in range(6):
                if kind == 0:
                    assert tok(ok, padding=True).input_ids == want_call
                elif kind == 1:
                    assert tok.encode_batch(ok) == want_ids
                else:
                    assert [e.offsets for e in tok.encode_batch_to_encoding(ok)] == [[tuple(o) for o in w] for w in want_off]
        except BaseException as e:       # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(k % 3,)) for k in range(6)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
