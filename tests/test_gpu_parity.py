"""Parity proper: the CUDA path, called through the C ABI (ctypes shim), against the oracle.
Bit-exact for ids and decoded bytes.  Needs a GPU: run with -m gpu."""
import unicodedata

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EDGE_TEXTS = [
    "", " ", "  ", "a", "Hello, world!", "don't", "'sup", "!'s", "a  b", "a b", "x ", " 's", "  's", "I'll've",
    "it's 12:30 o'clock", "tabs\tand\nnewlines\r\n\r\n  indented", "été À la carte — naïve café",
    "é (decomposed) vs é", "中文字符和日本語のテキスト。「引用」　全角",
    "emoji \U0001F600\U0001F603 \U0001F44D\U0001F3FD \U0001F468‍\U0001F469‍\U0001F467 end",
    "<s>literal specials</s><pad><unk><|endoftext|>", "Å 豈 ohm Ω", "x" * 31, "y" * 32, "z" * 33,
    "w" * 100 + " " + "v" * 257, " " * 40, "=-_*" * 30, "123456789012345678901234567890123456",
    "a b  ", "  non-breaking 　spaces here", "I'm he'd we're they've she'll 's't",
    "n't 'S 'RE upper'S", "''''s'''t", "mixed123abc456 7.5% (ok) [x] {y}", "\n", "\n\n\n", "a\n", " a", "a ",
    "\x00nul\x00\x00", "각 각 hangul", "ạ̇ ṩ ṩ",
]


def _oracle(path):
    import c_oracle
    return c_oracle.COracle.from_file(path)


def _tok(path):
    import complexity_tokenizer as ct
    return ct.Tokenizer.from_file(path)


@pytest.mark.parametrize('cfg', ['config1', 'config2', 'config3'])
def test_edge_texts_bit_exact(built_lib, tok_paths, cfg):
    tok, orc = _tok(tok_paths[cfg]), _oracle(tok_paths[cfg])
    got = tok.encode_batch(EDGE_TEXTS)
    want = orc.encode_batch(EDGE_TEXTS)
    twin = orc.twin.encode_batch(EDGE_TEXTS)
    assert want == twin
    for t, g, w in zip(EDGE_TEXTS, got, want):
        assert g == w, repr(t)
    for t in EDGE_TEXTS[:12]:                      # single-text entry point
        assert tok.encode(t) == orc.twin.encode(t)


@pytest.mark.parametrize('cfg,kind,seed,size', [('config1', 'english', 1001, 12 << 20), ('config2', 'ascii', 2002, 16 << 20),
                                                ('config3', 'mixed', 3003, 8 << 20)])
def test_corpus_bit_exact(built_lib, tok_paths, cfg, kind, seed, size):
    import synth
    tok, orc = _tok(tok_paths[cfg]), _oracle(tok_paths[cfg])
    text, offs = synth.gen_corpus(kind, seed, size, doc_median=1024 if cfg == 'config1' else 4096)
    ids, ioff = tok.encode_packed(text, offs)
    wids, woff = orc.encode_packed(text, offs)
    assert np.array_equal(ioff, woff)
    assert np.array_equal(ids, wids)
    # decode: cleaned (default) and raw, against the oracle; raw round trip == NFC(input)
    for skip, clean in ((False, True), (False, False), (True, True)):
        b, boff = tok.decode_packed(ids, ioff, skip, clean)
        wb, wboff = orc.decode_packed(ids, ioff, skip, clean)
        assert np.array_equal(boff, wboff)
        assert np.array_equal(b, wb)
    b, boff = tok.decode_packed(ids, ioff, False, False)
    docs = synth.split_docs(text, offs)
    raw = b.tobytes()
    for i in range(0, len(docs), 37):
        assert raw[int(boff[i]):int(boff[i + 1])] == unicodedata.normalize('NFC', docs[i].decode()).encode()


def test_readme_shape_tiny_texts(built_lib, tok_paths):
    """10 000 texts of 4-12 bytes (the reference README's usage shape, README.md:48-52)."""
    rng = np.random.default_rng(5)
    words = ['Hello', 'World', 'Foo', 'Bar', 'tokenizer', "don't", ' x ', '42', 'é', '中文']
    texts = [''.join(words[int(k)] for k in rng.integers(0, len(words), size=int(rng.integers(1, 3)))) for _ in range(10000)]
    tok, orc = _tok(tok_paths['config1']), _oracle(tok_paths['config1'])
    assert tok.encode_batch(texts) == orc.encode_batch(texts)


def test_empty_and_ragged_batches(built_lib, tok_paths):
    tok, orc = _tok(tok_paths['config2']), _oracle(tok_paths['config2'])
    assert tok.encode_batch([]) == []
    assert tok.encode_batch(['']) == [[]]
    assert tok.encode_batch(['', '', 'a', '']) == orc.encode_batch(['', '', 'a', ''])
    assert tok.decode_batch([]) == []
    assert tok.decode_batch([[]]) == ['']
    ids = orc.encode_batch(['Hello , world !', 'a  b\n\nc ', ' - - - ', 'say " hi " now', '  .  ,', '( a ) [ b ]'])
    assert tok.decode_batch(ids) == ['Hello, world!', 'a b c', '---', 'say"hi"now', '. ,', '(a) [b]']   # SURVEY 3.3 vectors
    assert tok.decode_batch(ids) == orc.decode_batch(ids)
    # unknown ids are dropped; invalid UTF-8 from split multi-byte tokens becomes U+FFFD
    weird = [[999999999, 5, 6], [tok.token_to_id('Ã') or 0], []]
    for opts in ((False, True), (False, False), (True, False)):
        assert tok.decode_batch_with_options(weird, *opts) == orc.decode_batch(weird, *opts)


def test_getters(built_lib, tok_paths):
    tok, orc = _tok(tok_paths['config3']), _oracle(tok_paths['config3'])
    assert tok.vocab_size == orc.twin.vocab_size == 100000
    assert tok.token_to_id('Ġt') == orc.twin.token_to_id('Ġt')
    assert tok.token_to_id('no such token ☃') is None
    assert tok.id_to_token(300) == orc.twin.id_to_token_str(300)
    assert tok.id_to_token(10 ** 9) is None
    assert tok.special_tokens == orc.twin.special_tokens == {'</s>': 0, '<pad>': 1, '<s>': 2, '<unk>': 3}


def test_golden_hf_vectors_on_the_gpu(built_lib):
    """The committed golden vectors (tests/golden/hf_crosscheck.json: ids from HuggingFace tokenizers 0.22.2 set up to
    coincide with the reference's semantics, tools/make_golden.py) straight against the CUDA path: 222 texts incl.
    pre-tokens of 12 .. 5000 bytes (cache, mid, warp-round and grid-round merge paths), as one batch and one by one."""
    import json
    import os
    import complexity_tokenizer as ct
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hf_crosscheck.json'), encoding='utf-8') as f:
        g = json.load(f)
    tok = ct.Tokenizer.from_str(json.dumps(g['tokenizer'], ensure_ascii=False))
    got = tok.encode_batch(g['texts'])
    bad = [i for i, (a, b) in enumerate(zip(got, g['ids'])) if a != b]
    assert not bad, (bad[:5], g['texts'][bad[0]][:60])
    for t, want in list(zip(g['texts'], g['ids']))[::7]:
        assert tok.encode(t) == want
    assert sum(len(t.encode()) for t in g['texts']) * 5 > (256 << 10)          # large enough for the per-class mid kernels
    assert tok.encode_batch(g['texts'] * 5) == g['ids'] * 5
    # decode of the golden ids gives the NFC text back (ByteLevel round trip, decoders.rs:94-119)
    back = tok.decode_batch_with_options(g['ids'], False, False)
    assert back == [unicodedata.normalize('NFC', t) for t in g['texts']]


def test_golden_hf_wide_vectors_on_the_gpu(built_lib):
    """tests/golden/hf_crosscheck_wide.json straight against the CUDA path: 7 816 texts over 13 pipelines (three tokenizers, NFC,
    "normalizer": null, add_prefix_space, Split stages of every behaviour), ids from HuggingFace tokenizers"""
    import json
    import complexity_tokenizer as ct
    from test_oracle_known_answers import _wide_cases
    n = 0
    for name, tj, texts, ids in _wide_cases():
        tok = ct.Tokenizer.from_str(json.dumps(tj, ensure_ascii=False))
        got = tok.encode_batch(texts)
        bad = [i for i, (a, b) in enumerate(zip(got, ids)) if a != b]
        assert not bad, (name, bad[:5], texts[bad[0]][:60])
        n += len(texts)
    assert n >= 7000
