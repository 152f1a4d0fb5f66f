"""Metaspace pipelines on the device (reference src/pretokenizers.rs:188-200, src/decoders.rs:121-131,
src/huggingface/parsing.rs:108-123, 279-293) against the oracle's Python twin: ids and decoded strings bit-exact.
Needs a GPU: run with -m gpu."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _corpus(n_docs=160):
    import synth
    out = []
    for kind, seed in (('english', 31), ('mixed', 32)):
        t, o = synth.gen_corpus(kind, seed, 96 << 10, doc_median=200, doc_min=8, doc_max=1200)
        out += [d.decode('utf-8') for d in synth.split_docs(t, o)][:n_docs // 2]
    return out


def make_metaspace_tokenizer(docs, rep='▁', aps=True, dec_aps=True, n_merges=350, pre_extra=None, added=None):
    """a Metaspace BPE tokenizer.json trained (oracle/py_trainer.py) on the Metaspace words of `docs`"""
    import unicodedata
    import py_oracle
    import py_trainer
    pre = {'type': 'Metaspace', 'replacement': rep, 'add_prefix_space': aps}
    if pre_extra:
        pre = {'type': 'Sequence', 'pretokenizers': pre_extra + [pre]}
    specials = ['<unk>', '<s>', '</s>']
    shell = {'model': {'type': 'BPE', 'vocab': {}, 'merges': []}, 'pre_tokenizer': pre, 'decoder': {'type': 'Metaspace', 'replacement': rep, 'add_prefix_space': dec_aps},
             'added_tokens': []}
    twin = py_oracle.OracleTokenizer(shell)
    words = []
    for d in docs:
        words += twin.pre_tokenize(unicodedata.normalize('NFC', d))
    n_chars = len(set(''.join(words)))
    vocab, merges = py_trainer.train_bpe(words, vocab_size=len(specials) + n_chars + n_merges, min_frequency=2, special_tokens=specials)
    tj = dict(shell)
    tj['model'] = {'type': 'BPE', 'vocab': vocab, 'merges': [a + ' ' + b for a, b in merges]}
    tj['added_tokens'] = [{'id': vocab[s], 'content': s, 'special': True, 'single_word': False, 'lstrip': False, 'rstrip': False, 'normalized': False} for s in specials]
    tj['added_tokens'] += added or []
    return tj


EDGE = ["", " ", "  ", "a", "Hello world", " leading", "trailing ", "two  spaces", "tab\tand\nnewline", "été À la carte", "é decomposed",
        "中文 字符　全角 space", "x" * 100, "word " * 40, "<s>inside</s> words<unk>", "a<s>b", "▁ already ▁there▁", "emoji \U0001F600 \U0001F44D\U0001F3FD",
        " line sep nbsp ogham", "unknown ჯ ⵣ chars", "12345 67.89", "</s>", "<s", "mixed ▁x y▁ z"]


@pytest.mark.parametrize('variant', ['default', 'no_prefix', 'underscore', 'split_digits', 'added_plain'])
def test_metaspace_encode_decode_bit_exact(built_lib, variant):
    import complexity_tokenizer as ct
    import py_oracle
    docs = _corpus()
    kw = {}
    if variant == 'no_prefix':
        kw = dict(aps=False, dec_aps=False)
    elif variant == 'underscore':
        kw = dict(rep='_')
    elif variant == 'split_digits':
        kw = dict(pre_extra=[{'type': 'Split', 'pattern': {'Regex': r'\d'}, 'behavior': 'Isolated', 'invert': False}])
    elif variant == 'added_plain':
        kw = dict(added=[{'id': 90000, 'content': 'ing', 'special': False, 'single_word': False, 'lstrip': False, 'rstrip': False, 'normalized': False},
                         {'id': 90001, 'content': '▁the', 'special': False, 'single_word': False, 'lstrip': True, 'rstrip': False, 'normalized': False},
                         {'id': 90002, 'content': 'ed', 'special': False, 'single_word': False, 'lstrip': False, 'rstrip': True, 'normalized': False}])
    tj = make_metaspace_tokenizer(docs, **kw)
    twin = py_oracle.OracleTokenizer(tj)
    tok = ct.Tokenizer.from_str(json.dumps(tj, ensure_ascii=False))
    texts = EDGE + docs
    got = tok.encode_batch(texts)
    want = twin.encode_batch(texts)
    for t, g, w in zip(texts, got, want):
        assert g == w, (variant, t[:60])
    assert sum(map(len, want)) > 5000
    for t in texts[::11]:
        assert tok.encode(t) == twin.encode(t)
    for skip, clean in ((False, True), (False, False), (True, False), (True, True)):
        assert tok.decode_batch_with_options(want, skip, clean) == twin.decode_batch(want, skip, clean), (variant, skip, clean)
    assert tok.decode_batch(want) == twin.decode_batch(want)
    assert tok.decode([10 ** 9] + want[4]) == twin.decode(want[4])
    # packed API, uint16 ids
    blobs = [t.encode('utf-8') for t in texts]
    offs = np.zeros(len(blobs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(b) for b in blobs])
    ids, ioff = tok.encode_packed(np.frombuffer(b''.join(blobs), dtype=np.uint8), offs)
    assert ids.tolist() == [i for w in want for i in w] and ioff.tolist() == np.concatenate([[0], np.cumsum([len(w) for w in want])]).tolist()


def test_metaspace_reference_known_answers_and_unsupported(built_lib):
    """pretokenizers.rs:633-640, decoders.rs:255-262; Encoding outputs and single_word added tokens are refused, not guessed"""
    import complexity_tokenizer as ct
    import py_oracle
    tj = make_metaspace_tokenizer(_corpus(40), n_merges=60)
    twin = py_oracle.OracleTokenizer(tj)
    assert twin.pre_tokenize("hello world")[0].startswith('▁')
    tok = ct.Tokenizer.from_str(json.dumps(tj, ensure_ascii=False))
    ids = [tj['model']['vocab'][c] for c in '▁Hello▁world' if c in tj['model']['vocab']]
    assert tok.decode_with_options(ids, False, False) == twin.decode(ids, False, False)
    assert twin.decode(ids, False, False) == ''.join(c for c in 'Hello world' if c in tj['model']['vocab'] or c == ' ')
    with pytest.raises(Exception):
        tok.encode_batch_to_encoding(["a b"])
    bad = json.loads(json.dumps(tj))
    bad['added_tokens'].append({'id': 5, 'content': 'foo', 'special': False, 'single_word': True, 'lstrip': False, 'rstrip': False, 'normalized': False})
    with pytest.raises(Exception):
        ct.Tokenizer.from_str(json.dumps(bad, ensure_ascii=False))
    with pytest.raises(py_oracle.Unsupported):
        py_oracle.OracleTokenizer(bad)
