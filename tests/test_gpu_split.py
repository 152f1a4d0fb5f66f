"""Split pre-tokenizer stages on the device (reference src/pretokenizers.rs:298-433 through the Sequence arm
:114-124) against the oracle: ids bit-exact, every behaviour, one and two stages, with NFC and add_prefix_space.
Needs a GPU: run with -m gpu."""
import json
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from test_gpu_parity import EDGE_TEXTS

GPT2_RX = r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+"
CASES = [   # (name, base config, split stages [(pattern, behavior, invert)], add_prefix_space)
    ('digit_isolated', 'config2', [(r'\d', 'Isolated', False)], False),
    ('num3_isolated', 'config1', [(r'\p{N}{1,3}', 'Isolated', False)], False),
    ('deepseek_like', 'config3', [(r'\p{N}{1,3}', 'Isolated', False), (r'[一-龥぀-ゟ゠-ヿ]+', 'Isolated', False)], False),
    ('ws_merged_next', 'config1', [(r'\s+', 'MergedWithNext', False)], False),
    ('ws_merged_prev', 'config3', [(r'\s', 'MergedWithPrevious', False)], False),
    ('punct_contiguous', 'config1', [(r'[^\s\p{L}\p{N}]', 'Contiguous', False)], True),
    ('gpt2_removed', 'config2', [(GPT2_RX, 'Removed', False)], False),
    ('word_removed', 'config3', [(r'\w+', 'Removed', False)], False),
    ('ws_removed_inverted', 'config1', [(r'\s+', 'Removed', True)], True),
    ('lazy_alt', 'config2', [(r"[a-z]+?[aeiou]|\d{2,}?", 'Isolated', False)], False),
    ('lookahead_passthrough', 'config2', [(r'\s+(?!\S)', 'Isolated', False), (r'\d', 'Isolated', False)], False),
]


def _make(tok_paths, base, stages, aps):
    with open(tok_paths[base], encoding='utf-8') as f:
        tj = json.load(f)
    seq = [{'type': 'Split', 'pattern': {'Regex': p}, 'behavior': b, 'invert': inv} for p, b, inv in stages]
    seq.append({'type': 'ByteLevel', 'add_prefix_space': aps, 'use_regex': False, 'trim_offsets': True})
    tj['pre_tokenizer'] = {'type': 'Sequence', 'pretokenizers': seq}
    fd, path = tempfile.mkstemp(suffix='.json')
    with os.fdopen(fd, 'w', encoding='utf-8') as f:
        json.dump(tj, f, ensure_ascii=False)
    return path


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_split_ids_bit_exact(built_lib, tok_paths, case):
    import c_oracle
    import complexity_tokenizer as ct
    import synth
    name, base, stages, aps = case
    path = _make(tok_paths, base, stages, aps)
    try:
        tok = ct.Tokenizer.from_file(path)
        orc = c_oracle.COracle.from_file(path)
        texts = list(EDGE_TEXTS) + ["a1b22c333d4444 55555 666666", "1", "12 ", " 3", "中文123かな45カナ", "x́ 9٣"]
        got = tok.encode_batch(texts)
        want = orc.encode_batch(texts)
        twin = orc.twin.encode_batch(texts)
        assert want == twin
        for t, g, w in zip(texts, got, want):
            assert g == w, (name, t)
        kind = {'config1': 'english', 'config2': 'ascii', 'config3': 'mixed'}[base]
        text, offs = synth.gen_corpus(kind, 77, 768 << 10)
        ids, ioff = tok.encode_packed(text, offs)
        wids, woff = orc.encode_packed(text, offs)
        assert np.array_equal(ioff, woff)
        assert np.array_equal(ids, wids)
    finally:
        os.unlink(path)


def test_split_equals_base_tokenizer_on_the_pieces(built_lib, tok_paths):
    """size-independent property at 64 MiB: with Split(\\d, Isolated) every document's ids are the concatenation of the base
    tokenizer's ids of its pieces (the ByteLevel stage runs per piece, pretokenizers.rs:114-124); the pieces come from NumPy"""
    import complexity_tokenizer as ct
    import synth
    path = _make(tok_paths, 'config2', [(r'\d', 'Isolated', False)], False)
    try:
        tok = ct.Tokenizer.from_file(path)
        base = ct.Tokenizer.from_file(tok_paths['config2'])
        text, offs = synth.gen_corpus('ascii', 2002, 64 << 20, doc_median=4096, doc_min=256, doc_max=65536)
        ids, ioff = tok.encode_packed(text, offs)
        digit = (text >= 0x30) & (text <= 0x39)
        cut = np.zeros(text.size + 1, dtype=bool)
        cut[:-1] |= digit                                     # a piece starts at every digit ...
        cut[1:] |= digit                                      # ... and after every digit
        cut[offs.astype(np.int64)] = True                     # and at every document start / the end
        poff = np.flatnonzero(cut).astype(np.uint64)
        pids, pioff = base.encode_packed(text, poff)
        assert np.array_equal(pids, ids)
        first = np.searchsorted(poff, offs)                   # first piece of every document
        assert np.array_equal(pioff[first], ioff)
        b, boff = tok.decode_packed(ids, ioff, False, False)
        assert np.array_equal(boff, offs) and b.tobytes() == text.tobytes()
    finally:
        os.unlink(path)


def test_split_unsupported_and_encoding_outputs(built_lib, tok_paths):
    import complexity_tokenizer as ct
    for rx in (r'(?i)a', r'^\d', r'a*', r'\p{Tamil}'):
        path = _make(tok_paths, 'config1', [(rx, 'Isolated', False)], False)
        try:
            with pytest.raises(Exception) as ei:
                ct.Tokenizer.from_file(path)
            assert 'subset' in str(ei.value) or 'unsupported' in str(ei.value).lower()
        finally:
            os.unlink(path)
    path = _make(tok_paths, 'config1', [(r'\d', 'Isolated', False)], False)
    try:
        tok = ct.Tokenizer.from_file(path)
        with pytest.raises(Exception):
            tok.encode_batch_to_encoding(["a1b"])
        assert tok.encode("") == [] and tok.encode_batch([]) == []
    finally:
        os.unlink(path)


def test_split_tokenizer_train_new_from_iterator(built_lib, tok_paths):
    """train_new_from_iterator (mod.rs:1231-1275) with a Split stage in the pipeline: the trainer's words are the pre-tokens of the pieces"""
    import complexity_tokenizer as ct
    import py_oracle
    import synth
    path = _make(tok_paths, 'config1', [(r'\p{N}{1,3}', 'Isolated', False)], False)
    try:
        tok = ct.Tokenizer.from_file(path)
        orc = py_oracle.OracleTokenizer.from_file(path)
        text, offs = synth.gen_corpus('english', 99, 256 << 10)
        raw = text.tobytes()
        docs = [raw[int(offs[i]):int(offs[i + 1])].decode('utf-8') for i in range(40)] + ["2024 12345 7 1000000 3.14159"] * 3
        new = tok.train_new_from_iterator(docs, 450)
        want_vocab, want_merges = orc.train_new_from_iterator(docs, 450)
        tj = new._config_json()
        assert tj['model']['vocab'] == want_vocab and tj['model']['merges'] == [a + ' ' + b for a, b in want_merges]
        assert new.encode_batch(docs[-2:]) == py_oracle.OracleTokenizer(tj).encode_batch(docs[-2:])
    finally:
        os.unlink(path)
