"""Pin the oracle: every known answer the reference's own unit tests hold for the hot path
(SURVEY.md section 4 / 8(c)), plus the committed HuggingFace-tokenizers cross-check vectors.
CPU only."""
import json
import os
import unicodedata

import pytest

import py_oracle
from py_oracle import OracleTokenizer

HERE = os.path.dirname(os.path.abspath(__file__))


def _tj(vocab, merges, **extra):
    tj = {'model': {'type': 'BPE', 'vocab': vocab, 'merges': merges}}
    tj.update(extra)
    return tj


def test_bpe_rs_220_250_hello_is_8():
    """/root/reference/src/bpe.rs:220-250  test_basic_encode_decode: encode("hello") == [8]"""
    vocab = {'h': 0, 'e': 1, 'l': 2, 'o': 3, 'he': 4, 'll': 5, 'hel': 6, 'hell': 7, 'hello': 8, 'lo': 9, 'llo': 10}
    merges = ['h e', 'he l', 'hel l', 'hell o', 'l l', 'l o', 'l lo']
    t = OracleTokenizer(_tj(vocab, merges))
    assert t.bpe_encode('hello') == [8]
    assert t.encode('hello') == [8]                 # through the whole pipeline: ByteLevel maps ASCII letters to themselves


def test_models_rs_961_space_is_G_dot():
    """/root/reference/src/models.rs:956-969 test_byte_level_mapping: 0x20 <-> U+0120, table is a bijection"""
    assert py_oracle.BYTE_ENCODER[0x20] == 'Ġ'
    assert len(py_oracle.BYTE_ENCODER) == 256 and len(set(py_oracle.BYTE_ENCODER.values())) == 256
    for b, c in py_oracle.BYTE_ENCODER.items():
        assert py_oracle.BYTE_DECODER[c] == b


def test_trainer_rs_660_667_byte_level_encoding():
    """/root/reference/src/trainer.rs:660-667: 256 entries; 'a'->'a', 'Z'->'Z'"""
    assert py_oracle.BYTE_ENCODER[ord('a')] == 'a' and py_oracle.BYTE_ENCODER[ord('Z')] == 'Z'


def test_normalizers_rs_224_230_nfc():
    """/root/reference/src/normalizers.rs:224-230 test_nfc: e + U+0301 -> é ; also through the C core"""
    import c_oracle
    assert unicodedata.normalize('NFC', 'é') == 'é'
    assert c_oracle.nfc('é'.encode()) == 'é'.encode()


def test_pretokenizers_rs_626_630_gpt2():
    """/root/reference/src/pretokenizers.rs:626-630 test_gpt2: "Hello, world!" splits into > 1 pieces (derived: 4)"""
    spans = py_oracle.gpt2_find_iter('Hello, world!')
    assert len(spans) > 1
    assert ['Hello, world!'[a:b] for a, b in spans] == ['Hello', ',', ' world', '!']


def test_mod_rs_1567_1592_load_minimal_json():
    """/root/reference/src/huggingface/mod.rs:1567-1592 test_load_tokenizer_json: 8-entry vocab, no merges"""
    vocab = {c: i for i, c in enumerate('abcdefgh')}
    t = OracleTokenizer(_tj(vocab, []))
    assert t.vocab_size == 8
    assert t.normalizer == 'nfc' and t.pre_stages == [('bytelevel', False)] and t.decoder == 'bytelevel'


def test_decoders_rs_275_280_byte_level_decode():
    """/root/reference/src/decoders.rs:275-280: decode(["ĠHello","Ġworld"]) contains "Hello" (exact: " Hello world")"""
    t = OracleTokenizer(_tj({'ĠHello': 0, 'Ġworld': 1}, []))
    assert t.decode([0, 1], False, False) == ' Hello world'
    assert t.decode([0, 1]) == 'Hello world'        # default clean-up trims


def test_vocab_rs_157_172_special_ids():
    """/root/reference/src/vocab.rs:157-172: <unk>,<s>,</s>,<pad> resolve"""
    vocab = {'<unk>': 0, '<s>': 1, '</s>': 2, '<pad>': 3, 'a': 4}
    added = [{'id': i, 'content': c, 'special': True} for c, i in list(vocab.items())[:4]]
    t = OracleTokenizer(_tj(vocab, [], added_tokens=added))
    assert t.special_tokens == {'<unk>': 0, '<s>': 1, '</s>': 2, '<pad>': 3}
    assert t.decode([4, 1, 4], skip_special_tokens=True) == 'aa'
    assert t.decode([4, 1, 4], skip_special_tokens=False, clean_up_tokenization_spaces=False) == 'a<s>a'


def test_survey_pretoken_examples():
    ex = {"don't": ['don', "'t"], "'sup": ["'s", 'up'], "!'s": ["!'", 's'], 'a  b': ['a', '  ', 'b'], 'a b': ['a', ' b'],
          'x ': ['x', ' '], " 's": [" '", 's'], "  's": ['  ', "'s"]}
    for s, want in ex.items():
        assert [s[a:b] for a, b in py_oracle.gpt2_find_iter(s)] == want


def test_cleanup_vectors():
    """SURVEY.md section 3.3 vectors for mod.rs:749-769"""
    cu = OracleTokenizer.clean_up
    assert cu('Hello , world !') == 'Hello, world!'
    assert cu('a  b\n\nc ') == 'a b c'
    assert cu(' - - - ') == '---'
    assert cu('say " hi " now') == 'say"hi"now'
    assert cu('  .  ,') == '. ,'
    assert cu('( a ) [ b ]') == '(a) [b]'


def test_added_tokens_are_matched_inside_words_only():
    """mod.rs:566-610: '<s>' can never match under ByteLevel; an all-letter added token does."""
    bs = {c: i for i, c in enumerate(py_oracle.BYTE_ENCODER.values())}
    vocab = dict(bs)
    vocab['<s>'] = 300
    vocab['hello'] = 301
    added = [{'id': 300, 'content': '<s>', 'special': True}, {'id': 301, 'content': 'hello', 'special': False}]
    t = OracleTokenizer(_tj(vocab, [], added_tokens=added))
    assert t.encode('<s>') == [bs['<'], bs['s'], bs['>']]
    assert t.encode('hello') == [301]
    assert t.encode(' helloworld') == [bs['Ġ'], 301] + [bs[c] for c in 'world']


def test_merge_rank_compaction_quirk():
    """bpe.rs:61-69 + :141: rank is the original index, merges vector is compacted."""
    vocab = {'a': 0, 'b': 1, 'c': 2, 'ab': 3, 'bc': 4}
    t = OracleTokenizer(_tj(vocab, ['a b', 'b c']))
    assert t.encode('abc') == [3, 2]
    # a skipped first line shifts the compacted vector: rank 1 now indexes past the end -> the reference panics
    t = OracleTokenizer(_tj(vocab, ['x y', 'a b']))
    with pytest.raises(py_oracle.ReferencePanic):
        t.encode('ab')
    # ... or silently takes another merge's new_id when it stays in bounds
    t = OracleTokenizer(_tj(vocab, ['x y', 'a b', 'b c']))
    assert t.merge_ranks == {(0, 1): 1, (1, 2): 2} and t.merge_ops == [3, 4]
    assert t.bpe_encode('ab') == [4]                # merges[1].new_id is 'bc''s id


def test_golden_hf_crosscheck_twin_and_core():
    """tests/golden/hf_crosscheck.json (tools/make_golden.py): ids from HuggingFace tokenizers 0.22.2 configured
    to coincide with the reference; both the Python twin and the C core must reproduce them."""
    import c_oracle
    with open(os.path.join(HERE, 'golden', 'hf_crosscheck.json'), encoding='utf-8') as f:
        g = json.load(f)
    twin = OracleTokenizer(g['tokenizer'])
    assert twin.encode_batch(g['texts']) == g['ids']
    core = c_oracle.COracle(twin)
    assert core.encode_batch(g['texts']) == g['ids']
    # decode raw round trip == NFC(text)
    back = core.decode_batch(g['ids'], False, False)
    assert back == [unicodedata.normalize('NFC', t) for t in g['texts']]
    assert back == twin.decode_batch(g['ids'], False, False)
    assert core.decode_batch(g['ids']) == twin.decode_batch(g['ids'])


def _wide_cases():
    with open(os.path.join(HERE, 'golden', 'hf_crosscheck_wide.json'), encoding='utf-8') as f:
        g = json.load(f)
    for c in g['cases']:
        tj = json.loads(json.dumps(g['tokenizers'][c['tokenizer']]))
        tj['normalizer'] = c['normalizer']
        tj['pre_tokenizer'] = c['pre_tokenizer']
        yield c['name'], tj, c['texts'], c['ids']


def test_golden_hf_wide_twin_and_core():
    """tests/golden/hf_crosscheck_wide.json (tools/make_golden_wide.py): 7 816 texts over 13 pipelines -- three tokenizers, NFC,
    "normalizer": null, add_prefix_space, and Split stages of every behaviour -- ids from HuggingFace tokenizers 0.22.2 set up to
    coincide with the reference; the Python twin and the C core must reproduce every one."""
    import c_oracle
    n = 0
    for name, tj, texts, ids in _wide_cases():
        twin = OracleTokenizer(tj)
        core = c_oracle.COracle(twin)
        assert core.encode_batch(texts) == ids, name
        assert twin.encode_batch(texts[::9]) == ids[::9], name
        n += len(texts)
    assert n >= 7000


def test_lossy_decode_matches_python_replace():
    """String::from_utf8_lossy (decoders.rs:118): maximal-subpart replacement == bytes.decode('utf-8','replace')"""
    import c_oracle
    import numpy as np
    bs = list(py_oracle.BYTE_ENCODER.items())
    vocab = {c: b for b, c in bs}                   # id == byte value
    core = c_oracle.COracle(OracleTokenizer(_tj(vocab, [])))
    rng = np.random.default_rng(3)
    pool = [0x41, 0x20, 0xC3, 0xA9, 0xE2, 0x82, 0xAC, 0xF0, 0x9F, 0x98, 0x80, 0xC0, 0xED, 0xA0, 0x80, 0xF4, 0x90, 0xFF, 0xE0, 0x9F]
    batch = [[pool[int(k)] for k in rng.integers(0, len(pool), size=int(rng.integers(0, 12)))] for _ in range(3000)]
    got = core.decode_batch(batch, False, False)
    assert got == [bytes(b).decode('utf-8', 'replace') for b in batch]


def test_metaspace_known_answers():
    """the reference's own tests for the Metaspace pre-tokenizer (pretokenizers.rs:633-640) and decoder (decoders.rs:255-262)"""
    vocab = {'<unk>': 0, '\u2581': 1, 'H': 2, 'e': 3, 'l': 4, 'o': 5, 'w': 6, 'r': 7, 'd': 8, 'h': 9, '\u2581Hello': 10, '\u2581world': 11}
    tj = {'model': {'type': 'BPE', 'vocab': vocab, 'merges': []}, 'pre_tokenizer': {'type': 'Metaspace', 'replacement': '\u2581', 'add_prefix_space': True},
          'decoder': {'type': 'Metaspace', 'replacement': '\u2581', 'add_prefix_space': True}, 'added_tokens': []}
    t = OracleTokenizer(tj)
    words = t.pre_tokenize('hello world')
    assert words[0].startswith('\u2581') and words == ['\u2581hello\u2581world']       # U+0020 is replaced before the split: one word
    assert t.pre_tokenize('a\tb  c\n') == ['\u2581a', 'b\u2581\u2581c']
    assert t.decode([10, 11], False, False) == 'Hello world'
    assert t.decode([11], False, False) == 'world' and t.decode([], False, False) == ''
    assert t.encode('') == [1]                                                          # the prefix alone is a word
    tj['pre_tokenizer'] = {'type': 'Metaspace'}                                         # defaults: parsing.rs:108-123
    assert OracleTokenizer(tj).pre_stages == [('metaspace', ('\u2581', True))]
