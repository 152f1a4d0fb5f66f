"""CPU tests of the rich-`Encoding` oracle (oracle/py_encoding.py) against the known answers the reference's own unit
tests hold (src/encoding.rs:464-576, src/postprocessors.rs:295-356), of the loader's post-processor reduction, and of the
host-side `Encoding` object of the shim (single-object pad / truncate / look-ups; no device needed)."""
import json

import numpy as np
import pytest


def test_from_ids_known_answer():                    # encoding.rs:467-476
    import py_encoding as pe
    e = pe.Enc.from_ids([1, 2, 3], ['a', 'b', 'c'])
    assert len(e) == 3 and e.attention_mask == [1, 1, 1] and e.type_ids == [0, 0, 0] and e.sequence_ids == [0, 0, 0]


def test_padding_known_answer():                     # encoding.rs:479-488
    import py_encoding as pe
    e = pe.Enc.from_ids([1, 2], ['a', 'b'])
    e.pad(5, 0, '<pad>', False)
    assert len(e) == 5 and e.attention_mask == [1, 1, 0, 0, 0] and e.sequence_ids == [0, 0, None, None, None]


def test_truncation_known_answer():                  # encoding.rs:491-501
    import py_encoding as pe
    e = pe.Enc.from_ids([1, 2, 3, 4, 5], list('abcde'))
    e.truncate(3)
    assert len(e) == 3 and len(e.overflowing) == 1 and len(e.overflowing[0]) == 2


def test_post_processors_known_answers():            # postprocessors.rs:299-323
    import py_encoding as pe
    assert pe.post_process(('bert', 101, 102), [1, 2, 3]) == [101, 1, 2, 3, 102]
    assert pe.post_process(('roberta', 0, 2), [1, 2, 3]) == [0, 1, 2, 3, 2]
    # default_postprocessor (postprocessors.rs:282-292) applied to a single sequence
    assert pe.template_process([7, 8], None, '<s> $A </s>', '<s> $A </s> $B </s>', {'<s>': 2, '</s>': 0}) == [2, 7, 8, 0]
    assert pe.template_process([7], [9], '<s> $A </s>', '<s> $A </s> $B </s>', {'<s>': 2, '</s>': 0}) == [2, 7, 0, 9, 0]


def test_shim_encoding_object_matches_reference_tests():
    """The shim's host-side Encoding (single-object methods) on the reference's own test cases (encoding.rs:467-576)."""
    from complexity_tokenizer import Encoding
    e = Encoding.from_ids([1, 2, 3], ['a', 'b', 'c'])
    assert len(e) == 3 and e.attention_mask == [1, 1, 1] and e.type_ids == [0, 0, 0] and e.sequence_ids == [0, 0, 0]
    e = Encoding.from_ids([1, 2], ['a', 'b'])
    e.pad(5, 0, '<pad>', False)
    assert len(e) == 5 and e.attention_mask == [1, 1, 0, 0, 0] and e.sequence_ids == [0, 0, None, None, None]
    assert e.tokens == ['a', 'b', '<pad>', '<pad>', '<pad>'] and e.special_tokens_mask == [0, 0, 1, 1, 1]
    e = Encoding.from_ids([1, 2, 3, 4, 5], list('abcde'))
    e.truncate(3)
    assert len(e) == 3 and e.n_overflowing == 1 and len(e.overflowing[0]) == 2 and e.overflowing[0].ids == [4, 5]
    e = Encoding.from_ids([1, 2, 3], ['hello', ' ', 'world'])
    e._offsets = np.array([(0, 5), (5, 6), (6, 11)], dtype=np.uint64)
    assert [e.char_to_token(p) for p in (0, 4, 5, 6, 11)] == [0, 0, 1, 2, None]
    assert e.token_to_chars(0) == (0, 5) and e.token_to_chars(3) is None
    e = Encoding.from_ids([1, 2, 3, 4], ['hel', 'lo', 'wor', 'ld'])
    e._word_ids = np.array([0, 0, 1, 1])
    e._offsets = np.array([(0, 3), (3, 5), (6, 9), (9, 11)], dtype=np.uint64)
    assert e.word_to_tokens(0) == (0, 2) and e.word_to_tokens(1) == (2, 4) and e.word_to_tokens(2) is None
    assert e.word_to_chars(0) == (0, 5) and e.word_to_chars(1) == (6, 11)
    e = Encoding.from_ids([1, 2, 3, 4, 5], list('abcde'))
    e._word_ids = np.array([0, 0, 1, 2, 2])
    assert e.n_words == 3
    # stride windows (encoding.rs:183-232) against the oracle's restatement
    import py_encoding as pe
    a, b = Encoding.from_ids(list(range(23)), ['t'] * 23), pe.Enc.from_ids(list(range(23)), ['t'] * 23)
    a.truncate_with_stride(8, 3)
    b.truncate_with_stride(8, 3)
    assert a.ids == b.ids and [o.ids for o in a.overflowing] == [o.ids for o in b.overflowing]


def _rich(tj):
    import py_encoding as pe
    import py_oracle
    return pe.RichOracle(py_oracle.OracleTokenizer(tj), tj)


TEMPLATE = {'type': 'TemplateProcessing',
            'single': [{'SpecialToken': {'id': '<s>', 'type_id': 0}}, {'Sequence': {'id': 'A', 'type_id': 0}},
                       {'SpecialToken': {'id': '</s>', 'type_id': 0}}],
            'pair': [{'SpecialToken': {'id': '<s>', 'type_id': 0}}, {'Sequence': {'id': 'A', 'type_id': 0}},
                     {'SpecialToken': {'id': '</s>', 'type_id': 0}}, {'Sequence': {'id': 'B', 'type_id': 1}},
                     {'SpecialToken': {'id': '</s>', 'type_id': 1}}]}


def test_offsets_follow_the_find_chain(small_tok_json):
    """mod.rs:447-478 by hand: words found from a running position; a whitespace word is never found (its mapped form is
    not in the text) and advances the position by the MAPPED length, after which later words are looked for too far right."""
    r = _rich(json.loads(small_tok_json))
    e = r.encode_to_encoding('ab cd')
    assert e.offsets[0][0] == 0 and e.offsets[-1][1] == 5
    assert e.word_ids == sorted(e.word_ids) and e.word_ids[0] == 0 and e.word_ids[-1] == 1
    spans = r.pre_tokenize_with_offsets('a\n\nb a', 'a\n\nb a')
    assert [(s, t) for _, s, t in spans] == [(0, 1), (1, 5), (5, 6), (6, 6)]   # "\n\n" -> 4 mapped bytes; "b" not found after 5; " a" not found
    with pytest.raises(Exception):
        r.pre_tokenize_with_offsets('\né\n', '\né\n')                              # running position lands inside 'é': panic


def test_loader_reduces_post_processors_like_the_oracle(built_lib, small_tok_json):
    """The C++ loader's reduction of the post-processor to items == what process(ids, None) does in the oracle."""
    import ctypes
    import py_encoding as pe
    tj = json.loads(small_tok_json)
    specials = {t['content']: t['id'] for t in tj['added_tokens'] if t['special']}
    cases = [None, TEMPLATE, {'type': 'RobertaProcessing'}, {'type': 'BertProcessing'}, {'type': 'ByteLevel'},
             {'type': 'TemplateProcessing'},
             {'type': 'TemplateProcessing', 'single': [{'Sequence': {'id': 'A'}}, {'SpecialToken': {'id': '<nope>'}},
                                                       {'Sequence': {'id': 'B'}}, {'Sequence': {'id': 'A'}}, {'SpecialToken': {'id': '<pad>'}}]}]
    lib = built_lib
    lib.ctk_debug_post_processor.restype = ctypes.c_int
    for pp in cases:
        tj2 = dict(tj)
        if pp is not None:
            tj2['post_processor'] = pp
        data = json.dumps(tj2).encode()
        items = (ctypes.c_int64 * 64)()
        n = ctypes.c_size_t()
        assert lib.ctk_debug_post_processor(data, len(data), items, 64, ctypes.byref(n)) == 0
        got = []
        for it in list(items)[:n.value]:
            got += [1000, 1001, 1002] if it < 0 else [int(it)]
        parsed = pe.parse_post_processor(tj2.get('post_processor'), specials)
        want = pe.post_process(parsed, [1000, 1001, 1002]) if parsed else [1000, 1001, 1002]
        assert got == want, pp
