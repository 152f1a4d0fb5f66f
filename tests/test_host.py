"""CPU-only tests of the product's host side: the C-ABI library loads and exports every symbol that
include/ctk.h declares, the tokenizer.json loader follows the reference's rules, and the device
start-predicate (compiled for the host through a debug hook) agrees with the oracle's regex
restatement.  No compute call is made on a GPU here."""
import ctypes
import json
import os
import random
import re

import numpy as np
import pytest

import py_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, 'include', 'ctk.h')).read()
    names = sorted(set(re.findall(r'\b(ctk_[a-z_0-9]+)\s*\(', hdr)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(built_lib, n), 'libctk.so does not export ' + n


def test_no_cpu_fallback_without_device(built_lib, small_tok_json):
    """Without a GPU, constructing a tokenizer must fail loudly (CTK_ERR_CUDA), not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    import complexity_tokenizer as ct
    with pytest.raises(RuntimeError, match='no CUDA device|CUDA'):
        ct.Tokenizer.from_str(small_tok_json)


def _load_only(lib, tj):
    data = json.dumps(tj, ensure_ascii=False).encode() if not isinstance(tj, (bytes, str)) else (tj.encode() if isinstance(tj, str) else tj)
    buf = ctypes.create_string_buffer(data, len(data))
    npairs, vs, nfc, mm = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int(), ctypes.c_int()
    rc = lib.ctk_debug_load_only(ctypes.addressof(buf), len(data), ctypes.byref(npairs), ctypes.byref(vs), ctypes.byref(nfc), ctypes.byref(mm))
    return rc, npairs.value, vs.value, nfc.value, mm.value, (lib.ctk_last_error() or b'').decode()


BASE = {'model': {'type': 'BPE', 'vocab': {'a': 0, 'b': 1, 'c': 2, 'ab': 3, 'bc': 4}, 'merges': ['a b', 'b c']}}


def test_loader_rules(built_lib):
    lib = built_lib
    rc, npairs, vs, nfc, mm, _ = _load_only(lib, BASE)
    assert (rc, npairs, vs, nfc, mm) == (0, 2, 5, 1, 0)                       # missing normalizer => NFC (parsing.rs:89)
    # array-form merges (mod.rs:85-91), lines that do not split into exactly two parts are dropped (:252-264)
    tj = json.loads(json.dumps(BASE)); tj['model']['merges'] = [['a', 'b'], 'b c', 'a b c', 'nospace', ['x'], 7]
    assert _load_only(lib, tj)[:2] == (0, 2)
    # "normalizer": null => NFC ; unknown type => none ; NFKC => unsupported (3)
    for norm, want_rc, want_nfc in ((None, 0, 1), ({'type': 'NFC'}, 0, 1), ({'type': 'Whatever'}, 0, 0),
                                    ({'type': 'Sequence', 'normalizers': [{'type': 'NFC'}]}, 0, 1), ({'type': 'NFKC'}, 3, None),
                                    ({'type': 'Sequence', 'normalizers': [{'type': 'Lowercase'}]}, 3, None)):
        tj = dict(BASE, normalizer=norm)
        r = _load_only(lib, tj)
        assert r[0] == want_rc, (norm, r)
        if want_nfc is not None:
            assert r[3] == want_nfc
    # pre-tokenizers: ByteLevel, Llama-3 style Sequence[Split(look-ahead), ByteLevel] ok; Whitespace etc. unsupported
    llama3 = {'type': 'Sequence', 'pretokenizers': [
        {'type': 'Split', 'pattern': {'Regex': r"(?i:'s|'t)|\s+(?!\S)|\s+"}, 'behavior': 'Isolated', 'invert': False},
        {'type': 'ByteLevel', 'add_prefix_space': False, 'trim_offsets': True, 'use_regex': False}]}
    assert _load_only(lib, dict(BASE, pre_tokenizer=llama3))[0] == 0
    assert _load_only(lib, dict(BASE, pre_tokenizer={'type': 'ByteLevel', 'add_prefix_space': False, 'use_regex': True}))[0] == 0
    assert _load_only(lib, dict(BASE, pre_tokenizer={'type': 'Whitespace'}))[0] == 3
    assert _load_only(lib, dict(BASE, pre_tokenizer={'type': 'Metaspace'}))[0] == 0                  # built: csrc/metaspace.cu
    assert _load_only(lib, dict(BASE, pre_tokenizer={'type': 'Metaspace', 'replacement': ' '}))[0] == 3
    assert _load_only(lib, dict(BASE, decoder={'type': 'Metaspace'}))[0] == 0
    assert _load_only(lib, dict(BASE, pre_tokenizer={'type': 'Unknown'}))[0] == 3
    compilable = {'type': 'Sequence', 'pretokenizers': [{'type': 'Split', 'pattern': {'Regex': r'\d'}, 'behavior': 'Isolated'},
                                                        {'type': 'ByteLevel'}]}
    assert _load_only(lib, dict(BASE, pre_tokenizer=compilable))[0] == 0          # compiled to a DFA at load (regex_dfa.cpp)
    assert _load_only(lib, dict(BASE, pre_tokenizer={'type': 'Split', 'pattern': {'Regex': r'\d'}, 'behavior': 'Isolated'}))[0] == 3   # no ByteLevel stage
    assert _load_only(lib, dict(BASE, pre_tokenizer={'type': 'Sequence', 'pretokenizers': [{'type': 'ByteLevel'}, compilable['pretokenizers'][0]]}))[0] == 3
    # Rust `regex` certainly rejects look-around and back-references (-> the Split stage passes text through); an ESCAPED
    # "(?=" or one inside a character class is no look-around: such a pattern compiles and the Split applies -- compiled when
    # it lies in the supported subset (0), else unsupported (3), never a silent pass-through
    import py_oracle
    for rx, want, rejected in ((r"\s+(?!\S)", 0, True), (r"(?<=a)b", 0, True), (r"(a)\1", 0, True), (r"(?<!x)y", 0, True), (r"a(?=b)", 0, True),
                               (r"\(\?=x\)", 0, False), (r"[(?=]+", 0, False), (r"[\1]", 3, False), (r"\\(?:a|b)", 0, False), (r"[^\]](?i:x)", 3, False),
                               (r"[]](?=x)", 0, True), (r"\\(?=x)", 0, True), (r"(?i)x", 3, False), (r"x*", 3, False)):
        pt = {'type': 'Sequence', 'pretokenizers': [{'type': 'Split', 'pattern': {'Regex': rx}, 'behavior': 'Isolated'}, {'type': 'ByteLevel'}]}
        assert _load_only(lib, dict(BASE, pre_tokenizer=pt))[0] == want, rx
        assert py_oracle._rust_regex_certainly_rejected(rx) == rejected, rx
    assert _load_only(lib, dict(BASE, decoder={'type': 'ByteLevel'}))[0] == 0
    assert _load_only(lib, dict(BASE, decoder={'type': 'WordPiece'}))[0] == 3
    # invalid data (2): not JSON, no model, vocab id not a u32, added token without `special` (mod.rs:107)
    assert _load_only(lib, b'{"model": ')[0] == 2
    assert _load_only(lib, {'version': '1.0'})[0] == 2
    assert _load_only(lib, {'model': {'vocab': {'a': -1}}})[0] == 2
    assert _load_only(lib, {'model': {'vocab': {'a': 1.5}}})[0] == 2
    assert _load_only(lib, dict(BASE, added_tokens=[{'id': 0, 'content': 'a'}]))[0] == 2
    # a merges table that would make the reference panic (bpe.rs:141) is rejected
    tj = json.loads(json.dumps(BASE)); tj['model']['merges'] = ['x y', 'a b']
    rc, *_, msg = _load_only(lib, tj)
    assert rc == 3 and 'panic' in msg
    # duplicate keys: last wins, like a HashMap insert
    assert _load_only(lib, b'{"model": {"vocab": {"a": 0, "a": 7, "b": 1}, "merges": []}}')[:3] == (0, 0, 2)
    # escapes
    assert _load_only(lib, b'{"model": {"vocab": {"\\u0120a": 0, "\\ud83d\\ude00": 1}, "merges": []}}')[:3] == (0, 0, 2)
    assert _load_only(lib, b'{"model": {"vocab": {"\\ud83d": 0}}}')[0] == 2


def test_added_token_may_match_analysis(built_lib):
    lib = built_lib
    for content, want in (('<s>', 0), ('</s>', 0), ('<|endoftext|>', 0), ('<pad>', 0), ('[MASK]', 0), ('hello', 1), ('Ġhello', 1),
                          (' hello', 0), ("'s", 1), ('...', 1), ('Ġ...', 1), ('a1', 0), ('Ġ', 1), ('日本', 0)):
        tj = dict(BASE, added_tokens=[{'id': 9, 'content': content, 'special': True}])
        rc, _, _, _, mm, _ = _load_only(lib, tj)
        assert rc == 0 and mm == want, content


def _starts_product(lib, docs):
    import synth
    text, offs = synth.pack(docs)
    n = text.size
    out = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
    lib.ctk_debug_starts_host(text.ctypes.data if n else None, n, offs.ctypes.data, len(docs), out.ctypes.data)
    bits = np.unpackbits(out.view(np.uint8), bitorder='little')[:n]
    return set(np.nonzero(bits)[0].tolist())


def _starts_oracle(docs):
    s, base = set(), 0
    for d in docs:
        t = d.decode()
        bidx = [0]
        for ch in t:
            bidx.append(bidx[-1] + len(ch.encode()))
        for a, _ in py_oracle.gpt2_find_iter(t):
            s.add(base + bidx[a])
        base += len(d)
    return s


def test_start_predicate_differential(built_lib):
    """The bounded-window local rule used on the device == leftmost-first regex matching (pretokenizers.rs:13)."""
    random.seed(7)
    alpha = ["a", "b", "s", "t", "r", "e", "v", "l", "m", "d", "'", " ", " ", "\n", "\t", "1", "9", ".", "!", "-", "é", "ü",
             "中", "あ", "　", " ", "\U0001F600", "́", "x", "'s", "'ll", "  ", "Ⅷ", "²", "_", " ", "٣", "ß"]
    bad = 0
    for _ in range(20000):
        docs = [''.join(random.choice(alpha) for _ in range(random.randint(0, 12))).encode() for _ in range(random.randint(1, 3))]
        if _starts_product(built_lib, docs) != _starts_oracle(docs):
            bad += 1
    assert bad == 0


def test_start_predicate_on_corpora(built_lib):
    import synth
    for kind, seed in (('english', 11), ('mixed', 12)):
        text, offs = synth.gen_corpus(kind, seed, 96 << 10, doc_median=700)
        docs = synth.split_docs(text, offs)
        assert _starts_product(built_lib, docs) == _starts_oracle(docs)


def _merge_props(lib, tj):
    data = json.dumps(tj, ensure_ascii=False).encode()
    buf = ctypes.create_string_buffer(data, len(data))
    mono, span = ctypes.c_int(), ctypes.c_uint32()
    rc = lib.ctk_debug_merge_props(ctypes.addressof(buf), len(data), ctypes.byref(mono), ctypes.byref(span))
    assert rc == 0
    return mono.value & 1, span.value


def test_merge_table_monotonicity(built_lib, tok_paths):
    """The round-parallel path for very long pre-tokens (encode_xlong.cuh) is only allowed for monotone tables:
    every pair that contains a merged token ranks after every merge producing that token."""
    lib = built_lib
    vocab = {'a': 0, 'b': 1, 'c': 2, 'ab': 3, 'abc': 4, 'bc': 5}
    mono = {'model': {'type': 'BPE', 'vocab': vocab, 'merges': ['a b', 'ab c']}}
    assert _merge_props(lib, mono) == (1, 3)
    # 'ab c' is listed before the merge that makes 'ab': a pair would outrank its own component
    assert _merge_props(lib, {'model': {'type': 'BPE', 'vocab': vocab, 'merges': ['ab c', 'a b']}})[0] == 0
    # two producers of 'abc', and a pair using 'abc' ranked between them
    v2 = dict(vocab, abca=6)
    assert _merge_props(lib, {'model': {'type': 'BPE', 'vocab': v2, 'merges': ['a b', 'ab c', 'abc a', 'b c', 'a bc']}})[0] == 0
    assert _merge_props(lib, {'model': {'type': 'BPE', 'vocab': v2, 'merges': ['a b', 'ab c', 'b c', 'a bc', 'abc a']}}) == (1, 4)
    # the trained fixtures are monotone, with the longest token well inside the kernel's window limit
    for cfg in ('config1', 'config2', 'config3'):
        with open(tok_paths[cfg], encoding='utf-8') as f:
            m, span = _merge_props(lib, json.load(f))
        assert m == 1 and 2 <= span <= 1024, (cfg, m, span)


def test_round_parallel_rule_matches_sequential_order():
    """tools/verify_window_merge.py: the rule encode_xlong.cuh applies equals one-merge-at-a-time for monotone tables."""
    import importlib.util
    import os
    import random
    spec = importlib.util.spec_from_file_location('vwm', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools', 'verify_window_merge.py'))
    vwm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vwm)
    rng = random.Random(11)
    bad = 0
    for _ in range(150):
        b, _r = vwm.trial(rng, rng.choice([1, 2, 3, 4]), rng.choice([2, 3, 4, 5, 6]), rng.choice([3, 6, 12, 30, 60]), 100, True)
        bad += b
    assert bad == 0


def test_arrow_packing_is_zero_copy_and_handles_slices():
    """Arrow string arrays are already the packed batch (offsets + bytes); slices, chunks, large
    offsets and empty arrays map to (text view, uint64 offsets) without per-string work."""
    import pyarrow as pa
    from complexity_tokenizer import arrow_lists_to_packed, arrow_strings_to_packed
    texts = ['hello', '', 'wörld ✓', 'x' * 70000, 'last']
    for arr in (pa.array(texts), pa.array(texts, type=pa.large_string()), pa.array(['skip me'] + texts)[1:],
                pa.chunked_array([pa.array(texts[:2]), pa.array(texts[2:])])):
        text, offs = arrow_strings_to_packed(arr)
        assert offs.dtype == np.uint64 and offs[0] == 0 and len(offs) == len(texts) + 1
        raw = text.tobytes()
        assert [raw[int(offs[i]):int(offs[i + 1])].decode() for i in range(len(texts))] == texts
    arr = pa.array(texts)
    text, _ = arrow_strings_to_packed(arr)
    assert text.ctypes.data == arr.buffers()[2].address                   # a view of Arrow's data buffer, not a copy
    text, offs = arrow_strings_to_packed(pa.array([], type=pa.string()))
    assert text.size == 0 and offs.tolist() == [0]
    with pytest.raises(ValueError):
        arrow_strings_to_packed(pa.array(['a', None]))
    with pytest.raises(TypeError):
        arrow_strings_to_packed(pa.array([1, 2]))
    ids, offs = arrow_lists_to_packed(pa.array([[1, 2, 3], [], [70000]], type=pa.list_(pa.uint32()))[1:])
    assert ids.tolist() == [70000] and offs.tolist() == [0, 0, 1]


def test_marshal_extension_round_trips(built_lib):
    """csrc/marshal.c: the compiled list <-> packed-buffer conversions of the shim (what PyO3 does for the reference,
    bindings/tokenizer.rs:203-238).  Same buffers as the NumPy conversions, same exceptions as PyO3's extraction."""
    import ctypes
    from complexity_tokenizer import _marshal, _pack_texts
    assert _marshal is not None, '_ctk_marshal is not built'
    texts = ['hello', '', 'wörld ✓ \U0001F600', 'x' * 5000, '\x00nul']
    text, off = _marshal.pack_strs(texts)
    buf, o = _pack_texts(texts)
    assert text == buf.tobytes() and np.frombuffer(off, dtype=np.uint64).tolist() == o.tolist()
    assert _marshal.pack_strs([]) == (b'', (0).to_bytes(8, 'little'))
    assert _marshal.pack_strs(tuple(texts))[0] == text                            # any sequence
    with pytest.raises(TypeError):
        _marshal.pack_strs(['ok', 5])
    with pytest.raises(UnicodeEncodeError):
        _marshal.pack_strs(['lone surrogate \ud800'])
    rows = [[1, 2, 3], [], [4294967295], list(range(1000))]
    ids, off = _marshal.pack_id_lists(rows)
    assert np.frombuffer(ids, dtype=np.uint32).tolist() == [x for r in rows for x in r]
    assert np.frombuffer(off, dtype=np.uint64).tolist() == [0, 3, 3, 4, 1004]
    for bad in ([[1, -1]], [[1 << 32]], [[1.5]], [['a']]):
        with pytest.raises((OverflowError, TypeError)):
            _marshal.pack_id_lists(bad)
    a_ids = ctypes.cast(ctypes.c_char_p(ids), ctypes.c_void_p).value
    a_off = ctypes.cast(ctypes.c_char_p(off), ctypes.c_void_p).value
    assert _marshal.unpack_ids(a_ids, a_off, len(rows)) == rows
    a_txt = ctypes.cast(ctypes.c_char_p(text), ctypes.c_void_p).value
    toff = _marshal.pack_strs(texts)[1]
    assert _marshal.unpack_strs(a_txt, ctypes.cast(ctypes.c_char_p(toff), ctypes.c_void_p).value, len(texts)) == texts


def test_staging_worker_pool(built_lib):
    """host_api.cu: TaskPool, the persistent workers that copy pageable input into page-locked staging buffers.
    Exercised here as a parallel memcpy: many jobs back to back, more parts than threads, from several caller threads."""
    import threading
    lib = built_lib
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, size=(8 << 20) + 123, dtype=np.uint8)
    for threads, parts in ((1, 1), (2, 7), (4, 16), (4, 1), (3, 64)):
        for _ in range(20):
            dst = np.zeros_like(src)
            assert lib.ctk_debug_parallel_copy(dst.ctypes.data, src.ctypes.data, src.size, threads, parts) == 0
            assert np.array_equal(dst, src)
    errs = []

    def hammer():
        try:
            for _ in range(30):
                d = np.zeros(1 << 20, dtype=np.uint8)
                assert lib.ctk_debug_parallel_copy(d.ctypes.data, src.ctypes.data, d.size, 4, 9) == 0
                assert np.array_equal(d, src[:d.size])
        except Exception as e:                                        # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=hammer) for _ in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs[:2]


def test_trainer_has_no_cpu_fallback(built_lib):
    """ctk_train_bpe fails loudly without a device (same rule as the encode path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a device is present')
    import complexity_tokenizer as ct
    with pytest.raises(RuntimeError, match='no CUDA device'):
        ct.BpeTrainer(vocab_size=100, min_frequency=1).train(['hello world'])
    assert ct.BpeTrainer(vocab_size=123, min_frequency=5).vocab_size == 123      # getters of src/bindings/trainers.rs:271-279
    assert ct.BpeTrainer(vocab_size=123, min_frequency=5).min_frequency == 5


def test_trainer_argument_errors(built_lib):
    """Argument checks of ctk_train_bpe come before any device work (include/ctk.h: CTK_ERR_ARG)."""
    lib = built_lib
    out = ctypes.c_void_p()
    assert lib.ctk_train_bpe(None, 0, None, None, 0, ctypes.byref(out)) == 5
    assert b'null' in (lib.ctk_last_error() or b'')
    from complexity_tokenizer import _TrainerConfig
    cfg = _TrainerConfig()
    cfg.vocab_size, cfg.min_frequency, cfg.n_special, cfg.limit_alphabet = 100, 2, 3, -1      # three specials announced, none given
    assert lib.ctk_train_bpe(ctypes.byref(cfg), 0, None, None, 0, ctypes.byref(out)) == 5
    off = (ctypes.c_uint64 * 2)(0, 4)
    cfg.n_special = 0
    assert lib.ctk_train_bpe(ctypes.byref(cfg), 0, None, ctypes.addressof(off), 1, ctypes.byref(out)) == 5   # 4 bytes announced, no text


def test_marshal_takes_numpy_ids_like_the_fallback():
    """decode's list marshalling accepts anything with __index__ (PyO3's Vec<u32> extraction does): ADVICE r1, marshal.c"""
    import numpy as np
    from complexity_tokenizer import _marshal
    assert _marshal is not None, '_ctk_marshal is not built'
    want_ids = np.array([1, 2, 3, 70000, 5], dtype=np.uint32).tobytes()
    want_off = np.array([0, 3, 5], dtype=np.uint64).tobytes()
    for batch in ([[1, 2, 3], [70000, 5]], [np.array([1, 2, 3]), np.array([70000, 5], dtype=np.int64)],
                  [[np.int64(1), np.uint8(2), 3], (np.uint32(70000), 5)]):
        ids, off = _marshal.pack_id_lists(batch)
        assert ids == want_ids and off == want_off
    import pytest
    with pytest.raises(OverflowError):
        _marshal.pack_id_lists([[-1]])
    with pytest.raises(OverflowError):
        _marshal.pack_id_lists([[np.int64(1 << 33)]])
    with pytest.raises(TypeError):
        _marshal.pack_id_lists([[1.5]])


def test_marshal_unpacks_narrow_ids_and_parts():
    """unpack_ids reads 16-bit and 32-bit ids and fills one output list part by part (ctk_result_part)"""
    import numpy as np
    from complexity_tokenizer import _marshal
    ids16 = np.array([7, 65535, 9, 10], dtype=np.uint16)
    ids32 = np.array([7, 65536, 9, 10], dtype=np.uint32)
    off = np.array([0, 1, 1, 4], dtype=np.uint64)
    assert _marshal.unpack_ids(ids16.ctypes.data, off.ctypes.data, 3, 2) == [[7], [], [65535, 9, 10]]
    assert _marshal.unpack_ids(ids32.ctypes.data, off.ctypes.data, 3) == [[7], [], [65536, 9, 10]]
    out = [None] * 5
    _marshal.unpack_ids(ids16.ctypes.data, off.ctypes.data, 3, 2, out, 2)
    off2 = np.array([0, 2, 4], dtype=np.uint64)
    _marshal.unpack_ids(ids32.ctypes.data, off2.ctypes.data, 2, 4, out, 0)
    assert out == [[7, 65536], [9, 10], [7], [], [65535, 9, 10]]
