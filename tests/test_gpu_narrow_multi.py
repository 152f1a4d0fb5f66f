"""Narrow (uint16) ids and the one-process / several-GPU tokenizer, through the C ABI, against the oracle.

The reference's encode_batch is one call that uses the whole machine (mod.rs:694-696): a tokenizer built with
devices=[...] must give exactly what the one-device tokenizer and the oracle give.  Needs a GPU: run with -m gpu
(the several-device tests skip on a box with one GPU)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _oracle(path):
    import c_oracle
    return c_oracle.COracle.from_file(path)


@pytest.mark.parametrize('cfg,kind,width', [('config2', 'ascii', 2), ('config1', 'english', 2), ('config3', 'mixed', 4)])
def test_narrow_ids_equal_wide_ids(built_lib, tok_paths, cfg, kind, width):
    """vocab <= 65 536 -> the device writes uint16 ids (config 1, 2); config 3 (100 000 ids) stays at uint32"""
    import complexity_tokenizer as ct
    import synth
    tok, orc = ct.Tokenizer.from_file(tok_paths[cfg]), _oracle(tok_paths[cfg])
    assert tok.id_width == width
    text, offs = synth.gen_corpus(kind, 4242, 6 << 20, doc_median=2048)
    native, noff = tok.encode_packed(text, offs, dtype=None)
    assert native.dtype == (np.uint16 if width == 2 else np.uint32)
    wide, woff = tok.encode_packed(text, offs)
    want, want_off = orc.encode_packed(text, offs)
    assert wide.dtype == np.uint32
    assert np.array_equal(noff, want_off) and np.array_equal(woff, want_off)
    assert np.array_equal(native.astype(np.uint32), want) and np.array_equal(wide, want)
    # the uint32 contract of ctk_encode_batch itself (no narrow entry point involved)
    lib = built_lib
    res = ctypes.c_void_p()
    t = np.ascontiguousarray(text)
    rc = lib.ctk_encode_batch(tok._h, t.ctypes.data, offs.ctypes.data, len(offs) - 1, ctypes.byref(res))
    assert rc == 0
    try:
        assert lib.ctk_result_id_width(res) == 4 and lib.ctk_result_parts(res) == 1
        got = np.ctypeslib.as_array(ctypes.cast(lib.ctk_result_ids(res), ctypes.POINTER(ctypes.c_uint32)), (want.size,))
        assert np.array_equal(got, want)
    finally:
        lib.ctk_result_free(res)
    # a narrow result read through the uint32 accessor is widened on demand
    rc = lib.ctk_encode_batch_narrow(tok._h, t.ctypes.data, offs.ctypes.data, len(offs) - 1, ctypes.byref(res))
    assert rc == 0
    try:
        assert lib.ctk_result_id_width(res) == width
        got = np.ctypeslib.as_array(ctypes.cast(lib.ctk_result_ids(res), ctypes.POINTER(ctypes.c_uint32)), (want.size,))
        assert np.array_equal(got, want)
    finally:
        lib.ctk_result_free(res)


def test_device_resident_narrow_output(built_lib, tok_paths):
    """ctk_encode_batch_device_ex with id_width = 2: same ids, half the bytes; id_width 2 is refused for a 100 000-id vocabulary"""
    import torch
    import complexity_tokenizer as ct
    import synth
    tok = ct.Tokenizer.from_file(tok_paths['config2'])
    text, offs = synth.gen_corpus('ascii', 99, 3 << 20, doc_median=4096)
    B, D = int(text.size), len(offs) - 1
    dev = torch.device('cuda', 0)
    d_text = torch.zeros(B + 64, dtype=torch.uint8, device=dev)
    d_text[:B] = torch.from_numpy(text.copy()).to(dev)
    d_off = torch.from_numpy(offs.astype(np.int64)).to(dev)
    cap = B + D + 16
    d16 = torch.zeros(cap, dtype=torch.int16, device=dev)
    d32 = torch.zeros(cap, dtype=torch.int32, device=dev)
    o16 = torch.zeros(D + 1, dtype=torch.int64, device=dev)
    o32 = torch.zeros(D + 1, dtype=torch.int64, device=dev)
    n16 = tok.encode_device(d_text.data_ptr(), d_off.data_ptr(), D, B, d16.data_ptr(), cap, o16.data_ptr(), id_width=2)
    n32 = tok.encode_device(d_text.data_ptr(), d_off.data_ptr(), D, B, d32.data_ptr(), cap, o32.data_ptr(), id_width=4)
    assert n16 == n32 and torch.equal(o16, o32)
    a = d16[:n16].cpu().numpy().view(np.uint16).astype(np.uint32)
    b = d32[:n32].cpu().numpy().view(np.uint32)
    assert np.array_equal(a, b)
    want, _ = _oracle(tok_paths['config2']).encode_packed(text, offs)
    assert np.array_equal(b, want)
    tok3 = ct.Tokenizer.from_file(tok_paths['config3'])
    with pytest.raises(ValueError):
        tok3.encode_device(d_text.data_ptr(), d_off.data_ptr(), D, B, d16.data_ptr(), cap, o16.data_ptr(), id_width=2)


def test_nfc_growth_does_not_overflow_the_id_buffer(built_lib, tok_paths):
    """text that NFC expands 3x with a near byte-level tokenizer needs more ids than input bytes / 1"""
    import json
    import complexity_tokenizer as ct
    import synth
    # byte-level vocabulary without merges: one id per byte of the NORMALISED text
    tj = synth.assemble_tokenizer([], specials_first=())
    tok = ct.Tokenizer.from_str(json.dumps(tj, ensure_ascii=False))
    import py_oracle
    twin = py_oracle.OracleTokenizer.from_str(json.dumps(tj, ensure_ascii=False))
    docs = ['\u0958' * 3000, '\uFB2C' * 2000 + ' x', '\U0001D160' * 1500]     # NFC: 3 -> 6, 3 -> 6, 4 -> 12 bytes per character
    got = tok.encode_batch(docs * 400)                                            # ~10 MB: the chunked host pipeline
    import unicodedata
    for d, g in zip(docs, got[:3]):
        assert len(g) == len(unicodedata.normalize('NFC', d).encode()), 'one id per normalised byte'
    assert got[:3] == twin.encode_batch(docs)
    assert got[3:6] == got[:3]


@pytest.mark.skipif('_n_gpus() < 2')
@pytest.mark.parametrize('cfg,kind', [('config2', 'ascii'), ('config3', 'mixed')])
def test_two_device_handle_equals_one_device_and_oracle(built_lib, tok_paths, cfg, kind, monkeypatch):
    import complexity_tokenizer as ct
    import synth
    monkeypatch.setenv('CTK_MULTI_MIN_MB', '1')
    one = ct.Tokenizer.from_file(tok_paths[cfg], device=0)
    two = ct.Tokenizer.from_file(tok_paths[cfg], devices=[0, 1])
    assert two.devices == [0, 1] and one.devices == [0]
    orc = _oracle(tok_paths[cfg])
    text, offs = synth.gen_corpus(kind, 777, 24 << 20, doc_median=4096)
    want, want_off = orc.encode_packed(text, offs)
    for t in (one, two):
        ids, ioff = t.encode_packed(text, offs)
        assert np.array_equal(ioff, want_off) and np.array_equal(ids, want)
    # parts: two of them, contiguous, byte-balanced
    lib = built_lib
    res = ctypes.c_void_p()
    tt = np.ascontiguousarray(text)
    assert lib.ctk_encode_batch_narrow(two._h, tt.ctypes.data, offs.ctypes.data, len(offs) - 1, ctypes.byref(res)) == 0
    try:
        parts = ct.Tokenizer._result_parts(lib, res)
        assert len(parts) == 2 and parts[0][0] == 0 and parts[1][0] == parts[0][1] and parts[0][1] + parts[1][1] == len(offs) - 1
        split = int(offs[parts[1][0]])
        assert abs(split - text.size / 2) < (1 << 20), 'document ranges balanced by bytes'
        flat = np.ctypeslib.as_array(ctypes.cast(lib.ctk_result_offsets(res), ctypes.POINTER(ctypes.c_uint64)), (len(offs),))
        assert np.array_equal(flat, want_off)
    finally:
        lib.ctk_result_free(res)
    # list API and decode through the same handle (decode shards the same way)
    docs = [bytes(text[int(offs[i]):int(offs[i + 1])]).decode() for i in range(0, 300)]
    assert two.encode_batch(docs) == one.encode_batch(docs)
    for skip, clean in ((False, True), (False, False)):
        b2, o2 = two.decode_packed(want, want_off, skip, clean)
        b1, o1 = one.decode_packed(want, want_off, skip, clean)
        assert np.array_equal(o1, o2) and np.array_equal(b1, b2)
    raw, roff = two.decode_packed(want, want_off, False, False)
    if cfg == 'config2':
        assert raw.tobytes() == text.tobytes() and np.array_equal(roff, offs)


@pytest.mark.skipif('_n_gpus() < 2')
def test_all_devices_handle_small_and_empty_batches(built_lib, tok_paths):
    import complexity_tokenizer as ct
    tok = ct.Tokenizer.from_file(tok_paths['config1'], devices='all')
    one = ct.Tokenizer.from_file(tok_paths['config1'])
    assert len(tok.devices) == _n_gpus()
    texts = ['', 'Hello, world!', "don't", ' ' * 40, 'x' * 5000]
    assert tok.encode_batch(texts) == one.encode_batch(texts)
    assert tok.encode_batch([]) == []
    assert tok.decode_batch(one.encode_batch(texts)) == one.decode_batch(one.encode_batch(texts))
