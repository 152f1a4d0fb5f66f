"""Boundary conventions of SURVEY.md section 8(b) on the GPU: concurrent calls from several host threads on one
tokenizer, two tokenizers side by side, the C ABI's argument errors, device-resident calls (capacity error,
alignment error, asynchronous form), and a hypothesis-driven differential test against the oracle."""
import ctypes
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle(path):
    import c_oracle
    return c_oracle.COracle.from_file(path)


def _tok(path):
    import complexity_tokenizer as ct
    return ct.Tokenizer.from_file(path)


def test_concurrent_calls_from_host_threads(built_lib, tok_paths):
    """`&self` is shared read-only across rayon workers in the reference (mod.rs:694-696); here a tokenizer may be
    called from several host threads at once (calls serialise on its device queue) and two tokenizers are independent."""
    import synth
    tok, tok3, orc, orc3 = _tok(tok_paths['config2']), _tok(tok_paths['config3']), _oracle(tok_paths['config2']), _oracle(tok_paths['config3'])
    jobs = []
    for seed in range(6):
        text, offs = synth.gen_corpus('ascii' if seed % 2 == 0 else 'mixed', 900 + seed, 3 << 20, doc_median=2048)
        t, o = (tok, orc) if seed % 2 == 0 else (tok3, orc3)
        jobs.append((t, text, offs, o.encode_packed(text, offs)))
    errors = []

    def work(k):
        try:
            for rep in range(3):
                t, text, offs, (wids, woff) = jobs[(k + rep) % len(jobs)]
                ids, ioff = t.encode_packed(text, offs)
                assert np.array_equal(ids, wids) and np.array_equal(ioff, woff)
                b, boff = t.decode_packed(ids, ioff, False, False)
                assert boff[-1] > 0
        except Exception as e:                                    # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:2]


def test_c_abi_argument_errors(built_lib, tok_paths):
    """Bad arguments come back as CTK_ERR_ARG with a message, never as a crash; load errors keep the reference's
    kinds (NotFound -> io error, malformed JSON -> invalid data)."""
    import complexity_tokenizer as ct
    lib = ct._lib()
    tok = _tok(tok_paths['config1'])
    res = ctypes.c_void_p()
    off = np.array([0, 3], dtype=np.uint64)
    txt = np.frombuffer(b'abc', dtype=np.uint8)
    assert lib.ctk_encode_batch(None, txt.ctypes.data, off.ctypes.data, 1, ctypes.byref(res)) == ct.CTK_ERR_ARG
    assert lib.ctk_encode_batch(tok._h, None, off.ctypes.data, 1, ctypes.byref(res)) == ct.CTK_ERR_ARG      # bytes promised, no buffer
    bad = np.array([1, 3], dtype=np.uint64)
    assert lib.ctk_encode_batch(tok._h, txt.ctypes.data, bad.ctypes.data, 1, ctypes.byref(res)) == ct.CTK_ERR_ARG
    assert b'text_off[0]' in lib.ctk_last_error()
    dec = np.array([5, 2, 9], dtype=np.uint64)                     # decreasing offsets
    assert lib.ctk_encode_batch(tok._h, txt.ctypes.data, dec.ctypes.data, 2, ctypes.byref(res)) == ct.CTK_ERR_ARG
    ids = np.array([1, 2, 3], dtype=np.uint32)
    assert lib.ctk_decode_batch(tok._h, ids.ctypes.data, dec.ctypes.data, 2, 0, 1, ctypes.byref(res)) == ct.CTK_ERR_ARG
    h = ctypes.c_void_p()
    assert lib.ctk_from_file(b'/nonexistent/tokenizer.json', 0, ctypes.byref(h)) == ct.CTK_ERR_IO
    assert b'os error 2' in lib.ctk_last_error()
    junk = ctypes.create_string_buffer(b'{"model": 5', 11)
    assert lib.ctk_from_json(ctypes.addressof(junk), 11, 0, ctypes.byref(h)) == ct.CTK_ERR_INVALID_DATA
    with pytest.raises(IOError):
        ct.Tokenizer.from_file('/nonexistent/tokenizer.json')
    with pytest.raises(TypeError):
        tok.encode_batch('a bare string is not a batch')


def test_device_resident_calls(built_lib, tok_paths):
    """ctk_encode_batch_device: too small an output buffer and an unaligned text pointer are reported; the asynchronous
    form (n_ids_host = NULL, the caller synchronises its stream) gives the same ids as the synchronous one."""
    import torch
    import complexity_tokenizer as ct
    import synth
    tok, orc = _tok(tok_paths['config2']), _oracle(tok_paths['config2'])
    text, offs = synth.gen_corpus('ascii', 4242, 4 << 20, doc_median=2048)
    B, D = text.size, len(offs) - 1
    wids, woff = orc.encode_packed(text, offs)
    buf = np.zeros(B + 80, dtype=np.uint8)
    buf[:B] = text
    d_text = torch.from_numpy(buf).cuda()
    d_off = torch.from_numpy(offs.astype(np.int64)).cuda()
    cap = B + D + 16
    d_ids = torch.empty(cap, dtype=torch.int32, device='cuda')
    d_ioff = torch.empty(D + 1, dtype=torch.int64, device='cuda')
    n = tok.encode_device(d_text.data_ptr(), d_off.data_ptr(), D, B, d_ids.data_ptr(), cap, d_ioff.data_ptr())
    assert n == wids.size
    assert np.array_equal(d_ids[:n].cpu().numpy().view(np.uint32), wids)
    assert np.array_equal(d_ioff.cpu().numpy().view(np.uint64), woff)
    # asynchronous form on a side stream
    d_ids2 = torch.zeros(cap, dtype=torch.int32, device='cuda')
    d_ioff2 = torch.zeros(D + 1, dtype=torch.int64, device='cuda')
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    lib = ct._lib()
    rc = lib.ctk_encode_batch_device(tok._h, d_text.data_ptr(), d_off.data_ptr(), D, B, d_ids2.data_ptr(), cap, d_ioff2.data_ptr(), None,
                                     ctypes.c_void_p(st.cuda_stream))
    assert rc == 0
    st.synchronize()
    assert torch.equal(d_ioff2, d_ioff) and torch.equal(d_ids2[:n], d_ids[:n])
    # too small an output
    with pytest.raises(Exception) as ei:
        tok.encode_device(d_text.data_ptr(), d_off.data_ptr(), D, B, d_ids.data_ptr(), n // 2, d_ioff.data_ptr())
    assert 'capacity' in str(ei.value)
    # unaligned text
    with pytest.raises(Exception) as ei:
        tok.encode_device(d_text.data_ptr() + 1, d_off.data_ptr(), D, B, d_ids.data_ptr(), cap, d_ioff.data_ptr())
    assert 'aligned' in str(ei.value)
    # and the tokenizer still works afterwards
    assert tok.encode('still alive') == orc.twin.encode('still alive')


def test_hypothesis_differential(built_lib, tok_paths):
    """Property-based differential test: arbitrary Unicode text (any valid str) encodes to the oracle's ids and decodes
    back to what the oracle decodes, for all three fixture tokenizers, as single texts and as one ragged batch."""
    from hypothesis import HealthCheck, given, settings, strategies as st
    toks = {c: (_tok(tok_paths[c]), _oracle(tok_paths[c])) for c in ('config1', 'config2', 'config3')}
    alphabet = st.one_of(st.characters(blacklist_categories=('Cs',)), st.sampled_from(list(" \n\t'.,-!?\"()[]0123456789sdmtrevl")))
    text = st.text(alphabet=alphabet, max_size=300)

    @settings(max_examples=120, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.lists(text, min_size=1, max_size=12), st.sampled_from(['config1', 'config2', 'config3']))
    def check(batch, cfg):
        tok, orc = toks[cfg]
        got, want = tok.encode_batch(batch), orc.encode_batch(batch)
        assert got == want
        assert tok.decode_batch(got) == orc.decode_batch(got)
        assert tok.decode_batch_with_options(got, True, False) == orc.decode_batch(got, True, False)

    check()


def test_persistent_cache_across_calls_of_different_sizes(built_lib, tok_paths):
    """ctk_set_cache_persistent(1) keeps the pre-token cache between calls; a call uses a prefix of the table sized to
    its input, so small and large calls alternate over differently sized prefixes (parts never cleared before must be
    cleared when first used).  Ids stay the oracle's whatever the order; turning persistence off again clears."""
    import synth
    tok, orc = _tok(tok_paths['config2']), _oracle(tok_paths['config2'])
    small = ['a few words only', 'another short text, with punctuation!', '']
    mid_text, mid_offs = synth.gen_corpus('ascii', 77, 300 << 10, doc_median=1024)
    big_text, big_offs = synth.gen_corpus('ascii', 78, 6 << 20, doc_median=2048)
    want_small = orc.encode_batch(small)
    want_mid, want_big = orc.encode_packed(mid_text, mid_offs), orc.encode_packed(big_text, big_offs)
    tok.set_cache_persistent(True)
    for rnd in range(2):
        assert tok.encode_batch(small) == want_small
        ids, off = tok.encode_packed(mid_text, mid_offs)
        assert np.array_equal(ids, want_mid[0]) and np.array_equal(off, want_mid[1])
        ids, off = tok.encode_packed(big_text, big_offs)
        assert np.array_equal(ids, want_big[0]) and np.array_equal(off, want_big[1])
        assert tok.encode_batch(small) == want_small
    tok.set_cache_persistent(False)
    ids, off = tok.encode_packed(mid_text, mid_offs)
    assert np.array_equal(ids, want_mid[0]) and np.array_equal(off, want_mid[1])


def test_arrow_in_arrow_out(built_lib, tok_paths):
    """encode_arrow / decode_arrow: same ids and strings as the list API, for plain, large-offset,
    sliced and chunked Arrow arrays."""
    import pyarrow as pa
    import synth
    tok, orc = _tok(tok_paths['config3']), _oracle(tok_paths['config3'])
    text, offs = synth.gen_corpus('mixed', 4711, 2 << 20, doc_median=700)
    docs = [d.decode() for d in synth.split_docs(text, offs)]
    want = orc.encode_batch(docs)
    for arr in (pa.array(docs), pa.array(docs, type=pa.large_string()), pa.array(['x'] + docs)[1:],
                pa.chunked_array([pa.array(docs[:100]), pa.array(docs[100:])])):
        got = tok.encode_arrow(arr)
        assert pa.types.is_large_list(got.type) and got.type.value_type == pa.uint32()
        assert got.to_pylist() == want
    back = tok.decode_arrow(tok.encode_arrow(pa.array(docs)), False, False)
    import unicodedata
    assert back.to_pylist() == [unicodedata.normalize('NFC', d) for d in docs]
    assert tok.decode_arrow(pa.array(want[:50], type=pa.list_(pa.uint32()))).to_pylist() == orc.decode_batch(want[:50])
