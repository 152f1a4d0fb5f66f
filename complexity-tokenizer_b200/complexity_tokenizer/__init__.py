"""complexity_tokenizer -- drop-in `Tokenizer` for the batched encode/decode hot path, running on a B200.

Mirrors the Python surface of the reference's PyO3 class for this path
(/root/reference/src/bindings/tokenizer.rs:19-23, 203-238, 271-289, 655-663 and the re-export in
python/complexity_tokenizer/__init__.py:16-18): same method names, argument meaning, defaults and
error behaviour (`from_file` raises IOError/OSError; encode/decode never raise for valid input).
All compute happens in libctk.so (CUDA, sm_100a) through the C ABI of include/ctk.h.  There is no
CPU fallback: importing works without a GPU, but constructing a Tokenizer raises if the library or
a device is missing.
"""
import ctypes
import os

import numpy as np

__version__ = '0.3.3+b200'
__all__ = ['Tokenizer', 'Trainer', '__version__']

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CTK_OK, CTK_ERR_IO, CTK_ERR_INVALID_DATA, CTK_ERR_UNSUPPORTED, CTK_ERR_CUDA, CTK_ERR_ARG = range(6)


class UnsupportedTokenizerError(IOError):
    """tokenizer.json uses a pipeline outside the ByteLevel-BPE hot path (ctk code 3)."""


def _lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get('CTK_LIB_VARIANT') or os.path.join(_HERE, 'libctk.so')   # variant: kernel A/B experiments (tools/variants.py)
    if not os.path.exists(path):
        raise ImportError('libctk.so is not built (run `python complexity-tokenizer_b200/build.py`); '
                          'this package has no CPU fallback')
    lib = ctypes.CDLL(path)
    P, S, I = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    U64 = ctypes.c_uint64
    sig = {
        'ctk_from_file': (I, [ctypes.c_char_p, I, ctypes.POINTER(P)]),
        'ctk_from_json': (I, [P, S, I, ctypes.POINTER(P)]),
        'ctk_free': (None, [P]),
        'ctk_vocab_size': (S, [P]),
        'ctk_token_to_id': (I, [P, ctypes.c_char_p, S, ctypes.POINTER(ctypes.c_uint32)]),
        'ctk_id_to_token': (P, [P, ctypes.c_uint32, ctypes.POINTER(S)]),
        'ctk_n_special_tokens': (S, [P]),
        'ctk_special_token': (P, [P, S, ctypes.POINTER(S), ctypes.POINTER(ctypes.c_uint32)]),
        'ctk_device': (I, [P]),
        'ctk_encode_batch': (I, [P, P, P, S, ctypes.POINTER(P)]),
        'ctk_decode_batch': (I, [P, P, P, S, I, I, ctypes.POINTER(P)]),
        'ctk_result_ids': (P, [P]),
        'ctk_result_offsets': (P, [P]),
        'ctk_result_bytes': (P, [P]),
        'ctk_result_count': (S, [P]),
        'ctk_result_free': (None, [P]),
        'ctk_encode_batch_device': (I, [P, P, P, S, U64, P, U64, P, ctypes.POINTER(U64), P]),
        'ctk_decode_batch_device': (I, [P, P, P, S, U64, I, I, P, U64, P, ctypes.POINTER(U64), P]),
        'ctk_decode_max_bytes': (S, [P]),
        'ctk_last_error': (ctypes.c_char_p, []),
        'ctk_kernel_launches': (U64, []),
        'ctk_set_cache_persistent': (None, [P, I]),
        'ctk_profile_enable': (None, [P, I]),
        'ctk_profile_report': (S, [P, ctypes.c_char_p, S]),
        'ctk_debug_use_general': (None, [P, I]),
        'ctk_debug_starts_host': (I, [P, U64, P, S, P]),
        'ctk_debug_starts_window_host': (I, [P, U64, P, S, P]),
        'ctk_debug_load_only': (I, [P, S, ctypes.POINTER(U64), ctypes.POINTER(U64), ctypes.POINTER(I), ctypes.POINTER(I)]),
        'ctk_debug_merge_props': (I, [P, S, ctypes.POINTER(I), ctypes.POINTER(ctypes.c_uint32)]),
        'ctk_debug_xlong_rounds': (I, [P]),
        'ctk_last_transfer_bytes': (None, [P, ctypes.POINTER(U64), ctypes.POINTER(U64)]),
        'ctk_debug_parallel_copy': (I, [P, P, S, I, I]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


try:                                    # compiled list <-> packed-buffer conversions (csrc/marshal.c); marshalling only
    from . import _ctk_marshal as _marshal
except ImportError:                     # not built: the NumPy conversions below do the same, slower
    _marshal = None


def _raise(rc):
    msg = (_lib().ctk_last_error() or b'').decode('utf-8', 'replace')
    if rc == CTK_ERR_IO:
        raise IOError(msg)                              # PyIOError in the reference
    if rc == CTK_ERR_INVALID_DATA:
        raise IOError(msg)                              # io::ErrorKind::InvalidData -> PyIOError
    if rc == CTK_ERR_UNSUPPORTED:
        raise UnsupportedTokenizerError(msg)
    if rc == CTK_ERR_ARG:
        raise ValueError(msg)
    raise RuntimeError(msg)


def _pack_texts(texts):
    blobs = [t.encode('utf-8') for t in texts]          # str.encode raises on lone surrogates, like PyO3's &str
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    if blobs:
        np.cumsum([len(b) for b in blobs], out=off[1:])
    buf = np.frombuffer(b''.join(blobs), dtype=np.uint8) if blobs else np.zeros(0, dtype=np.uint8)
    return buf, off


def arrow_strings_to_packed(strings):
    """(uint8 text view, uint64[n+1] offsets) of a pyarrow string array, zero-copy for the text; nulls are rejected
    (the reference takes list[str]).  Sliced and chunked arrays are handled."""
    import pyarrow as pa
    if isinstance(strings, pa.ChunkedArray):
        strings = strings.combine_chunks() if strings.num_chunks != 1 else strings.chunk(0)
    if not (pa.types.is_string(strings.type) or pa.types.is_large_string(strings.type)):
        raise TypeError('expected an Arrow string array, got %s' % strings.type)
    if strings.null_count:
        raise ValueError('null strings cannot be encoded')
    n = len(strings)
    _, off_buf, data_buf = strings.buffers()
    odt = np.int64 if pa.types.is_large_string(strings.type) else np.int32
    offs = np.frombuffer(off_buf, dtype=odt)[strings.offset:strings.offset + n + 1] if n or off_buf is not None else np.zeros(1, odt)
    if len(offs) == 0:
        offs = np.zeros(1, odt)
    lo, hi = int(offs[0]), int(offs[-1])
    text = np.frombuffer(data_buf, dtype=np.uint8)[lo:hi] if data_buf is not None and hi > lo else np.zeros(0, np.uint8)
    return text, (offs.astype(np.int64) - lo).astype(np.uint64)


def arrow_lists_to_packed(id_lists):
    """(uint32 ids, uint64[n+1] offsets) of a pyarrow (Large)ListArray of integer ids."""
    import pyarrow as pa
    if isinstance(id_lists, pa.ChunkedArray):
        id_lists = id_lists.combine_chunks() if id_lists.num_chunks != 1 else id_lists.chunk(0)
    if not (pa.types.is_list(id_lists.type) or pa.types.is_large_list(id_lists.type)):
        raise TypeError('expected an Arrow list array, got %s' % id_lists.type)
    if id_lists.null_count:
        raise ValueError('null id lists cannot be decoded')
    offs = id_lists.offsets.to_numpy().astype(np.int64)
    lo = int(offs[0]) if len(offs) else 0
    vals = id_lists.values.to_numpy(zero_copy_only=False)
    hi = int(offs[-1]) if len(offs) else 0
    ids = np.ascontiguousarray(vals[lo:hi]).astype(np.uint32, copy=False)
    return ids, (offs - lo).astype(np.uint64) if len(offs) else np.zeros(1, np.uint64)


class Tokenizer:
    """HuggingFace tokenizer.json byte-level BPE tokenizer; encode/decode run on the GPU."""

    def __init__(self, handle):
        self._h = handle

    # ---- constructors
    @staticmethod
    def from_file(path, device=None):
        lib = _lib()
        h = ctypes.c_void_p()
        dev = int(os.environ.get('CTK_DEVICE', '0')) if device is None else int(device)
        rc = lib.ctk_from_file(os.fsencode(path), dev, ctypes.byref(h))
        if rc != CTK_OK:
            _raise(rc)
        return Tokenizer(h)

    @staticmethod
    def from_str(json_text, device=None):
        lib = _lib()
        data = json_text.encode('utf-8') if isinstance(json_text, str) else bytes(json_text)
        buf = ctypes.create_string_buffer(data, len(data))
        h = ctypes.c_void_p()
        dev = int(os.environ.get('CTK_DEVICE', '0')) if device is None else int(device)
        rc = lib.ctk_from_json(ctypes.addressof(buf), len(data), dev, ctypes.byref(h))
        if rc != CTK_OK:
            _raise(rc)
        return Tokenizer(h)

    from_buffer = from_str

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h and _LIB is not None:
            try:
                _LIB.ctk_free(h)
            except Exception:
                pass

    # ---- packed (zero-object) API: numpy in, numpy out
    def encode_packed(self, text, offsets):
        """text: uint8 array (packed UTF-8), offsets: uint64[n+1] -> (ids uint32, ids_off uint64[n+1])"""
        lib = _lib()
        text = np.ascontiguousarray(text, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        res = ctypes.c_void_p()
        rc = lib.ctk_encode_batch(self._h, text.ctypes.data if text.size else None, offsets.ctypes.data, n, ctypes.byref(res))
        if rc != CTK_OK:
            _raise(rc)
        try:
            off = np.ctypeslib.as_array(ctypes.cast(lib.ctk_result_offsets(res), ctypes.POINTER(ctypes.c_uint64)), (n + 1,)).copy()
            tot = int(off[-1])
            ids = np.ctypeslib.as_array(ctypes.cast(lib.ctk_result_ids(res), ctypes.POINTER(ctypes.c_uint32)), (max(tot, 1),))[:tot].copy()
        finally:
            lib.ctk_result_free(res)
        return ids, off

    def decode_packed(self, ids, offsets, skip_special_tokens=False, clean_up_tokenization_spaces=True):
        """ids uint32 (packed), offsets uint64[n+1] -> (bytes uint8, byte_off uint64[n+1])"""
        lib = _lib()
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        res = ctypes.c_void_p()
        rc = lib.ctk_decode_batch(self._h, ids.ctypes.data if ids.size else None, offsets.ctypes.data, n,
                                  int(bool(skip_special_tokens)), int(bool(clean_up_tokenization_spaces)), ctypes.byref(res))
        if rc != CTK_OK:
            _raise(rc)
        try:
            off = np.ctypeslib.as_array(ctypes.cast(lib.ctk_result_offsets(res), ctypes.POINTER(ctypes.c_uint64)), (n + 1,)).copy()
            tot = int(off[-1])
            b = np.ctypeslib.as_array(ctypes.cast(lib.ctk_result_bytes(res), ctypes.POINTER(ctypes.c_uint8)), (max(tot, 1),))[:tot].copy()
        finally:
            lib.ctk_result_free(res)
        return b, off

    # ---- Arrow in / Arrow out (host ingestion without per-string Python objects)
    def encode_arrow(self, strings):
        """pyarrow StringArray / LargeStringArray / ChunkedArray of strings -> LargeListArray<uint32> of ids.
        The Arrow data and offsets buffers ARE the packed batch the C ABI takes; nothing is copied per string."""
        text, offs = arrow_strings_to_packed(strings)
        ids, ioff = self.encode_packed(text, offs)
        import pyarrow as pa
        return pa.LargeListArray.from_arrays(pa.array(ioff.astype(np.int64)), pa.array(ids))

    def decode_arrow(self, id_lists, skip_special_tokens=False, clean_up_tokenization_spaces=True):
        """pyarrow ListArray / LargeListArray of unsigned ids -> LargeStringArray (decode_batch_with_options semantics)."""
        import pyarrow as pa
        ids, offs = arrow_lists_to_packed(id_lists)
        b, boff = self.decode_packed(ids, offs, skip_special_tokens, clean_up_tokenization_spaces)
        return pa.LargeStringArray.from_buffers(len(offs) - 1, pa.py_buffer(boff.astype(np.int64)), pa.py_buffer(b))

    # ---- reference API (bindings/tokenizer.rs:203-238, 655-663)
    def encode(self, text):
        return self.encode_batch([text])[0]

    def encode_batch(self, texts):
        if isinstance(texts, str):
            raise TypeError("argument 'texts': Can't extract `str` to `Vec`")
        if _marshal is not None:
            lib = _lib()
            text, off = _marshal.pack_strs(texts if isinstance(texts, (list, tuple)) else list(texts))
            n = len(off) // 8 - 1
            res = ctypes.c_void_p()
            rc = lib.ctk_encode_batch(self._h, text if text else None, off, n, ctypes.byref(res))
            if rc != CTK_OK:
                _raise(rc)
            try:
                return _marshal.unpack_ids(lib.ctk_result_ids(res) or 0, lib.ctk_result_offsets(res), n)
            finally:
                lib.ctk_result_free(res)
        buf, off = _pack_texts(list(texts))
        ids, ioff = self.encode_packed(buf, off)
        lst = ids.tolist()
        o = ioff.tolist()
        return [lst[o[i]:o[i + 1]] for i in range(len(o) - 1)]

    def decode(self, ids):
        return self.decode_batch_with_options([ids], False, True)[0]

    def decode_with_options(self, ids, skip_special_tokens=False, clean_up_tokenization_spaces=True):
        return self.decode_batch_with_options([ids], skip_special_tokens, clean_up_tokenization_spaces)[0]

    def decode_batch(self, batch):
        return self.decode_batch_with_options(batch, False, True)

    def decode_batch_with_options(self, batch, skip_special_tokens=False, clean_up_tokenization_spaces=True):
        if _marshal is not None:
            lib = _lib()
            ids, off = _marshal.pack_id_lists(batch if isinstance(batch, (list, tuple)) else list(batch))
            n = len(off) // 8 - 1
            res = ctypes.c_void_p()
            rc = lib.ctk_decode_batch(self._h, ids if ids else None, off, n, int(bool(skip_special_tokens)),
                                      int(bool(clean_up_tokenization_spaces)), ctypes.byref(res))
            if rc != CTK_OK:
                _raise(rc)
            try:
                return _marshal.unpack_strs(lib.ctk_result_bytes(res) or 0, lib.ctk_result_offsets(res), n)
            finally:
                lib.ctk_result_free(res)
        batch = [list(b) for b in batch]
        off = np.zeros(len(batch) + 1, dtype=np.uint64)
        if batch:
            np.cumsum([len(b) for b in batch], out=off[1:])
        flat = np.fromiter((i for b in batch for i in b), dtype=np.uint32, count=int(off[-1]))   # OverflowError like PyO3's u32
        b, boff = self.decode_packed(flat, off, skip_special_tokens, clean_up_tokenization_spaces)
        raw = b.tobytes()
        o = boff.tolist()
        return [raw[o[i]:o[i + 1]].decode('utf-8') for i in range(len(o) - 1)]

    def batch_decode(self, sequences, skip_special_tokens=False, clean_up_tokenization_spaces=True):
        return self.decode_batch_with_options(sequences, skip_special_tokens, clean_up_tokenization_spaces)

    # ---- getters (bindings/tokenizer.rs:271-289)
    @property
    def vocab_size(self):
        return int(_lib().ctk_vocab_size(self._h))

    def token_to_id(self, token):
        out = ctypes.c_uint32()
        b = token.encode('utf-8')
        return int(out.value) if _lib().ctk_token_to_id(self._h, b, len(b), ctypes.byref(out)) else None

    def id_to_token(self, id):
        n = ctypes.c_size_t()
        p = _lib().ctk_id_to_token(self._h, int(id), ctypes.byref(n))
        return ctypes.string_at(p, n.value).decode('utf-8') if p else None

    @property
    def special_tokens(self):
        lib = _lib()
        out = {}
        for i in range(lib.ctk_n_special_tokens(self._h)):
            n, tid = ctypes.c_size_t(), ctypes.c_uint32()
            p = lib.ctk_special_token(self._h, i, ctypes.byref(n), ctypes.byref(tid))
            out[ctypes.string_at(p, n.value).decode('utf-8')] = int(tid.value)
        return out

    @property
    def device(self):
        return int(_lib().ctk_device(self._h))

    def profile_enable(self, on=True):
        _lib().ctk_profile_enable(self._h, int(bool(on)))

    def profile_report(self):
        """{kernel name: (total ms, launches)} measured with CUDA events on the launching stream"""
        buf = ctypes.create_string_buffer(1 << 16)
        _lib().ctk_profile_report(self._h, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, cnt = line.split('\t')
            out[name] = (float(ms), int(cnt))
        return out

    # ---- device-resident API (pointers are device pointers on self.device; see include/ctk.h)
    def encode_device(self, d_text, d_text_off, n_docs, total_bytes, d_ids, ids_cap, d_ids_off, stream=0, sync=True):
        lib = _lib()
        tot = ctypes.c_uint64(0)
        rc = lib.ctk_encode_batch_device(self._h, d_text, d_text_off, n_docs, total_bytes, d_ids, ids_cap, d_ids_off,
                                         ctypes.byref(tot) if sync else None, stream or None)
        if rc != CTK_OK:
            _raise(rc)
        return int(tot.value) if sync else None

    def decode_device(self, d_ids, d_ids_off, n_docs, total_ids, d_out, out_cap, d_out_off, skip_special_tokens=False,
                      clean_up_tokenization_spaces=True, stream=0, sync=True):
        lib = _lib()
        tot = ctypes.c_uint64(0)
        rc = lib.ctk_decode_batch_device(self._h, d_ids, d_ids_off, n_docs, total_ids, int(bool(skip_special_tokens)),
                                         int(bool(clean_up_tokenization_spaces)), d_out, out_cap, d_out_off,
                                         ctypes.byref(tot) if sync else None, stream or None)
        if rc != CTK_OK:
            _raise(rc)
        return int(tot.value) if sync else None

    def set_cache_persistent(self, flag):
        _lib().ctk_set_cache_persistent(self._h, int(bool(flag)))


class Trainer:
    """Present for import compatibility only: BPE training is outside the accelerated hot path."""

    def __init__(self, *a, **k):
        raise NotImplementedError('Trainer is outside the B200 encode/decode hot path (SURVEY.md section 2, rows 13-15)')
