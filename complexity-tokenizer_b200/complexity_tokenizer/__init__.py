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
import json
import os

import numpy as np

__version__ = '0.3.3+b200'
__all__ = ['Tokenizer', 'Encoding', 'BatchEncoding', 'Trainer', 'BpeTrainer', 'PanicException', '__version__']

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CTK_OK, CTK_ERR_IO, CTK_ERR_INVALID_DATA, CTK_ERR_UNSUPPORTED, CTK_ERR_CUDA, CTK_ERR_ARG, CTK_ERR_PANIC = range(7)


class PanicException(BaseException):
    """The reference panics on this input; PyO3 surfaces a Rust panic as pyo3_runtime.PanicException (a BaseException)."""


class _ResultOwner:
    """Frees a ctk_encodings result when the last NumPy view of it is gone."""

    def __init__(self, res):
        self._res = res

    def __del__(self):
        res, self._res = self._res, None
        if res and _LIB is not None:
            try:
                _LIB.ctk_encodings_free(res)
            except Exception:
                pass


class _EncodingOptions(ctypes.Structure):       # include/ctk.h: ctk_encoding_options
    _fields_ = [('add_special_tokens', ctypes.c_int), ('pair', ctypes.c_int), ('truncation', ctypes.c_int),
                ('max_length', ctypes.c_uint64), ('padding', ctypes.c_int), ('pad_to', ctypes.c_uint64),
                ('pad_left', ctypes.c_int), ('want_offsets', ctypes.c_int)]


class _TrainerConfig(ctypes.Structure):        # include/ctk.h: ctk_bpe_trainer_config
    _fields_ = [('vocab_size', ctypes.c_uint64), ('min_frequency', ctypes.c_uint32), ('special_tokens', ctypes.c_void_p),
                ('special_off', ctypes.c_void_p), ('n_special', ctypes.c_size_t), ('initial_alphabet', ctypes.c_void_p),
                ('n_alphabet', ctypes.c_size_t), ('limit_alphabet', ctypes.c_int64), ('continuing_subword_prefix', ctypes.c_void_p),
                ('prefix_len', ctypes.c_size_t), ('end_of_word_suffix', ctypes.c_void_p), ('suffix_len', ctypes.c_size_t)]


class _TrainStats(ctypes.Structure):           # include/ctk.h: ctk_train_stats
    _fields_ = [('n_bytes', ctypes.c_uint64), ('n_words', ctypes.c_uint64), ('n_unique_words', ctypes.c_uint64),
                ('n_symbols', ctypes.c_uint64), ('n_merges', ctypes.c_uint64), ('kernel_launches', ctypes.c_uint64),
                ('table_rebuilds', ctypes.c_uint32), ('cluster_size', ctypes.c_uint32), ('stop_reason', ctypes.c_uint32), ('ms_words', ctypes.c_double), ('ms_merges', ctypes.c_double), ('ms_words_kernels', ctypes.c_double)]


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
        'ctk_from_file_devices': (I, [ctypes.c_char_p, I, P, ctypes.POINTER(P)]),
        'ctk_from_json_devices': (I, [P, S, I, P, ctypes.POINTER(P)]),
        'ctk_n_devices': (S, [P]),
        'ctk_device_at': (I, [P, S]),
        'ctk_numa_node': (I, [P, S]),
        'ctk_device_numa_node': (I, [I]),
        'ctk_id_width': (I, [P]),
        'ctk_encode_batch_narrow': (I, [P, P, P, S, ctypes.POINTER(P)]),
        'ctk_result_id_width': (I, [P]),
        'ctk_result_ids_raw': (P, [P]),
        'ctk_result_parts': (S, [P]),
        'ctk_result_part': (I, [P, S, ctypes.POINTER(S), ctypes.POINTER(S), ctypes.POINTER(P), ctypes.POINTER(P)]),
        'ctk_encode_batch_device_ex': (I, [P, P, P, S, U64, P, U64, I, P, ctypes.POINTER(U64), P]),
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
        'ctk_debug_cp_classes': (I, [ctypes.c_uint32, ctypes.c_uint32, P]),
        'ctk_last_transfer_bytes': (None, [P, ctypes.POINTER(U64), ctypes.POINTER(U64)]),
        'ctk_debug_parallel_copy': (I, [P, P, S, I, I]),
        'ctk_encode_batch_to_encoding': (I, [P, P, P, S, ctypes.POINTER(_EncodingOptions), ctypes.POINTER(P)]),
        'ctk_encodings_rows': (S, [P]),
        'ctk_encodings_row_offsets': (P, [P]),
        'ctk_encodings_row_full_lengths': (P, [P]),
        'ctk_encodings_input_ids': (P, [P]),
        'ctk_encodings_attention_mask': (P, [P]),
        'ctk_encodings_type_ids': (P, [P]),
        'ctk_encodings_special_tokens_mask': (P, [P]),
        'ctk_encodings_token_offsets': (P, [P]),
        'ctk_encodings_token_ids': (P, [P]),
        'ctk_encodings_offsets': (P, [P]),
        'ctk_encodings_word_ids': (P, [P]),
        'ctk_encodings_free': (None, [P]),
        'ctk_post_processor_items': (S, [P, ctypes.POINTER(ctypes.c_int64), S]),
        'ctk_pad_token': (ctypes.c_uint32, [P, ctypes.POINTER(P), ctypes.POINTER(S)]),
        'ctk_train_bpe': (I, [ctypes.POINTER(_TrainerConfig), I, P, P, S, ctypes.POINTER(P)]),
        'ctk_train_new_from_texts': (I, [P, ctypes.POINTER(_TrainerConfig), P, P, S, ctypes.POINTER(P)]),
        'ctk_trained_symbols': (S, [P, ctypes.POINTER(P), ctypes.POINTER(P), ctypes.POINTER(P)]),
        'ctk_trained_merges': (S, [P, ctypes.POINTER(P)]),
        'ctk_trained_merge_counts': (P, [P]),
        'ctk_trained_stats': (None, [P, ctypes.POINTER(_TrainStats)]),
        'ctk_trained_free': (None, [P]),
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
    if rc == CTK_ERR_PANIC:
        raise PanicException(msg)
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
        self._model_max_length = 512        # mod.rs:243-245 (from_file)
        self._padding_side = 'right'        # mod.rs:325
        self._truncation_side = 'right'     # mod.rs:326

    # ---- constructors
    @staticmethod
    def _device_list(devices):
        """devices='all' or an iterable of device indices -> ctypes int array (None for all visible devices)"""
        if isinstance(devices, str):
            if devices != 'all':
                raise ValueError("devices must be 'all' or a list of device indices")
            return 0, None
        ids = [int(d) for d in devices]
        if not ids:
            raise ValueError('devices is empty')
        return len(ids), (ctypes.c_int * len(ids))(*ids)

    @staticmethod
    def from_file(path, device=None, devices=None):
        """from_file(path): the reference's constructor (bindings/tokenizer.rs:19-23).  `device` picks the GPU;
        `devices='all'` (or a list) makes ONE tokenizer that spreads every batch over several GPUs, the way the
        reference's encode_batch uses every core of the machine (mod.rs:694-696)."""
        lib = _lib()
        h = ctypes.c_void_p()
        if devices is not None:
            n, arr = Tokenizer._device_list(devices)
            rc = lib.ctk_from_file_devices(os.fsencode(path), n, arr, ctypes.byref(h))
        else:
            dev = int(os.environ.get('CTK_DEVICE', '0')) if device is None else int(device)
            rc = lib.ctk_from_file(os.fsencode(path), dev, ctypes.byref(h))
        if rc != CTK_OK:
            _raise(rc)
        t = Tokenizer(h)
        t._source, t._ctor = ('file', path), dict(device=device, devices=devices)
        return t

    @staticmethod
    def from_str(json_text, device=None, devices=None):
        lib = _lib()
        data = json_text.encode('utf-8') if isinstance(json_text, str) else bytes(json_text)
        buf = ctypes.create_string_buffer(data, len(data))
        h = ctypes.c_void_p()
        if devices is not None:
            n, arr = Tokenizer._device_list(devices)
            rc = lib.ctk_from_json_devices(ctypes.addressof(buf), len(data), n, arr, ctypes.byref(h))
        else:
            dev = int(os.environ.get('CTK_DEVICE', '0')) if device is None else int(device)
            rc = lib.ctk_from_json(ctypes.addressof(buf), len(data), dev, ctypes.byref(h))
        if rc != CTK_OK:
            _raise(rc)
        t = Tokenizer(h)
        t._source, t._ctor = ('str', data), dict(device=device, devices=devices)
        return t

    from_buffer = from_str

    # ---- training with this tokenizer's configuration
    def _config_json(self):
        kind, v = self._source
        if kind == 'file':
            with open(v, 'rb') as f:
                v = f.read()
        return json.loads(v.decode('utf-8'))

    def all_special_tokens(self):
        """mod.rs:1042-1057: bos, eos, pad, unk, sep, cls, mask (vocab.rs:18-30 defaults, overridden by the special added tokens
        in file order, mod.rs:284-303), then the rest of the special-token map -- in hash order in the reference, by content here."""
        roles = {'unk': '<unk>', 'bos': '<s>', 'eos': '</s>', 'pad': '<pad>', 'sep': None, 'cls': None, 'mask': None}
        special = {}
        for t in self._config_json().get('added_tokens') or []:
            if not t.get('special'):
                continue
            c = t['content']
            special[c] = t['id']
            lo = c.lower()
            if 'unk' in lo:
                roles['unk'] = c
            elif lo == '<s>' or 'bos' in lo:
                roles['bos'] = c
            elif lo == '</s>' or 'eos' in lo:
                roles['eos'] = c
            elif 'pad' in lo:
                roles['pad'] = c
            elif 'sep' in lo:
                roles['sep'] = c
            elif 'cls' in lo:
                roles['cls'] = c
            elif 'mask' in lo:
                roles['mask'] = c
        out = [roles[k] for k in ('bos', 'eos', 'pad', 'unk', 'sep', 'cls', 'mask') if roles[k] is not None]
        for c in sorted(special, key=lambda x: x.encode('utf-8')):
            if c not in out:
                out.append(c)
        return out

    def train_new_from_iterator(self, texts, vocab_size):
        """HuggingFaceTokenizer::train_new_from_iterator (mod.rs:1231-1322): same normaliser / pre-tokenizer / decoder /
        post-processor and special tokens, a new vocabulary trained on `texts`.  Normalisation, pre-tokenisation and training
        all run on the device (ctk_train_new_from_texts); returns the new Tokenizer."""
        specials = self.all_special_tokens()
        trainer = BpeTrainer(vocab_size=vocab_size, min_frequency=2, special_tokens=specials, show_progress=True)    # mod.rs:1244-1252
        buf, off = _pack_texts(list(texts))
        vocab, merges = trainer.train_packed(buf, off, tokenizer=self)
        self.last_train_stats = trainer.last_stats
        tj = self._config_json()
        tj.setdefault('model', {})
        tj['model']['vocab'] = vocab
        tj['model']['merges'] = [a + ' ' + b for a, b in merges]
        tj['added_tokens'] = [dict(id=vocab[t], content=t, special=True, single_word=False, lstrip=False, rstrip=False, normalized=False)
                              for t in dict.fromkeys(specials) if t in vocab]                                      # mod.rs:1283-1298
        new = Tokenizer.from_str(json.dumps(tj, ensure_ascii=False), **self._ctor)
        new._model_max_length, new._padding_side, new._truncation_side = self._model_max_length, self._padding_side, self._truncation_side
        return new

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h and _LIB is not None:
            try:
                _LIB.ctk_free(h)
            except Exception:
                pass

    # ---- packed (zero-object) API: numpy in, numpy out
    @staticmethod
    def _result_parts(lib, res):
        """[(first item, n items, data address, offsets address)] of a ctk_result (one per device that got items)"""
        out = []
        for i in range(int(lib.ctk_result_parts(res))):
            first, cnt, data, off = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_void_p(), ctypes.c_void_p()
            rc = lib.ctk_result_part(res, i, ctypes.byref(first), ctypes.byref(cnt), ctypes.byref(data), ctypes.byref(off))
            if rc != CTK_OK:
                _raise(rc)
            out.append((int(first.value), int(cnt.value), data.value or 0, off.value or 0))
        return out

    def encode_packed(self, text, offsets, dtype=np.uint32):
        """text: uint8 array (packed UTF-8), offsets: uint64[n+1] -> (ids, ids_off uint64[n+1]).
        ids are `dtype` (uint32 like the reference's Vec<u32>); dtype=None keeps the element type the device
        produced (uint16 when every id of the vocabulary fits: half the bytes over PCIe and in host memory)."""
        lib = _lib()
        text = np.ascontiguousarray(text, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        res = ctypes.c_void_p()
        rc = lib.ctk_encode_batch_narrow(self._h, text.ctypes.data if text.size else None, offsets.ctypes.data, n, ctypes.byref(res))
        if rc != CTK_OK:
            _raise(rc)
        try:
            width = int(lib.ctk_result_id_width(res))
            src_t = np.uint16 if width == 2 else np.uint32
            src_c = ctypes.c_uint16 if width == 2 else ctypes.c_uint32
            parts = self._result_parts(lib, res)
            off = np.empty(n + 1, dtype=np.uint64)
            totals = []
            base = 0
            for first, cnt, data, poff in parts:
                po = np.ctypeslib.as_array(ctypes.cast(poff, ctypes.POINTER(ctypes.c_uint64)), (cnt + 1,))
                off[first:first + cnt] = po[:cnt] + np.uint64(base)
                totals.append(int(po[cnt]))
                base += totals[-1]
            off[n] = base
            ids = np.empty(base, dtype=src_t if dtype is None else dtype)
            base = 0
            for (first, cnt, data, poff), tot in zip(parts, totals):
                if tot:
                    ids[base:base + tot] = np.ctypeslib.as_array(ctypes.cast(data, ctypes.POINTER(src_c)), (tot,))
                base += tot
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
            rc = lib.ctk_encode_batch_narrow(self._h, text if text else None, off, n, ctypes.byref(res))
            if rc != CTK_OK:
                _raise(rc)
            try:                        # one pass: device-width ids of every part -> Python ints (what PyO3 does with Vec<Vec<u32>>)
                width = int(lib.ctk_result_id_width(res))
                out = [None] * n
                for first, cnt, data, poff in self._result_parts(lib, res):
                    _marshal.unpack_ids(data, poff, cnt, width, out, first)
                return out
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

    # ---- rich Encoding outputs (bindings/tokenizer.rs:33-201, :259-368), computed in csrc/encoding.cu
    def _encode_rows(self, text_buf, text_off, pair=False, add_special_tokens=True, truncation=False, max_length=0, padding=0,
                     pad_to=0, pad_left=False, want_offsets=False, overflow_mode='single', stride=0):
        """Packed texts -> PackedEncodings (NumPy copies of every array the device produced)."""
        from .encoding import PackedEncodings
        lib = _lib()
        text_buf = np.ascontiguousarray(text_buf, dtype=np.uint8)
        text_off = np.ascontiguousarray(text_off, dtype=np.uint64)
        n = len(text_off) - 1
        opt = _EncodingOptions(int(bool(add_special_tokens)), int(bool(pair)), int(bool(truncation)), int(max_length), int(padding),
                               int(pad_to), int(bool(pad_left)), int(bool(want_offsets)))
        res = ctypes.c_void_p()
        rc = lib.ctk_encode_batch_to_encoding(self._h, text_buf.ctypes.data if text_buf.size else None, text_off.ctypes.data, n,
                                              ctypes.byref(opt), ctypes.byref(res))
        if rc != CTK_OK:
            _raise(rc)
        owner = _ResultOwner(res)                   # the arrays below are VIEWS of the library's page-locked result block

        def arr(fn, ctype, count, shape=None):
            p = fn(res)
            if not p or count == 0:
                return np.zeros(shape or (0,), dtype=ctype)
            raw = (ctypes.c_uint8 * (count * np.dtype(ctype).itemsize)).from_address(p)
            raw._owner = owner                      # keeps the result alive as long as any view of it
            a = np.frombuffer(raw, dtype=ctype)
            return a.reshape(shape) if shape else a
        rows = int(lib.ctk_encodings_rows(res))
        row_off = arr(lib.ctk_encodings_row_offsets, np.uint64, rows + 1)
        tok_off = arr(lib.ctk_encodings_token_offsets, np.uint64, n + 1)
        total, n_tok = int(row_off[-1]), int(tok_off[-1])
        ptok, plen = ctypes.c_void_p(), ctypes.c_size_t()
        pad_id = lib.ctk_pad_token(self._h, ctypes.byref(ptok), ctypes.byref(plen))
        offs = arr(lib.ctk_encodings_offsets, np.uint32, 2 * n_tok, (n_tok, 2)) if want_offsets else None
        wids = arr(lib.ctk_encodings_word_ids, np.uint32, n_tok) if want_offsets else None
        return PackedEncodings(
            self, pair=bool(pair), add_special_tokens=bool(add_special_tokens), truncation=bool(truncation),
            max_length=int(max_length), pad_left=bool(pad_left), overflow_mode=overflow_mode, stride=int(stride),
            pad_id=int(pad_id), pad_token=ctypes.string_at(ptok, plen.value).decode('utf-8'),
            text_buf=text_buf, text_off=text_off, row_off=row_off, tok_off=tok_off,
            row_full=arr(lib.ctk_encodings_row_full_lengths, np.uint64, rows),
            input_ids=arr(lib.ctk_encodings_input_ids, np.uint32, total),
            attention_mask=arr(lib.ctk_encodings_attention_mask, np.uint8, total),
            token_type_ids=arr(lib.ctk_encodings_type_ids, np.uint8, total),
            special_tokens_mask=arr(lib.ctk_encodings_special_tokens_mask, np.uint8, total),
            raw_ids=arr(lib.ctk_encodings_token_ids, np.uint32, n_tok), offsets=offs, word_ids=wids)

    @staticmethod
    def _interleave(texts, pairs):
        flat = []
        for a, b in zip(texts, pairs):
            flat.append(a)
            flat.append(b)
        return flat

    def __call__(self, text, text_pair=None, add_special_tokens=True, padding=None, truncation=False, max_length=None, stride=0,
                 return_attention_mask=True, return_token_type_ids=True, return_offsets_mapping=False,
                 return_special_tokens_mask=False):
        """bindings/tokenizer.rs:33-201"""
        from .encoding import BatchEncoding
        if isinstance(text, (list, tuple)) and all(isinstance(t, str) for t in text):
            texts = list(text)
            pairs = list(text_pair) if isinstance(text_pair, (list, tuple)) and all(isinstance(t, str) for t in text_pair) else None
        elif isinstance(text, str):
            texts = [text]
            pairs = [text_pair] if isinstance(text_pair, str) else None
        else:
            raise TypeError('Expected str or List[str]')
        flat = self._interleave(texts, pairs) if pairs is not None else texts
        max_len = self._model_max_length if max_length is None else int(max_length)
        if max_len < 0 or stride < 0:
            raise OverflowError("can't convert negative int to unsigned")
        pad_mode = 0 if padding is None else 2 if padding == 'max_length' else 1
        pad_left = padding == 'left' or self._padding_side == 'left'
        buf, off = _pack_texts(flat)
        if truncation and stride > 0 and stride >= max_len:
            # the reference only loops forever for a row that IS longer than max_length (encoding.rs:190-193); shorter rows pass
            probe = self._encode_rows(buf, off, pair=pairs is not None, add_special_tokens=add_special_tokens)
            if bool((probe.row_full > max_len).any()):
                raise ValueError('stride >= max_length: the reference loops forever (encoding.rs:190-193)')
            truncation = False
        p = self._encode_rows(buf, off, pair=pairs is not None, add_special_tokens=add_special_tokens, truncation=truncation,
                              max_length=max_len, padding=pad_mode, pad_to=max_len, pad_left=pad_left,
                              overflow_mode='stride' if stride > 0 else 'single', stride=stride)
        return BatchEncoding(p, return_attention_mask, return_token_type_ids, return_offsets_mapping, return_special_tokens_mask)

    def encode_to_encoding(self, text):
        return self.encode_batch_to_encoding([text])[0]

    def encode_pair_to_encoding(self, text, text_pair):
        return self.encode_batch_pairs_to_encoding([(text, text_pair)])[0]

    def encode_with_truncation(self, text, text_pair=None, max_length=512, stride=0):
        """mod.rs:349-357: truncate_with_stride whenever the row is longer than max_length (also for stride = 0)"""
        flat = [text] if text_pair is None else [text, text_pair]
        buf, off = _pack_texts(flat)
        if stride >= max_length:
            # the reference loops forever only when the row is longer than max_length (encoding.rs:190-193); otherwise nothing is cut
            p = self._encode_rows(buf, off, pair=text_pair is not None)
            if int(p.row_full[0]) > max_length:
                raise ValueError('stride >= max_length: the reference loops forever (encoding.rs:190-193)')
            return p.encoding(0)
        p = self._encode_rows(buf, off, pair=text_pair is not None, truncation=True, max_length=max_length,
                              overflow_mode='stride', stride=stride)
        return p.encoding(0)

    def encode_batch_to_encoding(self, texts):
        if isinstance(texts, str):
            raise TypeError("argument 'texts': Can't extract `str` to `Vec`")
        buf, off = _pack_texts(list(texts))
        p = self._encode_rows(buf, off)
        return [p.encoding(r) for r in range(p.n_rows)]

    def encode_batch_pairs_to_encoding(self, pairs):
        flat = self._interleave([a for a, _ in pairs], [b for _, b in pairs])
        buf, off = _pack_texts(flat)
        p = self._encode_rows(buf, off, pair=True)
        return [p.encoding(r) for r in range(p.n_rows)]

    def encode_batch_with_padding(self, texts, max_length=None, pad_left=False):
        """mod.rs:490-516: pad to max_length, or to the longest row; rows longer than that stay as they are"""
        buf, off = _pack_texts(list(texts))
        p = self._encode_rows(buf, off, padding=1 if max_length is None else 2, pad_to=max_length or 0, pad_left=pad_left)
        return [p.encoding(r) for r in range(p.n_rows)]

    def encode_batch_pairs_with_padding(self, pairs, max_length=None, pad_left=False):
        flat = self._interleave([a for a, _ in pairs], [b for _, b in pairs])
        buf, off = _pack_texts(flat)
        p = self._encode_rows(buf, off, pair=True, padding=1 if max_length is None else 2, pad_to=max_length or 0, pad_left=pad_left)
        return [p.encoding(r) for r in range(p.n_rows)]

    def encode_plus(self, text):
        return self.encode_to_encoding(text)

    def batch_encode_plus(self, texts):
        return self.encode_batch_to_encoding(texts)

    @property
    def model_max_length(self):
        return self._model_max_length

    @model_max_length.setter
    def model_max_length(self, value):
        self._model_max_length = int(value)

    @property
    def padding_side(self):
        return self._padding_side

    @padding_side.setter
    def padding_side(self, value):
        self._padding_side = str(value)

    @property
    def truncation_side(self):
        return self._truncation_side

    @truncation_side.setter
    def truncation_side(self, value):
        self._truncation_side = str(value)

    @property
    def post_processor_items(self):
        """What the loaded post-processor does to one sequence: list of items, -1 = the ids, else a literal id."""
        buf = (ctypes.c_int64 * 64)()
        k = _lib().ctk_post_processor_items(self._h, buf, 64)
        return [int(buf[i]) for i in range(min(k, 64))]

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

    @property
    def devices(self):
        """devices this tokenizer spreads its batches over (one entry unless built with devices=...)"""
        lib = _lib()
        return [int(lib.ctk_device_at(self._h, i)) for i in range(int(lib.ctk_n_devices(self._h)))]

    @property
    def id_width(self):
        """bytes per id the device produces: 2 when every id the tokenizer can emit is below 65 536, else 4"""
        return int(_lib().ctk_id_width(self._h))

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
    def encode_device(self, d_text, d_text_off, n_docs, total_bytes, d_ids, ids_cap, d_ids_off, stream=0, sync=True, id_width=4):
        """device pointers in, device pointers out; id_width = element size of d_ids (2 only if self.id_width == 2)"""
        lib = _lib()
        tot = ctypes.c_uint64(0)
        rc = lib.ctk_encode_batch_device_ex(self._h, d_text, d_text_off, n_docs, total_bytes, d_ids, ids_cap, int(id_width), d_ids_off,
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


from .encoding import BatchEncoding, Encoding  # noqa: E402


class Trainer:
    """Present for import compatibility only: BPE training is outside the accelerated hot path."""

    def __init__(self, *a, **k):
        raise NotImplementedError('Trainer is outside the B200 encode/decode hot path (SURVEY.md section 2, rows 13-15)')


class BpeTrainer:
    """BPE training on the GPU; mirrors the reference's PyO3 `BpeTrainer` (src/bindings/trainers.rs:218-280: same
    constructor arguments and defaults, `train(texts) -> (vocab, merges)`, getters) over `ctk_train_bpe`.
    `initial_alphabet` / `limit_alphabet` are the Rust-API knobs of BpeTrainerConfig (bpe_trainer.rs:24-26).
    Where the reference decides in hash-iteration order the choice is fixed (include/ctk.h)."""

    def __init__(self, vocab_size=30000, min_frequency=2, special_tokens=None, show_progress=True, end_of_word_suffix=None,
                 continuing_subword_prefix=None, initial_alphabet=None, limit_alphabet=None, device=None):
        self._vocab_size = int(vocab_size)
        self._min_frequency = int(min_frequency)
        self._special = list(special_tokens) if special_tokens is not None else ['<unk>', '<pad>', '<s>', '</s>']
        self._show_progress = bool(show_progress)      # accepted; nothing is printed
        self._suffix = end_of_word_suffix
        self._prefix = continuing_subword_prefix
        self._alphabet = None if initial_alphabet is None else [ord(c) for c in initial_alphabet]
        self._limit = limit_alphabet
        self._device = int(os.environ.get('LOCAL_RANK', 0)) if device is None else int(device)
        self.last_stats = None
        self.last_merge_counts = None

    @property
    def vocab_size(self):
        return self._vocab_size

    @property
    def min_frequency(self):
        return self._min_frequency

    def train(self, texts):
        """-> (dict token -> id, list of (left, right)); bpe_trainer.rs:100."""
        buf, off = _pack_texts(texts)
        return self.train_packed(buf, off)

    def train_packed(self, buf, off, tokenizer=None):
        """tokenizer: run the texts through that tokenizer's normaliser and pre-tokenizer on the device first
        (train_new_from_iterator, mod.rs:1257-1270)"""
        lib = _lib()
        sp, sp_off = _pack_texts(self._special)
        keep = [np.ascontiguousarray(buf, dtype=np.uint8), np.ascontiguousarray(off, dtype=np.uint64), sp, sp_off]
        cfg = _TrainerConfig()
        cfg.vocab_size, cfg.min_frequency = self._vocab_size, self._min_frequency
        cfg.special_tokens, cfg.special_off, cfg.n_special = sp.ctypes.data, sp_off.ctypes.data, len(self._special)
        if self._alphabet is not None:
            al = np.asarray(self._alphabet, dtype=np.uint32)
            keep.append(al)
            cfg.initial_alphabet, cfg.n_alphabet = (al.ctypes.data if len(al) else sp_off.ctypes.data), len(al)
        cfg.limit_alphabet = -1 if self._limit is None else int(self._limit)
        for name, ln, val in (('continuing_subword_prefix', 'prefix_len', self._prefix), ('end_of_word_suffix', 'suffix_len', self._suffix)):
            if val is not None:
                b = np.frombuffer(val.encode('utf-8') + b'\0', dtype=np.uint8)
                keep.append(b)
                setattr(cfg, name, b.ctypes.data)
                setattr(cfg, ln, len(b) - 1)
        h = ctypes.c_void_p()
        if tokenizer is not None:
            rc = lib.ctk_train_new_from_texts(tokenizer._h, ctypes.byref(cfg), keep[0].ctypes.data, keep[1].ctypes.data, len(keep[1]) - 1, ctypes.byref(h))
        else:
            rc = lib.ctk_train_bpe(ctypes.byref(cfg), self._device, keep[0].ctypes.data, keep[1].ctypes.data, len(keep[1]) - 1, ctypes.byref(h))
        if rc != CTK_OK:
            _raise(rc)
        try:
            pb, po, pv, pm = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
            ns = lib.ctk_trained_symbols(h, ctypes.byref(pb), ctypes.byref(po), ctypes.byref(pv))
            so = np.ctypeslib.as_array(ctypes.cast(po, ctypes.POINTER(ctypes.c_uint64)), (ns + 1,)).copy() if ns else np.zeros(1, np.uint64)
            sb = ctypes.string_at(pb, int(so[-1])) if ns else b''
            vid = np.ctypeslib.as_array(ctypes.cast(pv, ctypes.POINTER(ctypes.c_int64)), (ns,)).copy() if ns else np.zeros(0, np.int64)
            syms = [sb[int(so[i]):int(so[i + 1])].decode('utf-8') for i in range(ns)]
            nm = lib.ctk_trained_merges(h, ctypes.byref(pm))
            mp = np.ctypeslib.as_array(ctypes.cast(pm, ctypes.POINTER(ctypes.c_uint32)), (2 * nm,)).copy() if nm else np.zeros(0, np.uint32)
            pc = lib.ctk_trained_merge_counts(h)
            self.last_merge_counts = np.ctypeslib.as_array(ctypes.cast(pc, ctypes.POINTER(ctypes.c_uint32)), (nm,)).copy() if nm else np.zeros(0, np.uint32)
            st = _TrainStats()
            lib.ctk_trained_stats(h, ctypes.byref(st))
            self.last_stats = {k: getattr(st, k) for k, _ in _TrainStats._fields_}
        finally:
            lib.ctk_trained_free(h)
        vocab = {syms[i]: int(vid[i]) for i in range(ns) if vid[i] >= 0}
        merges = [(syms[int(mp[2 * i])], syms[int(mp[2 * i + 1])]) for i in range(nm)]
        return vocab, merges
