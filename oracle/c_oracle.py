"""ORACLE (test infrastructure, not product code) -- ctypes front end of oracle_core.c.

Loads tokenizer.json through the Python twin (oracle/py_oracle.py, which restates the reference's
load rules), flattens the tables and hands them to the multi-threaded C core.  Used by tests as
the checker at MB sizes and by bench.py as the all-core CPU baseline ("port").
"""
import ctypes
import os
import subprocess

import numpy as np

from py_oracle import OracleTokenizer, BYTE_ENCODER   # noqa: F401  (same directory)

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_lib(force=False):
    """gcc the C restatement into oracle/_build/liboracle.so (idempotent)."""
    global _LIB
    so = os.path.join(HERE, '_build', 'liboracle.so')
    srcs = [os.path.join(HERE, 'oracle_core.c'), os.path.join(HERE, 'unicode_ranges_gen.h')]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(['gcc', '-O3', '-std=c11', '-shared', '-fPIC', '-pthread', srcs[0], '-o', so])   # the reference ships opt-level 3
        _LIB = None
    if _LIB is None:
        lib = ctypes.CDLL(so)
        P = ctypes.c_void_p
        lib.orc_new.restype = P
        lib.orc_new.argtypes = [P, P, P, ctypes.c_size_t, P, ctypes.c_size_t, P, P, P, P, P, ctypes.c_size_t,
                                P, P, P, P, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
        lib.orc_free.argtypes = [P]
        lib.orc_encode_batch.restype = P
        lib.orc_encode_batch.argtypes = [P, P, P, ctypes.c_size_t, ctypes.c_int]
        lib.orc_decode_batch.restype = P
        lib.orc_decode_batch.argtypes = [P, P, P, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.orc_result_ids.restype = P
        lib.orc_result_ids.argtypes = [P]
        lib.orc_result_bytes.restype = P
        lib.orc_result_bytes.argtypes = [P]
        lib.orc_result_off.restype = P
        lib.orc_result_off.argtypes = [P]
        lib.orc_result_free.argtypes = [P]
        lib.orc_nfc.restype = P
        lib.orc_nfc.argtypes = [P, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        lib.orc_free_buf.argtypes = [P]
        _LIB = lib
    return _LIB


def _pack_strings(strs):
    blobs = [s.encode('utf-8') for s in strs]
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    if blobs:
        off[1:] = np.cumsum([len(b) for b in blobs], dtype=np.uint64)
    blob = np.frombuffer(b''.join(blobs) + b'\0', dtype=np.uint8).copy()
    return blob, off


def nfc(data: bytes) -> bytes:
    lib = build_lib()
    n = ctypes.c_size_t(0)
    buf = (ctypes.c_uint8 * max(1, len(data))).from_buffer_copy(data or b'\0')
    p = lib.orc_nfc(ctypes.addressof(buf), len(data), ctypes.byref(n))
    out = ctypes.string_at(p, n.value)
    lib.orc_free_buf(p)
    return out


class COracle:
    def __init__(self, twin: OracleTokenizer):
        self.twin = twin
        self.lib = build_lib()
        if len([s for s in twin.pre_stages if s[0] == 'bytelevel']) != 1 or twin.decoder != 'bytelevel':
            raise ValueError('C core handles exactly one ByteLevel stage and the ByteLevel decoder (Metaspace: use the Python twin)')
        aps = [s for s in twin.pre_stages if s[0] == 'bytelevel'][0][1]
        pa = np.array([k[0] for k in twin.merge_ranks], dtype=np.uint32)
        pb = np.array([k[1] for k in twin.merge_ranks], dtype=np.uint32)
        pr = np.array(list(twin.merge_ranks.values()), dtype=np.uint32)
        ops = np.array(twin.merge_ops, dtype=np.uint32)
        char_id = np.full(0x180, -1, dtype=np.int64)
        for tok, tid in twin.vocab.items():
            if len(tok) == 1 and ord(tok) < 0x180:
                char_id[ord(tok)] = tid
        added = list(twin.added_tokens.items())
        ablob, aoff = _pack_strings([a for a, _ in added])
        aid = np.array([i for _, i in added], dtype=np.uint32)
        afl = np.array([(1 if twin.added_cfg[a]['single_word'] else 0) | (2 if twin.added_cfg[a]['lstrip'] else 0)
                        | (4 if twin.added_cfg[a]['rstrip'] else 0) for a, _ in added], dtype=np.uint8)
        toks = list(twin.id_to_token.items())       # id -> token, after the twin resolved duplicate ids
        tblob, toff = _pack_strings([t for _, t in toks])
        tid = np.array([i for i, _ in toks], dtype=np.uint32)
        tsp = np.array([1 if t in twin.special_tokens else 0 for _, t in toks], dtype=np.uint8)
        self._keep = (pa, pb, pr, ops, char_id, ablob, aoff, aid, afl, tblob, toff, tid, tsp)

        def ptr(a):
            return a.ctypes.data if a.size else None
        self.h = self.lib.orc_new(ptr(pa), ptr(pb), ptr(pr), len(pr), ptr(ops), len(ops), ptr(char_id),
                                  ptr(ablob), ptr(aoff), ptr(aid), ptr(afl), len(added),
                                  ptr(tblob), ptr(toff), ptr(tid), ptr(tsp), len(toks),
                                  1 if twin.normalizer == 'nfc' else 0, 1 if aps else 0)

    @classmethod
    def from_file(cls, path):
        return cls(OracleTokenizer.from_file(path))

    @classmethod
    def from_str(cls, s):
        return cls(OracleTokenizer.from_str(s))

    def __del__(self):
        try:
            self.lib.orc_free(self.h)
        except Exception:
            pass

    def encode_packed(self, text: np.ndarray, offs: np.ndarray, threads=None):
        """text uint8, offs uint64[n+1] -> (ids uint32, ids_off uint64[n+1])"""
        text = np.ascontiguousarray(text, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        if any(s[0] == 'split' for s in self.twin.pre_stages):
            return self._encode_packed_split(text, offs, threads)
        tp = text.ctypes.data if text.size else ctypes.addressof(ctypes.create_string_buffer(1))
        r = self.lib.orc_encode_batch(self.h, tp, offs.ctypes.data, n, threads or os.cpu_count() or 1)
        off = np.ctypeslib.as_array(ctypes.cast(self.lib.orc_result_off(r), ctypes.POINTER(ctypes.c_uint64)), (n + 1,)).copy()
        tot = int(off[-1])
        ids = np.ctypeslib.as_array(ctypes.cast(self.lib.orc_result_ids(r), ctypes.POINTER(ctypes.c_uint32)), (max(tot, 1),))[:tot].copy()
        self.lib.orc_result_free(r)
        return ids, off

    def _encode_packed_split(self, text, offs, threads):
        """Split stages (pretokenizers.rs:298-433 through the Sequence arm :114-124) run in Python on the normalised document
        (mod.rs:553 normalises first); the pieces then go through the C core one by one -- the ByteLevel stage restarts its
        pattern at every piece -- and their ids are concatenated per document (mod.rs:562-612)."""
        import unicodedata
        import py_regex
        raw = text.tobytes()
        pieces, first = [], [0]
        for d in range(len(offs) - 1):
            doc = raw[int(offs[d]):int(offs[d + 1])].decode('utf-8')
            if self.twin.normalizer == 'nfc':
                doc = unicodedata.normalize('NFC', doc)
            words = [doc]
            for kind, arg in self.twin.pre_stages:
                if kind == 'split':
                    words = [p for w in words for p in py_regex.split_with_behavior(arg[0], w, arg[1], arg[2])]
            pieces += [w.encode('utf-8') for w in words]
            first.append(len(pieces))
        po = np.zeros(len(pieces) + 1, dtype=np.uint64)
        if pieces:
            po[1:] = np.cumsum([len(b) for b in pieces], dtype=np.uint64)
        pt = np.frombuffer(b''.join(pieces), dtype=np.uint8)
        tp = pt.ctypes.data if pt.size else ctypes.addressof(ctypes.create_string_buffer(1))
        r = self.lib.orc_encode_batch(self.h, tp, po.ctypes.data, len(pieces), threads or os.cpu_count() or 1)
        off = np.ctypeslib.as_array(ctypes.cast(self.lib.orc_result_off(r), ctypes.POINTER(ctypes.c_uint64)), (len(pieces) + 1,)).copy()
        tot = int(off[-1])
        ids = np.ctypeslib.as_array(ctypes.cast(self.lib.orc_result_ids(r), ctypes.POINTER(ctypes.c_uint32)), (max(tot, 1),))[:tot].copy()
        self.lib.orc_result_free(r)
        return ids, off[np.array(first, dtype=np.int64)]

    def decode_packed(self, ids: np.ndarray, offs: np.ndarray, skip_special_tokens=False,
                      clean_up_tokenization_spaces=True, threads=None):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        ip = ids.ctypes.data if ids.size else ctypes.addressof(ctypes.create_string_buffer(4))
        r = self.lib.orc_decode_batch(self.h, ip, offs.ctypes.data, n, int(skip_special_tokens),
                                      int(clean_up_tokenization_spaces), threads or os.cpu_count() or 1)
        off = np.ctypeslib.as_array(ctypes.cast(self.lib.orc_result_off(r), ctypes.POINTER(ctypes.c_uint64)), (n + 1,)).copy()
        tot = int(off[-1])
        b = np.ctypeslib.as_array(ctypes.cast(self.lib.orc_result_bytes(r), ctypes.POINTER(ctypes.c_uint8)), (max(tot, 1),))[:tot].copy()
        self.lib.orc_result_free(r)
        return b, off

    # list-of-str conveniences (mirror Tokenizer.encode_batch / decode_batch)
    def encode_batch(self, texts, threads=None):
        blobs = [t.encode('utf-8') if isinstance(t, str) else bytes(t) for t in texts]
        offs = np.zeros(len(blobs) + 1, dtype=np.uint64)
        if blobs:
            offs[1:] = np.cumsum([len(b) for b in blobs], dtype=np.uint64)
        text = np.frombuffer(b''.join(blobs), dtype=np.uint8)
        ids, off = self.encode_packed(text, offs, threads)
        return [ids[int(off[i]):int(off[i + 1])].tolist() for i in range(len(blobs))]

    def decode_batch(self, batch, skip_special_tokens=False, clean_up_tokenization_spaces=True, threads=None):
        offs = np.zeros(len(batch) + 1, dtype=np.uint64)
        if batch:
            offs[1:] = np.cumsum([len(b) for b in batch], dtype=np.uint64)
        ids = np.array([i for b in batch for i in b], dtype=np.uint32)
        b, off = self.decode_packed(ids, offs, skip_special_tokens, clean_up_tokenization_spaces, threads)
        raw = b.tobytes()
        return [raw[int(off[i]):int(off[i + 1])].decode('utf-8') for i in range(len(batch))]
