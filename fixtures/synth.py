"""Synthetic corpora + tokenizer.json files for the configs of BASELINE.json.

Not product code and not the oracle: plain deterministic input generators shared by tests and
bench.  The heavy lifting is in fixtures.cpp (built on demand with g++ into fixtures/_build/).
Generated tokenizer.json files are cached under fixtures/_cache/ (git-ignored).
"""
import ctypes
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, '_build')
CACHE = os.path.join(HERE, '_cache')
_LIB = None


def build_lib(force=False):
    global _LIB
    so = os.path.join(BUILD, 'libfixtures.so')
    src = os.path.join(HERE, 'fixtures.cpp')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(BUILD, exist_ok=True)
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-pthread', src, '-o', so])
        _LIB = None
    if _LIB is None:
        lib = ctypes.CDLL(so)
        lib.fx_gen_corpus.restype = ctypes.c_long
        lib.fx_gen_corpus.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_double,
                                      ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
        lib.fx_train_bpe.restype = ctypes.c_long
        lib.fx_train_bpe.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
        _LIB = lib
    return _LIB


KIND = {'english': 0, 'ascii': 1, 'mixed': 2}


def gen_corpus(kind, seed, target_bytes, doc_median=1024, doc_min=64, doc_max=16384, lexicon=50000,
               out=None, threads=None):
    """-> (text uint8[total], offsets uint64[n_docs+1]).  `out` may be a preallocated uint8 array
    (e.g. a view of pinned memory) of at least target_bytes."""
    lib = build_lib()
    target_bytes = int(target_bytes)
    if out is None:
        out = np.empty(target_bytes, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= target_bytes and out.flags.c_contiguous
    max_docs = target_bytes // max(1, doc_min) + 2
    offs = np.empty(max_docs + 1, dtype=np.uint64)
    n = lib.fx_gen_corpus(KIND[kind], seed, target_bytes, float(doc_median), doc_min, doc_max, lexicon,
                          out.ctypes.data, out.size, offs.ctypes.data, max_docs, threads or os.cpu_count() or 1)
    if n < 0:
        raise RuntimeError('fx_gen_corpus: capacity')
    offs = offs[:n + 1].copy()
    return out[:int(offs[-1])], offs


def split_docs(text, offs):
    b = text.tobytes()
    return [b[int(offs[i]):int(offs[i + 1])] for i in range(len(offs) - 1)]


def bytes_to_unicode():
    bs = list(range(ord('!'), ord('~') + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return bs, [chr(c) for c in cs]


def train_merges(sample, n_merges, min_frequency=2):
    """-> list of (left, right) in symbol space (0..255 bytes, 256+k = k-th merge)."""
    lib = build_lib()
    sample = np.ascontiguousarray(sample, dtype=np.uint8)
    pairs = np.zeros(2 * n_merges, dtype=np.uint32)
    made = lib.fx_train_bpe(sample.ctypes.data, sample.size, n_merges, min_frequency, pairs.ctypes.data)
    return [(int(pairs[2 * i]), int(pairs[2 * i + 1])) for i in range(made)]


def assemble_tokenizer(pairs, *, specials_first=(), specials_last=(), specials_in_vocab=True,
                       normalizer='absent', pre_tokenizer=None, decoder=None, added_extra=(), merges_as_arrays=False,
                       max_vocab=None):
    """Build a tokenizer.json dict.  ids: specials_first, then the 256 byte symbols (GPT-2 order),
    then one id per merge, then specials_last."""
    bs, cs = bytes_to_unicode()
    vocab = {}
    added = []
    nid = 0
    for s in specials_first:
        if specials_in_vocab:
            vocab[s] = nid
        added.append({'id': nid, 'content': s, 'special': True, 'single_word': False, 'lstrip': False,
                      'rstrip': False, 'normalized': False})
        nid += 1
    sym = {}
    for b, c in zip(bs, cs):
        vocab[c] = nid
        sym[b] = c
        nid += 1
    merges = []
    for k, (a, b) in enumerate(pairs):
        if max_vocab is not None and nid + len(specials_last) >= max_vocab:
            break
        sa, sb = sym[a], sym[b]
        tok = sa + sb
        if tok in vocab:            # same string reachable by two merges: keep the first id (as a dict would)
            sym[256 + k] = tok
            merges.append((sa, sb))
            continue
        vocab[tok] = nid
        sym[256 + k] = tok
        nid += 1
        merges.append((sa, sb))
    for s in specials_last:
        if specials_in_vocab:
            vocab[s] = nid
        added.append({'id': nid, 'content': s, 'special': True, 'single_word': False, 'lstrip': False,
                      'rstrip': False, 'normalized': False})
        nid += 1
    added += list(added_extra)
    tj = {'version': '1.0',
          'model': {'type': 'BPE', 'vocab': vocab,
                    'merges': [[a, b] for a, b in merges] if merges_as_arrays else ['%s %s' % m for m in merges]},
          'added_tokens': added}
    if normalizer != 'absent':
        tj['normalizer'] = normalizer
    tj['pre_tokenizer'] = pre_tokenizer if pre_tokenizer is not None else {'type': 'ByteLevel', 'add_prefix_space': False}
    tj['decoder'] = decoder if decoder is not None else {'type': 'ByteLevel'}
    return tj


def _cached(name, builder):
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, name)
    if not os.path.exists(path):
        tj = builder()
        tmp = path + '.tmp%d' % os.getpid()
        with open(tmp, 'w', encoding='utf-8') as f:
            json.dump(tj, f, ensure_ascii=False)
        os.replace(tmp, path)
    return path


def tokenizer_config1(vocab_size=32000):
    """32K: 4 specials + 256 byte symbols + merges trained on the first 2000 config-1 docs."""
    def build():
        text, offs = gen_corpus('english', 1001, 12 << 20)
        sample = text[:int(offs[min(2000, len(offs) - 1)])]
        pairs = train_merges(sample, vocab_size - 260 + 2000)
        return assemble_tokenizer(pairs, specials_first=('<unk>', '<pad>', '<s>', '</s>'), max_vocab=vocab_size)
    return _cached('config1_%d.json' % vocab_size, build)


def tokenizer_config2(n_merges=50000):
    """GPT-2 shape: 256 byte symbols + 50 000 merges + <|endoftext|> (50 257 entries), "normalizer": null."""
    def build():
        text, offs = gen_corpus('ascii', 2002, 32 << 20, doc_median=4096, doc_min=256, doc_max=65536)
        pairs = train_merges(text, n_merges + 2000)
        return assemble_tokenizer(pairs, specials_last=('<|endoftext|>',), normalizer=None, max_vocab=257 + n_merges)
    return _cached('config2_%d.json' % n_merges, build)


def tokenizer_config3(vocab_size=100000):
    """Pacific-Prime / INL-trainer save shape (src/trainer.rs:598-651): 4 specials ids 0..3 in
    added_tokens, no normalizer key, ByteLevel{add_prefix_space:false,use_regex:true}."""
    def build():
        text, offs = gen_corpus('mixed', 3003, 48 << 20, doc_median=4096, doc_min=256, doc_max=65536)
        pairs = train_merges(text, vocab_size - 260 + 4000)
        return assemble_tokenizer(pairs, specials_first=('</s>', '<pad>', '<s>', '<unk>'), max_vocab=vocab_size,
                                  pre_tokenizer={'type': 'ByteLevel', 'add_prefix_space': False, 'use_regex': True})
    return _cached('config3_%d.json' % vocab_size, build)


def gen_long_docs(seed=4004, doc_bytes=1 << 20, n_docs=64):
    """config 4: documents with very long pre-tokens."""
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, size=int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(200)]
    docs = []
    for d in range(n_docs):
        k = d % 4
        if k == 0:                                      # one run of [a-z], no separators
            parts, n = [], 0
            while n < doc_bytes:
                w = words[int(rng.integers(0, 200))]
                parts.append(w)
                n += len(w)
            docs.append(b''.join(parts)[:doc_bytes])
        elif k == 1:                                    # 64 KiB letter runs separated by single spaces
            parts, n = [], 0
            while n < doc_bytes:
                run, m = [], 0
                while m < 65536:
                    w = words[int(rng.integers(0, 200))]
                    run.append(w)
                    m += len(w)
                parts.append(b''.join(run)[:65535] + b' ')
                n += 65536
            docs.append(b''.join(parts)[:doc_bytes])
        elif k == 2:
            docs.append(b' ' * doc_bytes)
        else:
            docs.append(bytes(rng.choice(np.frombuffer(b'=-_*', dtype=np.uint8), size=doc_bytes)))
    return docs


def pack(docs):
    """list[bytes] -> (uint8 text, uint64 offsets)"""
    offs = np.zeros(len(docs) + 1, dtype=np.uint64)
    if docs:
        offs[1:] = np.cumsum([len(d) for d in docs], dtype=np.uint64)
    text = np.frombuffer(b''.join(docs), dtype=np.uint8).copy() if docs else np.zeros(0, dtype=np.uint8)
    return text, offs
