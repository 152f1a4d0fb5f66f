"""CPU tests of the Split pre-tokenizer (reference src/pretokenizers.rs:298-433, parsing.rs:145-167).

Three implementations of "compile the pattern, walk find_iter, apply the behaviour" are compared on the same inputs:
  * oracle/py_regex.py          backtracking matcher over the AST (the restated `regex` crate semantics)
  * Python's `regex` module     an independent engine with the same leftmost-first semantics (the pin of the oracle)
  * the product's host side     regex -> ordered-subset DFA (csrc/regex_dfa.cpp) walked by split_walk() (csrc/split_walk.cuh,
                                the function the device kernels run) through the test hook ctk_debug_split_pieces
No GPU is involved.
"""
import ctypes
import json
import random

import numpy as np
import pytest

import py_oracle
import py_regex

# characters assigned since Unicode <= 13 (so that every engine's tables agree), of many classes
POOL = list("abcXYZ019 \t\n.,!?'-_()é中あア한Ωж٣५") + [' ', '　', '́', '‍', '\U0001F600', ' ', '½', 'ǅ', 'ʰ', '€']

PATTERNS = [
    r"\d", r"\p{N}{1,3}", r"[0-9]+", r"\s+", r"\w+", r"[^\s\p{L}\p{N}]+", r"\p{L}+|\p{N}+", r" ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+",
    r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+", r"a|ab|abc", r"abc|ab|a", r"(?:a|ab)(?:c|bcd)", r"a+?b", r"a*?b", r"a{2,}?",
    r"[一-龥぀-ゟ゠-ヿ]+", r"\p{Han}+", r"\p{Lu}\p{Ll}*", r"[[:alpha:]]+[[:digit:]]", r"[\p{L}&&]", r".", r".+", r"[^a]", r"\.\s", r"\x41|\u{4e2d}|é",
    r"(?P<w>\p{L}+)(?:'\p{L}+)?", r"[a-c]{2}|[a-c]", r"\pL\pN", r"\P{L}+", r"\S+", r"\D+", r"\W", r"[\s\S]", r"(a|b)+c", r"(?:ab?)+", r"x{0,2}y", r"\n+|\t",
    r"[]a]+", r"[a\-z]+", r"[^\]]", r"\p{Greek}+|\p{Cyrillic}+", r"\p{Alphabetic}", r"[\d\s]+", r"a{3}", r"\p{Letter}+", r"\p{gc=Nd}", r"\p{^L}",
]


def _texts(rng, n, lo=0, hi=60):
    return [''.join(rng.choice(POOL) for _ in range(rng.randint(lo, hi))) for _ in range(n)]


def _regex_module_pattern(p):
    """the same pattern for Python's `regex`: Rust's \\s \\d \\w are Unicode properties there"""
    out, i, depth = [], 0, 0
    while i < len(p):
        c = p[i]
        if c == '\\' and i + 1 < len(p):
            n = p[i + 1]
            rep = {'s': r'\p{White_Space}', 'S': r'\P{White_Space}', 'd': r'\p{Nd}', 'D': r'\P{Nd}'}.get(n)
            w = r'\p{Alphabetic}\p{M}\p{Nd}\p{Pc}‌‍'
            if n == 'w':
                rep = w if depth else '[' + w + ']'
            elif n == 'W':
                rep = '[^' + w + ']'
            if n in 'pP' and i + 2 < len(p) and p[i + 2] != '{':
                rep = '\\' + n + '{' + p[i + 2] + '}'
                i += 1
            if n in 'ux' and i + 2 < len(p) and p[i + 2] == '{':
                j = p.index('}', i)
                rep = '\\U%08x' % int(p[i + 3:j], 16)
                i = j - 1
            out.append(rep if rep is not None else p[i:i + 2])
            i += 2
            continue
        if c == '[':
            depth += 1
        elif c == ']' and depth:
            depth -= 1
        out.append(c)
        i += 1
    return ''.join(out)


def _module_spans(p, text):
    import regex
    rx = regex.compile(_regex_module_pattern(p), regex.V1)
    return [m.span() for m in rx.finditer(text)]


def test_oracle_matcher_against_regex_module():
    pytest.importorskip('regex')
    rng = random.Random(5)
    texts = _texts(rng, 120) + ['', 'a', "it's we're I'll", '中文abc123  x', 'aaab', 'abcabcab']
    checked = 0
    for p in PATTERNS:
        if '[:' in p:
            continue                                           # POSIX classes are ASCII-only in the crate, Unicode-aware in the module
        try:
            node = py_regex.compile_pattern(p)
        except py_regex.Unsupported:
            continue
        for t in texts:
            assert py_regex.find_iter(node, t) == _module_spans(p, t), (p, t)
            checked += 1
    assert checked > 4000


def _rand_pattern(rng, depth=0):
    atoms = ['a', 'b', 'c', ' ', '1', 'é', '中', r'\p{L}', r'\p{N}', r'\s', r'\d', r'\w', '[a-c]', r'[^a\s]', '.', r'\S', r'[\p{L}\d]']
    r = rng.random()
    if depth > 2 or r < 0.35:
        node = rng.choice(atoms)
    elif r < 0.6:
        node = ''.join(_rand_pattern(rng, depth + 1) for _ in range(rng.randint(2, 3)))
    elif r < 0.8:
        node = '(?:' + '|'.join(_rand_pattern(rng, depth + 1) for _ in range(rng.randint(2, 3))) + ')'
    else:
        node = '(?:' + _rand_pattern(rng, depth + 1) + ')'
    if rng.random() < 0.45:
        node = '(?:' + node + ')' + rng.choice(['?', '*', '+', '{1,3}', '{2,}', '{2}', '+?', '*?', '??', '{1,2}?'])
    return node


def _host_pieces(lib, pattern, behavior, invert, text, segment=0):
    raw = text.encode('utf-8')
    cap = len(raw) + 4
    buf = (ctypes.c_uint64 * (2 * cap))()
    n = ctypes.c_size_t(0)
    ns, nc = ctypes.c_uint32(0), ctypes.c_uint32(0)
    tb = (ctypes.c_uint8 * max(1, len(raw))).from_buffer_copy(raw or b'\0')
    rc = lib.ctk_debug_split_pieces(pattern.encode('utf-8'), behavior, int(invert), tb, len(raw), buf, cap, ctypes.byref(n), ctypes.byref(ns), ctypes.byref(nc), segment)
    if rc != 0:
        return rc, None
    return 0, [raw[buf[2 * k]:buf[2 * k + 1]].decode('utf-8') for k in range(n.value)]


@pytest.fixture(scope='module')
def lib(built_lib):
    import complexity_tokenizer as ct
    lb = ct._lib()
    lb.ctk_debug_split_pieces.restype = ctypes.c_int
    lb.ctk_debug_split_pieces.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    return lb


def test_host_dfa_against_oracle_all_behaviours(lib):
    """the product's compiler + walk == the oracle's matcher + the reference's five behaviours, line by line"""
    rng = random.Random(11)
    texts = _texts(rng, 40) + ['', 'a', "it's we're I'll", '中文abc123  x', '12345678', '  ', 'a1b22c333']
    n_ok = 0
    for p in PATTERNS:
        try:
            node = py_regex.compile_pattern(p)
        except py_regex.Unsupported:
            rc, _ = _host_pieces(lib, p, 1, False, 'abc')
            assert rc == 3, p                                  # CTK_ERR_UNSUPPORTED, never a guess
            continue
        for bi, beh in enumerate(py_regex.BEHAVIORS):
            for inv in ((False, True) if bi == 0 else (False,)):
                for t in texts:
                    want = [w for w in py_regex.split_with_behavior(node, t, beh, inv) if w]       # (an empty text stays [""] in the reference: no piece)
                    for seg in (0, 1, 5):                      # one walk / walked in segments cut at safe starts, as the device kernel does
                        rc, got = _host_pieces(lib, p, bi, inv, t, seg)
                        assert rc == 0, (p, rc)
                        assert got == want, (p, beh, inv, t, seg)
                    n_ok += 1
    assert n_ok > 8000


def test_random_patterns_three_way(lib):
    """fuzz: random patterns of the subset -- oracle matcher == `regex` module == host DFA (leftmost-first, lazy and greedy)"""
    pytest.importorskip('regex')
    rng = random.Random(77)
    done = 0
    tries = 0
    while done < 250 and tries < 3000:
        tries += 1
        p = _rand_pattern(rng)
        try:
            node = py_regex.compile_pattern(p)
        except py_regex.Unsupported:
            rc, _ = _host_pieces(lib, p, 1, False, 'abc')
            assert rc == 3, p
            continue
        done += 1
        for t in _texts(rng, 12, 0, 30):
            spans = py_regex.find_iter(node, t)
            assert spans == _module_spans(p, t), (p, t)
            rc, got = _host_pieces(lib, p, 0, False, t)       # Removed, not inverted: the matches themselves
            assert rc == 0, (p, rc)
            want = [t[a:b] for a, b in spans] if spans else ([t] if t else [])
            assert got == want, (p, t)
    assert done == 250


def test_unsupported_and_rejected_agree(lib):
    for p, want in ((r"(?i)a", 3), (r"^a", 3), (r"a$", 3), (r"\bfoo", 3), (r"a*", 3), (r"", 3), (r"a|", 3), (r"(a*)+", 3), (r"[a&&b]", 3), (r"\p{Tamil}", 3),
                    (r"a{2000}", 3), (r"(?x) a", 3), (r"\<a", 3), (r"a{,3}", 3), (r"(?=a)b", -1), (r"(a)\1", -1), (r"x(?!y)", -1), (r"\d", 0), (r"\p{N}{1,3}", 0)):
        rc, _ = _host_pieces(lib, p, 1, False, 'abc 123')
        assert rc == want, (p, rc)
        if want == 3:
            with pytest.raises(py_regex.Unsupported):
                py_regex.compile_pattern(p)
        if want == -1:
            assert py_oracle._rust_regex_certainly_rejected(p)


def test_split_tokenizer_oracle_end_to_end():
    """a Sequence[Split, ByteLevel] tokenizer through the oracle: pieces restart the ByteLevel pattern"""
    vocab = {ch: i for i, ch in enumerate(sorted(set(py_oracle.BYTE_ENCODER.values())))}
    tj = {'model': {'type': 'BPE', 'vocab': vocab, 'merges': []}, 'normalizer': None,
          'pre_tokenizer': {'type': 'Sequence', 'pretokenizers': [{'type': 'Split', 'pattern': {'Regex': r'\d'}, 'behavior': 'Isolated', 'invert': False},
                                                                  {'type': 'ByteLevel', 'add_prefix_space': False}]}}
    t = py_oracle.OracleTokenizer(tj)
    assert t.pre_tokenize("ab12 c3") == ['ab', '1', '2', 'Ġc', '3']
    tj['pre_tokenizer']['pretokenizers'][0]['behavior'] = 'Removed'
    assert py_oracle.OracleTokenizer(tj).pre_tokenize("ab12 c3") == ['1', '2', '3']
    tj['pre_tokenizer']['pretokenizers'][0]['invert'] = True
    assert py_oracle.OracleTokenizer(tj).pre_tokenize("ab12 c3") == ['ab', 'Ġc']
    tj['pre_tokenizer']['pretokenizers'][0] = {'type': 'Split', 'pattern': {'Regex': r'\s+(?!\S)'}, 'behavior': 'Isolated'}
    assert py_oracle.OracleTokenizer(tj).pre_tokenize("ab12 c3") == ['ab', '12', 'Ġc', '3']     # Regex::new fails: passes through
    tj['pre_tokenizer'] = {'type': 'Split', 'pattern': {'Regex': r'\d'}, 'behavior': 'Isolated'}
    with pytest.raises(py_oracle.Unsupported):
        py_oracle.OracleTokenizer(tj)                          # no ByteLevel stage: outside the hot path
    assert json.dumps(tj)


def _rich_pattern(rng, depth=0):
    """a wider slice of the supported syntax than _rand_pattern: bracket classes with ranges / negation / escapes / nesting,
    negated properties, hex escapes, counted repetitions with zero lower bounds, named and capturing groups"""
    def klass():
        items = rng.sample(['a-c', 'x', '0-9', r'\d', r'\s', r'\p{L}', r'\p{Lu}', 'é', '中-和', r'\-', r'\]', '_', r'\x41', r'\u{3042}', '[b-d]', r'\P{N}'], rng.randint(1, 3))
        return '[' + ('^' if rng.random() < 0.3 else '') + ''.join(items) + ']'
    atoms = ['a', 'b', '1', ' ', '_', 'é', '中', r'\.', r'\p{L}', r'\P{L}', r'\pN', r'\s', r'\S', r'\w', r'\W', r'\d', r'\D', '.', r'\x62', r'\n']
    r = rng.random()
    if depth > 2 or r < 0.3:
        node = klass() if rng.random() < 0.4 else rng.choice(atoms)
    elif r < 0.55:
        node = ''.join(_rich_pattern(rng, depth + 1) for _ in range(rng.randint(2, 3)))
    elif r < 0.75:
        node = rng.choice(['(?:', '(', '(?P<n%d>' % rng.randint(0, 9)]) + '|'.join(_rich_pattern(rng, depth + 1) for _ in range(rng.randint(2, 3))) + ')'
    else:
        node = '(?:' + _rich_pattern(rng, depth + 1) + ')'
    if rng.random() < 0.4:
        node = '(?:' + node + ')' + rng.choice(['?', '*', '+', '{0,2}', '{1,3}', '{2,}', '{3}', '+?', '*?', '??', '{0,2}?', '{1,}?'])
    return node


def test_rich_patterns_oracle_against_host_dfa(lib):
    """400 random patterns of the wider grammar: the oracle's backtracking matcher and the product's DFA + walk (also cut at
    every safe start) produce the same pieces for every behaviour; where the `regex` module reads the pattern the same way, it agrees too"""
    rng = random.Random(2024)
    done = tries = agree_module = 0
    while done < 400 and tries < 6000:
        tries += 1
        p = _rich_pattern(rng)
        try:
            node = py_regex.compile_pattern(p)
        except py_regex.Unsupported:
            rc, _ = _host_pieces(lib, p, 1, False, 'abc')
            assert rc == 3, p
            continue
        rc, _ = _host_pieces(lib, p, 1, False, 'abc')
        if rc == 3:                                            # the product bounds its automaton (states, table size); the oracle has no such bound
            msg = ctypes.cast(lib.ctk_last_error(), ctypes.c_char_p).value or b''
            assert b'too many' in msg or b'too large' in msg, (p, msg)
            continue
        done += 1
        bi = rng.randrange(5)
        inv = bi == 0 and rng.random() < 0.5
        for t in _texts(rng, 8, 0, 40):
            want = [w for w in py_regex.split_with_behavior(node, t, py_regex.BEHAVIORS[bi], inv) if w]
            for seg in (0, 1):
                rc, got = _host_pieces(lib, p, bi, inv, t, seg)
                assert rc == 0, (p, rc)
                assert got == want, (p, py_regex.BEHAVIORS[bi], inv, t, seg)
            try:
                if '(?P<' not in p and py_regex.find_iter(node, t) == _module_spans(p, t):
                    agree_module += 1
            except Exception:
                pass
    assert done == 400 and agree_module > 2000
