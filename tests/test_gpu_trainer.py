"""BPE training on the GPU (csrc/train.cu through ctk_train_bpe) against the restated reference trainer
(oracle/py_trainer.py): vocabulary map and merge list must be equal, entry for entry."""
import json
import random

import numpy as np
import pytest

import py_trainer

pytestmark = pytest.mark.gpu


def both(texts, **kw):
    import complexity_tokenizer as ct
    want = py_trainer.train_bpe(texts, **kw)
    tr = ct.BpeTrainer(show_progress=False, **kw)
    got = tr.train(texts)
    if got[1] != want[1]:
        k = next((i for i, (a, b) in enumerate(zip(got[1], want[1])) if a != b), min(len(got[1]), len(want[1])))
        raise AssertionError('merges differ at %d: got %r want %r (%d / %d merges)' % (k, got[1][k:k + 3], want[1][k:k + 3], len(got[1]), len(want[1])))
    assert got[0] == want[0]
    return got, tr.last_stats


def test_reference_tests(built_lib):                     # bpe_trainer.rs:473-510
    (vocab, merges), _ = both(["hello world", "hello there", "world hello", "hello hello hello"], vocab_size=100, min_frequency=1)
    assert len(vocab) >= 4 and merges
    (vocab, _), _ = both(["hello world"], vocab_size=50, min_frequency=1, end_of_word_suffix="</w>")
    assert any('</w>' in k for k in vocab)


def test_edges(built_lib):
    both([], vocab_size=50)
    both(["", "   ", "\n\t"], vocab_size=50)
    both(["a"], vocab_size=50, min_frequency=1)
    both(["ab", "cd", "ab"], vocab_size=50, min_frequency=1)                  # words do not span texts
    both(["ab ab cd"], vocab_size=10, min_frequency=2, special_tokens=[])      # stops below min_frequency
    both(["ab ab ab b"], vocab_size=3, min_frequency=1, special_tokens=[])     # stops when the vocabulary is full
    both(["ab ab ab b"], vocab_size=2, min_frequency=1, special_tokens=[])     # full before the first merge
    both(["<s> <s> <s> <s> x<s>y"], vocab_size=12, min_frequency=1, special_tokens=['<s>', '<s>', 'q'])   # existing-string quirk, duplicate specials
    both(["a b\u3000c\u0085d\x1fe\u200bf \u00a0xy\u2003z\u2028\u1680w \u205fv\u202f"], vocab_size=40, min_frequency=1)   # White_Space, not str.split


def test_options(built_lib):
    texts = ["abab abab abc cab", "the cat sat on the mat", "hello hello world"]
    both(texts, vocab_size=60, min_frequency=1, continuing_subword_prefix='##')
    both(texts, vocab_size=60, min_frequency=1, end_of_word_suffix='</w>', continuing_subword_prefix='@@')
    both(texts, vocab_size=12, min_frequency=1, special_tokens=[], limit_alphabet=4)
    both(texts, vocab_size=80, min_frequency=1, initial_alphabet=[chr(c) for c in range(0x61, 0x7B)] + ['Ġ'])


def test_long_words(built_lib):
    rng = random.Random(5)
    texts = ["a" * 1000, "ab" * 333 + "a", "aaa" + "b" * 129 + "aa", " ".join("a" * k for k in range(60, 140)),
             "".join(rng.choice("abc") for _ in range(5000)), "x" + "a" * 64, "a" * 65, "a" * 64, "ba" * 33]
    both(texts, vocab_size=120, min_frequency=1)
    both(texts, vocab_size=300, min_frequency=2, special_tokens=[])


def test_unicode_text(built_lib):
    texts = ["café café naïve 日本語 日本 \U0001F600\U0001F600 \U0001F44D\U0001F3FD ok",
             "日本語　日本語 ééé \U0001F600"] * 3
    both(texts, vocab_size=90, min_frequency=1)


def _english(seed, nbytes):
    import synth
    text, offs = synth.gen_corpus('english', seed, nbytes)
    raw = text.tobytes()
    return [raw[int(offs[i]):int(offs[i + 1])].decode() for i in range(len(offs) - 1)]


def test_english_sample(built_lib):
    texts = _english(424, 96 << 10)
    (vocab, merges), stats = both(texts, vocab_size=700, min_frequency=2)
    assert len(merges) > 400 and stats['n_words'] > 10000 and stats['n_unique_words'] < stats['n_words']
    # byte-level shape: one text per pre-token-like piece, initial alphabet given
    pieces = [w for t in texts[:40] for w in t.replace(' ', ' Ġ').split(' ') if w]
    both(pieces, vocab_size=500, min_frequency=2, initial_alphabet=[chr(c) for c in range(0x21, 0x7F)] + ['Ġ'])


@pytest.mark.parametrize('seed', range(8))
def test_fuzz(built_lib, seed):
    rng = random.Random(seed)
    alpha = rng.choice(["ab", "abc", "abcdefgh", "abé日", "a<s>"])
    ws = [" ", "  ", "\n", " ", "\t", "\u3000"]
    texts = []
    for _ in range(rng.randrange(1, 40)):
        t = []
        for _ in range(rng.randrange(0, 30)):
            n = rng.choice([1, 2, 3, 5, 8, 70, 200]) if rng.random() < 0.9 else 400
            t.append("".join(rng.choice(alpha) for _ in range(n)))
            t.append(rng.choice(ws))
        texts.append("".join(t))
    both(texts, vocab_size=rng.choice([10, 40, 200]), min_frequency=rng.choice([1, 2, 3]),
         special_tokens=rng.choice([None, [], ['<s>', 'ab']]),
         continuing_subword_prefix=rng.choice([None, None, '##']), end_of_word_suffix=rng.choice([None, None, '</w>']))


def test_trained_table_feeds_the_encode_path(built_lib, tmp_path):
    """train on ByteLevel pieces -> tokenizer.json -> encode_batch on the GPU == the encode oracle on the same table."""
    import c_oracle
    import complexity_tokenizer as ct
    import synth
    docs = _english(99, 48 << 10)
    b2u = dict(zip(*synth.bytes_to_unicode()))
    mapped = ["".join(b2u[b] for b in w.encode()) for d in docs for w in d.replace(' ', '\n ').split('\n') if w]
    alphabet = [b2u[b] for b in range(256)]
    tr = ct.BpeTrainer(vocab_size=256 + 300, min_frequency=2, special_tokens=[], initial_alphabet=alphabet, show_progress=False)
    vocab, merges = tr.train(mapped)
    assert len(merges) == 300 and len(vocab) == 556 and sorted(vocab.values()) == list(range(556))
    tj = {"version": "1.0", "added_tokens": [], "normalizer": None,
          "pre_tokenizer": {"type": "ByteLevel", "add_prefix_space": False, "trim_offsets": True, "use_regex": True},
          "decoder": {"type": "ByteLevel"}, "post_processor": None,
          "model": {"type": "BPE", "vocab": vocab, "merges": ["%s %s" % m for m in merges]}}
    path = tmp_path / 'trained.json'
    path.write_text(json.dumps(tj, ensure_ascii=False), encoding='utf-8')
    tok = ct.Tokenizer.from_file(str(path), device=0)
    orc = c_oracle.COracle.from_file(str(path))
    text, offs = synth.pack([d.encode() for d in docs])
    ids, ioff = tok.encode_packed(text, offs)
    wids, woff = orc.encode_packed(text, offs)
    assert np.array_equal(ioff, woff) and np.array_equal(ids, wids)
    assert ids.size < text.size * 0.6                       # the trained merges are in use


@pytest.mark.parametrize('n_words', [60000, 130000])
def test_pair_table_growth(built_lib, n_words):
    """More distinct pairs than the first pair table holds: fill guard (60K) and overflow (130K) -> recount into a larger table."""
    rng = random.Random(n_words)
    alpha = [chr(0x4E00 + k) for k in range(400)]
    words = ["".join(rng.choice(alpha) for _ in range(2)) for _ in range(n_words)]
    texts = [" ".join(words[i:i + 50]) for i in range(0, n_words, 50)]
    _, stats = both(texts, vocab_size=4 + 400 + 25, min_frequency=1)
    assert stats['table_rebuilds'] >= 1


@pytest.mark.parametrize('size', ['8', '16'])
def test_cluster_merge_loop(built_lib, monkeypatch, size):
    """The opt-in single-cluster merge loop (CTK_TRAIN_CLUSTER) takes the same decisions as the three-kernel loop."""
    monkeypatch.setenv('CTK_TRAIN_CLUSTER', size)
    texts = _english(7, 64 << 10)
    _, stats = both(texts, vocab_size=500, min_frequency=2)
    assert stats['cluster_size'] in (8, 16)
    both(["a" * 1000, "ab" * 333 + "a", "<s> <s> x<s>y"], vocab_size=60, min_frequency=1, special_tokens=['<s>'])


def test_three_kernel_merge_loop(built_lib, monkeypatch):
    """Default: detect + apply fused, a thread per word (two launches per merge) when no word is longer than 256 symbols; the
    three-kernel loop (warp per listed word) otherwise, or with CTK_TRAIN_THREE_KERNELS: both equal the oracle."""
    texts = _english(9, 64 << 10)
    both(texts, vocab_size=500, min_frequency=2)                                    # fused path (short words)
    both(["a" * 1000, "ab" * 333 + "a", "<s> <s> x<s>y"], vocab_size=60, min_frequency=1, special_tokens=['<s>'])   # long words: three kernels
    monkeypatch.setenv('CTK_TRAIN_THREE_KERNELS', '1')
    both(texts, vocab_size=500, min_frequency=2)
    monkeypatch.setenv('CTK_TRAIN_NO_PDL', '1')
    both(texts[:40], vocab_size=300, min_frequency=2)


def test_properties_at_scale(built_lib):
    """64 MiB (11 M words), where the oracle cannot go: size-independent properties of the reference's loop."""
    import complexity_tokenizer as ct
    import synth
    text, offs = synth.gen_corpus('english', 2002, 64 << 20)
    tr = ct.BpeTrainer(vocab_size=3000, min_frequency=2, show_progress=False)
    vocab, merges = tr.train_packed(text, offs)
    counts, st = tr.last_merge_counts, tr.last_stats
    assert st['n_bytes'] == text.size and st['n_words'] > 10_000_000 and st['n_unique_words'] < st['n_words'] // 10
    assert len(vocab) == 3000 and sorted(vocab.values()) == list(range(3000))      # no string was produced twice here
    assert len(merges) == len(counts) == 3000 - (len(vocab) - len(merges))
    assert all(int(counts[i]) >= int(counts[i + 1]) for i in range(len(counts) - 1))   # a merge never creates a more frequent pair
    assert int(counts[-1]) >= 2
    for l, r in merges:
        assert l in vocab and r in vocab and (l + r) in vocab
    # the histogram: the word counts of the text's first megabyte, restated on the CPU, bound the first merge's count
    assert int(counts[0]) <= st['n_bytes']
    # the same input trained twice gives the same table (atomics only ever add: no order dependence)
    vocab2, merges2 = ct.BpeTrainer(vocab_size=3000, min_frequency=2, show_progress=False).train_packed(text, offs)
    assert merges2 == merges and vocab2 == vocab


def test_word_table_growth(built_lib, monkeypatch):
    """A word table that starts far too small (all words distinct) grows instead of probing a full table."""
    monkeypatch.setenv('CTK_TRAIN_WORD_TABLE_DIV', '4096')
    rng = random.Random(11)
    words = sorted({"".join(rng.choice("abcdefghij") for _ in range(8)) for _ in range(120000)})
    texts = [" ".join(words[i:i + 100]) for i in range(0, len(words), 100)]
    _, stats = both(texts, vocab_size=4 + 10 + 8, min_frequency=1)
    assert stats['n_unique_words'] == len(words) > 65536           # more than the first table holds


def test_train_new_from_iterator_on_device(built_lib, tok_paths):
    """mod.rs:1231-1322: normaliser + pre-tokenizer + trainer all on the device == the restated reference (oracle pre-tokens ->
    oracle/py_trainer.py); the new tokenizer keeps the pipeline and the special tokens and encodes like the oracle built from
    the same result"""
    import json
    import complexity_tokenizer as ct
    import py_oracle
    import synth
    for cfg, kind, n_docs, vs in (('config1', 'english', 60, 600), ('config3', 'mixed', 40, 700), ('config2', 'ascii', 30, 500)):
        tok = ct.Tokenizer.from_file(tok_paths[cfg])
        orc = py_oracle.OracleTokenizer.from_file(tok_paths[cfg])
        text, offs = synth.gen_corpus(kind, 4242, 1 << 20)
        raw = text.tobytes()
        docs = [raw[int(offs[i]):int(offs[i + 1])].decode('utf-8') for i in range(n_docs)]
        assert tok.all_special_tokens() == orc.all_special_tokens()
        new = tok.train_new_from_iterator(docs, vs)
        want_vocab, want_merges = orc.train_new_from_iterator(docs, vs)
        tj = new._config_json()
        assert tj['model']['vocab'] == want_vocab
        assert tj['model']['merges'] == [a + ' ' + b for a, b in want_merges]
        assert len(want_merges) > 100
        orc_new = py_oracle.OracleTokenizer(tj)
        assert new.encode_batch(docs[:10]) == orc_new.encode_batch(docs[:10])
        assert new.decode_batch_with_options(new.encode_batch(docs[:5]), False, False) == [unicodedata_nfc(d, orc) for d in docs[:5]]


def unicodedata_nfc(d, orc):
    import unicodedata
    return unicodedata.normalize('NFC', d) if orc.normalizer == 'nfc' else d
