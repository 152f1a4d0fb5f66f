"""The trainer restatement (oracle/py_trainer.py) against the reference's own tests (bpe_trainer.rs:473-510) and
hand-computed answers.  CPU only."""
import py_trainer


def test_reference_test_basic():                        # bpe_trainer.rs:473-494
    texts = ["hello world", "hello there", "world hello", "hello hello hello"]
    vocab, merges = py_trainer.train_bpe(texts, vocab_size=100, min_frequency=1)
    assert len(vocab) >= 4
    assert merges or len(vocab) <= 26
    for t in ('<unk>', '<pad>', '<s>', '</s>'):
        assert t in vocab
    assert 'hello' in vocab and 'world' in vocab          # every word ends up as one token


def test_reference_test_suffix():                       # bpe_trainer.rs:496-510
    vocab, _ = py_trainer.train_bpe(["hello world"], vocab_size=50, min_frequency=1, end_of_word_suffix="</w>")
    assert any('</w>' in k for k in vocab)


def test_hand_computed():
    vocab, merges = py_trainer.train_bpe(["ab ab ab b"], vocab_size=3, min_frequency=1, special_tokens=[])
    assert vocab == {'b': 0, 'a': 1, 'ab': 2} and merges == [('a', 'b')]
    # stops below min_frequency (:162-165): (c,d) occurs once
    vocab, merges = py_trainer.train_bpe(["ab ab cd"], vocab_size=10, min_frequency=2, special_tokens=[])
    assert merges == [('a', 'b')] and vocab == {'a': 0, 'b': 1, 'c': 2, 'd': 3, 'ab': 4}
    # ties: equal counts -> smallest (left index, right index); chars of equal frequency by code point
    vocab, merges = py_trainer.train_bpe(["xy ab"], vocab_size=6, min_frequency=1, special_tokens=[])
    assert [vocab[c] for c in 'abxy'] == [0, 1, 2, 3] and merges == [('a', 'b'), ('x', 'y')]


def test_existing_string_quirk():
    # a merged string equal to a special token overwrites its id with vocab.len() and the vocabulary does not grow (:168-169)
    vocab, merges = py_trainer.train_bpe(["<s> <s> <s> <s>"], vocab_size=9, min_frequency=1, special_tokens=['<s>'])
    assert merges == [('<', 's'), ('<s', '>')]
    assert vocab['<s>'] == 5                             # 1 special + 3 chars + '<s' = 5 entries when "<s>" is inserted again
    assert len(vocab) == 5


def test_white_space_is_the_unicode_property():
    # U+3000 and U+0085 split; U+001F and U+200B do not (str::split_whitespace, not Python's str.split)
    assert py_trainer.split_whitespace("a b\u3000c\u0085d\x1fe\u200bf") == ['a', 'b', 'c', 'd\x1fe\u200bf']


def test_prefix_and_limit():
    vocab, merges = py_trainer.train_bpe(["abab abab"], vocab_size=8, min_frequency=1, special_tokens=[], continuing_subword_prefix='##')
    assert set(vocab) >= {'a', 'b'} and merges[0] == ('a', '##b')
    vocab, _ = py_trainer.train_bpe(["aaa bb c"], vocab_size=2, min_frequency=1, special_tokens=[], limit_alphabet=2)
    assert vocab == {'a': 0, 'b': 1}
