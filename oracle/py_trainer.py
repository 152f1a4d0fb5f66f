"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's BPE trainer.

Follows /root/reference/src/bpe_trainer.rs line by line, with strings as symbols exactly like the reference:
    train                     bpe_trainer.rs:100-228
    build_word_frequencies    bpe_trainer.rs:241-275   (`split_whitespace` = Unicode White_Space)
    build_initial_vocab       bpe_trainer.rs:278-320
    split_word                bpe_trainer.rs:323-338
    count_pairs               bpe_trainer.rs:341-376   (full recount every iteration, u32 sums)
    merge_pair                bpe_trainer.rs:379-401

Only `tests/`, `__graft_entry__.smoke()` and bench.py's cpu_baseline leg may import this file.

PARITY UNPINNED beyond the reference's own property tests (bpe_trainer.rs:473-510: "some merges", "a token
contains </w>"): the reference takes two decisions in hash-iteration order, so its output is not a function of its
input.  This restatement fixes both, and the CUDA path (csrc/train.cu) takes the same ones:
  * the best pair (`max_by_key` over a HashMap, :152-155) -- highest count, ties by the smallest
    (symbol index of the left part, symbol index of the right part);
  * characters of equal frequency (`sort_by` on a Vec collected from a HashMap, :305-306) -- by code point.
Symbol indices: special tokens in the order given, then the initial alphabet, then the characters of the data by
(frequency descending, code point), then the prefixed continuation symbols in that same character order, then
merged strings in the order they first appear.  Every result the restatement can produce is one the reference
can produce (for some hash seed); the vocabulary ids and the quirks (an id given twice when a merged string already
exists, :168-169; duplicate special tokens, :283-286) are the reference's.
"""

WHITE_SPACE = frozenset([9, 10, 11, 12, 13, 0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000] + list(range(0x2000, 0x200B)))
DEFAULT_SPECIALS = ("<unk>", "<pad>", "<s>", "</s>")            # bpe_trainer.rs:38-43
M32 = 0xFFFFFFFF


def split_whitespace(text):
    out, cur = [], []
    for ch in text:
        if ord(ch) in WHITE_SPACE:
            if cur:
                out.append(''.join(cur))
                cur = []
        else:
            cur.append(ch)
    if cur:
        out.append(''.join(cur))
    return out


def train_bpe(texts, vocab_size=30000, min_frequency=2, special_tokens=None, initial_alphabet=None, limit_alphabet=None,
              continuing_subword_prefix=None, end_of_word_suffix=None, return_symbols=False):
    """-> (vocab: dict[str, int], merges: list[(str, str)])   (bpe_trainer.rs:100)"""
    special_tokens = list(DEFAULT_SPECIALS if special_tokens is None else special_tokens)
    # Step 1 (:241-275)
    word_freqs = {}
    for text in texts:
        for w in split_whitespace(text):
            if end_of_word_suffix is not None:
                w = w + end_of_word_suffix
            word_freqs[w] = (word_freqs.get(w, 0) + 1) & M32
    # Step 2 (:278-320)
    vocab, next_id = {}, 0
    index = {}                                                   # symbol string -> symbol index (tie-break order)

    def sym(s):
        if s not in index:
            index[s] = len(index)
        return index[s]

    for t in special_tokens:
        vocab[t] = next_id
        next_id += 1
        sym(t)
    if initial_alphabet is not None:
        for c in initial_alphabet:
            if c not in vocab:
                vocab[c] = next_id
                next_id += 1
            sym(c)
    char_freqs = {}
    for w, f in word_freqs.items():
        for c in w:
            char_freqs[c] = (char_freqs.get(c, 0) + f) & M32
    chars = sorted(char_freqs.items(), key=lambda kv: (-kv[1], ord(kv[0])))
    limit = len(chars) if limit_alphabet is None else limit_alphabet
    for k, (c, _) in enumerate(chars):
        if k < limit and c not in vocab:
            vocab[c] = next_id
            next_id += 1
        sym(c)
    if continuing_subword_prefix is not None:
        for c, _ in chars:
            sym(continuing_subword_prefix + c)

    # Step 3 (:323-338)
    def split_word(w):
        cs = list(w)
        if continuing_subword_prefix is not None and len(cs) > 1:
            return [cs[0]] + [continuing_subword_prefix + c for c in cs[1:]]
        return cs

    words = [(split_word(w), f) for w, f in word_freqs.items()]
    merges = []
    # Step 4 (:141-183)
    while len(vocab) < vocab_size:
        pf = {}
        for s, f in words:
            if len(s) < 2 or f == 0:
                continue
            for a, b in zip(s, s[1:]):
                pf[(a, b)] = (pf.get((a, b), 0) + f) & M32
        if not pf:
            break
        best = min(pf.items(), key=lambda kv: (-kv[1], index[kv[0][0]], index[kv[0][1]]))
        (a, b), freq = best
        if freq < min_frequency:
            break
        merged = a + b
        vocab[merged] = len(vocab)                               # :168-169 (len does not grow if `merged` exists)
        sym(merged)
        merges.append((a, b))
        new_words = []
        for s, f in words:                                       # :379-401
            r, i = [], 0
            while i < len(s):
                if i < len(s) - 1 and s[i] == a and s[i + 1] == b:
                    r.append(merged)
                    i += 2
                else:
                    r.append(s[i])
                    i += 1
            new_words.append((r, f))
        words = new_words
    if return_symbols:
        return vocab, merges, index
    return vocab, merges
