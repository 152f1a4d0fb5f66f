#!/usr/bin/env python3
"""Generate tests/golden/*.json -- small committed vectors that pin the oracle.

The reference (Rust) cannot be built or imported in this image, so golden vectors come from
  (1) the reference's own unit tests (known answers, cited in tests/test_oracle_known_answers.py), and
  (2) HuggingFace `tokenizers` (installed here, an independent rayon/Rust BPE implementation the
      reference's README compares itself with), configured so that its semantics coincide with the
      reference's: NFC + Sequence[Split(<reference pattern>, Isolated), ByteLevel(use_regex=False)] + BPE,
      no added tokens (HF extracts added tokens from raw text; the reference searches inside words).
This script produces (2): a small tokenizer.json, a list of texts and HF's ids for them.
Run from the repo root:  python tools/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fixtures'))
import synth  # noqa: E402

PATTERN = r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+"    # src/pretokenizers.rs:13


def hf_tokenizer(tj):
    from tokenizers import Regex, Tokenizer, models, normalizers, pre_tokenizers
    merges = [tuple(m.split(' ')) for m in tj['model']['merges']]
    tk = Tokenizer(models.BPE(vocab=tj['model']['vocab'], merges=merges))
    tk.normalizer = normalizers.NFC()
    tk.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(PATTERN), 'isolated'),
                                                pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    return tk


def main():
    import tokenizers
    text, offs = synth.gen_corpus('mixed', 4242, 3 << 20, doc_median=2048, doc_min=64, doc_max=8192)
    pairs = synth.train_merges(text, 3000)
    tj = synth.assemble_tokenizer(pairs, specials_first=('<unk>', '<pad>', '<s>', '</s>'))
    en, eoffs = synth.gen_corpus('english', 4243, 24 << 10, doc_median=300, doc_min=16, doc_max=2048)
    mx, moffs = synth.gen_corpus('mixed', 4244, 40 << 10, doc_median=300, doc_min=16, doc_max=2048)
    texts = [d.decode('utf-8') for d in synth.split_docs(en, eoffs)] + [d.decode('utf-8') for d in synth.split_docs(mx, moffs)]
    texts += ["", " ", "Hello, world!", "don't", "'sup", "!'s", "a  b", "a b", "x ", " 's", "  's", "I'll've been",
              "été été", "Å 豈 각", "tab\tnew\n\nline  two   spaces",
              "<s>hi</s> <pad> <unk>", "　full　width　", "1234 5.67 8,900 x2 3rd", "\U0001F600\U0001F44D\U0001F3FD ok"]
    # long pre-tokens: the mid (33..128 B, one lane each), warp-round (129..256 B) and grid-round (> 256 B) device paths
    import numpy as np
    rng = np.random.default_rng(4245)
    han = [chr(0x4E00 + int(i)) for i in rng.integers(0, 3500 * 5, size=400)]
    letters = 'etaoinshrdlucmfwypvbgkqjxz'
    def run(alpha, n):
        return ''.join(alpha[int(i)] for i in rng.integers(0, len(alpha), size=n))
    for n in (12, 17, 22, 33, 40, 43, 64, 86, 100, 129, 200, 257, 300, 1000, 5000):
        texts.append('start ' + run(letters, n) + ' end.')
        texts.append(run(han, max(4, n // 3)) + '。' + run(han, 5))
    texts += ['x' * 700, 'ab' * 400, ' ' * 300 + 'w', 'w' + ' ' * 300, '=' * 257, '-=' * 200, '9' * 129, 'é' * 150,
              ' '.join(run(letters, int(k)) for k in rng.integers(30, 140, size=12)),
              run(han, 43) + 'mixed' + run(han, 30) + ' tail ' + '\U0001F600' * 40]
    tk = hf_tokenizer(tj)
    ids = [e.ids for e in tk.encode_batch(texts, add_special_tokens=False)]
    out = {'generator': 'tools/make_golden.py', 'hf_tokenizers_version': tokenizers.__version__,
           'note': 'ids produced by HuggingFace tokenizers configured to coincide with the reference semantics',
           'tokenizer': tj, 'texts': texts, 'ids': ids}
    os.makedirs(os.path.join(ROOT, 'tests', 'golden'), exist_ok=True)
    path = os.path.join(ROOT, 'tests', 'golden', 'hf_crosscheck.json')
    with open(path, 'w', encoding='utf-8') as f:
        json.dump(out, f, ensure_ascii=False)
    print(path, os.path.getsize(path), 'bytes;', len(texts), 'texts;', sum(map(len, ids)), 'ids')


if __name__ == '__main__':
    main()
