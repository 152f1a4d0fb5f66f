#!/usr/bin/env python3
"""Generate tests/golden/hf_crosscheck_wide.json: >= 2000 texts x several pipelines, ids from HuggingFace `tokenizers`.

Complements tools/make_golden.py (one tokenizer, 222 texts).  Pipelines, each over its own small tokenizer (3 000 merges
trained on english / ascii / mixed text), set up so that HF's semantics coincide with the reference's (no added tokens: HF
extracts them from raw text, the reference searches inside words):
    nfc            NFC + Sequence[Split(reference pattern, Isolated), ByteLevel(use_regex=False)]
    null           "normalizer": null -- which the reference reads as its default, NFC (parsing.rs:89)
    prefix         ByteLevel{add_prefix_space: true}: the reference prepends one space to a non-empty text that does not start
                   with one (pretokenizers.rs:163-167); HF's flag prefixes every split piece instead, so the space is put into
                   the text handed to HF
    split_*        an extra Split stage in front (src/pretokenizers.rs:298-433): the reference applies the ByteLevel pattern
                   to each piece, i.e. Sequence[Split(user pattern, behaviour), Split(reference pattern, Isolated), ByteLevel].
                   Removed keeps the MATCHES in the reference unless `invert` (:313-331) -- HF's `removed` with the opposite flag.
                   MergedWithPrevious differs from HF on adjacent matches (the reference glues them, :362-366): those texts drop out.
A case is only written where HF agrees with the project's oracle at generation time (printed otherwise) -- the file pins the
oracle, it cannot prove anything about behaviours the two implement differently.
Run from the repo root:  python tools/make_golden_wide.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'fixtures'), os.path.join(ROOT, 'oracle')]
import numpy as np  # noqa: E402
import synth  # noqa: E402
import py_oracle  # noqa: E402

PATTERN = r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+"    # src/pretokenizers.rs:13
HF_BEH = {'Isolated': 'isolated', 'Removed': 'removed', 'MergedWithPrevious': 'merged_with_previous', 'MergedWithNext': 'merged_with_next',
          'Contiguous': 'contiguous'}


def hf_tokenizer(tj, split):
    from tokenizers import Regex, Tokenizer, models, normalizers, pre_tokenizers
    merges = [tuple(m.split(' ')) for m in tj['model']['merges']]
    tk = Tokenizer(models.BPE(vocab=tj['model']['vocab'], merges=merges))
    tk.normalizer = normalizers.NFC()
    seq = []
    if split:
        seq.append(pre_tokenizers.Split(Regex(split[0]), HF_BEH[split[1]], invert=(not split[2]) if split[1] == 'Removed' else split[2]))
    seq += [pre_tokenizers.Split(Regex(PATTERN), 'isolated'), pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)]
    tk.pre_tokenizer = pre_tokenizers.Sequence(seq)
    return tk


def main():
    import tokenizers
    toks = {}
    for name, kind, seed in (('english', 'english', 5101), ('ascii', 'ascii', 5102), ('mixed', 'mixed', 5103)):
        text, offs = synth.gen_corpus(kind, seed, 3 << 20, doc_median=2048, doc_min=64, doc_max=8192)
        toks[name] = synth.assemble_tokenizer(synth.train_merges(text, 3000), specials_first=('<unk>', '<pad>', '<s>', '</s>'))
    texts = {}
    for name, kind, seed in (('english', 'english', 5201), ('ascii', 'ascii', 5202), ('mixed', 'mixed', 5203)):
        t, o = synth.gen_corpus(kind, seed, 150 << 10, doc_median=160, doc_min=8, doc_max=1500)
        texts[name] = [d.decode('utf-8') for d in synth.split_docs(t, o)][:900]
    rng = np.random.default_rng(5300)
    pool = list("abcdefghij XYZ 0123456789 \t\n.,!?'-_()[]{}\"é à ü ñ 中文日本語かなカナ 한국어 Ωжд ٣५ ½ǅ € ") + ['　', '́', '̀', '‍', '\U0001F600', '\U0001F44D', '\U0001F3FD', ' ', "'s", "'ll", "n't"]
    fuzz = [''.join(pool[int(k)] for k in rng.integers(0, len(pool), size=int(rng.integers(0, 60)))) for _ in range(400)]
    cases = []
    plan = [('nfc', 'mixed', True, False, None), ('nfc', 'english', True, False, None), ('null', 'ascii', False, False, None), ('prefix', 'english', True, True, None),
            ('prefix', 'mixed', True, True, None),
            ('split_digit_isolated', 'ascii', False, False, (r'\d', 'Isolated', False)),
            ('split_num3_isolated', 'mixed', True, False, (r'\p{N}{1,3}', 'Isolated', False)),
            ('split_cjk_isolated', 'mixed', True, False, (r'[一-龥぀-ゟ゠-ヿ]+', 'Isolated', False)),
            ('split_ws_merged_next', 'english', True, False, (r'\s+', 'MergedWithNext', False)),
            ('split_ws_merged_prev', 'english', True, False, (r'\s', 'MergedWithPrevious', False)),
            ('split_punct_contiguous', 'ascii', False, False, (r'[^\s\p{L}\p{N}]', 'Contiguous', False)),
            ('split_ws_removed_inverted', 'english', True, False, (r'\s+', 'Removed', True)),
            ('split_words_removed', 'mixed', True, False, (r'\w+', 'Removed', False))]
    n_texts = 0
    for name, tname, nfc, aps, split in plan:
        tj = json.loads(json.dumps(toks[tname]))
        tj['normalizer'] = {'type': 'NFC'} if nfc else None
        seq = []
        if split:
            seq.append({'type': 'Split', 'pattern': {'Regex': split[0]}, 'behavior': split[1], 'invert': split[2]})
        seq.append({'type': 'ByteLevel', 'add_prefix_space': aps, 'use_regex': False, 'trim_offsets': True})
        tj['pre_tokenizer'] = {'type': 'Sequence', 'pretokenizers': seq} if split else seq[0]
        tx = (texts[tname][:260] if split else texts[tname][:600]) + fuzz[:120 if split else 400]
        hf = hf_tokenizer(tj, split)
        ids = [e.ids for e in hf.encode_batch([(' ' + t if aps and t and not t.startswith(' ') else t) for t in tx], add_special_tokens=False)]
        orc = py_oracle.OracleTokenizer(tj)
        keep = [i for i, t in enumerate(tx) if orc.encode(t) == ids[i]]
        if len(keep) != len(tx):
            print('case %s: HF and the oracle differ on %d of %d texts (dropped), e.g. %r' % (name, len(tx) - len(keep), len(tx), tx[[i for i in range(len(tx)) if i not in set(keep)][0]][:50]))
        cases.append({'name': name + ':' + tname, 'tokenizer': tname, 'normalizer': tj['normalizer'], 'pre_tokenizer': tj['pre_tokenizer'],
                      'texts': [tx[i] for i in keep], 'ids': [ids[i] for i in keep], 'dropped': len(tx) - len(keep)})
        n_texts += len(keep)
    out = {'generator': 'tools/make_golden_wide.py', 'hf_tokenizers_version': tokenizers.__version__,
           'note': 'ids produced by HuggingFace tokenizers configured to coincide with the reference semantics; cases where HF and the oracle differ are dropped and counted',
           'tokenizers': toks, 'cases': cases}
    path = os.path.join(ROOT, 'tests', 'golden', 'hf_crosscheck_wide.json')
    with open(path, 'w', encoding='utf-8') as f:
        json.dump(out, f, ensure_ascii=False, separators=(',', ':'))
    print(path, os.path.getsize(path), 'bytes;', n_texts, 'texts in', len(cases), 'cases')


if __name__ == '__main__':
    main()
