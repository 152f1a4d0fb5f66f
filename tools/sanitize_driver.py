#!/usr/bin/env python3
"""Small run of every device path, for compute-sanitizer (tools/sanitize.sh).  Results are checked against the oracle
so that a sanitizer-clean run is also a correct one."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('complexity-tokenizer_b200', 'oracle', 'fixtures'):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402
import c_oracle  # noqa: E402
import complexity_tokenizer as ct  # noqa: E402
import synth  # noqa: E402

SMALL = int(os.environ.get('CTK_SANITIZE_KIB', '192')) << 10


def check(name, tok, orc, text, offs):
    ids, ioff = tok.encode_packed(text, offs)
    want, woff = orc.encode_packed(text, offs)
    ok = np.array_equal(ioff, woff) and np.array_equal(ids, want)
    for skip, clean in ((False, False), (False, True)):
        b, boff = tok.decode_packed(ids, ioff, skip, clean)
        wb, wboff = orc.decode_packed(want, woff, skip, clean)
        ok = ok and np.array_equal(boff, wboff) and np.array_equal(b, wb)
    print('driver: %-28s %8d bytes %6d docs %8d ids  %s' % (name, text.size, len(offs) - 1, ids.size, 'ok' if ok else 'MISMATCH'), flush=True)
    return ok


def main():
    ok = True
    p2, p3 = synth.tokenizer_config2(), synth.tokenizer_config3()
    tok2, orc2 = ct.Tokenizer.from_file(p2), c_oracle.COracle.from_file(p2)
    tok3, orc3 = ct.Tokenizer.from_file(p3), c_oracle.COracle.from_file(p3)
    t, o = synth.gen_corpus('ascii', 11, SMALL, doc_median=700, doc_min=16, doc_max=4096)
    ok &= check('config2 ascii', tok2, orc2, t, o)
    t, o = synth.gen_corpus('mixed', 12, SMALL, doc_median=700, doc_min=16, doc_max=4096)
    ok &= check('config3 mixed (NFC, CJK runs)', tok3, orc3, t, o)
    docs = synth.gen_long_docs(doc_bytes=12 << 10, n_docs=4)
    t, o = synth.pack(docs)
    ok &= check('long pre-tokens', tok2, orc2, t, o)
    # added tokens that match inside words
    small_text, so = synth.gen_corpus('english', 77, 200 << 10)
    pairs = synth.train_merges(small_text, 300)
    tj = synth.assemble_tokenizer(pairs, specials_first=('<unk>',))
    nid = max(tj['model']['vocab'].values()) + 1
    for k, (content, flags) in enumerate([('the', {}), ('Ġand', {'single_word': True}), ('ing', {'rstrip': True}), ('42', {})]):
        t_ = {'id': nid + k, 'content': content, 'special': False, 'single_word': False, 'lstrip': False, 'rstrip': False, 'normalized': False}
        t_.update(flags)
        tj['added_tokens'].append(t_)
    js = json.dumps(tj, ensure_ascii=False)
    try:
        tka, orca = ct.Tokenizer.from_str(js), c_oracle.COracle.from_str(js)
        t, o = synth.gen_corpus('english', 78, SMALL // 2, doc_median=300, doc_min=16, doc_max=2048)
        ok &= check('in-word added tokens', tka, orca, t, o)
    except Exception as ex:
        print('driver: added-token tokenizer skipped:', repr(ex))
    # many tiny documents (document starts inside every slice), empty ones included
    rng = np.random.default_rng(5)
    words = [bytes(rng.integers(97, 123, size=int(k), dtype=np.uint8)) for k in rng.integers(0, 13, size=4000)]
    t, o = synth.pack(words)
    ok &= check('tiny documents', tok2, orc2, t, o)
    # trainer
    import py_trainer
    docs = [d.decode() for d in synth.split_docs(*synth.gen_corpus('english', 9, 48 << 10, doc_median=300, doc_min=16, doc_max=2048))]
    got = ct.BpeTrainer(vocab_size=400, min_frequency=2, show_progress=False).train(docs)
    tr_ok = got == py_trainer.train_bpe(docs, vocab_size=400, min_frequency=2)
    print('driver: trainer %d merges %s' % (len(got[1]), 'ok' if tr_ok else 'MISMATCH'), flush=True)
    ok &= tr_ok
    print('driver: ALL', 'ok' if ok else 'MISMATCH', flush=True)
    sys.exit(0 if ok else 3)


if __name__ == '__main__':
    main()
