"""BPE training on the GPU: timings of the two stages (word histogram, merge loop) on config-1-shaped input.

    python tools/diag_train.py [--mib 256] [--vocab 32000] [--oracle-merges 300]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('complexity-tokenizer_b200', 'oracle', 'fixtures'):
    sys.path.insert(0, os.path.join(ROOT, p))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mib', type=int, default=256)
    ap.add_argument('--vocab', type=int, default=32000)
    ap.add_argument('--oracle-merges', type=int, default=300)
    ap.add_argument('--skip-config1', action='store_true')
    a = ap.parse_args()
    import complexity_tokenizer as ct
    import synth
    out = {}
    # (1) config 1: the trainer's own sample, first 2000 documents of the English corpus
    text, offs = synth.gen_corpus('english', 1001, 12 << 20)
    offs = offs[:2001]
    text = text[:int(offs[-1])]
    tr = ct.BpeTrainer(vocab_size=a.vocab, min_frequency=2, show_progress=False)
    tr.train_packed(text[:4096], np.array([0, 4096], dtype=np.uint64))          # warm-up (context, module load)
    t0 = time.perf_counter()
    vocab, merges = tr.train_packed(text, offs) if not a.skip_config1 else ({}, [])
    wall = time.perf_counter() - t0
    s = tr.last_stats
    out['config1_sample'] = None if a.skip_config1 else dict(s, wall_s=wall, vocab=len(vocab), merges=len(merges),
                                 us_per_merge=1e3 * s['ms_merges'] / max(1, len(merges)),
                                 words_GBps=s['n_bytes'] / max(s['ms_words'], 1e-9) / 1e6)
    # (2) word histogram at scale: a few merges only
    if a.mib:
        big, boffs = synth.gen_corpus('english', 2002, a.mib << 20)
        tr2 = ct.BpeTrainer(vocab_size=400, min_frequency=2, show_progress=False)
        t0 = time.perf_counter()
        tr2.train_packed(big, boffs)
        wall = time.perf_counter() - t0
        s = tr2.last_stats
        out['word_histogram_%dMiB' % a.mib] = dict(s, wall_s=wall, words_GBps=s['n_bytes'] / max(s['ms_words'], 1e-9) / 1e6)
    # (3) CPU restatement (Python, strings as the reference has them) on a bounded sample
    if a.oracle_merges:
        import py_trainer
        n = 400
        raw = text.tobytes()
        docs = [raw[int(offs[i]):int(offs[i + 1])].decode() for i in range(n)]
        vs = 4 + 80 + a.oracle_merges
        t0 = time.perf_counter()
        want = py_trainer.train_bpe(docs, vocab_size=vs, min_frequency=2)
        cpu = time.perf_counter() - t0
        tr3 = ct.BpeTrainer(vocab_size=vs, min_frequency=2, show_progress=False)
        t0 = time.perf_counter()
        got = tr3.train(docs)
        gpu = time.perf_counter() - t0
        out['oracle_sample'] = dict(docs=n, merges=len(want[1]), equal=(got == want), oracle_s=cpu, gpu_wall_s=gpu,
                                    gpu_ms_merges=tr3.last_stats['ms_merges'], kind='port (Python restatement, 1 core)')
    print(json.dumps(out))


if __name__ == '__main__':
    main()
