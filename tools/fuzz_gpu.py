#!/usr/bin/env python3
"""Time-boxed differential fuzz of the CUDA path against the oracle (not part of the test-suite: a bug hunt).
   python tools/fuzz_gpu.py [minutes]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('complexity-tokenizer_b200', 'oracle', 'fixtures'): sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import complexity_tokenizer as ct, c_oracle, synth
budget = float(sys.argv[1]) * 60 if len(sys.argv) > 1 else 300
PIECES = ["a", "b", "e", "s", "t", "r", "v", "l", "m", "d", "'", " ", " ", " ", "\n", "\t", "1", "9", ".", ",", "!", "-", "é", "ü", "ñ", "中", "文", "あ", "　",
          " ", "\U0001F600", "\U0001F44D", "́", "̧", "x", "'s", "'ll", "  ", "Ⅷ", "²", "_", "٣", "ß", "'re", "'ve", "'d", "'m", "'t", "Å", "豈",
          "각", "ᄀ", "ᅡ", "ᆨ", "the", " the", " of", "ing", "tion", "\r\n", "\x00", "=-", "<s>", "</s>", "<pad>", "<|endoftext|>", "Ġ", "Ċ"]
cfgs = {'config1': synth.tokenizer_config1(), 'config2': synth.tokenizer_config2(), 'config3': synth.tokenizer_config3()}
toks = {k: (ct.Tokenizer.from_file(v), c_oracle.COracle.from_file(v)) for k, v in cfgs.items()}
rng = np.random.default_rng(int(time.time()) & 0xFFFF)
print('seed', rng.bit_generator.state['state']['state'] & 0xFFFF)
def rand_doc():
    kind = rng.integers(0, 10)
    if kind == 0: return ''
    if kind == 1:                                          # long runs of one class
        ch = PIECES[int(rng.integers(0, len(PIECES)))]
        return ch * int(rng.integers(1, 700)) + (' tail' if rng.random() < .5 else '')
    if kind == 2:                                          # length near the slice / chunk geometry
        n = int(rng.choice([15, 16, 17, 31, 32, 33, 431, 447, 448, 449, 463, 464, 479, 480, 495, 496, 497, 511, 512, 513, 895, 896, 897])) + int(rng.integers(-2, 3))
        return ''.join(PIECES[int(i)] for i in rng.integers(0, 24, size=max(n, 0)))[:max(n, 0)]
    if kind == 3:                                          # CJK-like run lengths
        han = [chr(0x4E00 + int(i)) for i in rng.integers(0, 3000, size=int(rng.integers(1, 60)))]
        return ''.join(han) + ('。' if rng.random() < .5 else ' ') + ''.join(PIECES[int(i)] for i in rng.integers(0, len(PIECES), size=int(rng.integers(0, 20))))
    k = int(rng.integers(0, 120 if kind < 8 else 2500))
    return ''.join(PIECES[int(i)] for i in rng.integers(0, len(PIECES), size=k))
t0 = time.time(); it = 0; bad = 0; nbytes = 0
while time.time() - t0 < budget:
    cfg = ['config1', 'config2', 'config3'][it % 3]
    tok, orc = toks[cfg]
    docs = [rand_doc() for _ in range(int(rng.integers(1, 400)))]
    nbytes += sum(len(d) for d in docs)
    got, want = tok.encode_batch(docs), orc.encode_batch(docs)
    if got != want:
        bad += 1
        i = next(i for i, (a, b) in enumerate(zip(got, want)) if a != b)
        print('ENCODE MISMATCH', cfg, 'doc', i, json.dumps(docs[i])[:300]); sys.stdout.flush()
        with open(os.path.join(ROOT, 'gpurun_out', 'fuzz_fail_%d.json' % bad), 'w') as f: json.dump({'cfg': cfg, 'doc': docs[i], 'got': got[i], 'want': want[i]}, f)
    for opts in ((False, True), (False, False), (True, True)):
        if tok.decode_batch_with_options(want, *opts) != orc.decode_batch(want, *opts):
            bad += 1; print('DECODE MISMATCH', cfg, opts); sys.stdout.flush()
    # ids shuffled / corrupted: decode robustness (unknown ids, split multi-byte tokens)
    junk = [[int(x) for x in rng.integers(0, 120000, size=int(rng.integers(0, 40)))] for _ in range(50)]
    for opts in ((False, True), (True, False)):
        if tok.decode_batch_with_options(junk, *opts) != orc.decode_batch(junk, *opts):
            bad += 1; print('DECODE(junk) MISMATCH', cfg, opts); sys.stdout.flush()
    it += 1
print('fuzz: %d batches, %.1f MB of text, %d mismatches, %.0f s' % (it, nbytes / 1e6, bad, time.time() - t0))
