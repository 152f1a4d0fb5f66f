"""Timing of the rich-`Encoding` path on the config-2 corpus, through the C ABI with host buffers:
  (a) tok(texts, add_special_tokens=False, padding='max_length', truncation=True, max_length=L)  -> dense ids + masks
  (b) the same with add_special_tokens=True (encode_to_encoding ids + post-processor + special mask)
  (c) encode_batch_to_encoding with offsets + word ids
Per-kernel device times come from the library's CUDA-event marks (ctk_profile_enable).
usage: python tools/diag_encoding.py [MiB=256] [max_length=1024]"""
import os, sys, time
sys.path.insert(0, 'complexity-tokenizer_b200'); sys.path.insert(0, 'fixtures'); sys.path.insert(0, 'oracle')
import numpy as np
import complexity_tokenizer as ct, synth

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
tok = ct.Tokenizer.from_file(synth.tokenizer_config2())
text, offs = synth.gen_corpus('ascii', 2002, mib << 20, doc_median=4096)
B, D = text.size, len(offs) - 1
print('corpus: %.1f MiB, %d docs' % (B / 2**20, D))


def run(name, reps=int(os.environ.get('DIAG_REPS', '3')), **kw):
    tok._encode_rows(text, offs, **kw)                       # warm-up (workspace growth, pinned pool)
    tok.profile_enable(True)
    t = time.perf_counter()
    for _ in range(reps):
        p = tok._encode_rows(text, offs, **kw)
    dt = (time.perf_counter() - t) / reps
    r = tok.profile_report()
    tok.profile_enable(False)
    kern = {k: round(v[0] / v[1], 3) for k, v in r.items()}
    dev = sum(v[0] / v[1] for k, v in r.items())
    print('%s: %.1f ms per call end to end (%.2f GB/s of text; includes the NumPy copies of the result), %.2f ms in marked kernels | rows %d, out %d | %s'
          % (name, dt * 1e3, B / dt / 1e9, dev, p.n_rows, p.input_ids.size, kern))
    return p


run('(a) ids+masks, padded/truncated to %d, add_special_tokens=False' % L, add_special_tokens=False, truncation=True, max_length=L, padding=2, pad_to=L)
run('(b) same, add_special_tokens=True', add_special_tokens=True, truncation=True, max_length=L, padding=2, pad_to=L)
p = run('(c) encode_batch_to_encoding + offsets + word ids', add_special_tokens=True, want_offsets=True)
# spot check against the oracle on a few documents
import json, py_oracle, py_encoding as pe
tj = json.load(open(synth.tokenizer_config2(), encoding='utf-8'))
orc = pe.RichOracle(py_oracle.OracleTokenizer(tj), tj)
raw = text.tobytes()
bad = 0
for d in range(0, D, max(1, D // 40)):
    w = orc.encode_to_encoding(raw[int(offs[d]):int(offs[d + 1])].decode())
    g = p.encoding(d)
    bad += not (g.ids == w.ids and g.offsets == [tuple(o) for o in w.offsets] and g.word_ids == w.word_ids)
print('spot check against the oracle: %d mismatching documents of ~40' % bad)
