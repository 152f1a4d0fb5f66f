"""Times a Metaspace pipeline (csrc/metaspace.cu) on English-like text: python tools/diag_metaspace.py [MiB]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'complexity-tokenizer_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'fixtures')]
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sys.argv = ['bench.py']
import numpy as np, torch, bench, synth
import complexity_tokenizer as ct
text, offs = synth.gen_corpus('english', 1001, mib << 20)
raw = text.tobytes()
sample = raw[:1 << 20].decode('utf-8', 'ignore')
# a Metaspace BPE tokenizer trained on the device: BpeTrainer splits on white space, so hand it the Metaspace words of the sample
words_text = '\n'.join(('▁' + line.replace(' ', '▁')) for line in sample.split('\n') if line)
vocab, merges = ct.BpeTrainer(vocab_size=4000, min_frequency=2, special_tokens=['<unk>', '<s>', '</s>'], show_progress=False).train([words_text])
tj = {'model': {'type': 'BPE', 'vocab': vocab, 'merges': [a + ' ' + b for a, b in merges]}, 'pre_tokenizer': {'type': 'Metaspace'}, 'decoder': {'type': 'Metaspace'},
      'added_tokens': [{'id': vocab[s], 'content': s, 'special': True, 'single_word': False, 'lstrip': False, 'rstrip': False, 'normalized': False} for s in ('<unk>', '<s>', '</s>')]}
tok = ct.Tokenizer.from_str(json.dumps(tj, ensure_ascii=False), device=0)
dev = torch.device('cuda:0')
ms, T, kern, ids, ioff = bench.device_encode_ms(tok, torch, text, offs, dev=dev)
lines = raw.count(b'\n') + len(offs) - 1
print('metaspace encode %d MiB: %.3f ms = %.1f MB/s, %d tokens, ~%d words (lines) of ~%d bytes; vocab %d, merges %d; kernels %s'
      % (mib, ms, text.size / (ms * 1e-3) / 1e6, T, lines, text.size // max(1, lines), len(vocab), len(merges), {k: round(v, 3) for k, v in kern.items()}))
b, boff = tok.decode_packed(ids, ioff, False, False)
doc0 = raw[int(offs[0]):int(offs[1])].decode()
print('first document decodes to its text with the other white space dropped:', bytes(b[int(boff[0]):int(boff[1])]).decode() == ''.join(doc0.split('\n')).replace('\t', ''))
