"""Times the Split stage on the headline corpus (device-resident): python tools/diag_split.py [MiB] [case ...]"""
import sys, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'complexity-tokenizer_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'fixtures')]
args = sys.argv[1:]
sys.argv = ['bench.py']
import numpy as np, torch, bench, synth
import complexity_tokenizer as ct
mib = int(args[0]) if args else 256
want = args[1:]
dev = torch.device('cuda:0')
tjs = json.load(open(synth.tokenizer_config2()))
CASES = (('num3', r'\p{N}{1,3}', 'Isolated'), ('ws_next', r'\s+', 'MergedWithNext'), ('letters_removed', r'\p{L}+', 'Removed'),
         ('gpt2_isolated', r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+", 'Isolated'))
for name, rx, beh in CASES:
    if want and name not in want:
        continue
    tjs['pre_tokenizer'] = {'type': 'Sequence', 'pretokenizers': [{'type': 'Split', 'pattern': {'Regex': rx}, 'behavior': beh, 'invert': False}, {'type': 'ByteLevel', 'add_prefix_space': False}]}
    tok = ct.Tokenizer.from_str(json.dumps(tjs), device=0)
    ts, B, offs = bench.make_corpus(mib << 20, 2002, pinned=False)
    ms, T, kern, ids, ioff = bench.device_encode_ms(tok, torch, ts.numpy()[:B], offs, dev=dev)
    print(name, mib, 'MiB', 'ms', round(ms, 3), {k: round(v, 3) for k, v in kern.items() if v >= 0.05}, flush=True)
