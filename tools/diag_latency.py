"""Latency of small calls through the Python shim (the reference README's usage shape: one short text per call)."""
import sys, time
sys.path.insert(0,'complexity-tokenizer_b200'); sys.path.insert(0,'fixtures'); sys.path.insert(0,'oracle')
import numpy as np
import complexity_tokenizer as ct, synth, c_oracle
tok=ct.Tokenizer.from_file(synth.tokenizer_config1()); orc=c_oracle.COracle.from_file(synth.tokenizer_config1())
def bench(label, fn, n=300):
    fn(); fn()
    t=time.perf_counter()
    for _ in range(n): fn()
    dt=(time.perf_counter()-t)/n
    print('%-52s %8.1f us/call'%(label, dt*1e6))
short='Hello, world! This is a test.'
para=' '.join(['The quick brown fox jumps over the lazy dog.']*40)
ids=tok.encode(para)
bench('encode(29-byte text)', lambda: tok.encode(short))
bench('encode(1.8 KB text)', lambda: tok.encode(para))
bench('encode_batch(100 x 29 bytes)', lambda: tok.encode_batch([short]*100))
bench('decode(400 ids)', lambda: tok.decode(ids))
bench('oracle C core: encode(29-byte text)', lambda: orc.encode_batch([short]))
tok.profile_enable(True)
for _ in range(50): tok.encode(short)
print({k:round(v[0]/v[1]*1e3,1) for k,v in tok.profile_report().items()}, '(us per kernel, device time)')
