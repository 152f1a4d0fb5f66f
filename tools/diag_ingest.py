"""BASELINE config 1 (10 000 English texts, 32K vocab) through the three host entry points of the shim:
list[str] -> list[list[int]] (the reference's signature), Arrow -> Arrow, packed numpy -> packed numpy."""
import sys, time
sys.path.insert(0,'complexity-tokenizer_b200'); sys.path.insert(0,'fixtures')
import numpy as np, pyarrow as pa
import complexity_tokenizer as ct, synth
tok=ct.Tokenizer.from_file(synth.tokenizer_config1())
text,offs=synth.gen_corpus('english',1001,12<<20,doc_median=1024,doc_min=64,doc_max=16384)
docs=[d.decode() for d in synth.split_docs(text,offs)]
arr=pa.array(docs)
B=text.size
def bench(label, fn, n=10):
    fn(); t=time.perf_counter()
    for _ in range(n): r=fn()
    dt=(time.perf_counter()-t)/n
    print('%-58s %8.2f ms  %8.1f MB/s'%(label, dt*1e3, B/dt/1e6)); return r
print(len(docs),'documents,',B,'bytes')
bench('encode_batch(list[str]) -> list[list[int]]', lambda: tok.encode_batch(docs))
bench('encode_arrow(StringArray) -> LargeListArray<uint32>', lambda: tok.encode_arrow(arr))
bench('encode_packed(numpy) -> numpy', lambda: tok.encode_packed(text, offs))
