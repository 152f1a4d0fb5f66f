"""Host-buffer encode of 1 GiB from pageable vs page-locked memory (what a NumPy / Arrow caller hands over)."""
import ctypes, sys, time
sys.path.insert(0,'complexity-tokenizer_b200'); sys.path.insert(0,'fixtures'); sys.path.insert(0,'.')
sys.argv=[sys.argv[0]]
import numpy as np, torch
import bench, complexity_tokenizer as ct, synth
tok=ct.Tokenizer.from_file(synth.tokenizer_config2())
h_text,B,offs=bench.make_corpus(1<<30,5000,pinned=True)
pinned=h_text.numpy()[:B]
pageable=np.array(pinned, copy=True)
lib=ct._lib(); D=len(offs)-1
def step(a):
    res=ctypes.c_void_p(); rc=lib.ctk_encode_batch(tok._h,a.ctypes.data,offs.ctypes.data,D,ctypes.byref(res)); assert rc==0; lib.ctk_result_free(res)
for name,a in (('pinned',pinned),('pageable',pageable)):
    step(a); step(a); ts=[]
    for _ in range(3):
        t=time.perf_counter(); step(a); ts.append(time.perf_counter()-t)
    print('%-9s %.1f ms  %.1f GB/s'%(name,min(ts)*1e3,B/min(ts)/1e9))
