import sys, os, time
sys.path.insert(0,'complexity-tokenizer_b200'); sys.path.insert(0,'fixtures')
import numpy as np, torch
import complexity_tokenizer as ct, synth
tok=ct.Tokenizer.from_file(synth.tokenizer_config2())
B=256<<20
t=torch.empty(B+64,dtype=torch.uint8).numpy()
text,offs=synth.gen_corpus('ascii',5000,B,doc_median=4096,doc_min=256,doc_max=65536,out=t)
D=len(offs)-1
d_text=torch.from_numpy(t).cuda(); d_off=torch.from_numpy(offs.astype(np.int64)).cuda()
d_ids=torch.empty(B+D+16,dtype=torch.int32,device='cuda'); d_ioff=torch.empty(D+1,dtype=torch.int64,device='cuda')
def run(n=5):
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): T=tok.encode_device(d_text.data_ptr(),d_off.data_ptr(),D,text.size,d_ids.data_ptr(),B+D+16,d_ioff.data_ptr())
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n, T
for _ in range(3): run(1)
print('cold cache per call: %.3f ms'%run()[0])
tok.set_cache_persistent(True)
run(1)
ms,T=run()
print('warm cache: %.3f ms  tokens %d'%(ms,T))
