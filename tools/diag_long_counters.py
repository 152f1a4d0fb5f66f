"""Round statistics of the warp-level round-parallel merge path on config 3 (CTK_ABLATE=9 debug counters)."""
import sys, os, ctypes
os.environ['CTK_ABLATE']='9'
sys.path.insert(0,'complexity-tokenizer_b200'); sys.path.insert(0,'fixtures')
import numpy as np, torch
import complexity_tokenizer as ct, synth
tok=ct.Tokenizer.from_file(synth.tokenizer_config3())
text,offs=synth.gen_corpus('mixed',3003,64<<20,doc_median=4096,doc_min=256,doc_max=65536)
B=text.size; D=len(offs)-1
buf=np.zeros(B+64,dtype=np.uint8); buf[:B]=text
d_text=torch.from_numpy(buf).cuda(); d_off=torch.from_numpy(offs.astype(np.int64)).cuda()
d_ids=torch.empty(2*B+D+16,dtype=torch.int32,device='cuda'); d_ioff=torch.empty(D+1,dtype=torch.int64,device='cuda')
lib=ct._lib(); lib.ctk_debug_counters.argtypes=[ctypes.c_void_p, ctypes.c_void_p]
tok.encode_device(d_text.data_ptr(),d_off.data_ptr(),D,B,d_ids.data_ptr(),2*B+D+16,d_ioff.data_ptr())
c=np.zeros(16,dtype=np.uint32); lib.ctk_debug_counters(tok._h,c.ctypes.data)
print('counters[16..31]:', c.tolist())
n,r,ch,tr=[int(x) for x in c[10:14]]
print('long pre-tokens via rounds: %d; rounds/pre-token %.1f; chunk-rounds/pre-token %.1f; max-lane loop trips/pre-token %.1f'%(n, r/max(n,1), ch/max(n,1), tr/max(n,1)))
