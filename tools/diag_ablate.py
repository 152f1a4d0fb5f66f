import sys, os
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
tok.set_cache_persistent(True)
for _ in range(3): tok.encode_device(d_text.data_ptr(),d_off.data_ptr(),D,text.size,d_ids.data_ptr(),B+D+16,d_ioff.data_ptr())
tok.profile_enable(True)
for _ in range(5): tok.encode_device(d_text.data_ptr(),d_off.data_ptr(),D,text.size,d_ids.data_ptr(),B+D+16,d_ioff.data_ptr())
r=tok.profile_report()
print('ABLATE',os.environ.get('CTK_ABLATE','0'),'warm cache, 256 MiB:', {k:round(v[0]/v[1],3) for k,v in r.items()})
