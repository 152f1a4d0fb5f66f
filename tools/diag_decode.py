import sys, time
sys.path.insert(0,'complexity-tokenizer_b200'); sys.path.insert(0,'fixtures')
import numpy as np, torch
import complexity_tokenizer as ct, synth
tok=ct.Tokenizer.from_file(synth.tokenizer_config2())
B=256<<20
text,offs=synth.gen_corpus('ascii',5000,B,doc_median=4096,doc_min=256,doc_max=65536)
ids,ioff=tok.encode_packed(text,offs)
T=ids.size; D=len(offs)-1
d_ids=torch.from_numpy(ids.view(np.int32)).cuda(); d_ioff=torch.from_numpy(ioff.astype(np.int64)).cuda()
d_out=torch.empty(B+1024,dtype=torch.uint8,device='cuda'); d_ooff=torch.empty(D+1,dtype=torch.int64,device='cuda')
for clean in (False, True):
    tok.profile_enable(True)
    for _ in range(3):
        n=tok.decode_device(d_ids.data_ptr(),d_ioff.data_ptr(),D,T,d_out.data_ptr(),B+1024,d_ooff.data_ptr(),False,clean)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(3): n=tok.decode_device(d_ids.data_ptr(),d_ioff.data_ptr(),D,T,d_out.data_ptr(),B+1024,d_ooff.data_ptr(),False,clean)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3
    print("decode clean=%s: %d ids -> %d bytes, %.2f ms = %.1f GB/s out"%(clean,T,n,dt*1e3,n/dt/1e9), {k:round(v[0]/v[1],3) for k,v in tok.profile_report().items()})
