import sys, os, ctypes
os.environ['CTK_ABLATE']='9'
sys.path.insert(0,'complexity-tokenizer_b200'); sys.path.insert(0,'fixtures')
import numpy as np, torch
import complexity_tokenizer as ct, synth
tok=ct.Tokenizer.from_file(synth.tokenizer_config2())
B=int(sys.argv[1]) if len(sys.argv)>1 else 256<<20
t=torch.empty(B+64,dtype=torch.uint8).numpy()
text,offs=synth.gen_corpus('ascii',5000,B,doc_median=4096,doc_min=256,doc_max=65536,out=t)
D=len(offs)-1
d_text=torch.from_numpy(t).cuda(); d_off=torch.from_numpy(offs.astype(np.int64)).cuda()
d_ids=torch.empty(B+D+16,dtype=torch.int32,device='cuda'); d_ioff=torch.empty(D+1,dtype=torch.int64,device='cuda')
lib=ct._lib(); lib.ctk_debug_counters.argtypes=[ctypes.c_void_p, ctypes.c_void_p]
names=['long>32','found@probe0(ovf/busy)','found@later probe','bpe miss<=16 (inserted)','bpe 17..32','bpe miss no-slot','?','?','pretokens']
def run(label):
    tok.encode_device(d_text.data_ptr(),d_off.data_ptr(),D,text.size,d_ids.data_ptr(),B+D+16,d_ioff.data_ptr())
    c=np.zeros(16,dtype=np.uint32); lib.ctk_debug_counters(tok._h,c.ctypes.data)
    print(label, {n:int(v) for n,v in zip(names,c[:9]) if n!='?'})
run('cold')
tok.set_cache_persistent(True)
run('2nd (cache kept)'); run('3rd')
