"""Device-resident timing of the other BASELINE configs (3: mixed UTF-8 / 100K vocab, 4: long pre-tokens)."""
import sys, time
sys.path.insert(0,'complexity-tokenizer_b200'); sys.path.insert(0,'fixtures'); sys.path.insert(0,'oracle')
import numpy as np, torch
import complexity_tokenizer as ct, synth
def run(name, tokpath, text, offs, reps=3, check=None):
    tok=ct.Tokenizer.from_file(tokpath)
    B=text.size; D=len(offs)-1
    buf=np.zeros(B+64,dtype=np.uint8); buf[:B]=text
    d_text=torch.from_numpy(buf).cuda(); d_off=torch.from_numpy(offs.astype(np.int64)).cuda()
    cap=3*B+D+16
    d_ids=torch.empty(cap,dtype=torch.int32,device='cuda'); d_ioff=torch.empty(D+1,dtype=torch.int64,device='cuda')
    T=tok.encode_device(d_text.data_ptr(),d_off.data_ptr(),D,B,d_ids.data_ptr(),cap,d_ioff.data_ptr())
    tok.profile_enable(True)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(reps): T=tok.encode_device(d_text.data_ptr(),d_off.data_ptr(),D,B,d_ids.data_ptr(),cap,d_ioff.data_ptr())
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/reps
    r=tok.profile_report()
    print('   xlong rounds of the last call:', ct._lib().ctk_debug_xlong_rounds(tok._h))
    print('%s: %.1f MiB, %d docs, %d ids: %.2f ms/call = %.1f GB/s | %s'%(name,B/2**20,D,T,dt*1e3,B/dt/1e9,{k:round(v[0]/v[1],3) for k,v in r.items()}))
    if check:
        import c_oracle
        orc=c_oracle.COracle.from_file(tokpath)
        n=min(D,check)
        wids,woff=orc.encode_packed(text[:int(offs[n])],offs[:n+1])
        ids=d_ids[:int(woff[-1])].cpu().numpy().view(np.uint32); ioff=d_ioff[:n+1].cpu().numpy().view(np.uint64)
        print('   parity on first %d docs:'%n, np.array_equal(ids,wids) and np.array_equal(ioff,woff))
which=sys.argv[1] if len(sys.argv)>1 else '3'
if which=='3':
    text,offs=synth.gen_corpus('mixed',3003,256<<20,doc_median=4096,doc_min=256,doc_max=65536)
    run('config3 mixed/100K', synth.tokenizer_config3(), text, offs, check=2000)
elif which=='4':
    size=int(sys.argv[2]) if len(sys.argv)>2 else 4096
    nd=int(sys.argv[3]) if len(sys.argv)>3 else 16
    docs=synth.gen_long_docs(doc_bytes=size, n_docs=nd)
    text,offs=synth.pack(docs)
    run('config4 long pre-tokens (%d docs of %d B)'%(nd,size), synth.tokenizer_config2(), text, offs, reps=2, check=16 if size<=16384 else None)
