"""Optional gather of every shard's ids onto rank 0's GPU over NVLink (sharding.gather_ids), under torchrun.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/diag_gather.py"""
import os, sys, time
sys.path.insert(0,'complexity-tokenizer_b200'); sys.path.insert(0,'fixtures'); sys.path.insert(0,'.')
sys.argv=[sys.argv[0]]
import numpy as np, torch, torch.distributed as dist
import bench, complexity_tokenizer as ct, synth
from complexity_tokenizer.sharding import exchange_shard_metadata, gather_ids
rank=int(os.environ['RANK']); world=int(os.environ['WORLD_SIZE']); local=int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dist.init_process_group('nccl', device_id=torch.device('cuda',local))
dev=torch.device('cuda',local)
tok=ct.Tokenizer.from_file(synth.tokenizer_config2(), device=local)
h_text,B,offs=bench.make_corpus(512<<20, 7000+rank, pinned=False)
D=len(offs)-1
d_text=h_text[:B+64].to(dev); d_off=torch.from_numpy(offs.astype(np.int64)).to(dev)
cap=B+D+16
d_ids=torch.empty(cap,dtype=torch.int32,device=dev); d_ioff=torch.empty(D+1,dtype=torch.int64,device=dev)
n=tok.encode_device(d_text.data_ptr(),d_off.data_ptr(),D,B,d_ids.data_ptr(),cap,d_ioff.data_ptr())
meta=exchange_shard_metadata(rank*D,D,n)
for it in range(3):
    dist.barrier(); torch.cuda.synchronize(); t=time.perf_counter()
    allids=gather_ids(d_ids,meta,rank,dst=0)
    torch.cuda.synchronize(); dist.barrier(); dt=time.perf_counter()-t
if rank==0:
    moved=sum(m['n_ids'] for r,m in enumerate(meta) if r!=0)*4
    ok=bool(torch.equal(allids[:n], d_ids[:n])) and allids.numel()==meta[-1]['ids_base']+meta[-1]['n_ids']
    print('gather_ids to rank 0: %d ranks, %.1f MB moved GPU->GPU in %.2f ms = %.1f GB/s, own shard in place: %s'%(world, moved/1e6, dt*1e3, moved/dt/1e9, ok))
dist.destroy_process_group()
