#!/usr/bin/env python3
"""Timeline of one host-buffer ctk_encode_batch call (CTK_TRACE=1): where the end-to-end time goes."""
import ctypes, os, sys, time
sys.path.insert(0, 'complexity-tokenizer_b200'); sys.path.insert(0, 'fixtures'); sys.path.insert(0, '.')
verbose = len(sys.argv) > 1
sys.argv = [sys.argv[0]]
import numpy as np, torch
import bench, complexity_tokenizer as ct, synth
tok = ct.Tokenizer.from_file(synth.tokenizer_config2())
h_text, B, offs = bench.make_corpus(1 << 30, 5000, pinned=True)
D = len(offs) - 1; h_np = h_text.numpy()[:B]; lib = ct._lib()
def step():
    res = ctypes.c_void_p()
    rc = lib.ctk_encode_batch(tok._h, h_np.ctypes.data, offs.ctypes.data, D, ctypes.byref(res)); assert rc == 0
    lib.ctk_result_free(res)
def timed(label):
    step(); step()
    ts = []
    for _ in range(4):
        t = time.perf_counter(); step(); ts.append((time.perf_counter() - t) * 1e3)
    print('%-40s best %.3f ms  (%.1f GB/s)  all %s' % (label, min(ts), B / min(ts) / 1e6, [round(x, 2) for x in ts]))
timed('default')
for mb in (16, 64, 128):
    os.environ['CTK_CHUNK_MB'] = str(mb); timed('steady chunk %d MiB' % mb)
del os.environ['CTK_CHUNK_MB']
os.environ['CTK_DIAG_NO_D2H'] = '1'; timed('no ids D2H'); del os.environ['CTK_DIAG_NO_D2H']
os.environ['CTK_ABLATE'] = '1'; timed('kernels stop after boundaries'); del os.environ['CTK_ABLATE']
if verbose:
    os.environ['CTK_TRACE'] = '1'; step()
