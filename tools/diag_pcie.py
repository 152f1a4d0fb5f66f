#!/usr/bin/env python3
"""PCIe ceilings of the box: pinned H2D alone, D2H alone, both at once (the e2e path's upper bound)."""
import torch, time
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device='cuda'); d_out = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunk=None):
    torch.cuda.synchronize(); t = time.perf_counter()
    c = chunk or n
    for o in range(0, n, c):
        if h2d:
            with torch.cuda.stream(s1): d_in[o:o + c].copy_(h_in[o:o + c], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out[o:o + c].copy_(d_out[o:o + c], non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t
for name, a in (('h2d', (True, False)), ('d2h', (False, True)), ('both', (True, True)), ('both 64MiB chunks', (True, True, 64 << 20))):
    run(*a); dt = min(run(*a) for _ in range(3))
    print('%-20s %.2f ms  %.1f GB/s per direction' % (name, dt * 1e3, n / dt / 1e9))
