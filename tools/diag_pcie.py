#!/usr/bin/env python3
"""PCIe ceilings of the box: pinned H2D alone, D2H alone, both at once (the e2e path's upper bound), and the
effect of chunking and of unaligned host / device addresses on the DMA rate."""
import torch, time
n = 1 << 30
pad = 4096
h_in = torch.empty(n + pad, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n + pad, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n + pad, dtype=torch.uint8, device='cuda'); d_out = torch.empty(n + pad, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunk=None, hoff=0, doff=0, jitter=0):
    torch.cuda.synchronize(); t = time.perf_counter()
    c = chunk or n
    o = 0; k = 0
    while o < n:
        e = min(n, o + c + (jitter * ((k * 7919) % 13) if jitter else 0))
        if h2d:
            with torch.cuda.stream(s1): d_in[o + doff:e + doff].copy_(h_in[o + hoff:e + hoff], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out[o + hoff:e + hoff].copy_(d_out[o + doff:e + doff], non_blocking=True)
        o = e; k += 1
    torch.cuda.synchronize(); return time.perf_counter() - t
M = 1 << 20
for name, a in (('h2d', dict(h2d=True, d2h=False)), ('d2h', dict(h2d=False, d2h=True)), ('both', dict(h2d=True, d2h=True)),
                ('both 64MiB chunks', dict(h2d=True, d2h=True, chunk=64 * M)), ('both 32MiB chunks', dict(h2d=True, d2h=True, chunk=32 * M)),
                ('both 32MiB, host +4 B', dict(h2d=True, d2h=True, chunk=32 * M, hoff=4)),
                ('both 32MiB, host +1 B', dict(h2d=True, d2h=True, chunk=32 * M, hoff=1)),
                ('both 32MiB, dev +4 B', dict(h2d=True, d2h=True, chunk=32 * M, doff=4)),
                ('both 32MiB, ragged sizes (+k*1001 B)', dict(h2d=True, d2h=True, chunk=32 * M, jitter=1001)),
                ('both 32MiB, ragged + host +4 + dev +4', dict(h2d=True, d2h=True, chunk=32 * M, jitter=1001, hoff=4, doff=4)),
                ('both 8MiB chunks', dict(h2d=True, d2h=True, chunk=8 * M))):
    run(**a); dt = min(run(**a) for _ in range(3))
    print('%-42s %.2f ms  %.1f GB/s per direction' % (name, dt * 1e3, n / dt / 1e9))
