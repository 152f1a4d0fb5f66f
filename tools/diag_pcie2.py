#!/usr/bin/env python3
"""What lowers the combined PCIe rate inside the encode pipeline?  Both directions in 32 MiB chunks, adding one
ingredient of the real pipeline at a time."""
import torch, time
n = 1 << 30; M = 1 << 20; c = 32 * M
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_small = torch.empty(65536, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device='cuda'); d_out = torch.empty(n, dtype=torch.uint8, device='cuda')
d_scr = torch.empty(256 * M, dtype=torch.uint8, device='cuda'); d_scr2 = torch.empty(256 * M, dtype=torch.uint8, device='cuda')
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def run(dep=False, tiny=False, kernels=False, hostsync=False):
    torch.cuda.synchronize(); t = time.perf_counter()
    for o in range(0, n, c):
        with torch.cuda.stream(s1):
            d_in[o:o + c].copy_(h_in[o:o + c], non_blocking=True)
            if tiny: d_in[o:o + 64].zero_()
            ev = torch.cuda.Event(); ev.record(s1)
    for o in range(0, n, c):
        if kernels or dep:
            with torch.cuda.stream(s3):
                if dep: s3.wait_stream(s1) if False else None
                if kernels:
                    d_scr2[:c].copy_(d_in[o:o + c]); d_scr2[c:2 * c].copy_(d_scr[:c])
                ev2 = torch.cuda.Event(); ev2.record(s3)
            if hostsync: ev2.synchronize()
        with torch.cuda.stream(s2):
            if kernels and not hostsync: s2.wait_event(ev2)
            h_out[o:o + c].copy_(d_out[o:o + c], non_blocking=True)
            if tiny: h_small.copy_(d_out[o:o + 65536], non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t
for name, a in (('plain both, 32 MiB', {}), ('+ tiny memset / 64 KiB copy per chunk', dict(tiny=True)),
                ('+ device kernels between', dict(tiny=True, kernels=True)),
                ('+ host sync per chunk before issuing D2H', dict(tiny=True, kernels=True, hostsync=True))):
    run(**a); dt = min(run(**a) for _ in range(3))
    print('%-48s %.2f ms  %.1f GB/s per direction' % (name, dt * 1e3, n / dt / 1e9))
# dependency-shaped: D2H of chunk k may only start when H2D of chunk k is done (as in the pipeline)
def run_dep():
    torch.cuda.synchronize(); t = time.perf_counter()
    evs = []
    for o in range(0, n, c):
        with torch.cuda.stream(s1):
            d_in[o:o + c].copy_(h_in[o:o + c], non_blocking=True)
            ev = torch.cuda.Event(); ev.record(s1); evs.append(ev)
    for k, o in enumerate(range(0, n, c)):
        with torch.cuda.stream(s2):
            s2.wait_event(evs[k])
            h_out[o:o + c].copy_(d_out[o:o + c], non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t
run_dep(); dt = min(run_dep() for _ in range(3))
print('%-48s %.2f ms  %.1f GB/s per direction' % ('D2H k waits for H2D k (lag of one chunk)', dt * 1e3, n / dt / 1e9))
