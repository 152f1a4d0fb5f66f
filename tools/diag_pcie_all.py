#!/usr/bin/env python3
"""The box's DMA floor: concurrent bidirectional cudaMemcpyAsync between page-locked host memory and N GPUs.

    python tools/diag_pcie_all.py [--gpus 1,2,4,8] [--mib 1024] [--d2h-frac 0.46] [--bind 0|1]

One process, one thread per GPU (the copies are asynchronous, the threads only issue them); every GPU moves `mib` MiB
host->device and d2h_frac x that device->host at the same time (the byte ratio of one encode_batch step with uint16 ids).
With --bind 1 every buffer is allocated by a thread running on the GPU's NUMA node (first touch), with --bind 0 wherever
the main thread runs.  Prints one JSON line per N: aggregate GB/s per direction, per-GPU times.  This is what
bench.py's e2e.floor measures inside the bench for its own N; this tool sweeps N in one go."""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'complexity-tokenizer_b200'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', default='1,2,4,8')
    ap.add_argument('--mib', type=int, default=1024)
    ap.add_argument('--d2h-frac', type=float, default=0.46)
    ap.add_argument('--bind', type=int, default=1)
    ap.add_argument('--reps', type=int, default=3)
    args = ap.parse_args()
    import torch
    from complexity_tokenizer import numa
    have = torch.cuda.device_count()
    nb = args.mib << 20
    nd = int(nb * args.d2h_frac)
    bufs = {}

    def alloc(g):
        if args.bind:
            cpus = numa.cpus_of_node(numa.node_of_device(g)) & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(threading.get_native_id(), cpus)
        h_in = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
        h_in.fill_(1)
        h_out = torch.empty(nd, dtype=torch.uint8, pin_memory=True)
        h_out.fill_(0)
        dev = torch.device('cuda', g)
        bufs[g] = (h_in, h_out, torch.empty(nb, dtype=torch.uint8, device=dev), torch.empty(nd, dtype=torch.uint8, device=dev),
                   torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))

    for n in [int(x) for x in args.gpus.split(',')]:
        if n > have:
            continue
        th = [threading.Thread(target=alloc, args=(g,)) for g in range(n) if g not in bufs]
        [t.start() for t in th]
        [t.join() for t in th]
        res = {}
        for mode in ('both', 'h2d', 'd2h'):
            times = [0.0] * n
            gate = threading.Barrier(n)

            def work(g):
                h_in, h_out, d_in, d_out, s1, s2 = bufs[g]
                torch.cuda.set_device(g)

                def once():
                    if mode != 'd2h':
                        with torch.cuda.stream(s1):
                            d_in.copy_(h_in, non_blocking=True)
                    if mode != 'h2d':
                        with torch.cuda.stream(s2):
                            h_out.copy_(d_out, non_blocking=True)
                    s1.synchronize()
                    s2.synchronize()
                once()
                gate.wait()
                t0 = time.perf_counter()
                for _ in range(args.reps):
                    once()
                times[g] = (time.perf_counter() - t0) / args.reps
            th = [threading.Thread(target=work, args=(g,)) for g in range(n)]
            [t.start() for t in th]
            [t.join() for t in th]
            worst = max(times)
            res[mode] = {'ms_max': worst * 1e3, 'ms_per_gpu': [round(t * 1e3, 2) for t in times],
                         'h2d_GBs_aggregate': 0.0 if mode == 'd2h' else n * nb / worst / 1e9,
                         'd2h_GBs_aggregate': 0.0 if mode == 'h2d' else n * nd / worst / 1e9}
        print(json.dumps({'n_gpus': n, 'mib_h2d_per_gpu': args.mib, 'mib_d2h_per_gpu': nd >> 20, 'bind': bool(args.bind),
                          'numa_nodes': [numa.node_of_device(g) for g in range(n)], 'host_cpus': os.cpu_count(), **res}), flush=True)


if __name__ == '__main__':
    main()
