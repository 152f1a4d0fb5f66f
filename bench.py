#!/usr/bin/env python3
"""bench.py -- encode_batch throughput of the B200-native path, next to the CPU restatement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--bytes B]

One "step" = one pass of the hot path (encode_batch) over one batch of synthetic input.
Workload at every N: BASELINE.json configs[1] -- GPT-2-style 50K ByteLevel BPE + the reference's
pre-token pattern over a synthetic ASCII corpus (1 GiB per GPU, ~4 KiB documents, 50 000-word Zipf
lexicon; fixtures/synth.py).  N > 1 (torchrun, one rank per GPU): documents shard across ranks with
no data-path collective (weak scaling: 1 GiB per rank, different seed per rank, config 5's shape);
only the per-shard id count crosses ranks, once, after the timed region.

Keys of the JSON line (see the task contract):
  value        whole-job input MB/s, inputs already resident in HBM, device-timed (CUDA events, max over ranks)
  e2e          same metric through the host-buffer C-ABI call (ctk_encode_batch_narrow): pinned host text in,
               pinned host ids out (uint16 when the vocabulary fits), H2D and D2H inside the timed region;
               e2e.floor = the same bytes moved by bare cudaMemcpyAsync in both directions at once on all
               ranks (what the box's DMA path can do), e2e.frac_of_floor = how close the call gets
  roofline     dominant kernel: algorithmic bytes (B + 4T + 16(D+1)) / its CUDA-event time
  cpu_baseline the oracle's C core ("port" of the reference algorithm) on all host cores, bounded sample
  single_process  one tokenizer handle over all N GPUs from ONE process (Tokenizer.from_file(devices=...)),
               what a drop-in caller of encode_batch gets; measured by rank 0 while the other ranks wait
  decode_batch extra: device-resident decode of the ids of the last step (round trip must be byte-exact)
  configs      (N = 1) BASELINE configs 1, 3, 4 and the README shape: device ms, MB/s, tokens/s, ids == oracle on
               a sample, CPU port on the same sample; config 4's CPU quadratic fit
  sweep        (N = 1) the headline workload as a function of the distinct-pre-token ratio (lexicon size; cache off)
  comparators  (N = 1) HF `tokenizers` encode_batch (labelled, secondary), the literal list API, pageable input
The pre-token cache is cleared inside every step (ctk default), so no step reuses work of another.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in ('complexity-tokenizer_b200', 'oracle', 'fixtures'):
    sys.path.insert(0, os.path.join(ROOT, p))

import numpy as np  # noqa: E402

METRIC = 'encode_batch_input_throughput'
UNIT = 'MB/s'
WORKLOAD = 'config2: GPT-2-style 50K ByteLevel BPE, synthetic ASCII corpus, 4 KiB docs'
PORT_NOTE = ('oracle C core = restatement of the reference algorithm, gcc -O3 (the Rust reference cannot be built here: no toolchain); '
             'unlike the crate it does not rebuild bytes_to_unicode per document (pretokenizers.rs:159), so it is not slower than the crate on short texts')


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def kernel_source_sha():
    """identifies the encode kernels' sources: a committed ncu traffic figure is only quoted for the same sources"""
    h = hashlib.sha256()
    for f in ('encode_fused.cu', 'start_window.cuh', 'device_common.cuh', 'encode_long.cuh'):
        with open(os.path.join(ROOT, 'complexity-tokenizer_b200', 'csrc', f), 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


class ClockSampler:
    """SM clocks and throttle reasons DURING the timed region, for every GPU of the job, from ONE poller started by rank 0.
    NVML (what nvidia-smi reads) is polled every ~5 ms by a separate process: the timed region of a default run lasts 20 ms,
    far below nvidia-smi's own 100-200 ms loop (the recipe's command line is the fallback when NVML cannot be loaded)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, indices):
        self.indices, self.rows, self.proc, self.nv, self.stop_flag = list(indices), [], None, None, False

    POLLER = r"""
import sys, time, pynvml as nv
nv.nvmlInit()
hs = [nv.nvmlDeviceGetHandleByIndex(int(i)) for i in sys.argv[1].split(',')]
mx = nv.nvmlDeviceGetMaxClockInfo(hs[0], nv.NVML_CLOCK_SM)
R = (nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown, nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap)
print('ready', flush=True)
while True:
    for h in hs:
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM); bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h); pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
        print('%.6f,%d,%d,%.1f,%s' % (time.time(), sm, mx, pw, ','.join('Active' if bits & m else 'Not Active' for m in R)), flush=True)
    time.sleep(0.004)
"""

    def start(self):
        try:                                                   # a separate process: the poller must not share this rank's interpreter or CUDA context
            self.proc = subprocess.Popen([sys.executable, '-c', self.POLLER, ','.join(str(i) for i in self.indices)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.nv = True
            self.ready = threading.Event()
            threading.Thread(target=self._read, daemon=True).start()
            for _ in range(80):                                 # (a poller that never reports must not hang the bench)
                if self.ready.wait(0.1):
                    return
                if self.proc.poll() is not None:
                    break
            self.proc.kill()
        except Exception:
            pass
        self.rows = []
        self.nv = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', ','.join(str(i) for i in self.indices), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(',')]
            if self.nv:                                         # the poller stamps its own rows
                if f[0] == 'ready':
                    self.ready.set()
                    continue
                if len(f) < 8:
                    continue
                self.rows.append((float(f[0]), f[1:]))
            else:
                self.rows.append((time.time(), f))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15 if not self.nv else 0.03)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        if not rows:
            return None
        try:
            sm = sorted(float(r[0]) for r in rows)
            reasons = []
            for name, col in (('hw_slowdown', 3), ('hw_thermal_slowdown', 4), ('sw_thermal_slowdown', 5), ('sw_power_cap', 6)):
                if any(str(r[col]).lower().startswith('active') for r in rows):
                    reasons.append(name)
            return {'sm_mhz': sm[len(sm) // 2], 'sm_mhz_min': sm[0], 'sm_max_mhz': float(rows[0][1]), 'reasons': reasons, 'samples': len(rows),
                    'power_w_max': max(float(r[2]) for r in rows), 'gpus': len(self.indices),
                    'source': 'NVML polled every ~5 ms by a separate process, samples inside the timed region only' if self.nv else 'nvidia-smi -lms 100'}
        except Exception:
            return None


def make_corpus(n_bytes, seed, pinned, kind='ascii', lexicon=50000, doc_median=4096, doc_min=256, doc_max=65536):
    """synthetic corpus straight into (pinned) host memory: (torch uint8 tensor, byte count, numpy offsets)"""
    import torch
    import synth
    t = torch.empty(n_bytes + 64, dtype=torch.uint8, pin_memory=pinned)
    view = t.numpy()
    text, offs = synth.gen_corpus(kind, seed, n_bytes, doc_median=doc_median, doc_min=doc_min, doc_max=doc_max, lexicon=lexicon, out=view)
    view[text.size:] = 0
    return t, text.size, offs


def cpu_port_rate(orc, text_np, offs, budget_s, cores, reps=2):
    """All-core run of the oracle's C core on a bounded prefix: (MB/s, tokens/s, docs, bytes, ids, ids_off)"""
    nd = len(offs) - 1
    base = int(offs[0])
    probe_docs = min(nd, max(16, nd // 64))
    t = time.perf_counter()
    orc.encode_packed(text_np[base:int(offs[probe_docs])], offs[:probe_docs + 1] - offs[0], threads=cores)
    dt = time.perf_counter() - t
    rate = (int(offs[probe_docs]) - base) / max(dt, 1e-6)
    want = min(int(offs[-1]) - base, int(rate * budget_s))
    k = int(np.searchsorted(offs, base + want, side='right')) - 1
    k = max(probe_docs, min(nd, k))
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        ids, ioff = orc.encode_packed(text_np[base:int(offs[k])], offs[:k + 1] - offs[0], threads=cores)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    nb = int(offs[k]) - base
    return nb / best / 1e6, ids.size / best, k, nb, ids, ioff


def cpu_baseline(tok_path, text_np, offs, budget_s=10.0):
    import c_oracle
    orc = c_oracle.COracle.from_file(tok_path)
    cores = os.cpu_count() or 1
    mbs, tps, k, nb, _, _ = cpu_port_rate(orc, text_np, offs, budget_s, cores)
    return {'value': mbs, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'tokens_per_s': tps,
            'sample': 'first %d docs (%.1f MiB) of the same corpus, best of 2; %s' % (k, nb / 2**20, PORT_NOTE)}, orc


def run_reference(args, json_out):
    """--impl reference: the reference's CPU algorithm (oracle port) with all host threads, same config/metric."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import synth
    import c_oracle
    tok_path = synth.tokenizer_config2()
    cores = os.cpu_count() or 1
    view = np.empty(args.bytes + 64, dtype=np.uint8)
    text, offs = synth.gen_corpus('ascii', 5000, args.bytes, doc_median=4096, doc_min=256, doc_max=65536, out=view)
    orc = c_oracle.COracle.from_file(tok_path)
    # the same bytes per step as the GPU arm when warm-up + steps then stay within ~4 minutes; else the largest prefix that does
    t = time.perf_counter()
    k0 = min(len(offs) - 1, 2048)
    orc.encode_packed(text[:int(offs[k0])], offs[:k0 + 1], threads=cores)
    rate = int(offs[k0]) / max(time.perf_counter() - t, 1e-6)
    per_step = min(int(offs[-1]), int(rate * 240.0 / max(1, args.steps + args.warmup)))
    k = len(offs) - 1 if per_step >= int(offs[-1]) else max(k0, int(np.searchsorted(offs, per_step, side='right')) - 1)
    nb = int(offs[k])
    for _ in range(args.warmup):
        orc.encode_packed(text[:nb], offs[:k + 1], threads=cores)
    t0 = time.perf_counter()
    ntok = 0
    for _ in range(args.steps):
        ids, _ = orc.encode_packed(text[:nb], offs[:k + 1], threads=cores)
        ntok = ids.size
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    v = nb / dt / 1e6
    full = nb == int(offs[-1])
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'tokens_per_s': ntok / dt,
        'config': {'workload': WORKLOAD, 'bytes_per_gpu': int(offs[-1]), 'docs_per_gpu': len(offs) - 1, 'lexicon_words': 50000, 'vocab': 50257,
                   'bytes_per_step': nb, 'docs_per_step': k, 'whole_corpus_per_step': full,
                   'parallelism': 'host threads (rayon-like, %d)' % cores},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d docs (%.1f MiB) per step%s; %s' % (k, nb / 2**20, '' if full else ' (bounded so that the run ends within minutes)', PORT_NOTE)},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}), file=json_out)
    json_out.flush()


def device_encode_ms(tok, torch, text_np, offs, steps=3, dev=None):
    """device-resident encode of one corpus: (ms per pass, tokens, per-kernel ms, ids as numpy, ids_off)"""
    B, D = int(text_np.size), len(offs) - 1
    d_text = torch.zeros(B + 64, dtype=torch.uint8, device=dev)
    d_text[:B] = torch.from_numpy(np.ascontiguousarray(text_np)).to(dev)
    d_off = torch.from_numpy(offs.astype(np.int64)).to(dev)
    cap = B + D + 16
    d_ids = torch.empty(cap, dtype=torch.int32, device=dev)
    d_ioff = torch.empty(D + 1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    run = lambda: tok.encode_device(d_text.data_ptr(), d_off.data_ptr(), D, B, d_ids.data_ptr(), cap, d_ioff.data_ptr(), stream=stream)  # noqa: E731
    for _ in range(2):
        T = run()
    torch.cuda.synchronize()
    tok.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        T = run()
    e1.record()
    torch.cuda.synchronize()
    prof = tok.profile_report()
    tok.profile_enable(False)
    ms = e0.elapsed_time(e1) / steps
    return ms, T, {k: v[0] / steps for k, v in sorted(prof.items())}, d_ids[:T].cpu().numpy().view(np.uint32), d_ioff.cpu().numpy().astype(np.uint64)


def dma_floor(torch, dev, h_text_t, B, d2h_bytes, reps, barrier):
    """bare cudaMemcpyAsync of the bytes one e2e step moves, both directions at once, pinned memory: ms (this rank)"""
    d_in = torch.empty(B, dtype=torch.uint8, device=dev)
    d_out = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=dev)
    h_out = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, pin_memory=True)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def once():
        with torch.cuda.stream(s1):
            d_in.copy_(h_text_t[:B], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
    once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    return (time.perf_counter() - t0) * 1e3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--bytes', type=int, default=1 << 30, help='corpus bytes per GPU')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip configs / sweep / comparators / encodings / train (kernel experiments)')
    ap.add_argument('--no-single', action='store_true', help='skip the single-process multi-device block (N > 1)')
    ap.add_argument('--no-encodings', action='store_true')
    ap.add_argument('--encodings-bytes', type=int, default=256 << 20)
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--train-bytes', type=int, default=3 << 20)
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: anything libraries print there (NCCL's version banner at N > 1) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        return run_reference(args, json_out)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    os.environ.pop('OMP_NUM_THREADS', None)                 # torchrun pins every rank to one thread; the helper threads here are not OpenMP, numpy is not on the timed path

    import torch
    import torch.distributed as dist
    import complexity_tokenizer as ct
    import synth
    from complexity_tokenizer import numa
    from complexity_tokenizer.sharding import exchange_shard_metadata

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product has no CPU path')
    torch.cuda.set_device(local)
    # one rank per GPU: run next to it -- this process's pages (the pinned corpus, the result buffers) then live on the GPU's NUMA node
    placement = numa.bind_process_to_device(local) if world > 1 else numa.describe(local)
    cpu_group = None
    if world > 1:
        # NCCL is plumbing here (barriers, the max over ranks): the data path has no collective.  At 8 GPUs NCCL would set up
        # NVLS (NVLink SHARP multicast), and a communicator that holds NVLS resources makes k_encode_slices 11 % slower on every
        # rank (3.86 vs 3.48 ms, measured: profiles/r2_v31_bench_n8.json vs r2_v33_nvls0.json; clocks, power and eight
        # independent processes are unaffected).  Nothing here reduces through the switch, so NVLS stays off unless asked for.
        os.environ.setdefault('NCCL_NVLS_ENABLE', '0')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        cpu_group = dist.new_group(backend='gloo')          # for waits that must not keep a GPU busy (an NCCL barrier spins in a kernel)
    dev = torch.device('cuda', local)

    tok_path = synth.tokenizer_config2()
    tok = ct.Tokenizer.from_file(tok_path, device=local)

    # ---- inputs: pinned host copy (for e2e) and a device-resident copy (for value)
    h_text, B, offs = make_corpus(args.bytes, 5000 + rank, pinned=True)
    D = len(offs) - 1
    h_off = torch.from_numpy(offs.astype(np.int64)).pin_memory()
    d_text = torch.empty(B + 64, dtype=torch.uint8, device=dev)
    d_text.copy_(h_text[:B + 64], non_blocking=True)
    d_off = h_off.to(dev)
    ids_cap = B + D + 16
    d_ids = torch.empty(ids_cap, dtype=torch.int32, device=dev)
    d_ids_off = torch.empty(D + 1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        return tok.encode_device(d_text.data_ptr(), d_off.data_ptr(), D, B, d_ids.data_ptr(), ids_cap, d_ids_off.data_ptr(), stream=stream)

    for _ in range(max(3, args.warmup)):
        T = step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region: exactly K steps
    tok.profile_enable(True)
    sampler = ClockSampler(range(world) if world > 1 else [local]) if rank == 0 else None     # rank 0 watches every GPU of the job
    if sampler:
        sampler.start()
    time.sleep(0.3)
    launches0 = ct._lib().ctk_kernel_launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        T = step()
    e1.record()
    barrier()
    w1 = time.time()
    launches = ct._lib().ctk_kernel_launches() - launches0
    clocks = sampler.stop(w0, w1) if sampler else None
    prof = tok.profile_report()
    tok.profile_enable(False)
    # the only cross-shard exchange of the path: per-shard (first_doc, n_docs, n_ids) -> global id offsets; nothing on the
    # data path waits for it, so it happens once here and not inside every step
    shard_meta = exchange_shard_metadata(rank * D, D, T) if world > 1 else None
    ms = e0.elapsed_time(e1)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(B), float(T), float(D)], dtype=torch.float64, device=dev)
    per_rank_ms = None
    if world > 1:
        allms = [torch.zeros_like(tms) for _ in range(world)]
        dist.all_gather(allms, tms)
        per_rank_ms = [float(x.item()) / args.steps for x in allms]
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms = float(tms.item())
    B_all, T_all, D_all = (float(x) for x in tot.tolist())
    ms_per_step = ms / args.steps
    value = B_all / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the host-buffer C-ABI call (pinned host in, pinned host out)
    h_np = h_text.numpy()[:B]
    lib = ct._lib()
    import ctypes

    def e2e_step(fn):
        res = ctypes.c_void_p()
        rc = fn(tok._h, h_np.ctypes.data, offs.ctypes.data, D, ctypes.byref(res))
        if rc != 0:
            ct._raise(rc)
        n = int(np.ctypeslib.as_array(ctypes.cast(lib.ctk_result_offsets(res), ctypes.POINTER(ctypes.c_uint64)), (D + 1,))[-1])
        w = int(lib.ctk_result_id_width(res))
        lib.ctk_result_free(res)
        return n, w

    def time_e2e(fn):
        e2e_step(fn)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            n, w = e2e_step(fn)
        torch.cuda.synchronize()
        ms_ = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
        et = torch.tensor([ms_], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        assert n == T
        hb, db = ctypes.c_uint64(), ctypes.c_uint64()
        lib.ctk_last_transfer_bytes(tok._h, ctypes.byref(hb), ctypes.byref(db))
        return float(et.item()), w, int(hb.value), int(db.value)

    e2e_ms, width, hbytes, dbytes = time_e2e(lib.ctk_encode_batch_narrow)
    e2e32_ms, _, _, dbytes32 = time_e2e(lib.ctk_encode_batch)
    floor_ms = dma_floor(torch, dev, h_text, B, dbytes, max(2, args.e2e_steps), barrier)
    ft = torch.tensor([floor_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ft, op=dist.ReduceOp.MAX)
    floor_ms = float(ft.item())
    e2e = {'value': B_all / (e2e_ms * 1e-3) / 1e6, 'unit': UNIT, 'ms_per_step': e2e_ms,
           'h2d_bytes_per_step': hbytes, 'd2h_bytes_per_step': dbytes,
           'api': 'ctk_encode_batch_narrow (C ABI, pinned host buffers in, uint%d ids in pinned host memory out; no host widening: '
                  'the binding reads the device-width ids when it builds its per-document vectors)' % (8 * width),
           'id_width': width, 'result_bytes_in_host_memory': int(width * T + 8 * (D + 1)),
           'uint32_api': {'api': 'ctk_encode_batch (uint32 ids)', 'value': B_all / (e2e32_ms * 1e-3) / 1e6, 'ms_per_step': e2e32_ms,
                          'd2h_bytes_per_step': dbytes32},
           'floor': {'value': B_all / (floor_ms * 1e-3) / 1e6, 'unit': UNIT, 'ms_per_step': floor_ms,
                     'what': 'bare cudaMemcpyAsync of the same h2d and d2h byte counts, both directions at once, pinned memory, all %d rank(s) together' % world},
           'frac_of_floor': floor_ms / e2e_ms, 'host_placement': placement}

    # ---- decode_batch on the ids just produced (device-resident; BASELINE config 5's round-trip shape):
    #      raw decode must give the input back byte for byte; the default decode (clean-up on) is timed next to it
    d_back = torch.empty(B + 1024, dtype=torch.uint8, device=dev)
    d_back_off = torch.empty(D + 1, dtype=torch.int64, device=dev)

    def dec(clean):
        return tok.decode_device(d_ids.data_ptr(), d_ids_off.data_ptr(), D, T, d_back.data_ptr(), B + 1024, d_back_off.data_ptr(), False, clean,
                                 stream=stream)

    dec_ms, dec_prof = {}, {}
    for clean in (False, True):
        dec(clean)
        barrier()
        tok.profile_enable(True)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(args.steps):
            nb_out = dec(clean)
        d1.record()
        barrier()
        dec_ms[clean] = d0.elapsed_time(d1) / args.steps
        dec_prof[clean] = {k: v[0] / args.steps for k, v in sorted(tok.profile_report().items())}
        tok.profile_enable(False)
        if not clean:
            roundtrip = bool(nb_out == B and torch.equal(d_back[:B], d_text[:B]) and torch.equal(d_back_off, d_off))
    dt_ = torch.tensor([dec_ms[False], dec_ms[True], 0.0 if roundtrip else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt_, op=dist.ReduceOp.MAX)
    peak, peak_src = peaks()
    dec_alg = 4 * T + B + 16 * (D + 1)
    decode = {'value': B_all / (float(dt_[0]) * 1e-3) / 1e6, 'unit': 'MB/s (decoded bytes, device-resident, clean_up_tokenization_spaces=False)',
              'ms_per_step': float(dt_[0]), 'roundtrip_exact': float(dt_[2]) == 0.0,
              'default_clean_up_ms_per_step': float(dt_[1]), 'default_clean_up_value': B_all / (float(dt_[1]) * 1e-3) / 1e6,
              'roofline_frac_raw': dec_alg / (dec_ms[False] * 1e-3) / 1e9 / peak, 'roofline_frac_clean_up': dec_alg / (dec_ms[True] * 1e-3) / 1e9 / peak,
              'algorithmic_bytes': int(dec_alg), 'kernels_ms_raw': dec_prof[False], 'kernels_ms_clean_up': dec_prof[True]}

    # ---- one process, all N GPUs, one tokenizer handle: what a drop-in caller of Tokenizer.encode_batch gets.
    #      Rank 0 runs it over every GPU of the job while the other ranks wait at a barrier (their GPUs are idle).
    single = None
    if world > 1:
        del d_back, d_back_off
        barrier()
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)                       # from here the other ranks wait on the CPU: their GPUs are free for rank 0's process
        if rank == 0 and not args.no_single:
            try:
                # this rank was bound to GPU 0's node: let the library's per-device threads place themselves
                numa.unbind_process()
                mt = ct.Tokenizer.from_file(tok_path, devices=list(range(world)))
                reps_text = [h_text]
                big_B = B
                # the same corpus shape, N times the bytes: N - 1 more pinned GiB (seeds of the other ranks)
                parts = [(h_np, offs)]
                for r in range(1, world):
                    t_r, B_r, o_r = make_corpus(args.bytes, 5000 + r, pinned=False)
                    parts.append((t_r.numpy()[:B_r], o_r))
                    reps_text.append(t_r)
                big = torch.empty(sum(p[0].size for p in parts) + 64, dtype=torch.uint8, pin_memory=True)
                big_np = big.numpy()
                big_off = [np.zeros(1, dtype=np.uint64)]
                pos = 0
                for tx, of in parts:
                    big_np[pos:pos + tx.size] = tx
                    big_off.append(of[1:] + np.uint64(pos))
                    pos += tx.size
                big_off = np.concatenate(big_off)
                big_B, big_D = pos, len(big_off) - 1
                del reps_text, parts

                def sp_step():
                    res = ctypes.c_void_p()
                    rc = lib.ctk_encode_batch_narrow(mt._h, big_np.ctypes.data, big_off.ctypes.data, big_D, ctypes.byref(res))
                    if rc != 0:
                        ct._raise(rc)
                    np_ = int(lib.ctk_result_parts(res))
                    lib.ctk_result_free(res)
                    return np_
                sp_step()
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    n_parts = sp_step()
                sp_ms = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
                single = {'value': big_B / (sp_ms * 1e-3) / 1e6, 'unit': UNIT, 'ms_per_step': sp_ms, 'bytes_per_step': int(big_B), 'docs_per_step': int(big_D),
                          'devices': mt.devices, 'result_parts': n_parts,
                          'numa_nodes': [int(lib.ctk_numa_node(mt._h, i)) for i in range(world)],
                          'api': 'Tokenizer.from_file(path, devices=[0..N-1]) -> one ctk_encode_batch_narrow call from one process, pinned host buffers'}
                del mt, big, big_np
            except Exception as ex:                       # never lose the headline line over the extra
                single = {'error': repr(ex)}
        dist.barrier(group=cpu_group)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events inside the library, same timed region)
    dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else (None, (0.0, 0))
    alg_bytes = B + 4 * T + 16 * (D + 1)
    roofline = None
    if dom[0]:
        avg_ms = dom[1][0] / max(1, dom[1][1])
        ach = alg_bytes / (avg_ms * 1e-3) / 1e9
        kernel_ms_total = sum(v[0] for v in prof.values())
        roofline = {'bound': 'hbm', 'kernel': dom[0], 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                    'traffic': None, 'peak_source': peak_src, 'kernel_ms_per_launch': avg_ms,
                    'kernel_share_of_device_time': dom[1][0] / max(kernel_ms_total, 1e-9),
                    'algorithmic_bytes_per_launch': int(alg_bytes),
                    'input_bandwidth_frac_whole_step': (B / (ms_per_step * 1e-3) / 1e9) / peak,
                    'input_bandwidth_frac_nominal_8TBs': (B / (ms_per_step * 1e-3) / 1e9) / 8000.0,
                    'all_kernels_ms_per_step': {k: v[0] / args.steps for k, v in sorted(prof.items())}}
        # DRAM bytes of the dominant kernel: from the committed ncu --set full capture, quoted only when it was taken from these very sources
        try:
            with open(os.path.join(ROOT, 'bench_traffic.json')) as f:
                ent = json.load(f).get(roofline['kernel'])
            if ent and int(ent['input_bytes']) == int(B) and ent.get('kernel_source_sha') == kernel_source_sha():
                roofline['traffic'] = int(ent['dram_bytes_per_launch'])
                roofline['traffic_source'] = ent.get('source')
            elif ent:
                roofline['traffic_note'] = 'committed capture (%s) is of other kernel sources or another input size: not quoted' % ent.get('source')
        except Exception:
            pass
    cpu = None
    orc2 = None
    if not args.no_cpu:
        cpu, orc2 = cpu_baseline(tok_path, h_np, offs)

    extras = {}
    if not args.no_extras and world == 1:
        del d_back, d_back_off
        try:
            extras = run_extras(args, tok, tok_path, torch, dev, ct, synth, h_np, offs, h_text, B, D, T, peak, orc2)
        except Exception as ex:
            import traceback
            extras = {'extras_error': repr(ex), 'trace': traceback.format_exc()[-1500:]}

    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
        'data': 'synthetic', 'tokens_per_s': T_all / (ms_per_step * 1e-3), 'ms_per_step_per_rank': per_rank_ms,
        'config': {'workload': WORKLOAD, 'bytes_per_gpu': int(B), 'docs_per_gpu': int(D), 'tokens_per_gpu': int(T),
                   'lexicon_words': 50000, 'vocab': 50257, 'l2': 'inputs (1 GiB) larger than L2 (126 MB); no flush needed',
                   'pretoken_cache': 'cleared inside every step', 'parallelism': 'documents sharded over %d GPU(s), no collective on the data path' % world,
                   'shard_metadata': 'exchanged once after the timed steps (24 bytes per rank)' if world > 1 else None,
                   'nccl': ('plumbing only (barrier, max over ranks); NCCL_NVLS_ENABLE=%s' % os.environ.get('NCCL_NVLS_ENABLE')) if world > 1 else None},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu,
        'decode_batch': decode, 'single_process': single}
    out.update(extras)
    print(json.dumps(out), file=json_out)
    json_out.flush()
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, tok, tok_path, torch, dev, ct, synth, h_np, offs, h_text, B, D, T, peak, orc2):
    """N = 1 only: the other BASELINE configs, the miss-rate sweep, the comparators, rich encodings, training"""
    import ctypes
    import c_oracle
    lib = ct._lib()
    cores = os.cpu_count() or 1
    out = {}

    def parity_sample(tok_, orc_, text, o, ids, ioff, every):
        """ids of every `every`-th document against the oracle"""
        nd = len(o) - 1
        pick = list(range(0, nd, max(1, every)))
        docs_t = [text[int(o[i]):int(o[i + 1])] for i in pick]
        po = np.zeros(len(pick) + 1, dtype=np.uint64)
        po[1:] = np.cumsum([d.size for d in docs_t])
        want, woff = orc_.encode_packed(np.concatenate(docs_t) if docs_t else np.zeros(0, np.uint8), po, threads=cores)
        ok = True
        for j, i in enumerate(pick):
            a = ids[int(ioff[i]):int(ioff[i + 1])]
            b = want[int(woff[j]):int(woff[j + 1])]
            if a.size != b.size or not np.array_equal(a, b):
                ok = False
                break
        return ok, len(pick)

    def one_config(name, tok_, orc_, text, o, budget_s, every, note):
        ms, Tn, kern, ids, ioff = device_encode_ms(tok_, torch, text, o, dev=dev)
        Bn, Dn = int(text.size), len(o) - 1
        alg = Bn + 4 * Tn + 16 * (Dn + 1)
        ent = {'workload': note, 'bytes': Bn, 'docs': Dn, 'tokens': int(Tn), 'device_ms': ms, 'MB_per_s': Bn / (ms * 1e-3) / 1e6,
               'tokens_per_s': Tn / (ms * 1e-3), 'roofline_achieved_GBs': alg / (ms * 1e-3) / 1e9, 'roofline_frac': alg / (ms * 1e-3) / 1e9 / peak,
               'kernels_ms': {k: v for k, v in kern.items() if v >= 0.01}}
        if orc_ is not None and not args.no_cpu:
            ok, n_s = parity_sample(tok_, orc_, text, o, ids, ioff, every)
            ent['ids_equal_oracle'] = ok
            ent['parity_sample_docs'] = n_s
            mbs, tps, k, nb, _, _ = cpu_port_rate(orc_, text, o, budget_s, cores, reps=1)
            ent['cpu_port'] = {'MB_per_s': mbs, 'tokens_per_s': tps, 'cores': cores, 'sample': 'first %d docs (%.1f MiB)' % (k, nb / 2**20)}
        return ent

    configs = {}
    # config 1: 10K English-like texts, 32K vocabulary; and the README shape (10 000 texts of 4-12 bytes)
    p1 = synth.tokenizer_config1()
    tok1 = ct.Tokenizer.from_file(p1, device=tok.device)
    orc1 = None if args.no_cpu else c_oracle.COracle.from_file(p1)
    t1, o1 = synth.gen_corpus('english', 1001, 12 << 20)
    configs['config1'] = one_config('config1', tok1, orc1, t1, o1, 6.0, 10, '10K synthetic English texts (12 MB), 32K byte-level BPE vocabulary')
    docs1 = [bytes(t1[int(o1[i]):int(o1[i + 1])]).decode() for i in range(len(o1) - 1)]
    tok1.encode_batch(docs1[:100])
    best_l = None
    for _ in range(3):                                      # best of 3: one shot once measured 328 ms against 91-100 ms (allocator / GC noise of 3.4 M Python ints)
        t0 = time.perf_counter()
        lst = tok1.encode_batch(docs1)
        dt_l = time.perf_counter() - t0
        best_l = dt_l if best_l is None else min(best_l, dt_l)
    configs['config1']['list_api'] = {'api': 'Tokenizer.encode_batch(list[str]) -> list[list[int]] (the literal drop-in call), best of 3', 'ms': best_l * 1e3,
                                      'MB_per_s': t1.size / best_l / 1e6, 'ids': int(sum(map(len, lst)))}
    rng = np.random.default_rng(7)
    words = [bytes(rng.integers(97, 123, size=int(k), dtype=np.uint8)).decode() for k in rng.integers(4, 13, size=10000)]
    tok1.encode_batch(words[:10])
    best = None
    for _ in range(5):
        t0 = time.perf_counter()
        r_ = tok1.encode_batch(words)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    configs['readme_shape'] = {'workload': 'README.md:70-72 shape: 10 000 texts of 4-12 bytes, Tokenizer.encode_batch(list[str])', 'ms_best_of_5': best * 1e3,
                               'reference_published_ms': 20.0, 'reference_published_note': 'README.md:72, hardware and text length not stated',
                               'texts_per_s': 10000 / best, 'ids': int(sum(map(len, r_)))}
    if orc1 is not None:
        t0 = time.perf_counter()
        want = orc1.encode_batch(words)
        configs['readme_shape']['cpu_port_ms'] = (time.perf_counter() - t0) * 1e3
        configs['readme_shape']['ids_equal_oracle'] = r_ == want
    del tok1, lst
    # config 3: 256 MiB mixed French / CJK / emoji, 100K vocabulary
    p3 = synth.tokenizer_config3()
    tok3 = ct.Tokenizer.from_file(p3, device=tok.device)
    orc3 = None if args.no_cpu else c_oracle.COracle.from_file(p3)
    t3, o3 = synth.gen_corpus('mixed', 3003, 256 << 20, doc_median=4096, doc_min=256, doc_max=65536)
    configs['config3'] = one_config('config3', tok3, orc3, t3, o3, 6.0, 100, '256 MiB mixed French/CJK/emoji UTF-8, 100K-entry tokenizer.json in the INL-trainer shape')
    del tok3, t3
    # config 4: 64 documents of 1 MiB with very long pre-tokens (config-2 tokenizer)
    t4, o4 = synth.pack(synth.gen_long_docs())
    ms4, T4, k4, _, _ = device_encode_ms(tok, torch, t4, o4, dev=dev)
    configs['config4'] = {'workload': '64 x 1 MiB documents, pre-tokens up to 2^20 bytes (block-level merge path)', 'bytes': int(t4.size), 'docs': 64, 'tokens': int(T4),
                          'device_ms': ms4, 'MB_per_s': t4.size / (ms4 * 1e-3) / 1e6, 'tokens_per_s': T4 / (ms4 * 1e-3),
                          'xlong_rounds': int(lib.ctk_debug_xlong_rounds(tok._h)), 'kernels_ms': {k: v for k, v in k4.items() if v >= 0.01}}
    if orc2 is not None:                                   # the reference's loop is O(n^2) per pre-token: time it at 4 / 16 / 64 KiB, fit, extrapolate
        pts = []
        run = t4[:1 << 20]                                  # document 0: one run of [a-z]
        for kib in (4, 16, 64):
            n = kib << 10
            if pts and pts[-1][1] * 16 > 45.0:              # the next point would take too long: the fit says so
                break
            oo = np.array([0, n], dtype=np.uint64)
            t0 = time.perf_counter()
            ids_c, _ = orc2.encode_packed(run[:n], oo, threads=1)
            pts.append((n, time.perf_counter() - t0))
            g_ids, g_off = tok.encode_packed(run[:n], oo)
            if not np.array_equal(g_ids, ids_c):
                configs['config4']['ids_equal_oracle'] = False
        configs['config4'].setdefault('ids_equal_oracle', True)
        a = float(np.mean([s / (n * n) for n, s in pts[1:] or pts]))
        configs['config4']['cpu_port_quadratic_fit'] = {'points_bytes_seconds': pts, 'seconds_per_byte_squared': a,
                                                        'extrapolated_seconds_for_one_1MiB_pretoken': a * float(1 << 20) ** 2,
                                                        'note': 'one thread per pre-token (the reference parallelises over documents only); the 1 MiB figure is EXTRAPOLATED'}
    out['configs'] = configs

    # ---- the headline workload as a function of the distinct-pre-token ratio: lexicon 50K / 500K / 5M words (256 MiB each)
    #      and the cache switched off (CTK_ABLATE=4: every pre-token goes through the merge loop)
    sweep = []
    sweep_bytes = 256 << 20
    for lex in (50000, 500000, 5000000):
        ts, Bs, os_ = make_corpus(sweep_bytes, 6000, pinned=False, lexicon=lex)
        ms, Ts, kern, ids_s, ioff_s = device_encode_ms(tok, torch, ts.numpy()[:Bs], os_, dev=dev)
        words_seen = None
        try:                                                  # distinct pre-tokens ~ distinct space-separated words (host count on a 32 MiB prefix)
            pref = bytes(ts.numpy()[:min(Bs, 32 << 20)])
            toks_ = pref.split()
            words_seen = {'prefix_MiB': len(pref) >> 20, 'words': len(toks_), 'distinct': len(set(toks_))}
        except Exception:
            pass
        ent = {'lexicon_words': lex, 'bytes': int(Bs), 'tokens': int(Ts), 'device_ms': ms, 'MB_per_s': Bs / (ms * 1e-3) / 1e6,
               'k_encode_slices_ms': kern.get('k_encode_slices'), 'host_word_count': words_seen}
        if orc2 is not None:
            ok, n_s = parity_sample(tok, orc2, ts.numpy()[:Bs], os_, ids_s, ioff_s, 200)
            ent['ids_equal_oracle'] = ok
        sweep.append(ent)
        del ts
    cache_off = None
    os.environ['CTK_ABLATE'] = '4'                            # every pre-token is merged where it stands (no lookup hit, no publication)
    try:
        ts, Bs, os_ = make_corpus(32 << 20, 6000, pinned=False)
        ms, Ts, kern, ids_s, ioff_s = device_encode_ms(tok, torch, ts.numpy()[:Bs], os_, steps=1, dev=dev)
        cache_off = {'bytes': int(Bs), 'device_ms': ms, 'MB_per_s': Bs / (ms * 1e-3) / 1e6, 'k_encode_slices_ms': kern.get('k_encode_slices')}
        if orc2 is not None:
            cache_off['ids_equal_oracle'] = parity_sample(tok, orc2, ts.numpy()[:Bs], os_, ids_s, ioff_s, 50)[0]
    finally:
        os.environ.pop('CTK_ABLATE', None)
    out['sweep'] = {'pretoken_cache_off': cache_off, 'what': 'config-2 generator at 256 MiB, Zipf(1.07) over a lexicon of N words: more distinct pre-tokens -> more first-occurrence merges', 'points': sweep}

    # ---- comparators: HF tokenizers (secondary, labelled), pageable input
    comp = {}
    try:
        from tokenizers import Regex, Tokenizer as HFTok, models, normalizers, pre_tokenizers
        import tokenizers as hf_mod
        with open(tok_path, encoding='utf-8') as f:
            tj = json.load(f)
        merges = [tuple(m.split(' ')) if isinstance(m, str) else tuple(m) for m in tj['model']['merges']]
        hf = HFTok(models.BPE(vocab=tj['model']['vocab'], merges=merges))
        hf.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+"), 'isolated'),
                                                    pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
        n_hf = int(np.searchsorted(offs, 48 << 20, side='right')) - 1
        docs = [bytes(h_np[int(offs[i]):int(offs[i + 1])]).decode() for i in range(n_hf)]
        hf.encode_batch(docs[:64], add_special_tokens=False)
        t0 = time.perf_counter()
        enc = hf.encode_batch(docs, add_special_tokens=False)
        dt = time.perf_counter() - t0
        ntok = sum(len(e.ids) for e in enc)
        g = tok.encode_batch(docs[:200])
        comp['hf_tokenizers'] = {'version': hf_mod.__version__, 'what': 'HuggingFace tokenizers encode_batch (rayon, all cores), same vocabulary and pattern, list[str] in, Encoding objects out '
                                 '(Python object conversion INCLUDED; secondary comparator, not the reference)', 'cores': cores,
                                 'sample': 'first %d docs (%.1f MiB) of the same corpus' % (n_hf, int(offs[n_hf]) / 2**20), 'MB_per_s': int(offs[n_hf]) / dt / 1e6,
                                 'tokens_per_s': ntok / dt, 'ids_equal_gpu_on_200_docs': [e.ids for e in enc[:200]] == g}
        del enc, docs
    except Exception as ex:
        comp['hf_tokenizers'] = {'error': repr(ex)}
    try:                                                      # the same 1 GiB from PAGEABLE memory (a NumPy array, a Rust Vec): the library stages it
        pg = np.array(h_np, copy=True)
        res = ctypes.c_void_p()
        for rep in range(2):
            t0 = time.perf_counter()
            rc = lib.ctk_encode_batch_narrow(tok._h, pg.ctypes.data, offs.ctypes.data, D, ctypes.byref(res))
            dt = time.perf_counter() - t0
            if rc != 0:
                ct._raise(rc)
            lib.ctk_result_free(res)
        comp['pageable_input'] = {'api': 'ctk_encode_batch_narrow, caller buffer NOT page-locked', 'ms': dt * 1e3, 'MB_per_s': B / dt / 1e6}
        del pg
    except Exception as ex:
        comp['pageable_input'] = {'error': repr(ex)}
    out['comparators'] = comp

    # ---- rich Encoding outputs on a bounded sample of the same corpus, through the C ABI with host buffers
    if not args.no_encodings:
        n_s = int(np.searchsorted(offs, min(B, args.encodings_bytes), side='right')) - 1
        s_off = offs[:n_s + 1].copy()
        s_B = int(s_off[-1])
        s_text = h_np[:s_B]
        encodings = {'sample': 'first %d docs (%.1f MiB) of the same corpus, host buffers in, page-locked result out' % (n_s, s_B / 2**20)}
        for key, kw in (('call_padded_1024', dict(add_special_tokens=True, truncation=True, max_length=1024, padding=2, pad_to=1024)),
                        ('to_encoding_offsets', dict(add_special_tokens=True, want_offsets=True))):
            r_ = tok._encode_rows(s_text, s_off, **kw)
            del r_
            tok.profile_enable(True)
            t0 = time.perf_counter()
            for _ in range(2):
                r_ = tok._encode_rows(s_text, s_off, **kw)
                n_out = int(r_.input_ids.size)
                del r_
            ms_ = (time.perf_counter() - t0) * 1e3 / 2
            pr = tok.profile_report()
            tok.profile_enable(False)
            encodings[key] = {'e2e_ms': ms_, 'e2e_MB_per_s': s_B / (ms_ * 1e-3) / 1e6, 'out_elements': n_out,
                              'device_ms': sum(v[0] / v[1] for v in pr.values()),
                              'device_MB_per_s': s_B / (sum(v[0] / v[1] for v in pr.values()) * 1e-3) / 1e6,
                              'kernels_ms': {k: v[0] / v[1] for k, v in sorted(pr.items())}}
        out['encodings'] = encodings

    # ---- BPE training on the GPU on the trainer's own shape
    if not args.no_train:
        n_t = int(np.searchsorted(offs, min(B, args.train_bytes), side='right')) - 1
        t_off = offs[:n_t + 1].copy()
        trn = ct.BpeTrainer(vocab_size=32000, min_frequency=2, show_progress=False, device=tok.device)
        trn.train_packed(h_np[:4096], np.array([0, 4096], dtype=np.uint64))
        t0 = time.perf_counter()
        tv, tm = trn.train_packed(h_np[:int(t_off[-1])], t_off)
        wall_ = time.perf_counter() - t0
        ts_ = trn.last_stats
        train = {'sample': 'first %d docs (%.1f MiB) of the same corpus, split on White_Space' % (n_t, int(t_off[-1]) / 2**20),
                 'vocab_size': len(tv), 'merges': len(tm), 'wall_s': wall_, 'device_ms_word_histogram': ts_['ms_words'],
                 'device_ms_merge_loop': ts_['ms_merges'], 'us_per_merge': 1e3 * ts_['ms_merges'] / max(1, len(tm)),
                 'words': ts_['n_words'], 'unique_words': ts_['n_unique_words'], 'symbols': ts_['n_symbols'],
                 'device_ms_word_histogram_kernels': ts_.get('ms_words_kernels'),
                 'kernel_launches': ts_['kernel_launches'], 'table_rebuilds': ts_['table_rebuilds']}
        if not args.no_cpu:             # CPU arm on a bounded sample: the restated reference trainer (Python, strings, one core)
            import py_trainer
            raw_ = h_np[:int(t_off[min(n_t, 120)])].tobytes()
            docs_ = [raw_[int(t_off[i]):int(t_off[i + 1])].decode() for i in range(min(n_t, 120))]
            t0 = time.perf_counter()
            want_ = py_trainer.train_bpe(docs_, vocab_size=4 + 80 + 200, min_frequency=2)
            cpu_s_ = time.perf_counter() - t0
            trn2 = ct.BpeTrainer(vocab_size=4 + 80 + 200, min_frequency=2, show_progress=False, device=tok.device)
            t0 = time.perf_counter()
            got_ = trn2.train(docs_)
            gpu_s_ = time.perf_counter() - t0
            train['cpu_baseline'] = {'kind': 'port', 'cores': 1, 'sample': '%d docs, %d merges (oracle/py_trainer.py)' % (len(docs_), len(want_[1])),
                                     'seconds': cpu_s_, 'gpu_wall_seconds': gpu_s_, 'gpu_device_ms_merge_loop': trn2.last_stats['ms_merges'],
                                     'equal': got_ == want_}
        out['train_bpe'] = train
        # ---- train_new_from_iterator (mod.rs:1231-1322) on config 1's own shape: normaliser, pre-tokenizer and trainer on the device
        try:
            p1 = synth.tokenizer_config1()
            tok1 = ct.Tokenizer.from_file(p1, device=tok.device)
            t1, o1 = synth.gen_corpus('english', 1001, 12 << 20)
            nd = min(2000, len(o1) - 1)
            raw1 = t1[:int(o1[nd])].tobytes()
            docs1 = [raw1[int(o1[i]):int(o1[i + 1])].decode() for i in range(nd)]
            tok1.train_new_from_iterator(docs1[:20], 400)
            t0 = time.perf_counter()
            new1 = tok1.train_new_from_iterator(docs1, 32000)
            wall1 = time.perf_counter() - t0
            st1 = tok1.last_train_stats
            got1 = new1.encode_batch(docs1[:50])
            out['train_new_from_iterator'] = {
                'workload': 'config 1: first %d docs (%.1f MiB) through the tokenizer\'s own NFC + ByteLevel pre-tokenizer on the device, then BpeTrainer to 32 000 entries' % (nd, len(raw1) / 2**20),
                'wall_s': wall1, 'merges': int(st1['n_merges']), 'words': int(st1['n_words']), 'unique_words': int(st1['n_unique_words']),
                'device_ms_word_histogram': st1['ms_words'], 'device_ms_merge_loop': st1['ms_merges'],
                'us_per_merge': 1e3 * st1['ms_merges'] / max(1, int(st1['n_merges'])), 'new_vocab_size': new1.vocab_size,
                'round_trip_exact_on_50_docs': new1.decode_batch_with_options(got1, False, False) == docs1[:50]}
        except Exception as ex:
            out['train_new_from_iterator'] = {'error': repr(ex)}
    # ---- Split stages: the headline corpus through Sequence[Split(\p{N}{1,3}, Isolated), ByteLevel]
    try:
        import tempfile
        with open(tok_path, encoding='utf-8') as f:
            tjs = json.load(f)
        tjs['pre_tokenizer'] = {'type': 'Sequence', 'pretokenizers': [{'type': 'Split', 'pattern': {'Regex': r'\p{N}{1,3}'}, 'behavior': 'Isolated', 'invert': False},
                                                                      {'type': 'ByteLevel', 'add_prefix_space': False, 'use_regex': False}]}
        toks = ct.Tokenizer.from_str(json.dumps(tjs), device=tok.device)
        n_s = int(np.searchsorted(offs, 256 << 20, side='right')) - 1
        s_off = offs[:n_s + 1].copy()
        s_text = h_np[:int(s_off[-1])]
        ms, Ts, kern, ids_s, ioff_s = device_encode_ms(toks, torch, s_text, s_off, dev=dev)
        ent = {'workload': 'first %d docs (%.1f MiB) of the headline corpus, pre_tokenizer Sequence[Split(\\p{N}{1,3}, Isolated), ByteLevel]' % (n_s, int(s_off[-1]) / 2**20),
               'device_ms': ms, 'MB_per_s': int(s_off[-1]) / (ms * 1e-3) / 1e6, 'tokens': int(Ts), 'kernels_ms': {k: v for k, v in kern.items() if v >= 0.01}}
        if not args.no_cpu:
            import c_oracle
            orcs = c_oracle.COracle.from_str(json.dumps(tjs))
            k = min(n_s, 60)
            wi, wo = orcs.encode_packed(s_text[:int(s_off[k])], s_off[:k + 1])
            ent['ids_equal_oracle'] = bool(np.array_equal(wo, ioff_s[:k + 1]) and np.array_equal(wi, ids_s[:int(ioff_s[k])]))
            ent['parity_sample_docs'] = k
        out['split_stage'] = ent
    except Exception as ex:
        out['split_stage'] = {'error': repr(ex)}
    return out


if __name__ == '__main__':
    main()
