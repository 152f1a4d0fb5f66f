#!/usr/bin/env python3
"""Kernel A/B experiments: build libctk variants with extra -D flags (here, on the CPU box), time them on the GPU box.

    python tools/variants.py build name=-DCTK_LB=5 name2="-DCTK_LB=5 -DCTK_PREFETCH=0" ...   # -> variants/libctk_<name>.so
    python tools/variants.py run [bytes]                                                      # on the GPU: bench each variant
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'complexity-tokenizer_b200')
VAR = os.path.join(PKG, 'variants')
sys.path.insert(0, PKG)

def build(specs):
    import build as b
    os.makedirs(VAR, exist_ok=True)
    for spec in specs:
        name, flags = spec.split('=', 1)
        objs = []
        procs = []
        for s in b.SOURCES:
            o = os.path.join(VAR, '%s_%s.o' % (name, s))
            objs.append(o)
            procs.append(subprocess.Popen([b.NVCC] + b.FLAGS + flags.split() + ['-c', os.path.join(b.CSRC, s), '-o', o],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        for pr in procs:
            out, _ = pr.communicate()
            if pr.returncode:
                print(out); raise SystemExit('nvcc failed for ' + name)
            for ln in out.splitlines():
                if 'k_encode_slices' in ln: want = True
                elif 'Used' in ln and locals().get('want'): print(name, ln.strip()); want = False
                elif 'spill' in ln and locals().get('want'): print(name, ln.strip())
        subprocess.check_call([b.NVCC, '-shared', '-o', os.path.join(VAR, 'libctk_%s.so' % name)] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
        for o in objs: os.remove(o)

def run(nbytes):
    libs = sorted(f for f in os.listdir(VAR) if f.endswith('.so'))
    for f in [None] + libs:
        env = dict(os.environ)
        if f: env['CTK_LIB_VARIANT'] = os.path.join(VAR, f)
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '5', '--warmup', '3', '--no-cpu', '--no-extras', '--e2e-steps', '1', '--bytes', str(nbytes)],
                             env=env, capture_output=True, text=True)
        for l in out.stdout.splitlines():
            if l.startswith('{'):
                d = json.loads(l)
                print('%-28s step %.3f ms  %s' % (f or 'libctk.so (default)', d['ms_per_step'],
                      {k: round(v, 3) for k, v in d['roofline']['all_kernels_ms_per_step'].items() if v > 0.05}), flush=True)
                break
        else:
            print(f, 'FAILED', out.stderr[-400:])

if __name__ == '__main__':
    if sys.argv[1] == 'build': build(sys.argv[2:])
    else: run(int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 30)
