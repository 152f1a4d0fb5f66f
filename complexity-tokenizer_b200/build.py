"""Build libctk.so (CUDA kernels + C ABI) in-tree for sm_100a.

    python complexity-tokenizer_b200/build.py [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'complexity_tokenizer', 'libctk.so')
SOURCES = ['api.cu', 'host_api.cu', 'encode_fused.cu', 'encode_general.cu', 'decode.cu', 'decode_clean.cu', 'nfc.cu', 'encoding.cu', 'train.cu', 'split.cu', 'metaspace.cu', 'loader.cpp', 'regex_dfa.cpp']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
         '-Xcompiler', '-fPIC,-Wall,-Wno-unused-function', '-Xptxas', '-v']


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for root, _, files in os.walk(CSRC):
        for f in files:
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    inc = os.path.join(os.path.dirname(HERE), 'include', 'ctk.h')
    return os.path.getmtime(inc) > t


def build(force=False, verbose=False):
    if not force and not needs_build():
        build_marshal()
        return OUT
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, 'build', s + '.o')
        objs.append(o)
        cmd = [NVCC] + FLAGS + ['-c', os.path.join(CSRC, s), '-o', o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append('== %s\n%s' % (s, out))
        if p.returncode != 0:
            sys.stderr.write('\n'.join(log))
            raise RuntimeError('nvcc failed on ' + s)
    with open(os.path.join(HERE, 'build', 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    subprocess.check_call([NVCC, '-shared', '-o', OUT] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    build_marshal()
    return OUT


def build_marshal():
    """_ctk_marshal: CPython helpers for list[str] <-> packed buffers (csrc/marshal.c); marshalling only."""
    import sysconfig
    out = os.path.join(HERE, 'complexity_tokenizer', '_ctk_marshal' + (sysconfig.get_config_var('EXT_SUFFIX') or '.so'))
    src = os.path.join(CSRC, 'marshal.c')
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    subprocess.check_call([os.environ.get('CC', 'gcc'), '-O2', '-shared', '-fPIC', '-Wall', '-I' + sysconfig.get_paths()['include'], src, '-o', out])
    return out


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
