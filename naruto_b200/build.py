"""Builds naruto_b200/libnaruto_b200.so (the C-ABI library of include/naruto_b200.h) with nvcc for sm_100a.

In-tree build: the .so sits next to this file so it travels with the repo snapshot to the GPU box.
    python -m naruto_b200.build [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libnaruto_b200.so')
STAMP = os.path.join(HERE, '.build_stamp')
SOURCES = ['api.cu', 'forward.cu', 'forward_tc.cu', 'forward_ws.cu', 'backward.cu', 'backward_tc.cu', 'backward_q.cu', 'optim.cu', 'peer.cu', 'sampler.cu', 'erp.cu', 'planner.cu', 'mcubes.cu', 'selftest.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-O3', '--expt-relaxed-constexpr']


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, 'include', 'naruto_b200.h')]
    for f in files:
        with open(f, 'rb') as fh:
            h.update(f.encode() + b'\0' + fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def build(force=False, verbose=False):
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for s in SOURCES:
        obj = os.path.join(HERE, 'build', s.replace('.cu', '.o'))
        cmd = [nvcc_path()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
              ['-I', os.path.join(ROOT, 'include'), '-I', CSRC, '-c', os.path.join(CSRC, s), '-o', obj]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f'--- nvcc {s} ---\n{out}\n')
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [nvcc_path(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    subprocess.check_call(cmd)
    with open(STAMP, 'w') as f:
        f.write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
