"""In-tree build of the sm_100a shared library (plain nvcc; no torch C++ ABI involved).

``python -m samplernn_pase_b200._build`` or ``__graft_entry__.build()``.  The resulting
``libsrnn_b200.so`` sits next to this file, is git-ignored and travels to the GPU box with
the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libsrnn_b200.so')
STAMP = os.path.join(HERE, '.libsrnn_b200.stamp')
SOURCES = ['core.cu', 'gemm.cu', 'gru.cu', 'elementwise.cu', 'generate.cu', 'precise.cu']
FLAGS = ['-std=c++17', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC',
         '--cudart', 'shared'] + os.environ.get('SRNN_NVCC_EXTRA', '').split()


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for name in sorted(os.listdir(root)):
            if name.endswith(('.cu', '.cuh', '.h')):
                with open(os.path.join(root, name), 'rb') as f:
                    h.update(name.encode() + f.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f'nvcc failed on {src}')
    cmd = [nvcc, '-shared', '--cudart', 'shared', '-o', LIB] + objs
    subprocess.check_call(cmd)
    with open(STAMP, 'w') as f:
        f.write(digest)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
