"""Build libfds_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m fds_b200.build [--force] [-v]

Every .cu is compiled to its own object (in parallel, only when it or a header changed) and the
objects are linked into the shared library next to this file.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, 'build')
SOURCES = ['l96_step.cu', 'boundary.cu', 'boundary_fast.cu', 'boundary_mma.cu', 'lss.cu', 'toy_ode.cu', 'diag.cu', 'ctx.cu']
OUT = os.path.join(HERE, 'libfds_b200.so')
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC']


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))] + \
           [os.path.join(os.path.dirname(HERE), 'include', 'fds_b200.h')]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', 'nvcc')
    hdr = _headers()
    if not force and not _newer(OUT, [os.path.join(CSRC, s) for s in SOURCES] + hdr):
        return OUT                      # up to date (the objects need not exist: the GPU box gets the .so only)
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + '.o')
        if force or _newer(obj, [src] + hdr):
            jobs.append([nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj])
    if jobs:
        with ThreadPoolExecutor(max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
            list(pool.map(subprocess.check_call, jobs))
    objs = [os.path.join(OBJ, s[:-3] + '.o') for s in SOURCES]
    if force or jobs or _newer(OUT, objs):
        subprocess.check_call([nvcc, '-shared', '-Wno-deprecated-gpu-targets', '-o', OUT] + objs + ['-ldl'])
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
