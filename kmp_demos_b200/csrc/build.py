"""Builds libkmpx.so (all sm_100a CUDA kernels + the C ABI) in-tree with nvcc.

    python -m kmp_demos_b200.csrc.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ['api.cu', 'kernel_build.cu', 'factor_small.cu', 'factor_large.cu', 'dgemm.cu', 'dgemm_tma.cu', 'predict.cu', 'predict_fast.cu', 'predict_tab.cu', 'predict_pers.cu', 'predict_stream.cu',
           'inverse_gj.cu', 'gmm.cu', 'kmeans.cu', 'gmr.cu', 'quat.cu', 'incremental.cu', 'predict_gemm.cu']
HEADERS = ['common.cuh', 'dmma_tile.cuh', 'kmpx_internal.cuh', 'factor_tma.cuh', 'factor_struct.cuh', 'factor_wpp.cuh', 'tma_util.cuh', os.path.join('..', '..', 'include', 'kmpx.h')]
LIB = os.path.join(HERE, 'libkmpx.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _stale(target, deps):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(HERE, h) for h in HEADERS]
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(HERE, 'build', src.replace('.cu', '.o'))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)

    def compile_one(job):
        s, o = job
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f'--- {os.path.basename(s)}\n{r.stdout}{r.stderr}\n')
            if r.returncode != 0:
                raise RuntimeError(f'nvcc failed on {s}')
    if jobs or force or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link of libkmpx.so failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
