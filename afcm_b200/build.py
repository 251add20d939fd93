"""Builds afcm_b200/libafcm_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m afcm_b200.build [--force]

The library links the CUDA runtime statically and resolves the one driver-API symbol it needs
(cuTensorMapEncodeTiled) at run time through cudaGetDriverEntryPoint, so it loads on a machine
without a GPU driver (the CPU test tier checks the exported symbols there).
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, 'libafcm_b200.so')
OBJ = os.path.join(HERE, 'csrc', '_obj')

SOURCES = ['common.cu', 'filtered_lrelu.cu', 'flr_tc.cu', 'flr_tcs.cu', 'flr_t5.cu', 'upfirdn2d.cu', 'bias_act.cu', 'small_ops.cu', 'conv2d_simt.cu',
           'conv2d_tc.cu', 'conv2d_bwd.cu', 'conv2d_wgrad_tc5.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC'] + os.environ.get('AFCM_NVCC_EXTRA', '').split()     # e.g. -DAFCM_FTC_MINB22=4 (tuning)


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, 'rb') as f:
            h.update(p.encode()); h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h'))]
    deps.append(os.path.join(ROOT, 'include', 'afcm_b200.h'))
    return deps


def build(force=False, verbose=False):
    stamp = os.path.join(OBJ, 'stamp.txt')
    digest = _digest(_deps())
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == digest:
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        if verbose and r.stderr:
            print(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, '-shared', '-o', OUT] + objs + ['-cudart', 'static', '-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    with open(stamp, 'w') as f:
        f.write(digest)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
