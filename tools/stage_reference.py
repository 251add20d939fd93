#!/usr/bin/env python
"""Stages the UNMODIFIED reference tree under baseline/_ref/AFCM (git-ignored; travels to the GPU box with gpurun) so that
GPU-side tools and tests can import the reference's own Python modules: tests/test_gpu_reference_network.py (the boundary
proof: the reference's networks_stylegan3.py running on this library's operators) and tools/ref_cuda_bench.py (the reference's
own CUDA plugins + cuDNN timed on the B200 as the kernel to beat).  Nothing under afcm_b200/ imports it.

    python tools/stage_reference.py [/root/reference]
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, 'baseline', '_ref', 'AFCM')


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    if not os.path.isdir(src):
        print('reference tree not found at', src)
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(src, DST, ignore=shutil.ignore_patterns('.git', '__pycache__', '*.pyc', '*.h5', '*.nii.gz', '*.pth'))
    n = sum(len(f) for _, _, f in os.walk(DST))
    print(f'staged {n} files under {DST}')
    return 0


if __name__ == '__main__':
    sys.exit(main())
