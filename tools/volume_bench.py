#!/usr/bin/env python
"""Volume-level throughput for BASELINE configs[2] / configs[3]: ADNI-style through-plane super-resolution (slice
thickness t, fractional slice position c) or CMSR over a synthetic volume, output slices sharded across the ranks of one
box with no data-path collective (afcm_b200/predictor.py).  Prints one JSON line on rank 0.

    python tools/volume_bench.py --slices 160 --thickness 5 --batch 32
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        tools/volume_bench.py --slices 256 --thickness 5 --batch 32

Timing: per rank, CUDA events around its whole block (host stack gather, uint8 upload, forward, fp32 read-back
included); the reported time is the max over ranks.  The final gather of the result on rank 0 is outside the timed
region (it is result collection, not the data path)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def synthetic_volume(D, seed=0):
    """Brain-like uint8 volume: smooth low-pass noise inside a centred ellipsoid, zero background."""
    rng = np.random.RandomState(seed)
    small = rng.rand(max(D // 8, 2), 32, 32).astype(np.float32)
    t = torch.nn.functional.interpolate(torch.from_numpy(small)[None, None], size=(D, 256, 256), mode='trilinear',
                                        align_corners=False)[0, 0].numpy()
    zz, yy, xx = np.meshgrid(np.linspace(-1, 1, D), np.linspace(-1, 1, 256), np.linspace(-1, 1, 256), indexing='ij')
    mask = (zz / 0.95) ** 2 + (yy / 0.8) ** 2 + (xx / 0.7) ** 2 <= 1
    return (np.clip(t, 0, 1) * 255 * mask).astype(np.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--slices', type=int, default=160)
    ap.add_argument('--thickness', type=float, default=5)
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--repeat', type=int, default=3)
    ap.add_argument('--precision', default='fast')
    args = ap.parse_args()
    import torch.distributed as dist
    from afcm_b200 import inference
    from afcm_b200.networks_stylegan3 import afcm_generator
    from afcm_b200.predictor import VolumePredictor
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)                                   # NCCL's version banner goes to stderr, stdout keeps the JSON line
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    inference.set_precision(args.precision)
    G = afcm_generator(seed=0, device=dev)
    t = int(args.thickness) if float(args.thickness).is_integer() else args.thickness
    vol = synthetic_volume(args.slices)
    pred = VolumePredictor(G, batch=args.batch, rank=rank, world_size=world)
    pred(vol, thickness=t, seed=0)                     # warm-up (weight preparation, tap caches)
    times = []
    for _ in range(args.repeat):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        y, blk = pred(vol, thickness=t, seed=0)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        times.append(float(ms.item()))
    full = pred.collect(y, blk, args.slices)
    if rank == 0:
        ms = float(np.median(times))
        print(json.dumps(dict(metric='volume_slices_per_sec_256x256', value=args.slices / (ms * 1e-3), unit='slices/s',
                              volumes_per_sec=1e3 / ms, n_gpus=world, ms_per_volume=ms,
                              config=dict(workload='synthetic volume, through-plane SR / CMSR inference', slices=args.slices,
                                          thickness=args.thickness, batch=args.batch, precision=args.precision,
                                          sharding='contiguous slice blocks, no data-path collective'),
                              checksum=float(full.double().abs().mean()))), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
