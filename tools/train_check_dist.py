#!/usr/bin/env python
"""Data-parallel training-step check under torchrun (NCCL): every rank runs the golden tiny-generator batch, the bucketed
all-reduce sums the gradients while backward is running, and after the 1/world average every rank must hold the
reference gradients (tests/golden/tiny_gen_grads.npz) and identical parameters after one fused Adam step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_check_dist.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TINY = dict(z_dim=64, c_dim=1, w_dim=64, img_resolution=32, mapping_layers=3, channel_base=512, channel_max=48,
            num_layers=6, skip_resolution=16)


def main():
    from afcm_b200.networks_stylegan3 import afcm_generator
    from afcm_b200.training import GeneratorTrainer
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'tiny_gen.npz'))
    gg = np.load(os.path.join(ROOT, 'tests', 'golden', 'tiny_gen_grads.npz'))
    G = afcm_generator(seed=0, device=None, **TINY)
    G.load_state_dict({k[2:]: torch.as_tensor(g[k]) for k in g.files if k.startswith('P.')}, strict=False)
    G = G.to(dev)
    tr = GeneratorTrainer(G, lr=1e-3, bucket_bytes=64 << 10)          # small buckets: several collectives in flight
    args = [torch.as_tensor(g[k], device=dev) for k in ('z', 'c', 'x')] + [torch.as_tensor(gg['target'], device=dev)]
    loss = tr.forward_backward(*args)
    worst = 0.0
    for n, p in G.named_parameters():
        ref = gg['G.' + n]
        if np.abs(ref).max() == 0:
            continue
        got = (p.grad / world).cpu().numpy()
        worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
    tr.opt.step(grad_scale=1.0 / world)
    chk = tr.flat.flat.double().sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(float(a) == float(allc[0]) for a in allc)
    res = dict(world=world, rank=rank, loss=float(loss), golden_loss=float(gg['loss']), worst_grad_rel_err=worst,
               buckets=len(tr.flat.buckets), collectives=len(tr.reducer.handles), params_identical_after_step=same)
    ok = worst < 2e-3 and same and abs(res['loss'] - res['golden_loss']) < 1e-5 and (world == 1 or res['collectives'] == res['buckets'])
    res['ok'] = bool(ok)
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
