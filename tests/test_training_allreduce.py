"""CPU tier (gloo, world_size 2): the data-parallel gradient path of afcm_b200/training.py -- flat parameter / gradient
buffers, bucketed all-reduce launched from post-accumulate-grad hooks while backward is still running -- gives every
rank the gradient of the GLOBAL batch: sum over ranks of the per-rank gradients == the single-process gradient over the
concatenated batch (torch CPU model standing in for the generator; the CUDA kernels are not involved here)."""
import os
import socket

import numpy as np
import pytest
import torch

from afcm_b200.training import FlatParams, GradAllReducer, slice_batch_partition


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(12, 33), torch.nn.LeakyReLU(0.2), torch.nn.Linear(33, 17),
                               torch.nn.LeakyReLU(0.2), torch.nn.Linear(17, 5))


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(10, 12, generator=g), torch.randn(10, 5, generator=g)


def test_partition():
    for B in (1, 7, 32, 64):
        for w in (1, 2, 3, 8):
            blocks = [slice_batch_partition(B, w, r) for r in range(w)]
            assert [i for lo, hi in blocks for i in range(lo, hi)] == list(range(B))


def test_flat_params_alias_and_buckets():
    m = _model()
    ref = [p.detach().clone() for p in m.parameters()]
    flat = FlatParams(m, bucket_bytes=1024)
    assert len(flat.buckets) > 1
    covered = sorted((s, e) for s, e, _ in flat.buckets)
    assert covered[0][0] == 0 and covered[-1][1] == flat.flat.numel()
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    for p, r in zip(m.parameters(), ref):
        assert torch.equal(p, r)
        assert p.data_ptr() >= flat.flat.data_ptr() and p.grad.data_ptr() >= flat.grad.data_ptr()
    x, t = _data()
    (m(x) - t).abs().mean().backward()
    assert flat.grad.abs().sum() > 0                       # autograd accumulated straight into the flat buffer
    flat.flat.mul_(0.5)                                    # an update of the flat buffer is an update of the model
    for p, r in zip(m.parameters(), ref):
        assert torch.allclose(p, r * 0.5)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        m = _model()
        flat = FlatParams(m, bucket_bytes=1024)
        red = GradAllReducer(flat)
        x, t = _data()
        lo, hi = slice_batch_partition(x.shape[0], world, rank)
        for _ in range(2):                                  # two steps: the countdown resets correctly
            flat.zero_grad()
            red.begin()
            # per-rank loss = sum over the rank's samples / GLOBAL batch: the rank gradients then sum to the global one
            ((m(x[lo:hi]) - t[lo:hi]).abs().sum() / (x.shape[0] * t.shape[1])).backward()
            nbytes = red.finish()
        assert nbytes == flat.grad.numel() * 4 and len(red.handles) == len(flat.buckets)
        q.put((rank, flat.grad.clone().numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_global_gradient():
    import torch.multiprocessing as mp
    m = _model()
    flat = FlatParams(m, bucket_bytes=1024)
    x, t = _data()
    (m(x) - t).abs().mean().backward()
    want = flat.grad.clone().numpy()
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got[0], got[1])                   # every rank holds the same reduced gradient
    assert np.allclose(got[0], want, rtol=1e-5, atol=1e-7)
