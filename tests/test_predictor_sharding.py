"""CPU tier: host logic either side of the hot path -- the 4-slice stack / slice-position arithmetic of
data/cmsr_dataset.py:121-151, the contiguous slice partition, and the world_size-2 (gloo) sharded volume run whose
gathered result must equal the single-rank result bit for bit (no collective on the data path)."""
import os
import socket

import numpy as np
import pytest
import torch

from afcm_b200.inference import slice_partition
from afcm_b200.predictor import VolumePredictor, build_stacks, stack_indices


def _ref_indices(idx, patch_count, thickness):
    """The reference arithmetic, restated literally (data/cmsr_dataset.py:131-136,151)."""
    idx_A = int((idx // thickness) * thickness)
    minus = idx_A - thickness if idx_A - thickness >= 0 else None
    plus = idx_A + thickness if idx_A + thickness <= patch_count - 1 else None
    plus2 = idx_A + thickness * 2 if idx_A + thickness * 2 <= patch_count - 1 else None
    return idx_A, [minus, idx_A, plus, plus2], np.array([idx - idx_A], dtype=np.float32) / thickness


@pytest.mark.parametrize('D,t', [(1, 1), (7, 1), (20, 5), (23, 5), (160, 5), (12, 3), (9, 2)])
def test_stack_indices_match_reference(D, t):
    for i in range(D):
        a, sl, c = stack_indices(i, D, t)
        ra, rsl, rc = _ref_indices(i, D, t)
        assert a == ra and sl == rsl
        assert np.float32(c) == rc[0]
    if t == 1:
        assert all(stack_indices(i, D, 1)[2] == 0 for i in range(D))


def test_partition_covers_every_slice_once():
    for D in (0, 1, 5, 64, 160, 161, 255):
        for world in (1, 2, 3, 4, 8):
            blocks = [slice_partition(D, world, r) for r in range(world)]
            flat = [i for lo, hi in blocks for i in range(lo, hi)]
            assert flat == list(range(D))
            assert max(hi - lo for lo, hi in blocks) == -(-D // world) or D == 0


def test_build_stacks_edges_and_positions():
    rng = np.random.RandomState(0)
    vol = rng.randint(1, 256, size=(11, 6, 8)).astype(np.uint8)
    x, c = build_stacks(vol, 0, 11, thickness=5)
    assert x.shape == (11, 4, 6, 8) and x.dtype == np.uint8
    assert (x[0, 0] == 0).all() and (x[0, 1] == vol[0]).all() and (x[0, 2] == vol[5]).all() and (x[0, 3] == vol[10]).all()
    assert (x[7, 0] == vol[0]).all() and (x[7, 1] == vol[5]).all() and (x[7, 2] == vol[10]).all() and (x[7, 3] == 0).all()
    np.testing.assert_allclose(c[:, 0], [0, .2, .4, .6, .8, 0, .2, .4, .6, .8, 0], rtol=1e-6)
    xf, _ = build_stacks(vol.astype(np.float32) / 255 * 2 - 1, 10, 11, thickness=1)
    assert (xf[0, 2] == -1).all() and (xf[0, 3] == -1).all()          # float volumes: raw zero == -1 after the transform


def _fake_run(z, c, x):
    # any deterministic per-slice function of (z, c, x): stands in for the generator in the CPU tier
    xf = x.float()
    return (xf[:, 1:2] * 0.5 + xf[:, 2:3] * 0.25 + xf[:, 0:1] * 0.125 + xf[:, 3:4] * 0.0625) / 255.0 \
        + c.reshape(-1, 1, 1, 1) + z[:, :1].reshape(-1, 1, 1, 1) * 1e-3


def _worker(rank, world, port, D, t, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        vol = np.random.RandomState(1).randint(0, 256, size=(D, 8, 8)).astype(np.uint8)
        pred = VolumePredictor(None, batch=3, rank=rank, world_size=world, run=_fake_run, device='cpu', z_dim=16)
        y, block = pred(vol, thickness=t, seed=7)
        full = pred.collect(y, block, D)
        if rank == 0:
            q.put(full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('D,t', [(13, 5), (8, 1)])
def test_two_rank_gloo_equals_single_rank(D, t):
    import torch.multiprocessing as mp
    vol = np.random.RandomState(1).randint(0, 256, size=(D, 8, 8)).astype(np.uint8)
    single, blk = VolumePredictor(None, batch=4, run=_fake_run, device='cpu', z_dim=16)(vol, thickness=t, seed=7)
    assert blk == (0, D)
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, D, t, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(full, single.numpy())
