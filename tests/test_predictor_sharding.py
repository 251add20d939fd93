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


# ---- patch-wise prediction: data/utils.py:85-124 + models/predictor.py:17-51,144-202 restated literally -------------------
def _ref_gen_indices(i, k, s):
    assert i >= k
    for j in range(0, i - k + 1, s):
        yield j
    if j + k < i:
        yield i - k


def _ref_remove_halo(patch, index, shape, patch_halo):
    def _new_slices(slicing, max_size, pad):
        if slicing.start == 0:
            p_start, i_start = 0, 0
        else:
            p_start, i_start = pad, slicing.start + pad
        if slicing.stop == max_size:
            p_stop, i_stop = None, max_size
        else:
            p_stop, i_stop = (-pad if pad != 0 else 1), slicing.stop - pad
        return slice(p_start, p_stop), slice(i_start, i_stop)
    D, H, W = shape
    i_c, i_z, i_y, i_x = index
    p_z, i_z = _new_slices(i_z, D, patch_halo[0])
    p_y, i_y = _new_slices(i_y, H, patch_halo[1])
    p_x, i_x = _new_slices(i_x, W, patch_halo[2])
    return patch[(slice(0, patch.shape[0]), p_z, p_y, p_x)], (i_c, i_z, i_y, i_x)


@pytest.mark.parametrize('H,W,patch,stride,halo', [(40, 56, 32, 8, 4), (32, 32, 32, 8, 4), (70, 33, 32, 16, 8), (48, 48, 32, 16, 0)])
def test_patchwise_prediction_matches_the_reference_procedure(H, W, patch, stride, halo):
    from afcm_b200.predictor import gen_indices, patch_grid
    assert gen_indices(H, patch, stride) == list(_ref_gen_indices(H, patch, stride))
    rng = np.random.RandomState(1)
    D = 5
    vol = rng.randint(0, 256, size=(D, H, W)).astype(np.uint8)
    pred = VolumePredictor(None, batch=7, run=_fake_run, device='cpu', z_dim=4)
    y, blk = pred.predict_patches(vol, thickness=1, seed=3, patch=(patch, patch), stride=(stride, stride), halo=(halo, halo))
    assert blk == (0, D) and y.shape == (D, 1, H, W)
    # the reference loops: accumulate un-haloed patches, count visits, divide
    x_full, c = build_stacks(vol, 0, D, 1)
    z = pred.latents(0, D, 3)
    pm = np.zeros((1, D, H, W), np.float32); nm = np.zeros((1, D, H, W), np.uint8)
    for zi in range(D):
        for y0 in _ref_gen_indices(H, patch, stride):
            for x0 in _ref_gen_indices(W, patch, stride):
                out = _fake_run(z[zi:zi + 1], torch.from_numpy(c[zi:zi + 1]),
                                torch.from_numpy(x_full[zi:zi + 1, :, y0:y0 + patch, x0:x0 + patch])).numpy()[0]      # [1,p,p]
                index = (slice(0, 1), slice(zi, zi + 1), slice(y0, y0 + patch), slice(x0, x0 + patch))
                u, ui = _ref_remove_halo(out[:, None], index, (D, H, W), (0, halo, halo))
                pm[ui] += u; nm[ui] += 1
    assert nm.min() >= 1
    np.testing.assert_allclose(y.numpy()[:, 0], (pm / nm)[0], rtol=1e-6, atol=1e-7)


def test_noninteger_thickness_matches_reference_arithmetic():
    """thickness 2.5: idx_A = int((i // t) * t) and c = (i - idx_A) / t as the reference computes them."""
    D, t = 23, 2.5
    for i in range(D):
        a, sl, c = stack_indices(i, D, t)
        ra, rsl, rc = _ref_indices(i, D, t)
        assert a == ra and np.float32(c) == rc[0]
        assert [None if v is None else int(v) for v in rsl] == sl


def test_latents_do_not_depend_on_the_sharding():
    pred = VolumePredictor(None, run=_fake_run, device='cpu', z_dim=512)
    full = pred.latents(0, 40, 7)
    parts = torch.cat([pred.latents(0, 13, 7), pred.latents(13, 14, 7), pred.latents(14, 40, 7)])
    assert torch.equal(full, parts)
    assert not torch.equal(full, pred.latents(0, 40, 8))
    assert abs(float(full.mean())) < 0.03 and abs(float(full.std()) - 1.0) < 0.03 and torch.isfinite(full).all()
