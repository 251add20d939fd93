"""Volume-level inference around the generator forward: the 4-slice input stack and the
fractional slice position `c` of every output slice, batching, slice sharding across ranks, result collection.

Reference behaviour restated here (paths relative to the reference repo):
  * data/cmsr_dataset.py:121-151 -- for output slice i and slice thickness t: idx_A = int((i // t) * t); the input is
    the stack of raw slices (idx_A - t, idx_A, idx_A + t, idx_A + 2t), a raw-zero slice where the index falls outside
    [0, D-1]; `slice_idx` c = (i - idx_A) / t.  t = 1 is plain cross-modality translation (c = 0), t > 1 is
    through-plane super-resolution, and t need not be an integer.
  * models/comodgan_model.py:101-108 -- z ~ N(0, I) per slice, c = slice_idx.
  * models/predictor.py:144-202 with data/utils.py:85-124 -- the test loader cuts every slice into patches of
    `patch_shape` (1 x 256 x 256) at `stride_shape` (1 x 32 x 32; the last patch of an axis is aligned to its end),
    each prediction loses `patch_halo` (0, 8, 8) voxels on the sides that are not volume borders (remove_halo,
    models/predictor.py:17-51), the remainders are ACCUMULATED into the prediction map while a normalisation mask counts
    the visits, and the map is divided by the mask at the end.  A 256 x 256 volume has exactly one patch per slice and no
    halo is removed, so there the procedure is a plain scatter (the fast path of VolumePredictor); larger in-plane
    sizes take predict_patches() below, which restates the accumulate / normalise procedure.
Slices are independent, so ranks own contiguous blocks of output slices (inference.slice_partition) and exchange
nothing while computing; the only communication is the final gather of results (`collect`), which is not on the data
path and works with any torch.distributed backend (gloo in the CPU tests, nccl on the GPUs).
"""
import numpy as np
import torch

from .inference import slice_partition


def stack_indices(i, num_slices, thickness):
    """-> (idx_A, [four raw slice indices or None], c) for output slice i (data/cmsr_dataset.py:131-136,151)."""
    t = thickness
    idx_a = int((i // t) * t)
    cand = [idx_a - t, idx_a, idx_a + t, idx_a + 2 * t]
    sl = []
    for k, j in enumerate(cand):
        if k == 0:
            ok = j >= 0
        elif k == 1:
            ok = True
        else:
            ok = j <= num_slices - 1
        sl.append(int(j) if ok else None)
    return idx_a, sl, np.float32(i - idx_a) / np.float32(t)


def build_stacks(volume, lo, hi, thickness=1):
    """Input stacks and slice positions for output slices [lo, hi) of a raw volume [D,H,W] (uint8 codes or float32
    already normalised).  Missing neighbours are raw zero -- for uint8 that is code 0, which the generator's input
    table maps to -1 exactly like the reference transform of a zero slice.  -> x [n,4,H,W], c [n,1] float32."""
    vol = np.asarray(volume)
    D, H, W = vol.shape
    n = max(hi - lo, 0)
    fill = 0 if vol.dtype == np.uint8 else -1.0
    x = np.full((n, 4, H, W), fill, dtype=vol.dtype)
    c = np.zeros((n, 1), dtype=np.float32)
    for r, i in enumerate(range(lo, hi)):
        _, sl, ci = stack_indices(i, D, thickness)
        for k, j in enumerate(sl):
            if j is not None:
                x[r, k] = vol[j]
        c[r, 0] = ci
    return x, c


def gen_indices(i, k, s):
    """Patch origins along one axis (data/utils.py:119-124): 0, s, 2s, ... while the patch fits, then one patch aligned to
    the end of the axis if the last regular one stops short of it."""
    assert i >= k, 'Sample size has to be bigger than the patch size'
    out = list(range(0, i - k + 1, s))
    if out[-1] + k < i:
        out.append(i - k)
    return out


def patch_grid(H, W, patch=(256, 256), stride=(32, 32)):
    """(y, x) origins of the in-plane patches of one slice, in the reference's iteration order (data/utils.py:100-116)."""
    return [(y, x) for y in gen_indices(H, patch[0], stride[0]) for x in gen_indices(W, patch[1], stride[1])]


def halo_window(start, size, full, pad):
    """One axis of remove_halo (models/predictor.py:22-36): the part of a patch [start, start + size) that is kept and where it
    lands in the volume.  -> (patch slice, volume slice).  A side that touches the volume border keeps its voxels."""
    stop = start + size
    p0, i0 = (0, 0) if start == 0 else (pad, start + pad)
    if stop == full:
        p1, i1 = size, full
    else:
        p1, i1 = (size - pad, stop - pad) if pad != 0 else (1, stop)        # (pad == 0 away from the border keeps ONE voxel: the
    return slice(p0, p1), slice(i0, i1)                                    #  reference's `-pad if pad != 0 else 1`)


def latents_for(lo, hi, seed, z_dim):
    """z ~ N(0, I) of output slices [lo, hi): a function of (seed, slice index, component) only -- a counter-based generator
    (splitmix64 hash -> two uniforms -> Box-Muller), vectorised over the whole block, so the values do not depend on how the
    volume is sharded or batched."""
    n = max(hi - lo, 0)
    if n == 0:
        return torch.empty([0, z_dim], dtype=torch.float32)
    idx = (np.arange(lo, hi, dtype=np.uint64)[:, None] * np.uint64(z_dim) + np.arange(z_dim, dtype=np.uint64)[None, :]) * np.uint64(2)

    def mix(v):
        with np.errstate(over='ignore'):
            v = v + np.uint64(0x9E3779B97F4A7C15) * (np.uint64(seed) + np.uint64(1))
            v = (v ^ (v >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            v = (v ^ (v >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return v ^ (v >> np.uint64(31))
    u1 = ((mix(idx) >> np.uint64(11)).astype(np.float64) + 1.0) / 9007199254740993.0          # (0, 1)
    u2 = (mix(idx + np.uint64(1)) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0   # [0, 1)
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return torch.from_numpy(z.astype(np.float32))


class VolumePredictor:
    """Runs the generator over this rank's block of output slices of a volume.

        pred = VolumePredictor(G, batch=32, rank=rank, world_size=world)
        y_local, (lo, hi) = pred(volume_u8, thickness=5, seed=0)
        full = pred.collect(y_local, (lo, hi), D)       # rank 0 gets [D,1,H,W], others None

    `run` is any callable (z, c, x) -> y on the generator's device: the module itself, or an
    inference.GraphedGenerator for fixed batch sizes."""

    def __init__(self, G, batch=32, rank=0, world_size=1, run=None, device=None, z_dim=None):
        self.G = G
        self.batch = int(batch)
        self.rank, self.world_size = int(rank), int(world_size)
        self.run = run if run is not None else (lambda z, c, x: G(z, c, x, noise_mode='const'))
        if device is None:
            device = next(G.parameters()).device if G is not None else torch.device('cpu')
        self.device = torch.device(device)
        self.z_dim = z_dim if z_dim is not None else getattr(G, 'z_dim', 512)

    def latents(self, lo, hi, seed):
        """z of every output slice, a function of (seed, slice index) only, so the result does not depend on how
        the volume is sharded."""
        return latents_for(lo, hi, int(seed), self.z_dim)

    def predict_patches(self, volume, thickness=1, seed=0, patch=(256, 256), stride=(32, 32), halo=(8, 8)):
        """The reference's patch-wise prediction of this rank's block of slices for volumes whose slices are larger than the
        generator's field (models/predictor.py:144-202): every slice is cut into patches (patch_grid), every patch of every
        slice goes through the generator (its 4-slice stack is cut at the same window), the halo is removed on the sides that
        are not volume borders, the remainders are accumulated and the visit counts normalise the sum.  All patches of a
        slice share the slice's z and c.  -> (y [n,1,H,W] float32, (lo, hi))."""
        vol = np.asarray(volume)
        D, H, W = vol.shape
        lo, hi = slice_partition(D, self.world_size, self.rank)
        n = max(hi - lo, 0)
        grid = patch_grid(H, W, patch, stride)
        acc = np.zeros((n, 1, H, W), dtype=np.float32)
        cnt = np.zeros((n, 1, H, W), dtype=np.uint8)
        x_full, c = build_stacks(vol, lo, hi, thickness)
        z = self.latents(lo, hi, seed)
        items = [(r, y0, x0) for r in range(n) for (y0, x0) in grid]
        with torch.no_grad():
            for b0 in range(0, len(items), self.batch):
                chunk = items[b0:b0 + self.batch]
                xb = np.stack([x_full[r, :, y0:y0 + patch[0], x0:x0 + patch[1]] for r, y0, x0 in chunk])
                rows = [r for r, _, _ in chunk]
                yb = self.run(z[rows].to(self.device), torch.from_numpy(c[rows]).to(self.device),
                              torch.from_numpy(xb).to(self.device)).float().cpu().numpy()
                for (r, y0, x0), pred in zip(chunk, yb):
                    py, iy = halo_window(y0, patch[0], H, halo[0])
                    px, ix = halo_window(x0, patch[1], W, halo[1])
                    acc[r, :, iy, ix] += pred[:, py, px]
                    cnt[r, :, iy, ix] += 1
        return torch.from_numpy(acc / np.maximum(cnt, 1)), (lo, hi)

    def __call__(self, volume, thickness=1, seed=0):
        D = int(np.asarray(volume).shape[0])
        lo, hi = slice_partition(D, self.world_size, self.rank)
        res = getattr(getattr(self.G, 'synthesis', None), 'img_resolution', None)
        if res is not None and tuple(np.asarray(volume).shape[1:]) != (res, res):
            # slices larger than the generator's field: the reference's patch / halo / averaging procedure
            return self.predict_patches(volume, thickness, seed, patch=(res, res))
        if self.device.type == 'cuda' and hi > lo:
            return self._run_pipelined(np.asarray(volume), lo, hi, thickness, seed), (lo, hi)
        x, c = build_stacks(volume, lo, hi, thickness)
        z = self.latents(lo, hi, seed)
        outs = []
        with torch.no_grad():
            for b0 in range(0, hi - lo, self.batch):
                b1 = min(b0 + self.batch, hi - lo)
                xb = torch.from_numpy(x[b0:b1]).to(self.device, non_blocking=True)
                cb = torch.from_numpy(c[b0:b1]).to(self.device, non_blocking=True)
                zb = z[b0:b1].to(self.device, non_blocking=True)
                outs.append(self.run(zb, cb, xb).float().cpu())
        H, W = np.asarray(volume).shape[1:]
        y = torch.cat(outs) if outs else torch.empty([0, 1, H, W])
        return y, (lo, hi)

    def _pinned(self, name, shape, dtype):
        """Pinned host staging buffers, kept between calls (pinning is a driver call of milliseconds)."""
        cache = self.__dict__.setdefault('_pin', {})
        key = (name, tuple(shape), dtype)
        if key not in cache:
            cache[key] = torch.empty(shape, dtype=dtype).pin_memory()
        return cache[key]

    def _run_pipelined(self, vol, lo, hi, thickness, seed):
        """GPU path: the host gathers the stacks of batch k+1 into pinned memory while the device runs batch k; uploads
        go through one copy stream, results come back through another into one pinned output block.  Nothing is
        synchronised between batches; the call returns after the last download.  The result lives in a pinned buffer that
        the next call with the same block shape reuses (clone it to keep it)."""
        n, B = hi - lo, self.batch
        H, W = vol.shape[1:]
        dev = self.device
        np_dt = torch.uint8 if vol.dtype == np.uint8 else torch.float32
        hx = self._pinned('x', [n, 4, H, W], np_dt)
        hc = self._pinned('c', [n, 1], torch.float32)
        hz = self._pinned('z', [n, self.z_dim], torch.float32)
        hy = self._pinned('y', [n, 1, H, W], torch.float32)
        cur = torch.cuda.current_stream(dev)
        if '_streams' not in self.__dict__:
            self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        up, down = self._streams
        up.wait_stream(cur)
        down.wait_stream(cur)
        keep = []                                            # device tensors stay alive until the final synchronisation
        with torch.no_grad():
            for b0 in range(0, n, B):
                b1 = min(b0 + B, n)
                x, c = build_stacks(vol, lo + b0, lo + b1, thickness)
                hx[b0:b1].copy_(torch.from_numpy(x)); hc[b0:b1].copy_(torch.from_numpy(c))
                hz[b0:b1].copy_(self.latents(lo + b0, lo + b1, seed))
                with torch.cuda.stream(up):
                    xb = hx[b0:b1].to(dev, non_blocking=True)
                    cb = hc[b0:b1].to(dev, non_blocking=True)
                    zb = hz[b0:b1].to(dev, non_blocking=True)
                    ev_up = torch.cuda.Event(); ev_up.record(up)
                cur.wait_event(ev_up)
                yb = self.run(zb, cb, xb).float()
                ev_y = torch.cuda.Event(); ev_y.record(cur)
                with torch.cuda.stream(down):
                    down.wait_event(ev_y)
                    hy[b0:b1].copy_(yb, non_blocking=True)
                keep.append((xb, cb, zb, yb))
        cur.wait_stream(down)
        down.synchronize()
        return hy                    # pinned; valid until the next call with a block of the same shape

    def collect(self, y_local, block, num_slices):
        """Gathers the per-rank blocks on rank 0 (plain scatter into the volume: every voxel is produced once)."""
        import torch.distributed as dist
        if self.world_size == 1:
            return y_local
        per = -(-num_slices // self.world_size)
        pad = torch.zeros([per] + list(y_local.shape[1:]), dtype=y_local.dtype)
        pad[:y_local.shape[0]] = y_local
        backend = dist.get_backend()
        buf = pad.to(self.device) if backend == 'nccl' else pad
        parts = [torch.empty_like(buf) for _ in range(self.world_size)] if self.rank == 0 else None
        dist.gather(buf, parts, dst=0)
        if self.rank != 0:
            return None
        out = torch.empty([num_slices] + list(y_local.shape[1:]), dtype=y_local.dtype)
        for r, part in enumerate(parts):
            lo, hi = slice_partition(num_slices, self.world_size, r)
            out[lo:hi] = part[:hi - lo].cpu()
        return out
