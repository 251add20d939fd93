"""Inference-side host runtime of the hot path: precision selection, CUDA-graph replay of the generator
forward, and the slice partition used to shard a volume (or a batch of slices) across the GPUs of one box.

The reference runs inference slice by slice through `StyleGAN3Model.forward` -> `netG_ema(z, c, real_A)`
(models/stylegan3_model.py:67-86, evaluate.py:43-104) on one device (`nn.DataParallel` at most).  Here
every rank owns a contiguous block of slices and runs the same generator on it; nothing crosses NVLink on
the data path.
"""
import torch

from .torch_utils.ops import conv2d_gradfix

PRECISIONS = ('fp32', 'tc', 'fast')


def set_precision(mode):
    """'fp32': exact SIMT kernels everywhere (parity path, <= 1e-4 of the reference).
    'tc'  : tcgen05 convolutions with fp16 operands / fp32 accumulation, fp32 activation storage, exact
            filtered_lrelu.
    'fast': 'tc' + tensor-core filtered_lrelu + fp16 activation storage between the operators (the
            benchmarked inference path; tolerance stated in tests/test_gpu_generator.py)."""
    assert mode in PRECISIONS, mode
    if mode == 'fp32':
        conv2d_gradfix.set_conv_impl('f32')
    elif mode == 'tc':
        conv2d_gradfix.set_conv_impl('tc', torch.float16)
    else:
        conv2d_gradfix.set_conv_impl('tc', torch.float16, act=torch.float16)


def get_precision():
    if conv2d_gradfix.conv_impl == 'f32':
        return 'fp32'
    return 'fast' if conv2d_gradfix.fast_path() else 'tc'


class GraphedGenerator:
    """Replays the generator forward for a fixed batch size as ONE CUDA graph launch (~250 kernel launches
    per forward otherwise; at small batches the forward is launch-bound).  Inputs are copied into static
    device buffers on the current stream, the graph is replayed, the static output is returned (valid until
    the next call).  Host-side filter taps, tensor maps and the addresses of the prepared weights are baked into the
    captured kernel parameters: the runner re-captures by itself when the precision changes, when a parameter has been
    updated in place (optimizer step, load_state_dict) or when the prepared-weight cache has released anything
    (conv2d_gradfix.prep_generation()), so a replay never reads freed or stale weights."""

    def __init__(self, G, batch, device=None, noise_mode='const', warmup=2, input_dtype=torch.float32):
        self.G = G
        self.batch = int(batch)
        p = next(G.parameters())
        self.device = torch.device(device) if device is not None else p.device
        self.noise_mode = noise_mode
        self.warmup = warmup
        self.z = torch.zeros([self.batch, G.z_dim], dtype=torch.float32, device=self.device)
        self.c = torch.zeros([self.batch, max(G.c_dim, 1)], dtype=torch.float32, device=self.device)
        S = G.synthesis
        assert input_dtype in (torch.float32, torch.uint8)      # uint8 codes are normalised by the pad kernel
        self.x = torch.zeros([self.batch, S.img_channels_in, S.img_resolution, S.img_resolution], dtype=input_dtype,
                             device=self.device)
        self.y = None
        self.graph = None
        self.precision = None
        self.kernels_per_replay = 0
        self._weights_tag = None

    def _tag(self):
        # parameter versions move on every in-place update; the generation moves on cache clears / evictions / re-preparation
        return (conv2d_gradfix.prep_generation(), tuple(p._version for p in self.G.parameters()))

    def capture(self):
        self.precision = get_precision()
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):               # JIT-free, but fills the weight / tap caches before capture
                self.G(self.z, self.c, self.x, noise_mode=self.noise_mode)
        torch.cuda.current_stream(self.device).wait_stream(side)
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.y = self.G(self.z, self.c, self.x, noise_mode=self.noise_mode)
        self.kernels_per_replay = _lib.launch_count() - n0        # library kernels recorded in the graph
        self._weights_tag = self._tag()
        return self

    def __call__(self, z, c, x):
        if self.graph is None or self.precision != get_precision() or self._weights_tag != self._tag():
            self.capture()
        assert z.shape[0] == self.batch, f'graph captured for batch {self.batch}, got {z.shape[0]}'
        self.z.copy_(z, non_blocking=True)
        self.c.copy_(c.reshape(self.batch, -1), non_blocking=True)
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y


class PipelinedGenerator:
    """Host-to-host inference with the copies hidden behind the forward: batches arrive in pinned HOST buffers and results
    leave into pinned HOST buffers, like the reference's per-slice loop (evaluate.py:43-104), but the upload of batch k+1
    (copy stream) and the download of batch k-1 (second copy stream) run while the graph of batch k replays.  Two device
    staging sets per direction; events order upload -> forward -> download per set, so a set is reused only after its
    consumer has finished.  `submit()` returns immediately; `finish()` makes the current stream wait for every download."""

    def __init__(self, runner, depth=2):
        assert isinstance(runner, GraphedGenerator) and depth >= 2
        self.runner, self.depth, self.k = runner, depth, 0
        dev = runner.device
        self.up, self.down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        mk = lambda t: [torch.empty_like(t) for _ in range(depth)]
        self.sz, self.sc, self.sx = mk(runner.z), mk(runner.c), mk(runner.x)
        if runner.graph is None:
            runner.capture()
        self.sy = mk(runner.y)
        ev = lambda: [torch.cuda.Event() for _ in range(depth)]
        self.uploaded, self.consumed, self.computed, self.downloaded = ev(), ev(), ev(), ev()

    def submit(self, hz, hc, hx, hy):
        """hz, hc, hx: pinned host inputs of one batch; hy: pinned host buffer that receives the result."""
        r, i = self.runner, self.k % self.depth
        cur = torch.cuda.current_stream(r.device)
        with torch.cuda.stream(self.up):
            self.up.wait_event(self.consumed[i])                 # the forward that read this staging set is done
            self.sz[i].copy_(hz, non_blocking=True)
            self.sc[i].copy_(hc.reshape(r.batch, -1), non_blocking=True)
            self.sx[i].copy_(hx, non_blocking=True)
            self.uploaded[i].record(self.up)
        cur.wait_event(self.uploaded[i])
        y = r(self.sz[i], self.sc[i], self.sx[i])                # device-to-device into the graph's buffers + replay
        self.consumed[i].record(cur)
        cur.wait_event(self.downloaded[i])                       # the previous result of this set has left the device
        self.sy[i].copy_(y, non_blocking=True)
        self.computed[i].record(cur)
        with torch.cuda.stream(self.down):
            self.down.wait_event(self.computed[i])
            hy.copy_(self.sy[i], non_blocking=True)
            self.downloaded[i].record(self.down)
        self.k += 1

    def finish(self):
        cur = torch.cuda.current_stream(self.runner.device)
        cur.wait_stream(self.down)
        cur.wait_stream(self.up)


def slice_partition(num_slices, world_size, rank):
    """Contiguous block [lo, hi) of slice indices owned by `rank` (ceil(D / world) per rank)."""
    assert 0 <= rank < world_size and num_slices >= 0
    per = -(-num_slices // world_size)
    lo = min(rank * per, num_slices)
    return lo, min(lo + per, num_slices)
