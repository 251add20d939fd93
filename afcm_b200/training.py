"""Generator training step of the AFCM path (BASELINE config 5): forward + backward of the
generator on this rank's slices, data-parallel all-reduce of the generator gradients, fused Adam.

Reference: StyleGAN3Model.optimize_parameters (models/stylegan3_model.py:113-135) trains G on ONE GPU and wraps
only D / G_ema in nn.DataParallel; the gradient scrub and the Adam step live in train.py:67-77 and
models/base_model.py.  Here every rank owns a replica and a disjoint block of slices; the only exchange step of the
whole path is the gradient all-reduce, so that is the only collective (torch.distributed: NCCL over NVLink on the
GPU box, gloo in the CPU tests):

  * all parameters live in ONE flat fp32 buffer, all gradients in another, cut into buckets in REVERSE registration
    order (the order backward produces them);  `p.grad` is a view into its bucket, so autograd accumulates in place
    and nothing is copied before the collective;
  * a post-accumulate-grad hook per parameter counts a bucket down and launches its asynchronous all-reduce as soon
    as it is complete, overlapping the remaining backward kernels;
  * `step()` waits for the handles and runs one fused Adam kernel (afcm_adam_step) over the flat buffers, with the
    1/world_size average and the reference's nan_to_num gradient scrub folded in.
"""
import torch
import torch.distributed as dist

from . import _lib


def slice_batch_partition(global_batch, world_size, rank):
    """Contiguous block of a global batch owned by `rank`: blocks differ by at most one sample (the first
    `global_batch % world_size` ranks hold one more).  NB this is NOT inference.slice_partition, which hands out
    ceil(D / world) slices per rank.  With unequal blocks the plain 1/world average of per-rank MEAN losses that
    GeneratorTrainer applies is a mean of means; pass `loss_weight = local_n * world / global_n` to
    GeneratorTrainer.step() to obtain the global-batch mean gradient (bench.py and the tests use equal blocks)."""
    base, rem = divmod(int(global_batch), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatParams:
    """Moves the parameters of `module` into one flat fp32 buffer (views keep names and shapes) and gives every
    parameter a gradient view into a second flat buffer.  Order: reverse registration order, so that the first
    bucket is the one backward completes first."""

    def __init__(self, module, bucket_bytes=64 << 20):
        params = [p for p in module.parameters() if p.requires_grad]
        assert params, 'no trainable parameters'
        dev = params[0].device
        assert all(p.dtype == torch.float32 and p.device == dev for p in params)
        order = list(reversed(params))
        # 64-element alignment keeps every view 256-byte aligned
        offs, total = [], 0
        for p in order:
            offs.append(total)
            total += (p.numel() + 63) // 64 * 64
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.params, self.offsets = order, offs
        with torch.no_grad():
            for p, o in zip(order, offs):
                self.flat[o:o + p.numel()].copy_(p.reshape(-1))
                p.data = self.flat[o:o + p.numel()].view(p.shape)
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
        # buckets: [start, end) element ranges of the flat buffers + the parameters inside
        self.buckets = []
        start, members = 0, []
        per = max(1, int(bucket_bytes) // 4)
        for idx, (p, o) in enumerate(zip(order, offs)):
            members.append(idx)
            end = offs[idx + 1] if idx + 1 < len(order) else total
            if end - start >= per or idx + 1 == len(order):
                self.buckets.append((start, end, members))
                start, members = end, []
        self.bucket_of = {}
        for b, (_, _, mem) in enumerate(self.buckets):
            for idx in mem:
                self.bucket_of[idx] = b

    def zero_grad(self):
        self.grad.zero_()


class GradAllReducer:
    """Bucketed, overlapped gradient all-reduce (sum; the average is folded into the optimiser step)."""

    def __init__(self, flat, group=None):
        self.flat = flat
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.handles = []
        self.launched_bytes = 0
        self._pending = None
        self._hooks = []
        if self.world > 1:
            for idx, p in enumerate(flat.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(idx)))
        self.begin()

    def begin(self):
        """Call before backward: resets the per-bucket countdown."""
        self._pending = [len(mem) for (_, _, mem) in self.flat.buckets]
        self.handles = []
        self.launched_bytes = 0

    def _make_hook(self, idx):
        def hook(_p):
            b = self.flat.bucket_of[idx]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        s, e, _ = self.flat.buckets[b]
        buf = self.flat.grad[s:e]
        self.handles.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.launched_bytes += buf.numel() * 4
        self._pending[b] = -1

    def finish(self):
        """Call after backward: launches buckets that never completed (parameters without a gradient this step) and
        waits for every handle.  Returns the number of bytes reduced."""
        if self.world > 1:
            for b, left in enumerate(self._pending):
                if left >= 0:
                    self._launch(b)
            for h in self.handles:
                h.wait()
        return self.launched_bytes

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


class FusedAdam:
    """torch.optim.Adam semantics (no weight decay, no amsgrad) as ONE launch over the flat parameter buffer."""

    def __init__(self, flat, lr=0.0025, betas=(0.0, 0.99), eps=1e-8, scrub=True):
        self.flat = flat
        self.lr, self.betas, self.eps, self.scrub = float(lr), (float(betas[0]), float(betas[1])), float(eps), bool(scrub)
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        self.t = 0

    def step(self, grad_scale=1.0):
        self.t += 1
        f = self.flat
        _lib.require_cuda(f.flat)
        with torch.no_grad():
            _lib.check(_lib.lib().afcm_adam_step(_lib.ptr(f.flat), _lib.ptr(f.grad), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                                 f.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.t,
                                                 float(grad_scale), int(self.scrub), _lib.stream_ptr(f.flat.device)))
        # the kernel updated the weights behind autograd's version counters: drop the prepared-weight cache of the
        # inference path (conv2d_gradfix.prepare_weight keys on the version)
        from .torch_utils.ops import conv2d_gradfix
        conv2d_gradfix.clear_prepared_weights()


def ema_beta(batch_size, ema_kimgs=10, total_iters=None, ramp=None):
    """train.py:69-73: ema_nimg = ema_kimgs * 1000, optionally ramped with the iteration count; beta = 0.5 ** (batch / ema_nimg)."""
    ema_nimg = ema_kimgs * 1000
    if ramp is not None:
        ema_nimg = min(ema_nimg, total_iters * ramp)
    return 0.5 ** (batch_size / max(ema_nimg, 1e-8))


class GeneratorEMA:
    """G_ema of the reference (models/comodgan_model.py:16, train.py:67-77): a deep copy of the generator in eval mode whose
    parameters follow p_ema = lerp(p, p_ema, beta) after every optimizer step and whose buffers are copied.  Both parameter
    sets live in flat fp32 buffers (FlatParams), so the update is ONE kernel (afcm_ema_lerp) over 58.5 M values."""

    def __init__(self, G, flat):
        import copy
        assert isinstance(flat, FlatParams)
        self.src, self.src_flat = G, flat
        self.G_ema = copy.deepcopy(G).eval()
        for p in self.G_ema.parameters():
            p.requires_grad_(True)                    # FlatParams collects the trainable parameters: same order as the source
        self.flat = FlatParams(self.G_ema, 64 << 20)
        self.flat.flat.copy_(flat.flat)
        for p in self.G_ema.parameters():
            p.requires_grad_(False)
        assert self.flat.flat.numel() == flat.flat.numel()

    def update(self, beta):
        _lib.require_cuda(self.flat.flat)
        with torch.no_grad():
            _lib.check(_lib.lib().afcm_ema_lerp(_lib.ptr(self.flat.flat), _lib.ptr(self.src_flat.flat), self.flat.flat.numel(),
                                                float(beta), _lib.stream_ptr(self.flat.flat.device)))
            for b_ema, b in zip(self.G_ema.buffers(), self.src.buffers()):           # train.py:76-77
                b_ema.copy_(b)
        from .torch_utils.ops import conv2d_gradfix
        conv2d_gradfix.clear_prepared_weights()       # the kernel moved the weights behind autograd's version counters


class GeneratorTrainer:
    """One data-parallel generator training step: loss = mean |G(z, c, x) - target| (L1, the reconstruction term of
    models/stylegan3_model.py:124-131; the discriminator is outside the north-star path)."""

    def __init__(self, G, lr=0.0025, betas=(0.0, 0.99), bucket_bytes=64 << 20, group=None):
        self.G = G
        for p in G.parameters():
            p.requires_grad_(True)
        self.flat = FlatParams(G, bucket_bytes)
        self.reducer = GradAllReducer(self.flat, group)
        self.opt = FusedAdam(self.flat, lr=lr, betas=betas)
        self.world = self.reducer.world

    def forward_backward(self, z, c, x, target, loss_weight=1.0):
        self.flat.zero_grad()
        self.reducer.begin()
        y = self.G(z, c, x, noise_mode='const')
        loss = (y - target).abs().mean()
        (loss * loss_weight if loss_weight != 1.0 else loss).backward()
        self.reducer.finish()
        return loss.detach()

    def step(self, z, c, x, target, loss_weight=1.0):
        """`loss_weight`: local_n * world / global_n when the ranks hold unequal blocks (see slice_batch_partition)."""
        loss = self.forward_backward(z, c, x, target, loss_weight)
        self.opt.step(grad_scale=1.0 / self.world)
        return loss
