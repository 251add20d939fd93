#!/usr/bin/env python
"""bench.py -- AFCM generator forward throughput (slices/s @256^2) on B200, per the driver contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one generator forward (mapping + 14 encoder layers + co-modulation code + 15 synthesis
layers) over one batch of synthetic slices: BASELINE.json configs[1], "IXI T1->T2 synthesis generator
forward, synthetic 256x256 slices, batch 64, 1x B200".  For N > 1 every rank runs its own batch of 64
slices (weak scaling, slice-sharded, no data-path collective -- SURVEY.md 8(e)).

Own arm (default): every kernel on the path is hand-written sm_100a code reached through
libafcm_b200.so; convolutions run on the tcgen05/TMEM implicit GEMM with fp16 operands and fp32
accumulation.  `value` = slices/s with inputs resident in HBM; `e2e` = the same through the public
generator call with pinned HOST buffers (H2D of cond_img/z/c and D2H of the result inside the timed
region).  `roofline` is the dominant kernel (the implicit-GEMM convolution, tensor bound); `rooflines`
adds filtered_lrelu against the measured HBM copy bandwidth.  `cpu_baseline` is the oracle port of the
reference's CPU `_ref` path on this box's host cores (rank 0, N=1 only, bounded sample).

Reference arm (--impl reference): times that CPU port alone with all host threads, one slice per step.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'generator_slices_per_sec_256x256'
UNIT = 'slices/s'
WORKLOAD = 'IXI T1->T2 synthesis generator forward, synthetic 256x256 slices, batch 64 per GPU'


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d['hbm_gbs']), tf_burst=float(d['bf16_tflops']),
                    tf_sustained=float(d.get('bf16_tflops_sustained', d['bf16_tflops'])), source='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback')


def synthetic_inputs(batch, seed=0, as_uint8=False):
    """uint8-quantised slices mapped to [-1,1] like data/augment/transforms.py:604-616 of the reference;
    4-slice stacks; fractional slice position c; z ~ N(0,1).  as_uint8: return the codes themselves (the
    generator normalises them on the GPU with the same float64-derived table)."""
    import numpy as np
    import torch
    rng = np.random.RandomState(seed)
    u8 = rng.randint(0, 256, size=(batch, 4, 256, 256)).astype(np.uint8)
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(batch, 512, generator=g)
    c = torch.zeros(batch, 1)
    if as_uint8:
        return z, c, torch.from_numpy(u8)
    x = np.clip(2.0 * u8.astype(np.float64) / 255.0 - 1.0, -1, 1).astype(np.float32)
    return z, c, torch.from_numpy(x)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons during the timed region (NVML, falls back to nvidia-smi)."""

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._halt = threading.Event()
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    NAMES = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown',
             0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
             0x80: 'hw_power_brake_slowdown', 0x100: 'display_clock_setting'}

    def _sample(self):
        if self._h is not None:
            nv = self._nv
            self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
            try:
                fn = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = int(fn(self._h))
                for bit, name in self.NAMES.items():
                    if r & bit and name != 'gpu_idle':
                        self.reasons.add(name)
            except Exception:
                pass
        else:
            import subprocess
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=clocks.sm,clocks.max.sm,'
                                      'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
                                      'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap',
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                f = [t.strip() for t in out.strip().split(',')]
                self.samples.append(int(f[0])); self.sm_max = int(f[1])
                for name, val in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[2:]):
                    if val.lower().startswith('active'):
                        self.reasons.add(name)
            except Exception:
                pass

    def run(self):
        while not self._halt.is_set():
            self._sample()
            self._halt.wait(self.period)

    def finish(self):
        self._halt.set()
        self.join(timeout=5)
        import statistics
        return dict(sm_mhz=(statistics.median(self.samples) if self.samples else None), sm_max_mhz=self.sm_max,
                    reasons=sorted(self.reasons), samples=len(self.samples))


def cpu_reference_slices_per_sec(steps, warmup, seed=0):
    """The oracle's torch-CPU restatement of the reference `_ref` generator forward (oracle/afcm_oracle.py
    generator_forward), all host threads, ONE slice per step.  Returns (slices/s, ms/step, cores, sample)."""
    import torch
    from oracle import afcm_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = orc.init_params(seed=0)
    z, c, x = synthetic_inputs(1, seed)
    for _ in range(warmup):
        orc.generator_forward(P, z, c, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.generator_forward(P, z, c, x)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return 1.0 / dt, dt * 1e3, cores, f'{steps} timed + {warmup} warm-up forwards of 1 slice (batch 1) of the same workload'


def cpu_reference_train_slices_per_sec(seed=0):
    """The oracle's torch-CPU restatement differentiated by torch autograd: one forward + backward of the L1 training loss
    on ONE slice, all host threads (bounded sample of the training workload)."""
    import torch
    from oracle import afcm_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and not k.endswith(('_filter', 'magnitude_ema', 'w_avg')) else v)
         for k, v in orc.init_params(seed=0).items()}
    z, c, x = synthetic_inputs(1, seed)
    tgt = torch.rand(1, 1, 256, 256, generator=torch.Generator().manual_seed(100)) * 2 - 1
    t0 = time.perf_counter()
    (orc.generator_forward(P, z, c, x, grad=True) - tgt).abs().mean().backward()
    dt = time.perf_counter() - t0
    return 1.0 / dt, cores, '1 forward + backward of 1 slice (batch 1) of the same workload, no warm-up'


def run_reference(args, rank):
    if rank != 0:
        return
    sps, ms, cores, sample = cpu_reference_slices_per_sec(args.steps, args.warmup)
    line = dict(metric=METRIC, value=sps, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference',
                config=dict(workload=WORKLOAD, batch_per_gpu=64, global_batch=64 * args.gpus, resolution=256, precision='fp32',
                            sample='one slice of the batch per step (bounded sample of the same workload)',
                            note='CPU port of the reference _ref path'),
                cpu_baseline=dict(value=sps, unit=UNIT, cores=cores, kind='port', sample=sample),
                e2e=dict(value=sps, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def golden_check(forward, dz, dc, dx, dev, precision):
    """Parity of the benchmarked call itself: slices 0-1 of the batch are replaced by the inputs of the committed golden file
    (tests/golden/full_gen.npz: the reference's own fp32 `_ref` forward at B = 2, seeded weights == afcm_generator(seed=0)), the
    batch runs through the SAME callable that is timed (graph replay at the benchmark batch size), and rows 0-1 of the result
    are compared with the golden output.  Bounds as in tests/test_gpu_generator.py: fp32 path 1e-4, 16-bit paths 6e-3 / 58 dB."""
    import numpy as np
    import torch
    path = os.path.join(ROOT, 'tests', 'golden', 'full_gen.npz')
    if not os.path.exists(path) or dz.shape[0] < 2:
        return dict(checked=False, why='golden file missing or batch < 2')
    g = np.load(path)
    z, c, x = dz.clone(), dc.clone(), dx.clone()
    z[:2] = torch.as_tensor(g['z'], device=dev)
    c[:2] = torch.as_tensor(g['c'], device=dev).reshape(2, -1)
    x[:2] = torch.as_tensor(g['x_u8'], device=dev).to(x.dtype) if x.dtype == torch.uint8 else \
        (torch.as_tensor(g['x_u8']).float() * (2.0 / 255.0) - 1.0).clamp(-1, 1).to(dev)
    y = forward(z, c, x)[:2].float().cpu().numpy().astype(np.float64)
    ref = g['y'].astype(np.float64)
    err = float(np.abs(y - ref).max() / np.abs(ref).max())
    psnr = float(10 * np.log10(np.abs(ref).max() ** 2 / max(np.mean((y - ref) ** 2), 1e-300)))
    tol = 1e-4 if precision == 'fp32' else 6e-3
    ok = err < tol and (precision == 'fp32' or psnr > 58.0)
    if not ok:
        raise RuntimeError(f'bench.py: the benchmarked forward does not match the golden output: rel err {err:.3e} (bound {tol:g}), '
                           f'PSNR {psnr:.1f} dB')
    return dict(checked=True, against='tests/golden/full_gen.npz (reference _ref forward, B=2) as slices 0-1 of the timed batch',
                rel_err=err, psnr_db=psnr, bound=tol)


def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    from afcm_b200 import _lib, inference
    from afcm_b200.networks_stylegan3 import afcm_generator

    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    _lib.check(_lib.lib().afcm_device_check())
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there when the communicator is created
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    inference.set_precision(args.precision)
    G = afcm_generator(seed=0, device=dev)
    z, c, x = synthetic_inputs(B, seed=rank, as_uint8=True)       # slices travel as bytes, normalised on the GPU
    hz, hc, hx = z.pin_memory(), c.pin_memory(), x.pin_memory()
    hy = torch.empty([B, 1, 256, 256], dtype=torch.float32).pin_memory()
    dz, dc, dx = hz.to(dev), hc.to(dev), hx.to(dev)

    runner = inference.GraphedGenerator(G, batch=B, input_dtype=torch.uint8).capture() if args.graph else None

    def eager(z_, c_, x_):
        with torch.no_grad():
            return G(z_, c_, x_, noise_mode='const')

    def step_resident():
        # inputs already in HBM (the graph runner copies them device-to-device into its static buffers)
        return runner(dz, dc, dx) if runner is not None else eager(dz, dc, dx)

    pipe = inference.PipelinedGenerator(runner) if runner is not None else None

    def step_e2e():
        # pinned host buffers in, pinned host buffer out, all inside the timed region; with the graph runner the upload of
        # the next batch and the download of the previous one overlap the forward (two staging sets, three streams)
        if pipe is not None:
            pipe.submit(hz, hc, hx, hy)
        else:
            y = eager(hz.to(dev, non_blocking=True), hc.to(dev, non_blocking=True), hx.to(dev, non_blocking=True))
            hy.copy_(y, non_blocking=True)

    def timed_region(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if pipe is not None:
            pipe.finish()                      # every download has landed before the closing event
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    parity = golden_check(lambda z_, c_, x_: (runner(z_, c_, x_) if runner is not None else eager(z_, c_, x_)), dz, dc, dx, dev,
                          args.precision)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.launch_count()
    ms_total = timed_region(step_resident, args.steps)
    launches = _lib.launch_count() - n0
    if runner is not None:
        launches = runner.kernels_per_replay * args.steps      # kernels of this library inside the replayed graphs
    clocks = sampler.finish()

    for _ in range(2):
        step_e2e()
    ms_e2e = timed_region(step_e2e, args.steps)

    # the exact-fp32 parity path on the same workload (one B200, eager launches, a bounded number of steps): reported next
    # to the headline so the reduced-precision number is never read as the fp32 path's
    fp32_path = None
    if world == 1 and args.precision == 'fast' and not args.no_fp32_leg:
        fp32_path = {}
        xf = torch.from_numpy(G.synthesis.u8_lut()[dx.cpu().numpy()]).to(dev)
        for mode, note in (('fp32', 'exact fp32 SIMT kernels everywhere (<= 1e-4 of the reference), eager launches'),
                           ('tc', 'tcgen05 convolutions (fp16 operands, fp32 accumulation), exact fp32 filtered_lrelu, fp32 storage')):
            inference.set_precision(mode)
            try:
                eager(dz, dc, xf)
                ms32 = timed_region(lambda: eager(dz, dc, xf), 2)
                fp32_path[mode] = dict(value=float(B * 2) / (ms32 * 1e-3), unit=UNIT, ms_per_step=ms32 / 2, steps=2, note=note)
            finally:
                inference.set_precision(args.precision)
        del xf

    # per-kernel roofline leg: the same forward launched eagerly, CUDA events around each heavy launch
    _lib.profile_begin()
    for _ in range(min(args.steps, 3)):
        eager(dz, dc, dx)
    prof = _lib.profile_end()
    peaks = load_peaks()

    slices = float(B * world * args.steps)
    value = slices / (ms_total * 1e-3)
    e2e = slices / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    def ncu_traffic(pattern):
        """DRAM read+write bytes per launch of the kernels whose name contains `pattern`, from the committed ncu pass over
        this same command at batch 64 (profiles/r02_traffic_bench_b64.json, written by tools/summarize_profiles.py from the launch
        list profiles/r02_launches_final.txt)."""
        path = os.path.join(ROOT, 'profiles', 'r02_traffic_bench_b64.json')
        if B != 64 or args.precision != 'fast' or not os.path.exists(path):
            return None
        ent = [v for k, v in json.load(open(path)).items() if pattern in k]
        n = sum(v['launches'] for v in ent)
        return sum(v['launches'] * v['dram_bytes_per_launch'] for v in ent) / n if n else None

    def roof(name, bound):
        d = prof.get(name)
        if not d or d['ms'] <= 0:
            return None
        if bound == 'tensor':
            ach = d['work'] / d['ms'] / 1e9          # TFLOP/s
            peak, unit = peaks['tf_sustained'], 'TFLOP/s'
        else:
            ach = d['work'] / d['ms'] / 1e6          # GB/s
            peak, unit = peaks['hbm'], 'GB/s'
        return dict(kernel=name, bound=bound, achieved=ach, peak=peak, unit=unit, frac=ach / peak,
                    traffic=ncu_traffic({'conv2d_tc': 'conv2d_tc_kernel', 'filtered_lrelu': 'flr_tc_kernel',
                                         'conv_tc_pack': 'tc_pack_pairs_kernel'}.get(name, name)),
                    peak_source=peaks['source'] + (' (sustained bf16 cuBLAS)' if bound == 'tensor' else ' (copy)'),
                    launches_per_step=d['launches'] // max(min(args.steps, 3), 1),
                    ms_per_step=d['ms'] / max(min(args.steps, 3), 1),
                    work_per_launch=d['work'] / max(d['launches'], 1))

    r_conv = roof('conv2d_tc', 'tensor')
    r_flr = roof('filtered_lrelu', 'hbm')
    r_pack = roof('conv_tc_pack', 'hbm')
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms_total / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype={'fast': 'f16 operands and activation storage / f32 accumulate', 'tc': 'f16 conv operands / f32 '
                       'accumulate, f32 storage', 'fp32': 'f32'}[args.precision], data='synthetic',
                config=dict(workload=WORKLOAD, batch_per_gpu=B, global_batch=B * world, resolution=256,
                            precision=args.precision, cuda_graph=bool(args.graph),
                            sharding='slices across ranks, no data-path collective',
                            l2='working set per step (>5 GB of activations) exceeds the 126 MB L2; no explicit flush'),
                clocks=clocks, gpu_launches=int(launches), parity=parity, fp32_path=fp32_path,
                e2e=dict(value=e2e, unit=UNIT, h2d_bytes_per_step=int(hz.nbytes + hc.nbytes + hx.nbytes),
                         d2h_bytes_per_step=int(hy.nbytes), ms_per_step=ms_e2e / args.steps),
                roofline=max([r for r in (r_conv, r_flr, r_pack) if r], key=lambda r: r['ms_per_step'], default=None),
                rooflines=dict(conv2d_tc=r_conv, filtered_lrelu=r_flr, conv_tc_pack=r_pack))
    if world == 1 and not args.no_cpu_baseline:
        sps, ms, cores, sample = cpu_reference_slices_per_sec(1, 1)
        line['cpu_baseline'] = dict(value=sps, unit=UNIT, cores=cores, kind='port', sample=sample)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train(args, rank, world):
    """BASELINE configs[4]: generator training step (forward + backward + gradient all-reduce + fused Adam), synthetic
    256x256 slices, batch 32 per GPU, weak scaling.  `python bench.py --workload train [--batch 32]`."""
    import torch
    import torch.distributed as dist
    from afcm_b200 import _lib
    from afcm_b200.networks_stylegan3 import afcm_generator
    from afcm_b200.torch_utils.ops import conv2d_gradfix
    from afcm_b200.training import GeneratorTrainer

    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    _lib.check(_lib.lib().afcm_device_check())
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch if args.batch != 64 else 32                       # config 5: batch 32 per GPU
    from afcm_b200.torch_utils.ops import filtered_lrelu as flr_op
    if args.precision == 'fp32':
        conv2d_gradfix.set_conv_impl('f32')
    else:
        conv2d_gradfix.set_conv_impl('tc', torch.bfloat16)           # bf16 operands (gradient range), fp32 accumulation/storage
        flr_op.set_train_impl('exact' if args.precision == 'tc' else args.flr_train)   # 'fast': tensor-core filtered_lrelu with sign tensor
    G = afcm_generator(seed=0, device=dev).train()
    tr = GeneratorTrainer(G, lr=0.0025, betas=(0.0, 0.99))
    z, c, x = synthetic_inputs(B, seed=rank, as_uint8=True)
    tgt = torch.rand(B, 1, 256, 256, generator=torch.Generator().manual_seed(100 + rank)) * 2 - 1
    hz, hc, hx, ht = z.pin_memory(), c.pin_memory(), x.pin_memory(), tgt.pin_memory()
    hl = torch.empty([1], dtype=torch.float32).pin_memory()
    dz, dc, dx, dt = hz.to(dev), hc.to(dev), hx.to(dev), ht.to(dev)

    def step_resident():
        return tr.step(dz, dc, dx, dt)

    def step_e2e():
        loss = tr.step(hz.to(dev, non_blocking=True), hc.to(dev, non_blocking=True), hx.to(dev, non_blocking=True),
                       ht.to(dev, non_blocking=True))
        hl.copy_(loss.reshape(1), non_blocking=True)

    def timed_region(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    losses = []
    for _ in range(max(args.warmup, 3)):
        losses.append(float(step_resident()))
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.launch_count()
    ms_total = timed_region(step_resident, args.steps)
    launches = _lib.launch_count() - n0
    clocks = sampler.finish()
    step_e2e()
    ms_e2e = timed_region(step_e2e, args.steps)
    nprof = min(args.steps, 2)
    _lib.profile_begin()
    for _ in range(nprof):
        step_resident()
    prof = _lib.profile_end()
    peak_mem = torch.cuda.max_memory_allocated(dev)
    peaks = load_peaks()
    slices = float(B * world * args.steps)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    def roof(name, bound):
        d = prof.get(name)
        if not d or d['ms'] <= 0:
            return None
        if bound == 'tensor':
            ach, peak, unit = d['work'] / d['ms'] / 1e9, peaks['tf_sustained'], 'TFLOP/s'
        else:
            ach, peak, unit = d['work'] / d['ms'] / 1e6, peaks['hbm'], 'GB/s'
        return dict(kernel=name, bound=bound, achieved=ach, peak=peak, unit=unit, frac=ach / peak, traffic=None,
                    peak_source=peaks['source'] + (' (sustained bf16 cuBLAS)' if bound == 'tensor' else ' (copy)'),
                    launches_per_step=d['launches'] // nprof, ms_per_step=d['ms'] / nprof,
                    work_per_launch=d['work'] / max(d['launches'], 1))

    rf = {k: roof(k, b) for k, b in (('conv2d_tc', 'tensor'), ('conv2d_tc_dgrad', 'tensor'), ('conv2d_wgrad_tc', 'tensor'),
                                     ('filtered_lrelu', 'hbm'), ('conv_tc_pack', 'hbm'))}
    line = dict(metric='generator_train_slices_per_sec_256x256', value=slices / (ms_total * 1e-3), unit=UNIT, n_gpus=world,
                steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_total / args.steps, higher_is_better=True,
                scaling='weak', vs_baseline=None,
                dtype={'fast': 'bf16 conv operands, fp16/bf16 filtered_lrelu operands / f32 accumulate, f32 storage', 'tc': 'bf16 conv operands / f32 accumulate, f32 storage and filtered_lrelu', 'fp32': 'f32'}[args.precision],
                data='synthetic',
                config=dict(workload='AFCM generator training step (forward + backward + gradient all-reduce + Adam), synthetic '
                                     '256x256 slices, batch 32 per GPU', batch_per_gpu=B, global_batch=B * world, resolution=256,
                            loss='L1 against a synthetic target (discriminator outside the path)', parameters=tr.flat.flat.numel(),
                            allreduce_bytes_per_step=(tr.flat.grad.numel() * 4 if world > 1 else 0),
                            allreduce='torch.distributed NCCL sum, %d buckets launched from post-accumulate-grad hooks' % len(tr.flat.buckets),
                            peak_memory_gb=peak_mem / 2 ** 30, first_losses=losses[:3],
                            l2='working set per step (tens of GB of saved activations) exceeds the 126 MB L2; no explicit flush'),
                clocks=clocks, gpu_launches=int(launches),
                e2e=dict(value=slices / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=int(hz.nbytes + hc.nbytes + hx.nbytes + ht.nbytes),
                         d2h_bytes_per_step=4, ms_per_step=ms_e2e / args.steps),
                roofline=max([r for r in rf.values() if r], key=lambda r: r['ms_per_step'], default=None), rooflines=rf)
    if world == 1 and not args.no_cpu_baseline:
        sps, cores, sample = cpu_reference_train_slices_per_sec()
        line['cpu_baseline'] = dict(value=sps, unit=UNIT, cores=cores, kind='port', sample=sample)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='forward', choices=['forward', 'train'],
                    help='forward = BASELINE configs[1] (the headline metric); train = configs[4] (training step)')
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--flr-train', default='tc', choices=['tc', 'tcs'], help="training workload: 'tc' = register-chained filtered_lrelu with sign tensor, 'tcs' = the shared-memory tiled kernel")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-fp32-leg', action='store_true', help='skip the extra fp32-path timing of the forward workload')
    ap.add_argument('--precision', default='fast', choices=['fast', 'tc', 'fp32'])
    ap.add_argument('--graph', type=int, default=1, help='replay the forward as one CUDA graph (0 = eager launches)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank)
    elif args.workload == 'train':
        run_train(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == '__main__':
    main()
