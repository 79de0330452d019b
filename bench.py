#!/usr/bin/env python
"""Benchmark of the fused tri-plane render (BASELINE.json metric: rays/s at
128x128, 64 coarse + 64 fine samples per ray).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of config 2
(p3d_car geometry, batch 32 per GPU): planes [32,3,32,256,256] fp32 as the
synthesis network leaves them -> channel-last re-layout -> fused forward render
-> rgb/depth/mask.  Inputs (805 MB of planes + 268 MB of noise per GPU) are far
larger than the 126 MB L2, so consecutive steps cannot be served from cache.

  value      rays/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the C-ABI host entry point
             nfi_render_forward_host with PINNED HOST buffers: H2D of every
             input and D2H of rgb/depth/mask inside the timed region
  roofline   dominant kernel (render_forward_*) against the algorithmic
             tensor-FLOP bound of SURVEY.md section 8d
  cpu_baseline   the oracle port (oracle/render_oracle.py, same torch ops as the
             reference) on this box's host cores, bounded sample, rank 0, N=1

--impl reference times only the CPU oracle port (the reference's own CPU
PyTorch path cannot travel to the GPU box) on the same config.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'rays_per_sec_128x128_64+64spp'
UNIT = 'rays/s'
CFG = dict(batch=32, height=128, width=128, samples=64, plane_res=256, attention_values=10,
           dataset='p3d_car')
FLOPS_PER_POINT = 2 * 32 * 64 + 2 * 64 * 11 + 2 * 10 * 3  # 5,564


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p['hbm_gbs'], tflops=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    source='measured (MEASURED_PEAKS.json, bf16 sustained)')
    return dict(hbm_gbs=6650.0, tflops=1590.0, source='fallback (B200_PROFILING.md)')


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every
    ~10 ms; falls back to nvidia-smi if pynvml is missing)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    n = self.nvml
                    mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                    bits = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    self.rows.append((mhz, bits))
                    time.sleep(0.01)
                else:
                    q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.active'
                    out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits'], capture_output=True,
                                         text=True, timeout=5).stdout.strip().split(',')
                    self.max_mhz = int(out[1])
                    self.rows.append((int(out[0]), int(out[2], 16)))
            except Exception:
                time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=['unsampled'])
        sm = sorted(r[0] for r in self.rows)
        # NVML clocks-event-reason bits
        names = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown',
                 0x4: 'sw_power_cap'}
        seen = 0
        for _, bits in self.rows:
            seen |= bits
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.max_mhz,
                    reasons=[n for b, n in names.items() if seen & b], samples=len(self.rows))


def best_cpu_threads():
    """The reference's CPU path is plain PyTorch; its intra-op scaling saturates
    well before all cores of a big host, so probe a few thread counts on a small
    sample of the same workload and keep the fastest (reported as `cores`)."""
    from fixtures import synthetic
    from oracle import render_oracle as O
    ds = synthetic.DATASET_CONFIGS[CFG['dataset']]
    scene = synthetic.make_scene(7, 1, plane_res=64, scene_range=ds['scene_range'])
    cams = synthetic.make_cameras(7, 1, radius=ds['radius'])
    nt, nu = synthetic.make_noise(7, 1, 32, 32, 32)
    total = os.cpu_count() or 1
    best, best_t = total, None
    for th in sorted({t for t in (8, 16, 32, 64, total) if t <= total}):
        torch.set_num_threads(th)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            with torch.no_grad():
                O.render_oracle(scene['planes'], scene['w1'], scene['b1'], scene['w2'],
                                scene['b2'], scene['palette'], scene['beta'], scene['alpha'],
                                cams['c2w'], cams['focal'], None, None, 32, 32, 32, nt, nu,
                                scene_range=scene['scene_range'])
            ts.append(time.perf_counter() - t0)
        if best_t is None or min(ts) < best_t:
            best, best_t = th, min(ts)
    return best


def cpu_oracle_rate(n_images, threads=None, grad=False):
    """Times the oracle port on the host cores; returns (rays/s, seconds, outputs, inputs)."""
    from fixtures import synthetic
    from oracle import render_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    ds = synthetic.DATASET_CONFIGS[CFG['dataset']]
    scene = synthetic.make_scene(1234, n_images, plane_res=CFG['plane_res'],
                                 attention_values=CFG['attention_values'],
                                 scene_range=ds['scene_range'],
                                 white_background=ds['white_background'])
    cams = synthetic.make_cameras(1234, n_images, radius=ds['radius'])
    nt, nu = synthetic.make_noise(1234, n_images, CFG['height'], CFG['width'], CFG['samples'])
    t0 = time.perf_counter()
    with torch.no_grad():
        out = O.render_oracle(scene['planes'], scene['w1'], scene['b1'], scene['w2'],
                              scene['b2'], scene['palette'], scene['beta'], scene['alpha'],
                              cams['c2w'], cams['focal'], None, None, CFG['height'],
                              CFG['width'], CFG['samples'], nt, nu,
                              scene_range=scene['scene_range'],
                              white_background=scene['white_background'])
    dt = time.perf_counter() - t0
    rays = n_images * CFG['height'] * CFG['width']
    return rays / dt, dt, out, (scene, cams, nt, nu)


def run_reference(args):
    """--impl reference: the CPU path of the reference (oracle port), bounded sample."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_img = 1
    threads = best_cpu_threads()
    for _ in range(args.warmup):
        cpu_oracle_rate(n_img, threads)
    times = []
    for _ in range(args.steps):
        _, dt, _, _ = cpu_oracle_rate(n_img, threads)
        times.append(dt)
    rays = n_img * CFG['height'] * CFG['width']
    total = sum(times)
    value = rays * len(times) / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        # same workload as the b200 arm; each step times a bounded sample of it (see cpu_baseline)
        'config': {'workload': 'config 2: p3d_car render 128x128, 64 coarse + 64 fine '
                               'samples/ray, batch %d per GPU, 256^2x32ch fp32 tri-planes given '
                               '(channel-first) -> rgb/depth/mask' % CFG['batch'],
                   'global_batch': args.gpus * CFG['batch'],
                   'parallelism': 'host CPU cores (reference CPU path), rank 0 only',
                   'randomize': True},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'host_cores': os.cpu_count(), 'kind': 'port',
                         'sample': '%d image(s) of the 32-image batch per step, torch CPU fp32, '
                                   'no_grad' % n_img},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=CFG['batch'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-backward', action='store_true')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--mlp-mode', type=lambda x: int(x, 0), default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)

    import torch.distributed as dist
    from nerf_from_image_b200 import _lib, fused, parallel
    from fixtures import synthetic

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the fused renderer has no CPU path')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    B, H, W, S = args.batch, CFG['height'], CFG['width'], CFG['samples']
    ds = synthetic.DATASET_CONFIGS[CFG['dataset']]
    # every rank renders its own 32 images (weak scaling; images are independent)
    scene = synthetic.make_scene(1234 + rank, B, plane_res=CFG['plane_res'],
                                 attention_values=CFG['attention_values'],
                                 scene_range=ds['scene_range'],
                                 white_background=ds['white_background'], device=dev)
    cams = synthetic.make_cameras(1234 + rank, B, radius=ds['radius'], device=dev)
    nt, nu = synthetic.make_noise(1234 + rank, B, H, W, S, device=dev)
    cfg = fused.RenderConfig(scene_range=scene['scene_range'],
                             white_background=scene['white_background'],
                             attention_values=CFG['attention_values'], mlp_mode=args.mlp_mode)

    kernel_events = []

    def step(time_kernel=False):
        with torch.no_grad():
            if time_kernel:
                fused.KERNEL_EVENTS = kernel_events
            out = full = None
            if world > 1:
                # the path's one exchange step, in place: the kernel writes this rank's tiles
                # into its slice of the full-batch buffers, NCCL all-gathers them where they lie
                full = parallel.gathered_buffers(world * B, H, W, dev)
                out = parallel.shard_views(full, world * B, world, rank)
            rgb, depth, mask, _ = fused.fused_render(
                scene['planes'], scene['w1'], scene['b1'], scene['w2'], scene['b2'],
                scene['palette'], scene['beta'], scene['alpha'], cams['c2w'], cams['focal'],
                None, None, cfg, H, W, S, nt, nu, out=out)
            fused.KERNEL_EVENTS = None
            if world > 1:
                rgb, depth, mask = parallel.all_gather_inplace(full, world * B)
        return rgb, depth, mask

    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = step(time_kernel=True)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.summary() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = t.item()
    ms_step = ms_total / args.steps
    rays_step = world * B * H * W
    value = rays_step / (ms_step * 1e-3)
    k_ms = sum(a.elapsed_time(b) for a, b in kernel_events) / max(1, len(kernel_events))

    # ------------------------------------------------------------ e2e (host buffers)
    e2e = None
    if not args.no_e2e:
        host = {}
        pin = lambda t: t.detach().cpu().contiguous().pin_memory()
        for k in ('planes', 'w1', 'b1', 'w2', 'b2', 'palette', 'beta', 'alpha'):
            host[k] = pin(scene[k])
        host['c2w'], host['focal'] = pin(cams['c2w']), pin(cams['focal'])
        host['rgb'] = torch.empty(B, H, W, 3).pin_memory()
        host['depth'] = torch.empty(B, H, W).pin_memory()
        host['mask'] = torch.empty(B, H, W).pin_memory()
        p = _lib.RenderParams()
        p.batch, p.height, p.width, p.num_samples = B, H, W, S
        p.plane_res, p.n_attention = CFG['plane_res'], CFG['attention_values']
        p.scene_range = scene['scene_range']
        p.white_background = int(scene['white_background'])
        # The two random draws of the path (torch.rand_like / torch.rand on the device in the
        # reference, lib/nerf_utils.py:112,201) are generated on the device from a seed:
        # they are not inputs that exist on the host in the reference either.
        p.use_sdf, p.fine_sampling, p.noise_mode = 1, 1, _lib.NOISE_PHILOX
        p.noise_seed = 1234 + rank
        p.mlp_mode = args.mlp_mode
        in_keys = ('planes', 'w1', 'b1', 'w2', 'b2', 'palette', 'beta', 'alpha', 'c2w', 'focal')
        for k in in_keys + ('rgb', 'depth', 'mask'):
            setattr(p, k, ctypes.c_void_p(host[k].data_ptr()))
        h2d = sum(host[k].numel() * 4 for k in in_keys)
        d2h = sum(host[k].numel() * 4 for k in ('rgb', 'depth', 'mask'))
        _lib.check(lib.nfi_render_forward_host(ctypes.byref(p), local_rank))  # warm-up
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            _lib.check(lib.nfi_render_forward_host(ctypes.byref(p), local_rank))
        dt = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = t.item()
        # same render through the device path with the same device-generated noise
        nt2, nu2 = torch.empty_like(nt), torch.empty_like(nu)
        st_ = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.nfi_fill_uniform(ctypes.c_void_p(nt2.data_ptr()), nt2.numel(),
                                        p.noise_seed, 0, 0, st_))
        _lib.check(lib.nfi_fill_uniform(ctypes.c_void_p(nu2.data_ptr()), nu2.numel(),
                                        p.noise_seed, 1, 0, st_))
        with torch.no_grad():
            rgb_d = fused.fused_render(
                scene['planes'], scene['w1'], scene['b1'], scene['w2'], scene['b2'],
                scene['palette'], scene['beta'], scene['alpha'], cams['c2w'], cams['focal'],
                None, None, cfg, H, W, S, nt2, nu2)[0]
        e2e_err = (host['rgb'] - rgb_d[:B].cpu()).abs().max().item()
        del nt2, nu2, rgb_d
        e2e = {'value': rays_step / dt, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': d2h, 'ms_per_step': dt * 1e3,
               'max_abs_diff_vs_device_path': e2e_err,
               'api': 'nfi_render_forward_host (C ABI, pinned host buffers; planes, decoder, '
                      'palette, cameras H2D every step, the two uniform draws generated on the '
                      'device from a seed (NFI_NOISE_PHILOX), rgb/depth/mask D2H)'}

    # ------------------------------------------------------------ forward + backward
    # (the inversion loop's use of the path, run.py:2202-2299: grads to planes, palette and
    # cameras with the decoder frozen).  Reported beside the headline, not instead of it.
    fwd_bwd = None
    if not args.no_backward:
        planes_g = scene['planes'].detach().clone().requires_grad_()
        pal_g = scene['palette'].detach().clone().requires_grad_()
        c2w_g = cams['c2w'].detach().clone().requires_grad_()

        def step_bwd():
            rgb, depth, mask, _ = fused.fused_render(
                planes_g, scene['w1'], scene['b1'], scene['w2'], scene['b2'], pal_g,
                scene['beta'], scene['alpha'], c2w_g, cams['focal'], None, None, cfg, H, W, S,
                nt, nu)
            (rgb.square().mean() + mask.mean()).backward()
            planes_g.grad = pal_g.grad = c2w_g.grad = None

        # start from a clean caching-allocator state (the legs above leave blocks of other
        # sizes behind), then warm up: the timed steps must not contain cudaMalloc / cudaFree
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        for _ in range(3):
            step_bwd()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nb = max(2, min(5, args.steps))
        b0.record()
        for _ in range(nb):
            step_bwd()
        b1.record()
        torch.cuda.synchronize()
        ms_b = b0.elapsed_time(b1) / nb
        if world > 1:
            t = torch.tensor([ms_b], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_b = t.item()
        fwd_bwd = {'value': rays_step / (ms_b * 1e-3), 'unit': UNIT, 'ms_per_step': ms_b,
                   'what': 'fused_render forward + backward through the autograd.Function '
                           '(grads to planes, palette, tform_cam2world; decoder frozen), '
                           'incl. both re-layouts'}
        del planes_g, pal_g, c2w_g
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    flops_launch = float(B * H * W) * (2 * S) * FLOPS_PER_POINT
    hbm_bytes_launch = float(B) * 3 * 32 * CFG['plane_res'] ** 2 * 4 + B * H * W * 20.0
    achieved_tf = flops_launch / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    roofline = {
        'bound': 'tensor', 'achieved': achieved_tf, 'peak': pk['tflops'], 'unit': 'TFLOP/s',
        'frac': achieved_tf / pk['tflops'], 'traffic': TRAFFIC_BYTES_PER_IMAGE * B,
        'kernel': 'render_forward (dominant kernel of the step)', 'kernel_ms': k_ms,
        'kernel_share_of_step': k_ms / ms_step if ms_step > 0 else None,
        'algorithmic_flops_per_launch': flops_launch,
        'algorithmic_hbm_bytes_per_launch': hbm_bytes_launch,
        'hbm_achieved_gbs': hbm_bytes_launch / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0,
        'hbm_peak_gbs': pk['hbm_gbs'], 'peak_source': pk['source'],
    }

    cpu_baseline = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        n_img = 2
        cores = best_cpu_threads()
        # bounded sample: the same 2 images rendered repeatedly until >= 10 s of CPU work
        # (first pass untimed: thread pool and allocator warm-up)
        _, _, ref, (sc_c, cm_c, nt_c, nu_c) = cpu_oracle_rate(n_img, cores)
        reps, dt = 0, 0.0
        while dt < 10.0 and reps < 8:
            dt += cpu_oracle_rate(n_img, cores)[1]
            reps += 1
        rate = reps * n_img * H * W / dt
        cpu_baseline = {'value': rate, 'unit': UNIT, 'cores': cores, 'host_cores': os.cpu_count(), 'kind': 'port',
                        'sample': '%d of the 32 images (same geometry) x %d passes, %.1f s, torch '
                                  'CPU fp32 no_grad' % (n_img, reps, dt)}
        # parity of the CUDA path on exactly those images
        sc_g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc_c.items()}
        with torch.no_grad():
            rgb_g, dep_g, msk_g, _ = fused.fused_render(
                sc_g['planes'], sc_g['w1'], sc_g['b1'], sc_g['w2'], sc_g['b2'], sc_g['palette'],
                sc_g['beta'], sc_g['alpha'], cm_c['c2w'].to(dev), cm_c['focal'].to(dev), None,
                None, cfg, H, W, S, nt_c.to(dev), nu_c.to(dev))
        from oracle.render_oracle import psnr, rel_l2
        parity = {'rgb_rel_l2': rel_l2(rgb_g.cpu(), ref['rgb']),
                  'rgb_psnr_db': psnr(rgb_g.cpu(), ref['rgb']),
                  'mask_rel_l2': rel_l2(msk_g.cpu(), ref['mask']),
                  'depth_rel_l2': rel_l2(dep_g.cpu(), ref['depth']),
                  'against': 'CPU oracle, identical injected noise, %d images' % n_img}

    # ------------------------------------------------------------ PyTorch eager on this GPU
    # The same torch ops as the reference's render (the oracle port), fp32 with TF32 off like
    # run.py:59-60, on a sample that fits (the unfused path materialises ~2 GB per image).
    eager = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import render_oracle as O
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            n_img = 2
            sc_e = synthetic.make_scene(77, n_img, plane_res=CFG['plane_res'],
                                        attention_values=CFG['attention_values'],
                                        scene_range=ds['scene_range'],
                                        white_background=ds['white_background'], device=dev)
            cm_e = synthetic.make_cameras(77, n_img, radius=ds['radius'], device=dev)
            nt_e, nu_e = synthetic.make_noise(77, n_img, H, W, S, device=dev)

            def eager_step():
                with torch.no_grad():
                    return O.render_oracle(sc_e['planes'], sc_e['w1'], sc_e['b1'], sc_e['w2'],
                                           sc_e['b2'], sc_e['palette'], sc_e['beta'],
                                           sc_e['alpha'], cm_e['c2w'], cm_e['focal'], None, None,
                                           H, W, S, nt_e, nu_e, scene_range=sc_e['scene_range'],
                                           white_background=sc_e['white_background'])
            for _ in range(2):
                eager_step()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(3):
                eager_step()
            g1.record()
            torch.cuda.synchronize()
            ms_e = g0.elapsed_time(g1) / 3
            eager = {'value': n_img * H * W / (ms_e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e,
                     'sample': '%d images per step (same geometry), torch eager fp32 on this GPU, '
                               'no_grad, oracle port of the reference ops' % n_img}
            del sc_e, cm_e, nt_e, nu_e
            torch.cuda.empty_cache()
        except Exception as exc:  # out of memory on a smaller part: report, do not fail the bench
            eager = {'unavailable': repr(exc)[:200]}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'config 2: p3d_car render 128x128, 64 coarse + 64 fine '
                               'samples/ray, batch %d per GPU, 256^2x32ch fp32 tri-planes given '
                               '(channel-first) -> rgb/depth/mask' % B,
                   'global_batch': world * B, 'parallelism': 'images sharded over %d GPU(s)'
                   % world + (', all_gather of [rgb,depth,mask] tiles' if world > 1 else ''),
                   'cache': 'inputs (1.07 GB per GPU) larger than L2; no flush needed',
                   'randomize': True},
        'clocks': clocks, 'e2e': e2e,
        # per step: planes_to_cl_kernel, prep_weight_image, render_forward_pipe
        'gpu_launches': 3 * args.steps,
        'roofline': roofline, 'fwd_bwd': fwd_bwd, 'cpu_baseline': cpu_baseline,
        'torch_eager_gpu': eager, 'parity': parity,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of the render kernel from the
# committed `ncu --set full` capture (profiles/), per launch; None until captured.
# (269.0 + 118.2) MB for the 8 images of profiles/r1_ncu_v5_pipe_B8.txt; a launch moves this per image
TRAFFIC_BYTES_PER_IMAGE = (269012736 + 118194688) / 8

if __name__ == '__main__':
    main()
