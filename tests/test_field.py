"""The two neighbours of the render path: the `sampler` seam (nfi_sample_field) and
pose_to_matrix (nfi_pose_to_matrix[_backward]).

CPU: the oracle reproduces the fixtures the reference generated
(tests/golden/make_golden_field.py).  GPU: the CUDA kernels reproduce the same fixtures
through the C ABI, and agree with the oracle on larger seeded inputs.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import pose_oracle as PO
from oracle import render_oracle as O
from tests import helpers as Hh

FIELD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'field')
SAMPLER_FILES = sorted(glob.glob(os.path.join(FIELD_DIR, 'sampler_*.npz')))
POSE_FILES = sorted(glob.glob(os.path.join(FIELD_DIR, 'pose_*.npz')))
ids = lambda files: [os.path.basename(f)[:-4] for f in files]

# tolerances: fp32 SIMT kernels against the fp32 reference
SAMPLER_TOL = 2e-4   # relative L2 per output
POSE_TOL = 1e-5      # absolute (entries are O(1))


def load_sampler(path, device='cpu'):
    z = np.load(path)
    t = lambda k: torch.from_numpy(z[k]).to(device) if k in z.files else None
    scene = {k: t('in_' + k) for k in ('planes', 'w1', 'b1', 'w2', 'b2', 'palette', 'beta', 'alpha')}
    meta = {k[5:]: z[k].item() for k in z.files if k.startswith('meta_')}
    expect = {k[4:]: t(k) for k in z.files if k.startswith('out_')}
    request = list(expect) + (['coords'] if meta['coords'] else [])
    return scene, t('points'), meta, expect, request


def test_fixture_set_is_complete():
    assert len(SAMPLER_FILES) >= 4 and len(POSE_FILES) == 4


@pytest.mark.parametrize('path', SAMPLER_FILES, ids=ids(SAMPLER_FILES))
def test_oracle_sampler_reproduces_reference_golden(path):
    scene, x, m, exp, request = load_sampler(path)
    out = O.sampler(x, scene['planes'], scene['w1'], scene['b1'], scene['w2'], scene['b2'],
                    scene['palette'], scene['beta'], scene['alpha'], m['scene_range'],
                    request=request, use_sdf=bool(m['use_sdf']), bbox_debug=bool(m['bbox']))
    for k, v in exp.items():
        assert out[k].shape == v.shape, k
        assert (out[k].detach() - v).abs().max().item() < (1e-4 if k == 'normals' else 2e-5), k
    if m['coords']:
        assert out['coords'] is x


def load_pose(path, device='cpu'):
    z = np.load(path)
    return {k: torch.from_numpy(z[k]).to(device) for k in z.files if not k.startswith('meta_')}, \
        bool(z['meta_flipped'].item())


def pose_loss(mat, focal, d):
    return (mat * d['wm']).sum() + ((focal * d['wf']).sum() if focal is not None else 0)


def check_pose(fn, d, flipped, tol):
    names = [n for n in ('z0', 't2', 's', 'q') if n in d]
    leaves = {n: d[n].clone().requires_grad_() for n in names}
    mat, focal = fn(leaves.get('z0'), leaves['t2'], leaves['s'], leaves['q'], flipped)
    assert (mat - d['mat']).abs().max().item() < tol
    assert (focal is None) == ('focal' not in d)
    if focal is not None:
        assert (focal - d['focal']).abs().max().item() < tol
    grads = torch.autograd.grad(pose_loss(mat, focal, d), [leaves[n] for n in names])
    for n, g in zip(names, grads):
        assert (g - d['grad_' + n]).abs().max().item() < 20 * tol, n


@pytest.mark.parametrize('path', POSE_FILES, ids=ids(POSE_FILES))
def test_oracle_pose_reproduces_reference_golden(path):
    d, flipped = load_pose(path)
    check_pose(PO.pose_to_matrix, d, flipped, 1e-6)


def test_fused_entry_points_fail_loudly_without_a_gpu():
    from nerf_from_image_b200 import _lib
    from nerf_from_image_b200.pose import pose_to_matrix
    from nerf_from_image_b200.sampler import FusedSampler
    scene, _ = Hh.make_case('p3d_plain', batch=1, plane_res=8)
    with pytest.raises(_lib.NfiError):
        FusedSampler(scene['planes'], scene['w1'], scene['b1'], scene['w2'], scene['b2'],
                     scene['palette'], scene['beta'], scene['alpha'], scene['scene_range'])
    with pytest.raises(_lib.NfiError):
        pose_to_matrix(None, torch.zeros(1, 2), torch.ones(1), torch.tensor([[1., 0, 0, 0]]), False)
    lib = _lib.load()
    assert lib.nfi_sample_field(None, None) != 0 and b'NULL' in lib.nfi_last_error()
    assert lib.nfi_pose_to_matrix(None, None, None, None, 0, 0, None, None, None) != 0


# ------------------------------------------------------------------------------ GPU
def make_sampler(scene, use_sdf=True, bbox_debug=False, device='cuda'):
    from nerf_from_image_b200.sampler import FusedSampler
    sc = Hh.to_device(scene, device)
    return FusedSampler(sc['planes'], sc['w1'], sc['b1'], sc['w2'], sc['b2'], sc['palette'],
                        sc['beta'] if use_sdf else None, sc['alpha'] if use_sdf else None,
                        float(scene['scene_range']), use_sdf=use_sdf, bbox_debug=bbox_debug)


@pytest.mark.gpu
@pytest.mark.parametrize('path', SAMPLER_FILES, ids=ids(SAMPLER_FILES))
def test_cuda_sampler_reproduces_reference_golden(cuda_lib, path):
    scene, x, m, exp, request = load_sampler(path)
    scene['scene_range'] = m['scene_range']
    sampler = make_sampler(scene, bool(m['use_sdf']), bool(m['bbox']))
    with torch.no_grad():
        out = sampler(x.cuda(), request)
    assert sorted(out) == sorted(request)
    for k, v in exp.items():
        assert out[k].shape == v.shape, k
        assert Hh.rel_l2(out[k].cpu(), v) < SAMPLER_TOL, k
    if m['bbox']:
        assert (out['sigma'] > 50).any()


@pytest.mark.gpu
@pytest.mark.parametrize('A,use_sdf,n', [(10, True, 31 ** 3), (0, True, 1), (15, True, 257),
                                         (3, False, 128), (10, False, 4099)])
def test_cuda_sampler_matches_oracle(cuda_lib, A, use_sdf, n):
    """Ragged point counts (1 .. the 31^3 regulariser grid), every decoder width."""
    scene, _ = Hh.make_case('p3d_plain', seed=5, batch=3, plane_res=32, attention_values=A)
    g = torch.Generator().manual_seed(n)
    x = (torch.rand(3, n, 3, generator=g) * 2 - 1) * 1.1 * scene['scene_range']
    request = ['sdf_distance', 'sigma', 'rgb'] + (['semantics'] if A > 0 else []) + \
        (['normals'] if use_sdf else [])
    o = O.sampler(x, scene['planes'], scene['w1'], scene['b1'], scene['w2'], scene['b2'],
                  scene['palette'], scene['beta'], scene['alpha'], scene['scene_range'],
                  request=request, use_sdf=use_sdf)
    with torch.no_grad():
        out = make_sampler(scene, use_sdf)(x.cuda(), request)
    for k in request:
        assert out[k].shape == o[k].shape, k
        assert Hh.rel_l2(out[k].cpu(), o[k].detach()) < SAMPLER_TOL, k


@pytest.mark.gpu
def test_cuda_sampler_agrees_with_the_render_kernel(cuda_lib):
    """A ray composited from sampler outputs equals the fused render (no fine pass)."""
    scene, cams = Hh.make_case('p3d_plain', seed=7, batch=2, plane_res=32)
    H = W = 16
    S = 16
    rgb, depth, mask, _ = Hh.run_cuda(scene, cams, H, W, S, None, None, fine_sampling=False,
                                      mlp_mode=1)[:4]
    o, d = O.ray_bundle(H, W, cams['focal'], cams['c2w'], cams['bbox'], cams['center'])
    dn = torch.nn.functional.normalize(d, dim=-1)
    near, far, _ = O.near_far_planes(o, dn, scene['scene_range'])
    t = O.coarse_depths(near, far, S, None)
    pts = o[..., None, :] + dn[..., None, :] * t[..., None]
    with torch.no_grad():
        out = make_sampler(scene)(pts.cuda(), ['sigma', 'rgb'])
    comp = O.composite(out['sigma'].cpu().view(2, H, W, S), out['rgb'].cpu().view(2, H, W, S, 3),
                       dn, t, scene['white_background'])
    assert Hh.rel_l2(rgb.cpu(), comp[0]) < SAMPLER_TOL
    assert Hh.rel_l2(mask.cpu(), comp[2]) < SAMPLER_TOL


@pytest.mark.gpu
def test_cuda_sampler_errors(cuda_lib):
    from nerf_from_image_b200 import _lib
    scene, _ = Hh.make_case('p3d_plain', seed=5, batch=2, plane_res=16)
    sampler = make_sampler(scene)
    x = torch.zeros(2, 8, 3, device='cuda')
    with pytest.raises(AssertionError):
        sampler(x, ['density'])                      # unknown output name (generator.py:588)
    with pytest.raises(_lib.NfiError):
        sampler(x.requires_grad_(), ['sigma'])       # forward-only: no silent constant
    with pytest.raises(_lib.NfiError):
        with torch.no_grad():
            sampler(torch.zeros(3, 8, 3, device='cuda'), ['sigma'])  # batch mismatch
    with torch.no_grad():
        assert sampler(x, ['coords'])['coords'] is x  # nothing to launch


@pytest.mark.gpu
@pytest.mark.parametrize('path', POSE_FILES, ids=ids(POSE_FILES))
def test_cuda_pose_reproduces_reference_golden(cuda_lib, path):
    from nerf_from_image_b200.pose import pose_to_matrix
    d, flipped = load_pose(path, 'cuda')
    check_pose(pose_to_matrix, d, flipped, POSE_TOL)


@pytest.mark.gpu
def test_cuda_pose_feeds_render_and_backpropagates(cuda_lib):
    """pose -> (c2w, focal) -> fused render -> loss: gradients reach the pose parameters and
    match the oracle chain (oracle pose -> oracle render)."""
    from nerf_from_image_b200.pose import pose_to_matrix
    scene, cams = Hh.make_case('p3d_plain', seed=9, batch=2, plane_res=32)
    B, H, W, S = 2, 12, 12, 8
    g = torch.Generator().manual_seed(3)
    q0 = torch.nn.functional.normalize(torch.tensor([[0.9, 0.1, 0.3, -0.2], [0.7, -0.4, 0.2, 0.5]]), dim=-1)
    z0 = torch.tensor([0.4, 0.7])
    s = torch.tensor([0.85, 0.95])
    t2 = torch.tensor([[0.02, -0.03], [-0.05, 0.04]])
    wr = torch.randn(B, H, W, 3, generator=g)

    def run(pose_fn, render_fn, dev):
        leaves = [t.clone().to(dev).requires_grad_() for t in (z0, t2, s, q0)]
        mat, focal = pose_fn(*leaves, False)
        c = dict(c2w=mat, focal=focal, center=None, bbox=None)
        rgb = render_fn(c)
        loss = (rgb * wr.to(dev)).sum()
        return rgb.detach().cpu(), [x.cpu() for x in torch.autograd.grad(loss, leaves)]

    rgb_o, g_o = run(PO.pose_to_matrix,
                     lambda c: Hh.run_oracle(scene, c, H, W, S, None, None)['rgb'], 'cpu')
    rgb_c, g_c = run(pose_to_matrix,
                     lambda c: Hh.run_cuda(scene, c, H, W, S, None, None)[0], 'cuda')
    assert Hh.rel_l2(rgb_c, rgb_o) < 2e-4
    for a, b, n in zip(g_c, g_o, ('z0', 't2', 's', 'q')):
        assert Hh.rel_l2(a, b) < 2e-3, n
