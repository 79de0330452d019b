"""--use_viewdir on the CUDA path (run.py:216-217,
models/generator.py:189-253,376-377,662-663): the decoder emits 1 + 32 values and every sample's
colour logits are  w3 . leaky_relu(view_features[ray] + features, 0.2) + b3.

The kernels (render_forward_simt / render_backward_simt with VD = true, csrc/nfi_viewdir.cu) are
compared through the C ABI with the oracle restatement -- itself pinned to the reference's
ViewDirectionMapper and render() by tests/test_oracle_vs_reference.py and the two
``*viewdir*.npz`` golden fixtures of tests/test_golden.py -- forward on every camera model and
output mode, backward (float64 oracle autograd as ground truth) for every leaf including the
mapper trunk's parameters and the cameras THROUGH the view directions.
"""
import pytest
import torch

from fixtures import synthetic
from nerf_from_image_b200 import _lib
from tests import helpers as Hh

pytestmark = pytest.mark.gpu

H, W, S = 24, 40, 16


def make(case, A=10, batch=2, res=32, seed=3):
    scene, cams = Hh.make_case(case, seed=seed, batch=batch, plane_res=res, attention_values=A)
    return synthetic.add_view_mapper(scene), cams


@pytest.mark.parametrize('case,A,kw', [
    ('p3d_bbox', 10, {}),
    ('cub_ortho', 10, {}),
    ('chairs_white_center', 0, {}),
    ('p3d_plain', 15, {}),
    ('p3d_plain', 3, dict(compute_semantics=True)),
    ('p3d_plain', 10, dict(compute_semantics=True)),
    ('p3d_plain', 10, dict(compute_coords=True)),
    ('p3d_plain', 10, dict(fine_sampling=False)),
    ('p3d_plain', 10, dict(use_sdf=False)),
    ('p3d_plain', 10, dict(randomize=False)),
])
def test_forward_matches_oracle(cuda_lib, case, A, kw):
    scene, cams = make(case, A)
    kw = dict(kw)
    randomize = kw.pop('randomize', True)
    nt, nu = synthetic.make_noise(7, 2, H, W, S) if randomize else (None, None)
    want = Hh.run_oracle(scene, cams, H, W, S, nt, nu, **kw)
    extra_mode = (_lib.EXTRA_COORDS if kw.get('compute_coords') else
                  _lib.EXTRA_SEMANTICS if kw.get('compute_semantics') else _lib.EXTRA_NONE)
    with torch.no_grad():
        rgb, depth, mask, extra = Hh.run_cuda(scene, cams, H, W, S, nt, nu,
                                              use_sdf=kw.get('use_sdf', True),
                                              fine_sampling=kw.get('fine_sampling', True),
                                              extra_mode=extra_mode)
    tol = 1e-3 if not randomize else 2e-4   # randomize=False: sample 0 sits on the cube face
    assert Hh.rel_l2(rgb.cpu(), want['rgb']) < tol
    assert Hh.rel_l2(mask.cpu(), want['mask']) < tol
    assert Hh.rel_l2(depth.cpu(), want['depth']) < tol
    if extra_mode:
        assert Hh.rel_l2(extra.cpu(), want['semantics']) < tol


def test_normals_with_viewdir(cuda_lib):
    scene, cams = make('p3d_plain')
    nt, nu = synthetic.make_noise(9, 2, H, W, S)
    want = Hh.run_oracle(scene, cams, H, W, S, nt, nu, compute_normals=True)
    with torch.no_grad():
        out = Hh.run_cuda(scene, cams, H, W, S, nt, nu, compute_normals=True)
    assert Hh.rel_l2(out[0].cpu(), want['rgb']) < 2e-4
    assert Hh.rel_l2(out[4].cpu(), want['normals']) < 1e-3


def test_the_view_direction_matters(cuda_lib):
    """Not a vacuous test: the same scene without the ray's mapper features renders differently."""
    scene, cams = make('p3d_plain')
    nt, nu = synthetic.make_noise(7, 2, H, W, S)
    with torch.no_grad():
        a = Hh.run_cuda(scene, cams, H, W, S, nt, nu)[0]
        flat = dict(scene, view_mapper={k: torch.zeros_like(v) for k, v in scene['view_mapper'].items()})
        b = Hh.run_cuda(flat, cams, H, W, S, nt, nu)[0]
    assert Hh.rel_l2(a, b) > 1e-2


@pytest.mark.parametrize('case,A,wgrad', [('p3d_plain', 10, True), ('p3d_plain', 10, False),
                                          ('cub_ortho_bbox', 0, True), ('chairs_white_center', 3, True)])
def test_backward_matches_float64_oracle(cuda_lib, case, A, wgrad):
    scene, cams = make(case, A)
    nt, nu = synthetic.make_noise(13, 2, H, W, S)
    names = ['planes', 'w3', 'b3'] if wgrad else ['planes']
    names += ['w1', 'b1', 'w2', 'b2'] if wgrad else []
    names += ['palette'] if A > 0 else []
    names += ['beta', 'alpha']
    vm_names = ['fc0_w', 'fc2_w', 'norm4_w', 'fc6_b']
    g = torch.Generator().manual_seed(1)
    wr, wm = torch.randn(2, H, W, 3, generator=g), torch.randn(2, H, W, generator=g)

    def leaves(sc, cm):
        sc = {k: (v.clone().requires_grad_() if k in names else v) for k, v in sc.items()}
        sc['view_mapper'] = {k: (v.clone().requires_grad_() if k in vm_names else v)
                             for k, v in sc['view_mapper'].items()}
        cm = dict(cm, c2w=cm['c2w'].clone().requires_grad_())
        return sc, cm, [sc[n] for n in names] + [sc['view_mapper'][n] for n in vm_names] + [cm['c2w']]

    dbl = lambda d: {k: (v.double() if torch.is_tensor(v) else
                         ({a: b.double() for a, b in v.items()} if isinstance(v, dict) else v))
                     for k, v in d.items()}
    sc, cm, lv = leaves(dbl(scene), dbl(cams))
    ref = Hh.run_oracle(sc, cm, H, W, S, nt.double(), nu.double())
    want = torch.autograd.grad((ref['rgb'] * wr.double()).sum() + (ref['mask'] * wm.double()).sum(), lv)

    dev = lambda d: {k: (v.cuda() if torch.is_tensor(v) else
                         ({a: b.cuda() for a, b in v.items()} if isinstance(v, dict) else v))
                     for k, v in d.items()}
    sc, cm, lv = leaves(dev(scene), dev(cams))
    rgb, depth, mask, _ = Hh.run_cuda(sc, cm, H, W, S, nt, nu)
    have = torch.autograd.grad((rgb * wr.cuda()).sum() + (mask * wm.cuda()).sum(), lv)
    assert Hh.rel_l2(rgb.detach().cpu(), ref['rgb'].float()) < 2e-4
    for n, a, b in zip(names + ['vm_' + n for n in vm_names] + ['c2w'], have, want):
        tol = 5e-3 if n in ('beta', 'alpha') else 1e-3
        assert Hh.rel_l2(a.cpu().double(), b) < tol, (n, Hh.rel_l2(a.cpu().double(), b))


def test_force_no_cam_grad_cuts_the_view_directions_too(cuda_lib):
    scene, cams = make('p3d_plain')
    nt, nu = synthetic.make_noise(13, 2, H, W, S)
    sc = Hh.to_device(scene, 'cuda')
    sc['view_mapper'] = {k: v.cuda() for k, v in scene['view_mapper'].items()}
    sc['planes'] = sc['planes'].requires_grad_()
    cm = Hh.to_device(cams, 'cuda')
    cm['c2w'] = cm['c2w'].requires_grad_()
    rgb = Hh.run_cuda(sc, cm, H, W, S, nt, nu, cam_grad=False)[0]
    gp, gc = torch.autograd.grad(rgb.square().sum(), [sc['planes'], cm['c2w']], allow_unused=True)
    assert gc is None and gp.abs().sum() > 0


def test_errors(cuda_lib):
    scene, cams = make('p3d_plain')
    nt, nu = synthetic.make_noise(7, 2, H, W, S)
    bad = dict(scene, w3=scene['w3'][:5])
    with pytest.raises(AssertionError):
        Hh.run_cuda(bad, cams, H, W, S, nt, nu)
    with pytest.raises(AssertionError):          # 33 decoder rows are required with a view
        Hh.run_cuda(dict(scene, w2=scene['w2'][:11], b2=scene['b2'][:11]), cams, H, W, S, nt, nu)


def test_sampler_seam_serves_the_geometry_of_a_view_conditioned_model(cuda_lib):
    """FusedSampler.from_generator on a --use_viewdir model: 'sigma' / 'sdf_distance' / 'normals'
    (row 0 of the 33-row decoder) as the oracle's sampler, colours refused (they need the rays)."""
    import types
    from nerf_from_image_b200.sampler import FusedSampler
    from oracle import render_oracle as O
    scene, cams = make('p3d_plain')
    sc = Hh.to_device(scene, 'cuda')
    g = torch.Generator().manual_seed(5)
    x = ((torch.rand(2, 7, 9, 3, generator=g) * 2 - 1) * scene['scene_range'] * 0.9)
    model = types.SimpleNamespace(use_viewdir=True, use_sdf=True, scene_range=scene['scene_range'])
    field = dict(planes=sc['planes'], palette=sc['palette'], w1=sc['w1'], b1=sc['b1'], w2=sc['w2'],
                 b2=sc['b2'], beta=sc['beta'], alpha=sc['alpha'])
    fs = FusedSampler.from_generator(model, field)
    with torch.no_grad():
        got = fs(x.cuda(), ['sigma', 'sdf_distance', 'normals'])
        with pytest.raises(NotImplementedError):
            fs(x.cuda(), ['sigma', 'rgb'])
    want = O.sampler(x, scene['planes'], scene['w1'], scene['b1'], scene['w2'][:11], scene['b2'][:11],
                     scene['palette'], scene['beta'], scene['alpha'], scene['scene_range'],
                     request=('sigma', 'sdf_distance', 'normals'))
    for k in ('sigma', 'sdf_distance', 'normals'):
        assert Hh.rel_l2(got[k].cpu().reshape(want[k].shape), want[k].detach()) < 1e-4, k
