"""Pins the oracle against the reference's own code, run here (CPU).

Skipped where /root/reference is absent (the GPU box); there
tests/test_golden.py carries the same information as committed vectors.
"""
import pytest
import torch

from oracle import reference_lift as RL
from oracle import render_oracle as O
from tests import helpers as Hh

pytestmark = pytest.mark.skipif(not RL.available(), reason='/root/reference not mounted')


@pytest.mark.parametrize('case', list(Hh.CASES))
@pytest.mark.parametrize('randomize', [True, False])
def test_render_matches_reference(case, randomize):
    scene, cams = Hh.make_case(case, seed=2, batch=2, plane_res=24)
    out, nt, nu = RL.reference_render(scene, cams, 12, 20, 12, seed=4, randomize=randomize)
    o = Hh.run_oracle(scene, cams, 12, 20, 12, nt, nu)
    for ref, k in zip(out[:3], ('rgb', 'depth', 'mask')):
        assert (ref - o[k]).abs().max().item() < 2e-5, k


@pytest.mark.parametrize('kw', [dict(fine_sampling=False), dict(use_sdf=False),
                                dict(compute_semantics=True), dict(compute_coords=True),
                                dict(compute_normals=True), dict(force_no_cam_grad=True)])
def test_render_variants_match_reference(kw):
    scene, cams = Hh.make_case('p3d_bbox', seed=3, batch=2, plane_res=24)
    out, nt, nu = RL.reference_render(scene, cams, 12, 12, 10, seed=5, **kw)
    o = Hh.run_oracle(scene, cams, 12, 12, 10, nt, nu, **kw)
    for ref, k in zip(out[:3], ('rgb', 'depth', 'mask')):
        assert (ref - o[k]).abs().max().item() < 2e-5, k
    if out[3] is not None:
        assert (out[3] - o['normals']).abs().max().item() < 1e-4
    if out[4] is not None:
        assert (out[4] - o['semantics']).abs().max().item() < 2e-5


def test_stage_functions_match_reference():
    nerf_utils, _ = RL._import_reference()
    scene, cams = Hh.make_case('p3d_bbox', seed=6, batch=2)
    for cam_case in ('p3d_bbox', 'cub_ortho_bbox', 'chairs_white_center'):
        _, cams = Hh.make_case(cam_case, seed=6, batch=2)
        ro, rd = nerf_utils.get_ray_bundle(9, 13, cams['focal'], cams['c2w'], cams['bbox'],
                                           cams['center'])
        o, d = O.ray_bundle(9, 13, cams['focal'], cams['c2w'], cams['bbox'], cams['center'])
        assert (ro - o).abs().max() < 1e-6 and (rd - d).abs().max() < 1e-6
        dn = torch.nn.functional.normalize(rd, dim=-1)
        rn, rf = nerf_utils.compute_near_far_planes(ro, dn, scene['scene_range'])
        n, f, _ = O.near_far_planes(o, dn, scene['scene_range'])
        assert torch.equal(rn, n) and torch.equal(rf, f)
    torch.manual_seed(0)
    w = torch.rand(50, 16)
    bins = torch.sort(torch.rand(50, 15), dim=-1).values
    u = torch.rand(50, 16)
    torch.manual_seed(1)
    ref = nerf_utils.sample_pdf(bins, w[..., 1:-1], 16, deterministic=False)
    torch.manual_seed(1)
    u = torch.rand(50, 16)
    assert torch.allclose(ref, O.inverse_cdf_samples(bins, w[..., 1:-1], u), atol=1e-6)
    ref = nerf_utils.sample_pdf(bins, w[..., 1:-1], 16, deterministic=True)
    assert torch.allclose(ref, O.inverse_cdf_samples(bins, w[..., 1:-1],
                                                     O.deterministic_u(16, 50)), atol=1e-6)


def test_missed_rays_do_not_need_the_global_fallback():
    """The CUDA path skips lib/nerf_utils.py:258-259; images must not change."""
    scene, cams = Hh.make_case('p3d_plain', seed=8, batch=2)
    cams['focal'] = torch.tensor([0.35, 0.4])  # wide field of view: corner rays miss the cube
    from nerf_from_image_b200 import synthetic
    nt, nu = synthetic.make_noise(1, 2, 16, 16, 8)
    a = Hh.run_oracle(scene, cams, 16, 16, 8, nt, nu, global_near_far_fallback=True)
    b = Hh.run_oracle(scene, cams, 16, 16, 8, nt, nu, global_near_far_fallback=False)
    o, d = O.ray_bundle(16, 16, cams['focal'], cams['c2w'], cams['bbox'], None)
    hit = O.near_far_planes(o, torch.nn.functional.normalize(d, dim=-1), scene['scene_range'])[2]
    assert (~hit).any() and hit.any()
    for k in ('rgb', 'depth', 'mask'):
        assert torch.equal(a[k], b[k]), k
