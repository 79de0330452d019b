"""Seeded synthetic scenes and cameras for tests, smoke and bench.

No datasets or checkpoints exist offline, so the render path is exercised on
an analytic stand-in for a trained generator: tri-planes
whose channel 0 carries a quadratic that the 2-layer decoder turns into a
sphere-like signed distance (so the Laplace-CDF density of
/root/reference/models/generator.py:629-636 has a real surface, a mask
coverage well away from 0 and 1, and an empty margin near the cube faces),
and smooth low-frequency content in the other 31 channels that drives the
softmax-over-palette colour of :668-679.  Cameras follow the reference's
conventions: OpenGL axes (camera looks down -z, lib/nerf_utils.py:60),
look-at-origin poses, focal in [1.0, 1.6] for the perspective sets and
``c2w[3,3] = 1/scale`` for the orthographic CUB model (:66-89).
"""

import math

import torch
import torch.nn.functional as F

HIDDEN = 64
PLANE_CHANNELS = 32

# per-dataset constants of /root/reference/data/loaders.py:23-87
DATASET_CONFIGS = {
    'p3d_car': dict(scene_range=1.4, white_background=False, ortho=False,
                    radius=3.0, object_radius=0.7),
    'cub': dict(scene_range=2.0, white_background=False, ortho=True,
                radius=3.0, object_radius=0.45),
    'shapenet_chairs': dict(scene_range=0.55, white_background=True,
                            ortho=False, radius=1.6, object_radius=0.8),
}


def make_scene(seed, batch, plane_res=256, attention_values=10,
               scene_range=1.4, white_background=False, object_radius=0.7,
               device='cpu', channels=PLANE_CHANNELS):
    """Returns a dict: planes [B,3,C,R,R], w1 [64,C], b1 [64], w2 [1+A,64],
    b2 [1+A] (EFFECTIVE weights, EqualizedLinear gains already applied),
    palette [B,A,3] | None, beta [1], alpha [1], scene_range, white_background.
    """
    gen = torch.Generator().manual_seed(seed)
    B, C, R = batch, channels, plane_res
    A = attention_values
    n_out = 1 + (A if A > 0 else 3)
    low = max(4, R // 8)
    noise = torch.randn(B * 3, C, low, low, generator=gen)
    shape_noise = torch.randn(B * 3, 1, 4, 4, generator=gen)
    w1 = torch.randn(HIDDEN, C, generator=gen) / math.sqrt(C)
    b1 = 0.1 * torch.randn(HIDDEN, generator=gen)
    w2 = torch.randn(n_out, HIDDEN, generator=gen) / math.sqrt(HIDDEN)
    b2 = 0.1 * torch.randn(n_out, generator=gen)
    palette = None
    if A > 0:
        palette = torch.rand(B, A, 3, generator=gen) * 2 - 1

    a, c = 4.0, 6.0
    rho = object_radius
    g = scene_range / (2 * rho)
    w1[0].zero_()
    w1[0, 0] = a
    b1[0] = c
    w2[0] = 0.01 * torch.randn(HIDDEN, generator=gen)
    w2[0, 0] = 3 * g / a
    b2[0] = -(3 * g / a) * c - g * rho * rho

    noise = noise.to(device)
    shape_noise = shape_noise.to(device)
    planes = 0.5 * F.interpolate(noise, size=(R, R), mode='bilinear',
                                 align_corners=True)
    lin = torch.linspace(-1, 1, R, device=device)
    quad = 0.5 * (lin[None, :] ** 2 + lin[:, None] ** 2)
    planes[:, 0] = quad[None] + 0.03 * F.interpolate(
        shape_noise, size=(R, R), mode='bicubic', align_corners=True)[:, 0]
    planes = planes.view(B, 3, C, R, R).contiguous()

    to = lambda t: t.to(device) if t is not None else None
    return dict(planes=planes, w1=to(w1), b1=to(b1), w2=to(w2), b2=to(b2),
                palette=to(palette), beta=torch.tensor([0.1], device=device),
                alpha=torch.tensor([1.0], device=device),
                scene_range=float(scene_range),
                white_background=bool(white_background))


def add_view_mapper(scene, seed=5):
    """Turns ``scene`` into a view-direction-conditioned one (--use_viewdir, CARLA:
    models/generator.py:189-253,376-377): the decoder's second layer gets 1 + 32 outputs (row 0,
    the distance, is kept) and the scene gains ``view_mapper`` -- EFFECTIVE weights of a
    ViewDirectionMapper trunk (fc0 .. fc6, four LayerNorms; oracle.render_oracle.view_mapper_trunk)
    -- and ``w3`` [A,32] / ``b3`` [A] of its output layer (random instead of the reference's zero
    initialisation, so that the conditioning is visible).  Returns a new dict."""
    gen = torch.Generator().manual_seed(seed)
    dev = scene['w2'].device
    A = scene['palette'].shape[1] if scene['palette'] is not None else 3
    rn = lambda *shape: torch.randn(*shape, generator=gen)
    w2 = torch.cat((scene['w2'][:1].cpu(), rn(32, HIDDEN) / math.sqrt(HIDDEN)), dim=0)
    b2 = torch.cat((scene['b2'][:1].cpu(), 0.1 * rn(32)), dim=0)
    m = {'fc0_w': rn(64, 3) / math.sqrt(3), 'fc0_b': 0.1 * rn(64)}
    for i in range(1, 5):
        m['fc%d_w' % i] = rn(64, 64) / 8
        m['norm%d_w' % i] = 1 + 0.1 * rn(64)
        m['norm%d_b' % i] = 0.1 * rn(64)
    m['fc5_w'], m['fc5_b'] = rn(64, 64) / 8, 0.1 * rn(64)
    m['fc6_w'], m['fc6_b'] = rn(32, 64) / 8, 0.1 * rn(32)
    out = dict(scene)
    out.update(w2=w2.to(dev), b2=b2.to(dev), view_mapper={k: v.to(dev) for k, v in m.items()},
               w3=(rn(A, 32) / math.sqrt(32)).to(dev), b3=(0.1 * rn(A)).to(dev))
    return out


def make_cameras(seed, batch, ortho=False, radius=3.0, with_bbox=False,
                 with_center=False, device='cpu'):
    """Look-at-origin cameras: dict(c2w [B,4,4], focal [B]|None, center, bbox)."""
    gen = torch.Generator().manual_seed(seed + 7919)
    az = torch.rand(batch, generator=gen) * 2 * math.pi
    el = torch.rand(batch, generator=gen) * 0.8 - 0.3
    back = torch.stack((torch.cos(el) * torch.sin(az), torch.sin(el),
                        torch.cos(el) * torch.cos(az)), dim=-1)
    up = torch.tensor([0., 1., 0.]).expand(batch, 3)
    right = F.normalize(torch.linalg.cross(up, back), dim=-1)
    true_up = torch.linalg.cross(back, right)
    c2w = torch.zeros(batch, 4, 4)
    c2w[:, :3, 0] = right
    c2w[:, :3, 1] = true_up
    c2w[:, :3, 2] = back
    c2w[:, :3, 3] = back * radius
    c2w[:, 3, 3] = 1.0
    focal = None
    if ortho:
        scale = 0.9 + 0.2 * torch.rand(batch, generator=gen)
        c2w[:, 3, 3] = 1.0 / scale
    else:
        focal = 1.0 + 0.6 * torch.rand(batch, generator=gen)
    bbox = None
    if with_bbox:
        start = -1.0 + 0.3 * torch.rand(batch, 1, 2, generator=gen)
        rng = 1.6 + 0.4 * torch.rand(batch, 1, 2, generator=gen)
        bbox = torch.cat((start, rng), dim=1)
    center = None
    if with_center and not ortho:
        center = 0.45 + 0.1 * torch.rand(batch, 2, generator=gen)
    to = lambda t: t.to(device) if t is not None else None
    return dict(c2w=to(c2w), focal=to(focal), center=to(center), bbox=to(bbox))


def make_noise(seed, batch, height, width, num_samples, fine=True,
               device='cpu'):
    """Explicit stratified jitter [B,H,W,S] and inverse-CDF uniforms [B*H*W,S]."""
    gen = torch.Generator().manual_seed(seed + 104729)
    noise_t = torch.rand(batch, height, width, num_samples, generator=gen)
    noise_u = torch.rand(batch * height * width, num_samples,
                         generator=gen) if fine else None
    return noise_t.to(device), (noise_u.to(device) if fine else None)


def make_synthesis_params(seed, res, channels, w_dim, device='cpu'):
    """Parameters of a synthesis network (models/stylegan.py:438-490) under the reference's
    state_dict names, drawn directly from a seed (no module needed): the flat dict
    ``oracle.synthesis_oracle`` and ``synthesis.FusedSynthesis.from_params`` take."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g).to(device)
    p, layers = {}, {}
    resolutions = [4 << i for i in range(len(channels))]
    assert resolutions[-1] == res
    for i, (r, c) in enumerate(zip(resolutions, channels)):
        pre = 'b%d' % r
        if i == 0:
            p[pre + '.const'] = rn(c, 4, 4)
        for name, cin in (('conv0', channels[i - 1] if i else None), ('conv1', c)):
            if cin is None:
                continue
            key = pre + '.' + name
            p[key + '.weight'] = rn(c, cin, 3, 3)
            p[key + '.affine.weight'] = rn(cin, w_dim)
            p[key + '.affine.bias'] = 1 + 0.1 * rn(cin)
            p[key + '.bias'] = 0.2 * rn(c)
            p[key + '.noise_strength'] = torch.tensor(0.07, device=device)
            p[key + '.noise_const'] = rn(r, r)
            layers[key] = dict(use_noise=True, up=(name == 'conv0'))
        key = pre + '.torgb'
        p[key + '.weight'] = rn(96, c, 1, 1)
        p[key + '.affine.weight'] = rn(c, w_dim)
        p[key + '.affine.bias'] = 1 + 0.1 * rn(c)
        p[key + '.bias'] = 0.2 * rn(96)
    p['meta'] = dict(img_resolution=res, img_channels=96, w_dim=w_dim, resolutions=resolutions,
                     layers=layers)
    return p
