"""CPU oracle for the fused tri-plane volume render  --  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch fp32 restatement of the reference's per-ray render
path with every random draw turned into an explicit input tensor.  It exists
to check the CUDA kernels in ``nerf_from_image_b200/csrc``; nothing in the
product package may import it (only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs do).

Parity status: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against the reference ITSELF:
``tests/test_oracle_vs_reference.py`` runs the reference's own ``render()``
(lifted from /root/reference/run.py by AST, see ``oracle/reference_lift.py``)
in this container and compares, and ``tests/golden/*.npz`` hold input/output
vectors produced by the reference (``tests/golden/make_golden.py``).

Reference functions restated here (all paths relative to /root/reference):
  ray_bundle            lib/nerf_utils.py:28-91   get_ray_bundle
  near_far_planes       lib/nerf_utils.py:225-273 compute_near_far_planes
  coarse_depths         lib/nerf_utils.py:94-120  compute_query_points_from_rays
  triplane_decoder      models/generator.py:288-331 TriplanarDecoder.forward
                        models/stylegan.py:148-180  EqualizedLinear (gains are
                        folded into the "effective" weights by the caller)
  field                 models/generator.py:587-681 sampler closure
  view_mapper_trunk     models/generator.py:223-240 ViewDirectionMapper.forward (per-ray part)
  view_head             models/generator.py:242-251 mapper_closure (per-sample part, --use_viewdir)
  composite_weights     lib/nerf_utils.py:164-180 render_volume_density_weights_only
  smooth_weights        run.py:266-272
  inverse_cdf_samples   lib/nerf_utils.py:183-222 sample_pdf
  composite             lib/nerf_utils.py:123-161 render_volume_density
  render_oracle         run.py:176-350           render
"""

import math
from typing import Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# rays
# --------------------------------------------------------------------------
def ray_bundle(height: int, width: int, focal: Optional[torch.Tensor],
               c2w: torch.Tensor, bbox: Optional[torch.Tensor],
               center: Optional[torch.Tensor] = None, rows=None):
    """Ray origins / (unnormalised) directions, [B,H,W,3] each.

    Pixel (h, w) sits at (w/W, h/H): there is no half-pixel offset
    (lib/nerf_utils.py:36-39).  ``focal is None`` selects the orthographic
    model (lib/nerf_utils.py:66-89), otherwise perspective (:40-65).
    """
    dev = c2w.device
    # (dtype follows the cameras: the fp64 runs of tests/ use the oracle as ground truth)
    # rows = (row_offset, full_height): rows [row_offset, row_offset + height) of a taller image
    # (the row-tile split of the multi-GPU path; not a reference feature)
    r0, hfull = rows if rows is not None else (0, height)
    u = (torch.arange(width, device=dev, dtype=c2w.dtype) / width).view(1, 1, width)
    v = ((torch.arange(height, device=dev, dtype=c2w.dtype) + r0) / hfull).view(1, height, 1)
    u = u.expand(1, height, width)
    v = v.expand(1, height, width)
    rot = c2w[:, None, None, :3, :3]
    trans = c2w[:, None, None, :3, 3]
    if focal is not None:
        if center is not None:
            u = u - 0.5 * (2 * center[:, 0, None, None] - 1) - 0.5
            v = v - 0.5 * (2 * center[:, 1, None, None] - 1) - 0.5
        else:
            u = u - 0.5
            v = v - 0.5
        if bbox is not None:
            u = (bbox[:, 1:2, 0].unsqueeze(-1) * (u + 0.5) +
                 bbox[:, 0:1, 0].unsqueeze(-1)) * 0.5
            v = -(bbox[:, 1:2, 1].unsqueeze(-1) * (-v + 0.5) +
                  bbox[:, 0:1, 1].unsqueeze(-1)) * 0.5
        f = focal.view(-1, 1, 1)
        u = u / f
        v = v / f
        d_cam = torch.stack((u, -v, -torch.ones_like(u)), dim=-1)
        dirs = (d_cam[..., None, :] * rot).sum(dim=-1)
        origins = trans.expand(dirs.shape)
    else:
        u = (u - 0.5) * 2
        v = (v - 0.5) * 2
        if bbox is not None:
            u = (bbox[:, 1:2, 0].unsqueeze(-1) * (u / 2 + 0.5) +
                 bbox[:, 0:1, 0].unsqueeze(-1))
            v = -(bbox[:, 1:2, 1].unsqueeze(-1) * (-v / 2 + 0.5) +
                  bbox[:, 0:1, 1].unsqueeze(-1))
        u = u.expand(c2w.shape[0], height, width)
        v = v.expand(c2w.shape[0], height, width)
        o_cam = torch.stack((u, -v, torch.zeros_like(u)), dim=-1)
        d_cam = torch.stack((torch.zeros_like(u), torch.zeros_like(u),
                             -torch.ones_like(u)), dim=-1)
        origins = (o_cam[..., None, :] * rot).sum(dim=-1) + trans
        dirs = (d_cam[..., None, :] * rot).sum(dim=-1) / \
            c2w[:, None, None, 3, 3].unsqueeze(-1)
    return origins, dirs


def near_far_planes(origins: torch.Tensor, dirs: torch.Tensor,
                    scene_range: float, global_fallback: bool = True):
    """Slab test against the cube [-scene_range, scene_range]^3.

    lib/nerf_utils.py:225-273.  Rays that miss take the global min(near) /
    max(far) of the rays that hit (:258-259) when ``global_fallback`` is on;
    the CUDA path skips that step (such rays never enter the cube, so every
    sample is masked and all outputs are 0 / background either way) and the
    tests check both variants give the same images.
    """
    shape = origins.shape[:-1]
    o = origins.detach().reshape(-1, 3)
    d = dirs.detach().reshape(-1, 3)
    inv = 1 / d
    neg = inv < 0
    r = torch.full_like(o, scene_range)
    lo = (torch.where(neg, r, -r) - o) * inv
    hi = (torch.where(neg, -r, r) - o) * inv
    hit = ~((lo[:, 0] > hi[:, 1]) | (lo[:, 1] > hi[:, 0]))
    near = torch.max(lo[:, 0], lo[:, 1])
    far = torch.min(hi[:, 0], hi[:, 1])
    hit = hit & ~((near > hi[:, 2]) | (lo[:, 2] > far))
    near = torch.max(near, lo[:, 2])
    far = torch.min(far, hi[:, 2])
    if global_fallback:
        near = torch.where(hit, near, near[hit].min())
        far = torch.where(hit, far, far[hit].max())
    near = near.clamp(min=0.1)
    far = far.clamp(min=0.1)
    far = torch.where((far - near) < 1e-3, near + 1e-3, far)
    return near.reshape(shape), far.reshape(shape), hit.reshape(shape)


def coarse_depths(near: torch.Tensor, far: torch.Tensor, num_samples: int,
                  noise_t: Optional[torch.Tensor]):
    """t_i = lerp(near, far, i/S) (+ noise_t * (far-near)/S), [B,H,W,S].

    lib/nerf_utils.py:101-114.  ``noise_t`` replaces ``torch.rand_like``.
    """
    frac = torch.arange(num_samples, device=near.device, dtype=near.dtype) / num_samples
    t = torch.lerp(near.unsqueeze(-1), far.unsqueeze(-1), frac)
    if noise_t is not None:
        t = t + noise_t * ((far - near).unsqueeze(-1) / num_samples)
    return t


# --------------------------------------------------------------------------
# radiance field
# --------------------------------------------------------------------------
def triplane_decoder(planes: torch.Tensor, coords: torch.Tensor,
                     w1, b1, w2, b2):
    """planes [B,3,C,R,R], coords [B,N,3] in [-1,1] -> [B,N,1+A].

    models/generator.py:312-331: bilinear / border / align_corners fetch of
    plane 0 at (x,y), plane 1 at (x,z), plane 2 at (y,z) (first coordinate of
    each pair indexes the plane's width), mean of the three, then
    Linear(C->64) - Softplus - Linear(64->1+A) with EFFECTIVE weights.
    """
    g = coords.unsqueeze(2)  # [B,N,1,3]
    feats = 0
    for p, (a, b) in enumerate(((0, 1), (0, 2), (1, 2))):
        feats = feats + F.grid_sample(planes[:, p], g[..., [a, b]],
                                      mode='bilinear', padding_mode='border',
                                      align_corners=True)
    feats = feats / 3
    x = feats.view(feats.shape[0], feats.shape[1], -1).transpose(-2, -1)
    h = F.softplus(F.linear(x, w1, b1))
    return F.linear(h, w2, b2)


def view_mapper_trunk(viewdir, m):
    """ViewDirectionMapper.forward up to the closure (models/generator.py:223-240): the per-RAY
    features x [..., 32] from unit view directions [..., 3].  ``m`` holds EFFECTIVE weights:
    fc0_w/fc0_b, fc1_w .. fc4_w (no bias), norm1_w/norm1_b .. norm4_*, fc5_w/fc5_b, fc6_w/fc6_b."""
    lrelu = lambda t: F.leaky_relu(t, 0.2)
    ln = lambda t, i: F.layer_norm(t, (t.shape[-1],), m['norm%d_w' % i], m['norm%d_b' % i])
    scale = math.sqrt(2) / 2
    x = lrelu(F.linear(viewdir, m['fc0_w'], m['fc0_b']))
    for a, b in ((1, 2), (3, 4)):
        shortcut = x
        x = lrelu(ln(F.linear(x, m['fc%d_w' % a]), a))
        x = lrelu(ln(F.linear(x, m['fc%d_w' % b]), b))
        x = (x + shortcut) * scale
    x = lrelu(F.linear(x, m['fc5_w'], m['fc5_b']))
    return F.linear(x, m['fc6_w'], m['fc6_b'])


def view_head(feat, view_features, w3, b3):
    """mapper_closure (models/generator.py:242-251): ``feat`` [B, rays*S, 32] decoder features,
    ``view_features`` [B, rays, 32] -> colour logits [B, rays*S, A]."""
    B, n, c = feat.shape
    rays = view_features.shape[1]
    y = F.leaky_relu(view_features.unsqueeze(2) + feat.view(B, rays, n // rays, c), 0.2)
    return F.linear(y.view(B, n, c), w3, b3)


def field(points, planes, w1, b1, w2, b2, palette, beta, alpha, scene_range,
          use_sdf=True, want_normals=False, view=None):
    """sigma [B,N], rgb [B,N,3], probs [B,N,A]|None, normals [B,N,3]|None.

    models/generator.py:587-681.  ``points`` is [B,N,3] in world units.  ``view`` =
    (view_features [B,rays,32], w3, b3) switches on the view-direction conditioning
    (:662-663): N = rays * samples, ray-major.
    """
    if want_normals:
        points = points.detach().requires_grad_()
    x = points / scene_range
    with torch.no_grad():
        outside = (x.abs() > 1).any(dim=-1).to(x.dtype)
    out = triplane_decoder(planes, x, w1, b1, w2, b2)
    dist = out[..., 0]
    feat = out[..., 1:]
    normals = None
    if want_normals:
        grad, = torch.autograd.grad(dist.sum(), points, create_graph=False)
        normals = F.normalize(grad, dim=-1)
        dist = dist.detach()
        feat = feat.detach()
    if view is not None:
        feat = view_head(feat, *view)
    if use_sdf:
        nd = -dist
        cdf = 0.5 + 0.5 * torch.sign(nd) * (1 - torch.exp(-nd.abs() / beta))
        sigma = (1 / alpha) * (cdf * (1 - outside))
    else:
        sigma = F.softplus(dist - 1) * (1 - outside)
    probs = None
    if palette is not None:
        probs = F.softmax(feat, dim=-1)
        rgb = torch.matmul(probs, palette)
    else:
        rgb = torch.sigmoid(feat) * 2.004 - 1.002
    return sigma, rgb, probs, normals


def sampler(x_in, planes, w1, b1, w2, b2, palette, beta, alpha, scene_range,
            request=('sigma', 'rgb'), use_sdf=True, bbox_debug=False):
    """The generator's ``sampler`` closure (models/generator.py:587-681) as a dict
    of flat outputs: 'sdf_distance' [B,N,1], 'sigma' [B,N], 'rgb' [B,N,3],
    'semantics' [B,N,A], 'normals' (x_in's shape), 'coords' (= x_in).
    ``x_in`` is [B, ..., 3] in world units."""
    out = {}
    bs = x_in.shape[0]
    pts = x_in.reshape(bs, -1, 3)
    if 'normals' in request:
        pts = pts.detach().requires_grad_()
    x = pts / scene_range
    with torch.no_grad():
        outside = (x.abs() > 1).any(dim=-1).to(x.dtype)
    dec = triplane_decoder(planes, x, w1, b1, w2, b2)
    dist, feat = dec[..., :1], dec[..., 1:]
    if 'normals' in request:
        grad, = torch.autograd.grad(dist[..., -1].sum(), pts, create_graph=False)
        out['normals'] = F.normalize(grad, dim=-1).reshape(x_in.shape)
        dist, feat = dist.detach(), feat.detach()
    if 'sdf_distance' in request:
        out['sdf_distance'] = dist
    if 'sigma' in request:
        if use_sdf:
            nd = -dist[..., -1]
            cdf = 0.5 + 0.5 * torch.sign(nd) * (1 - torch.exp(-nd.abs() / beta))
            sigma = (1 / alpha) * (cdf * (1 - outside))
        else:
            sigma = F.softplus(dist[..., -1] - 1) * (1 - outside)
        if bbox_debug and 'coords' in request:  # generator.py:640-657
            a = pts.detach().abs()
            inside = a < scene_range - 5e-2
            m = torch.ones_like(sigma)
            for i, j in ((0, 1), (0, 2), (1, 2), (1, 2)):
                m = m * (1 - (inside[..., i] & inside[..., j]).to(m.dtype))
            sigma = sigma + 100 * (m * (1 - outside))
        out['sigma'] = sigma
    if 'coords' in request:
        out['coords'] = x_in
    if 'rgb' in request or 'semantics' in request:
        if palette is not None:
            probs = F.softmax(feat, dim=-1)
            if 'semantics' in request:
                out['semantics'] = probs
            if 'rgb' in request:
                out['rgb'] = torch.matmul(probs, palette)
        elif 'rgb' in request:
            out['rgb'] = torch.sigmoid(feat) * 2.004 - 1.002
    return out


# --------------------------------------------------------------------------
# quadrature
# --------------------------------------------------------------------------
def composite_weights(sigma, dirs_unit, depths):
    """w_i = alpha_i * prod_{j<i}(1 - alpha_j + 1e-10); last delta is 0.

    lib/nerf_utils.py:164-180 (and :135-141).
    """
    delta = torch.cat((depths[..., 1:] - depths[..., :-1],
                       torch.zeros_like(depths[..., :1])), dim=-1)
    delta = delta * dirs_unit.norm(p=2, dim=-1, keepdim=True)
    a = 1. - torch.exp(-sigma * delta)
    trans = torch.cumprod(1. - a[..., :-1] + 1e-10, dim=-1)
    trans = torch.cat((torch.ones_like(trans[..., :1]), trans), dim=-1)
    return a * trans


def smooth_weights(w):
    """run.py:266-272: max over (i-1, i), mean over (i, i+1), + 0.01."""
    w = F.max_pool1d(w.unsqueeze(1), 2, 1, padding=1)
    w = F.avg_pool1d(w, 2, 1).squeeze(1)
    return w + 0.01


def inverse_cdf_samples(bins, weights, u):
    """lib/nerf_utils.py:183-222 with ``u`` given ([N,S] in [0,1])."""
    weights = weights + 1e-5
    pdf = weights / weights.sum(dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    cdf = torch.cat((torch.zeros_like(cdf[..., :1]), cdf), dim=-1).contiguous()
    u = u.contiguous()
    idx = torch.searchsorted(cdf, u, right=True)
    below = (idx - 1).clamp(min=0)
    above = idx.clamp(max=cdf.shape[-1] - 1)
    c0 = torch.gather(cdf, 1, below)
    c1 = torch.gather(cdf, 1, above)
    z0 = torch.gather(bins, 1, below)
    z1 = torch.gather(bins, 1, above)
    denom = c1 - c0
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return z0 + (u - c0) / denom * (z1 - z0)


def deterministic_u(num_samples, n_rays, device=None):
    return torch.linspace(0.0, 1.0, steps=num_samples,
                          device=device).expand(n_rays, num_samples)


def composite(sigma, rgb, dirs_unit, depths, white_background,
              normals=None, extra=None):
    """lib/nerf_utils.py:123-161."""
    w = composite_weights(sigma, dirs_unit, depths)
    rgb_map = (w[..., None] * rgb).sum(dim=-2)
    depth_map = (w.detach() * depths.detach()).sum(dim=-1)
    mask = w.sum(-1)
    normal_map = None
    if normals is not None:
        normal_map = (w[..., None].detach() * normals).sum(dim=-2)
    extra_map = None
    if extra is not None:
        extra_map = (w[..., None] * extra).sum(dim=-2)
    if white_background:
        rgb_map = rgb_map + (1. - mask[..., None])
        if normal_map is not None:
            normal_map = normal_map + (1. - mask[..., None])
    return rgb_map, depth_map, mask, normal_map, extra_map


# --------------------------------------------------------------------------
# the whole path
# --------------------------------------------------------------------------
def render_oracle(planes, w1, b1, w2, b2, palette, beta, alpha,
                  c2w, focal, center, bbox, height, width, num_samples,
                  noise_t=None, noise_u=None, *, scene_range,
                  white_background=False, use_sdf=True, fine_sampling=True,
                  compute_normals=False, compute_semantics=False,
                  compute_coords=False, force_no_cam_grad=False,
                  global_near_far_fallback=True, view_features=None, w3=None, b3=None,
                  rows=None):
    """run.py:176-350 with the planes / palette given instead of produced by
    ``target_model`` and with the two random draws passed in:

    noise_t  [B,H,W,S]  stratified jitter (``torch.rand_like`` at
                        lib/nerf_utils.py:112); None -> ``randomize=False``
    noise_u  [B*H*W,S]  inverse-CDF uniforms (``torch.rand`` at :201); None ->
                        the deterministic ``linspace(0,1,S)`` of :194-199

    ``view_features`` [B,H,W,32] (+ ``w3`` [A,32], ``b3`` [A]; then w2 is [33,64]) is the
    ViewDirectionMapper trunk output of the rays (``--use_viewdir``, run.py:216-217: the
    reference feeds the unit directions, detached under ``force_no_cam_grad``, to the model).

    Returns a dict with rgb [B,H,W,3], depth, mask [B,H,W], normals, semantics
    (``coords`` overwrite ``semantics`` as at run.py:337-338), plus z_fine.
    """
    B = planes.shape[0]
    S = num_samples
    origins, dirs = ray_bundle(height, width, focal, c2w, bbox, center, rows)
    dirs = F.normalize(dirs, dim=-1)
    with torch.no_grad():
        near, far, _ = near_far_planes(origins, dirs, scene_range,
                                       global_near_far_fallback)
    depths = coarse_depths(near, far, S, noise_t)
    points = origins[..., None, :] + dirs[..., None, :] * depths[..., :, None]
    if force_no_cam_grad:
        points = points.detach()
        depths = depths.detach()
        dirs = dirs.detach()

    view = None
    if view_features is not None:
        view = (view_features.reshape(B, height * width, -1), w3, b3)

    def run_field(pts):
        shp = pts.shape[:-1]
        s, c, p, n = field(pts.reshape(B, -1, 3), planes, w1, b1, w2, b2,
                           palette, beta, alpha, scene_range, use_sdf,
                           compute_normals, view)
        s = s.view(*shp)
        c = c.view(*shp, 3)
        n = n.view(*shp, 3) if n is not None else None
        e = None
        if compute_coords:
            e = pts if not compute_normals else pts.detach()
        elif compute_semantics:
            e = p.view(*shp, -1)
        return s, c, n, e

    sigma, rgb, normals, extra = run_field(points)
    z_fine = None
    if fine_sampling:
        with torch.no_grad():
            w = composite_weights(sigma, dirs, depths).flatten(0, 2)
            w = smooth_weights(w)
            mid = .5 * (depths[..., 1:] + depths[..., :-1])
            u = noise_u if noise_u is not None else deterministic_u(
                S, w.shape[0], w.device)
            z_fine = inverse_cdf_samples(mid.flatten(0, 2), w[..., 1:-1], u)
            z_fine = z_fine.view(*depths.shape[:3], S)
        z_all, order = torch.sort(torch.cat((depths, z_fine), dim=-1), dim=-1)
        pts_f = origins[..., None, :] + dirs[..., None, :] * z_fine[..., :, None]
        sigma_f, rgb_f, normals_f, extra_f = run_field(pts_f)

        def merge(a, b):
            if a.dim() == order.dim():
                return torch.cat((a, b), dim=-1).gather(-1, order)
            idx = order.unsqueeze(-1).expand(-1, -1, -1, -1, a.shape[-1])
            return torch.cat((a, b), dim=-2).gather(-2, idx)

        sigma = merge(sigma, sigma_f)
        rgb = merge(rgb, rgb_f)
        if normals is not None:
            normals = merge(normals, normals_f)
        if extra is not None:
            extra = merge(extra, extra_f)
        depths = z_all
    rgb_map, depth_map, mask, normal_map, extra_map = composite(
        sigma, rgb, dirs, depths, white_background, normals, extra)
    return {'rgb': rgb_map, 'depth': depth_map, 'mask': mask,
            'normals': normal_map, 'semantics': extra_map, 'z_fine': z_fine,
            'near': near, 'far': far}


def effective_view_mapper_weights(mapper):
    """(trunk dict for ``view_mapper_trunk``, w3, b3) of a reference ViewDirectionMapper with the
    EqualizedLinear gains folded in (models/generator.py:189-221, stylegan.py:175-177)."""
    m = {}
    for i in range(7):
        fc = getattr(mapper, 'fc%d' % i)
        m['fc%d_w' % i] = fc.weight * fc.weight_gain
        if fc.bias is not None:
            m['fc%d_b' % i] = fc.bias * fc.bias_gain
    for i in range(1, 5):
        norm = getattr(mapper, 'norm%d' % i)
        m['norm%d_w' % i], m['norm%d_b' % i] = norm.weight, norm.bias
    out = mapper.output
    return m, out.weight * out.weight_gain, out.bias * out.bias_gain


def effective_decoder_weights(decoder):
    """EqualizedLinear gains folded in (models/stylegan.py:175-177)."""
    l1, l2 = decoder.net[0], decoder.net[2]
    return (l1.weight * l1.weight_gain, l1.bias * l1.bias_gain,
            l2.weight * l2.weight_gain, l2.bias * l2.bias_gain)


def psnr(a, b, data_range=2.0):
    """Unclamped PSNR on the [-1,1] image range (the reference's
    lib/metrics.py:30-45 works on (img/2+0.5) with a 60 dB cap)."""
    mse = ((a - b) ** 2).mean().item()
    return float('inf') if mse == 0 else 10 * math.log10(data_range ** 2 / mse)


def rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
