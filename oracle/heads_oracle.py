"""CPU/torch restatement of the regulariser heads of ``Generator.forward`` -- TEST INFRASTRUCTURE.

/root/reference/models/generator.py:520-585: on (nstrata-1)^3 = 29,791
stratified points per image the decoder's first output (SDF or pre-density) feeds
    sdf_eikonal_loss      mean (|grad_x sdf| - 1)^2           (double backward through the fetch)
    sdf_distance_loss     mse(sdf, |x| - 1)                    (SDF pre-training target: unit sphere)
    total_variation_loss  mean |cdf(-sdf(x)) - cdf(-sdf(x + 0.004 n))|
    entropy_loss          mean laplace_pdf(-sdf)
Restated over explicit tensors (planes, effective decoder weights, the stratified points and the
perturbation noise) so that ``nerf_from_image_b200.heads`` (one fused kernel pair for
``(sdf, grad sdf)`` and their backward) can be compared value by value and gradient by gradient.
Pinned to the reference in tests/test_heads_oracle.py (reference Generator, stubbed synthesis).
Only tests/ import this module.
"""
import torch
import torch.nn.functional as F


def stratified_points(batch, nstrata, scene_range, noise):
    """lib/ops.py:18-25 with the ``torch.rand_like`` draw passed in: noise [B,n,n,n,3] in [0,1),
    n = nstrata - 1.  Returns [B, n^3, 3] world coordinates."""
    n = nstrata - 1
    r = torch.arange(n, device=noise.device)
    bins = torch.stack(torch.meshgrid(r, r, r, indexing='xy'), dim=-1).to(noise.dtype)
    bins = bins.unsqueeze(0).expand(batch, -1, -1, -1, -1)
    bins = (bins + noise) / n * 2 - 1
    return bins.flatten(1, 3) * scene_range


def bilinear_border(image, gx, gy):
    """lib/ops.py:58-120: bilinear fetch, border clamp of the INDICES (not of the weights),
    align_corners -- twice differentiable.  image [B,C,R,R], gx/gy [B,N] in [-1,1]."""
    B, C, ih, iw = image.shape
    ix = ((gx + 1) / 2) * (iw - 1)
    iy = ((gy + 1) / 2) * (ih - 1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    wx1, wy1 = ix - x0, iy - y0
    wx0, wy0 = (x0 + 1) - ix, (y0 + 1) - iy
    cl = lambda v, hi: v.long().clamp(0, hi)
    flat = image.reshape(B, C, ih * iw)

    def tap(yy, xx):
        idx = (cl(yy, ih - 1) * iw + cl(xx, iw - 1)).unsqueeze(1).expand(-1, C, -1)
        return torch.gather(flat, 2, idx)
    return (tap(y0, x0) * (wx0 * wy0).unsqueeze(1) + tap(y0, x0 + 1) * (wx1 * wy0).unsqueeze(1)
            + tap(y0 + 1, x0) * (wx0 * wy1).unsqueeze(1) + tap(y0 + 1, x0 + 1) * (wx1 * wy1).unsqueeze(1))


def decoder_first_output(planes, w1, b1, w2, b2, coords):
    """TriplanarDecoder (generator.py:302-331) on coords [B,N,3] in [-1,1]: mean of the three
    plane fetches -> 32->64 softplus -> first of the 1+A outputs.  planes [B,3,32,R,R]."""
    e = (bilinear_border(planes[:, 0], coords[..., 0], coords[..., 1])
         + bilinear_border(planes[:, 1], coords[..., 0], coords[..., 2])
         + bilinear_border(planes[:, 2], coords[..., 1], coords[..., 2])) / 3
    h = F.softplus(F.linear(e.transpose(1, 2), w1, b1))
    return F.linear(h, w2[:1], b2[:1])[..., 0]


def laplace_cdf(x, beta):
    return 0.5 + 0.5 * torch.sign(x) * (1 - torch.exp(-x.abs() / beta))


def laplace_pdf(x, beta):
    return 0.5 * torch.exp(-x.abs() / beta) / beta


def heads(planes, w1, b1, w2, b2, beta, scene_range, points, request, perturb=None, use_sdf=True):
    """generator.py:520-585 -> dict of per-image losses [B].  ``points`` [B,N,3] world units
    (stratified_points), ``perturb`` [B,N,3] the ``randn_like`` draw of :553-555."""
    out = {}
    pts = points
    if 'sdf_eikonal_loss' in request:
        pts = points.detach().requires_grad_()
    coords = pts / scene_range
    d = decoder_first_output(planes, w1, b1, w2, b2, coords)
    if 'sdf_eikonal_loss' in request:
        g, = torch.autograd.grad(d.sum(), pts, create_graph=True)
        out['sdf_eikonal_loss'] = ((g.norm(dim=-1) - 1) ** 2).mean(dim=1)
    if 'sdf_distance_loss' in request:
        target = points.detach().norm(dim=-1) - 1
        out['sdf_distance_loss'] = F.mse_loss(d, target, reduction='none').mean(dim=1)
    tv = 'total_variation_loss' in request
    if tv:
        d2 = decoder_first_output(planes, w1, b1, w2, b2, coords.detach() + perturb * 0.004)
    if use_sdf:
        if tv:
            out['total_variation_loss'] = (laplace_cdf(-d, beta) - laplace_cdf(-d2, beta)).abs().mean(dim=1)
        if 'entropy_loss' in request:
            out['entropy_loss'] = laplace_pdf(-d, beta).mean(dim=1)
    else:
        t = torch.sigmoid(d - 1)
        if tv:
            out['total_variation_loss'] = (t - torch.sigmoid(d2 - 1)).abs().mean(dim=1)
        if 'entropy_loss' in request:
            out['entropy_loss'] = (t * (1 - t)).mean(dim=1)
    return out
