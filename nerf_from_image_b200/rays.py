"""Host-side (torch) ray generation, used ONLY to chain ray gradients to the
camera parameters in the backward pass.

The forward kernels generate rays themselves (csrc/nfi_common.cuh setup_ray).
The backward kernel returns dL/d(origin) and dL/d(unit direction) per ray;
turning those into dL/d(tform_cam2world, focal, bbox, center) is a few
elementwise ops on [B,H,W,3] tensors, so it is left to autograd over this
function (same camera model as /root/reference/lib/nerf_utils.py:28-91 plus
the F.normalize of run.py:196).
"""

import torch
import torch.nn.functional as F


def unit_rays(height, width, c2w, focal, center, bbox, rows=None):
    """``rows`` = (row_offset, full_height): the rays of rows [row_offset, row_offset + height)
    of images full_height rows tall (parallel.render_row_sharded)."""
    B = c2w.shape[0]
    dev, dt = c2w.device, c2w.dtype
    r0, hfull = rows if rows is not None else (0, height)
    u = (torch.arange(width, device=dev, dtype=dt) / width).view(1, 1, width)
    v = ((torch.arange(height, device=dev, dtype=dt) + r0) / hfull).view(1, height, 1)
    u = u.expand(B, height, width)
    v = v.expand(B, height, width)
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
            u = (bbox[:, 1, 0, None, None] * (u + 0.5) + bbox[:, 0, 0, None, None]) * 0.5
            v = -(bbox[:, 1, 1, None, None] * (-v + 0.5) + bbox[:, 0, 1, None, None]) * 0.5
        f = focal.view(-1, 1, 1)
        cam = torch.stack((u / f, -(v / f), -torch.ones_like(u)), dim=-1)
        dirs = (cam[..., None, :] * rot).sum(-1)
        origins = trans.expand(B, height, width, 3)
    else:
        u = (u - 0.5) * 2
        v = (v - 0.5) * 2
        if bbox is not None:
            u = bbox[:, 1, 0, None, None] * (u / 2 + 0.5) + bbox[:, 0, 0, None, None]
            v = -(bbox[:, 1, 1, None, None] * (-v / 2 + 0.5) + bbox[:, 0, 1, None, None])
        cam_o = torch.stack((u, -v, torch.zeros_like(u)), dim=-1)
        origins = (cam_o[..., None, :] * rot).sum(-1) + trans
        dirs = (-c2w[:, None, None, :3, 2] / c2w[:, None, None, 3, 3].unsqueeze(-1)
                ).expand(B, height, width, 3)
    return origins, F.normalize(dirs, dim=-1)
