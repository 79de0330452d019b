"""CPU restatement of the reference's pose representation -- TEST INFRASTRUCTURE.

Only tests/ may import this.  Follows /root/reference/lib/pose_utils.py:
``quaternion_rotate_vector`` (:32-38), ``quaternion_to_matrix`` (:41-45) and
``pose_to_matrix`` (:48-70).  Pinned to the reference by
tests/test_oracle_vs_reference.py (run here, where the reference imports) and
by the committed fixture tests/golden/pose_cases.npz it generated.
"""

import torch


def quaternion_rows(q):
    """[B,4] (w, x, y, z) -> [B,3,3] whose row i is the rotated unit vector e_i."""
    w = q[:, :1].unsqueeze(1)
    qv = q[:, 1:].unsqueeze(1).expand(-1, 3, -1)
    eye = torch.eye(3, dtype=q.dtype, device=q.device).unsqueeze(0).expand(q.shape[0], -1, -1)
    uv = torch.cross(qv, eye, dim=2)
    uuv = torch.cross(qv, uv, dim=2)
    return eye + 2 * (w * uv + uuv)


def pose_to_matrix(z0, t2, s, q, camera_flipped):
    rows = quaternion_rows(q)
    B = q.shape[0]
    bottom = torch.tensor([0., 0., 0., 1.], dtype=q.dtype, device=q.device).expand(B, 1, 4)
    if z0 is not None:
        f = 1 + z0.exp()
        t3 = torch.cat((t2 / s.unsqueeze(-1), (f / s).unsqueeze(-1)), dim=-1)
    else:
        f = None
        t3 = torch.cat((t2, 10 * torch.ones_like(t2[:, :1])), dim=-1)
    trans = (t3[:, None, :] * rows).sum(dim=-1, keepdim=True)
    top = torch.cat((rows, trans), dim=-1)
    if camera_flipped:
        top = top * torch.tensor([1., -1., -1., -1.], dtype=q.dtype, device=q.device)
    mat = torch.cat((top, bottom), dim=1)
    if z0 is not None:
        return mat, f / 2
    return mat / s[:, None, None], None
