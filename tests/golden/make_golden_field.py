"""Generates tests/golden/field/*.npz by running the UNMODIFIED reference here.

    python tests/golden/make_golden_field.py

sampler_*.npz: inputs (planes, effective decoder weights, palette, beta/alpha,
points) and what the reference Generator's own ``sampler`` closure
(models/generator.py:587-681) returned for them.  pose_*.npz: pose parameters,
``lib/pose_utils.py:48-70`` ``pose_to_matrix`` outputs and reference-autograd
gradients of a fixed random functional.  The GPU box has no /root/reference;
there these files pin the oracle and the CUDA kernels to the reference.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_lift as RL  # noqa: E402
from tests import helpers as Hh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'field')

SAMPLER = {
    # name: (attention_values, use_sdf, requested outputs, bbox_debug)
    'sampler_palette': (10, True, ['sdf_distance', 'sigma', 'rgb', 'semantics', 'normals'], False),
    'sampler_direct_rgb': (0, True, ['sdf_distance', 'sigma', 'rgb'], False),
    'sampler_density': (10, False, ['sigma', 'rgb', 'semantics'], False),
    'sampler_bbox': (10, True, ['sigma', 'coords'], True),
}


def points(scene, batch, n, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, 4, n // 4, 3, generator=g) * 2 - 1) * 1.15 * scene['scene_range']


def pose_inputs(seed, batch, persp):
    g = torch.Generator().manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn(batch, 4, generator=g), dim=-1)
    t2 = 0.3 * torch.randn(batch, 2, generator=g)
    s = 0.8 + 0.5 * torch.rand(batch, generator=g)
    z0 = 0.5 * torch.randn(batch, generator=g) if persp else None
    return z0, t2, s, q


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    for name, (A, use_sdf, request, bbox) in SAMPLER.items():
        scene, _ = Hh.make_case('p3d_plain', seed=31, batch=2, plane_res=16, attention_values=A)
        x = points(scene, 2, 600, 17)
        ref = RL.reference_sampler(scene, x.clone(), request, use_sdf=use_sdf, bbox_debug=bbox)
        arrays = {'out_' + k: v for k, v in ref.items() if k != 'coords'}
        arrays['points'] = x
        for k in ('planes', 'w1', 'b1', 'w2', 'b2', 'palette', 'beta', 'alpha'):
            if scene[k] is not None:
                arrays['in_' + k] = scene[k]
        meta = dict(scene_range=scene['scene_range'], A=A, use_sdf=int(use_sdf), bbox=int(bbox),
                    coords=int('coords' in request))
        np.savez_compressed(os.path.join(OUT, name + '.npz'),
                            **{k: v.detach().numpy().astype(np.float32) for k, v in arrays.items()},
                            **{'meta_' + k: np.array(v) for k, v in meta.items()})
        print(name, {k: tuple(v.shape) for k, v in ref.items()})

    pu = RL.reference_pose_utils()
    B = 7
    for persp in (True, False):
        for flipped in (True, False):
            z0, t2, s, q = pose_inputs(41, B, persp)
            leaves = [t.clone().requires_grad_() for t in (z0, t2, s, q) if t is not None]
            a = (leaves[0], leaves[1], leaves[2], leaves[3]) if persp else (None, *leaves)
            mat, focal = pu.pose_to_matrix(*a, flipped)
            g = torch.Generator().manual_seed(43)
            wm, wf = torch.randn(B, 4, 4, generator=g), torch.randn(B, generator=g)
            loss = (mat * wm).sum() + ((focal * wf).sum() if persp else 0)
            grads = torch.autograd.grad(loss, leaves)
            names = ['z0', 't2', 's', 'q'] if persp else ['t2', 's', 'q']
            arrays = dict(t2=t2, s=s, q=q, mat=mat, wm=wm, wf=wf)
            if persp:
                arrays.update(z0=z0, focal=focal)
            for n, gr in zip(names, grads):
                arrays['grad_' + n] = gr
            name = 'pose_%s_%s' % ('persp' if persp else 'ortho', 'flipped' if flipped else 'plain')
            np.savez_compressed(os.path.join(OUT, name + '.npz'),
                                **{k: v.detach().numpy().astype(np.float32) for k, v in arrays.items()},
                                meta_flipped=np.array(int(flipped)))
            print(name)


if __name__ == '__main__':
    main()
