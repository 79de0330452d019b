"""Generates tests/golden/*.npz by running the UNMODIFIED reference here.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py
Each file holds the inputs (planes, effective decoder weights, palette,
beta/alpha, cameras, the two noise tensors the reference drew) and what the
reference's own render() (run.py:176-350, lifted by oracle/reference_lift.py)
returned for them, plus reference-autograd gradients of a fixed random
functional of (rgb, mask).  The GPU box has no /root/reference: there the
fixtures are what pins the CUDA path and the oracle to the reference.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_lift as RL  # noqa: E402
from tests import helpers as Hh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# name: (case, kwargs)
GOLDEN = {
    'p3d_bbox_rand': ('p3d_bbox', dict(randomize=True)),
    'p3d_plain_det': ('p3d_plain', dict(randomize=False)),
    'cub_ortho_rand': ('cub_ortho', dict(randomize=True)),
    'chairs_white_rand': ('chairs_white_center', dict(randomize=True)),
    'p3d_nofine_rand': ('p3d_bbox', dict(randomize=True, fine_sampling=False)),
    'p3d_semantics': ('p3d_plain', dict(randomize=True, compute_semantics=True)),
    'p3d_coords': ('p3d_plain', dict(randomize=True, compute_coords=True)),
    'p3d_direct_rgb': ('p3d_plain', dict(randomize=True, attention_values=0)),
    'p3d_density': ('p3d_plain', dict(randomize=True, use_sdf=False)),
    # --use_viewdir (the CARLA models): view-direction-conditioned colour, generator.py:189-253
    'chairs_viewdir': ('chairs_white_center', dict(randomize=True, viewdir=True)),
    'p3d_viewdir_direct_rgb': ('p3d_bbox', dict(randomize=True, viewdir=True, attention_values=0)),
}
B, R, H, W, S = 2, 16, 12, 20, 8


def loss_weights(shape_rgb, shape_mask):
    g = torch.Generator().manual_seed(99)
    return torch.randn(shape_rgb, generator=g), torch.randn(shape_mask, generator=g)


def main():
    torch.set_num_threads(4)
    for name, (case, kw) in GOLDEN.items():
        kw = dict(kw)
        A = kw.pop('attention_values', 10)
        scene, cams = Hh.make_case(case, seed=11, batch=B, plane_res=R, attention_values=A)
        viewdir = kw.pop('viewdir', False)
        if viewdir:
            from fixtures import synthetic
            scene = synthetic.add_view_mapper(scene)
        planes = scene['planes'].clone().requires_grad_()
        palette = scene['palette'].clone().requires_grad_() if A > 0 else None
        cams_l = dict(cams)
        cams_l['c2w'] = cams['c2w'].clone().requires_grad_()
        gen = RL.build_reference_generator(scene, use_sdf=kw.get('use_sdf', True))
        out, nt, nu = RL.reference_render(scene, cams_l, H, W, S, seed=21, generator=gen,
                                          planes=planes, palette=palette, **kw)
        rgb, depth, mask, normals, extra, _ = out
        wr, wm = loss_weights(rgb.shape, mask.shape)
        loss = (rgb * wr).sum() + (mask * wm).sum()
        dec = gen.decoder.net
        leaves = [planes, dec[0].weight, dec[0].bias, dec[2].weight, dec[2].bias, cams_l['c2w']]
        names = ['planes', 'w1', 'b1', 'w2', 'b2', 'c2w']
        if A > 0:
            leaves.append(palette); names.append('palette')
        if kw.get('use_sdf', True):
            leaves += [gen.beta, gen.alpha]; names += ['beta', 'alpha']
        gains = {'w1': dec[0].weight_gain, 'b1': dec[0].bias_gain,
                 'w2': dec[2].weight_gain, 'b2': dec[2].bias_gain}
        if viewdir:
            vm = gen.viewdir_mapper
            leaves += [vm.output.weight, vm.output.bias, vm.fc0.weight, vm.fc4.weight, vm.norm3.bias]
            names += ['w3', 'b3', 'vm_fc0_w', 'vm_fc4_w', 'vm_norm3_b']
            gains.update(w3=vm.output.weight_gain, b3=vm.output.bias_gain,
                         vm_fc0_w=vm.fc0.weight_gain, vm_fc4_w=vm.fc4.weight_gain)
        grads = torch.autograd.grad(loss, leaves)
        arrays = dict(rgb=rgb, depth=depth, mask=mask)
        if extra is not None:
            arrays['extra'] = extra
        for n, g in zip(names, grads):
            # the fixture stores gradients w.r.t. the EFFECTIVE weights
            arrays['grad_' + n] = g / gains[n] if n in gains else g
        for k in ('planes', 'w1', 'b1', 'w2', 'b2', 'palette', 'beta', 'alpha', 'w3', 'b3'):
            if scene.get(k) is not None:
                arrays['in_' + k] = scene[k]
        for k, v in scene.get('view_mapper', {}).items():
            arrays['in_vm_' + k] = v
        for k, v in cams.items():
            if v is not None:
                arrays['cam_' + k] = v
        if nt is not None:
            arrays['noise_t'] = nt
        if nu is not None:
            arrays['noise_u'] = nu
        meta = dict(H=H, W=W, S=S, scene_range=scene['scene_range'],
                    white_background=int(scene['white_background']), A=A,
                    use_sdf=int(kw.get('use_sdf', True)),
                    fine_sampling=int(kw.get('fine_sampling', True)),
                    compute_semantics=int(kw.get('compute_semantics', False)),
                    compute_coords=int(kw.get('compute_coords', False)))
        np.savez_compressed(os.path.join(OUT, name + '.npz'),
                            **{k: v.detach().numpy().astype(np.float32) for k, v in arrays.items()},
                            **{'meta_' + k: np.array(v) for k, v in meta.items()})
        print(name, 'rgb', tuple(rgb.shape), 'mask mean %.3f' % mask.mean().item())


if __name__ == '__main__':
    main()
