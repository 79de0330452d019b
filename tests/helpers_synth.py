"""Loading of the synthesis-network fixtures (tests/golden/synth/*.npz)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'synth')
CASES = ['synth_c64_const', 'synth_mixed_nonoise']


def load_case(name, device='cpu'):
    """-> (params dict as oracle.synthesis_oracle.extract_params makes it, ws, img, noise_mode)."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    p = {k[len('param:'):]: torch.from_numpy(z[k]).to(device) for k in z.files
         if k.startswith('param:')}
    layers = {str(k): dict(use_noise=bool(u), up=str(k).endswith('conv0'))
              for k, u in zip(z['meta_layers'], z['meta_use_noise'])}
    res, ch, wd = (int(x) for x in z['meta_dims'])
    p['meta'] = dict(img_resolution=res, img_channels=ch, w_dim=wd,
                     resolutions=[int(r) for r in z['meta_resolutions']], layers=layers)
    mode = str(z['noise_mode'])
    return p, torch.from_numpy(z['ws']).to(device), torch.from_numpy(z['img']).to(device), mode


def const_noises(p):
    """The ``noise_const * noise_strength`` tensors of noise_mode='const' (stylegan.py:341-343)."""
    out = {}
    for key, info in p['meta']['layers'].items():
        if info['use_noise'] and float(p[key + '.noise_strength']) != 0.0:
            out[key] = (p[key + '.noise_const'] * p[key + '.noise_strength'])[None, None]
    return out
