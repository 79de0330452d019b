"""A/B timing of render-kernel builds on one box: kernel-only CUDA-event times of the forward at
config-2 geometry for the library NFI_LIB_PATH points at (default: the in-tree build), and a
checksum of the image so that builds that must be bit-identical can be seen to be.
Usage: [NFI_LIB_PATH=...] python tools/ab_forward.py [batch] [steps] [S] [res]"""
import hashlib
import sys
import torch
sys.path.insert(0, '.')
from nerf_from_image_b200 import fused, _lib
from fixtures import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
H = W = int(sys.argv[4]) if len(sys.argv) > 4 else 128
ds = synthetic.DATASET_CONFIGS['p3d_car']
sc = synthetic.make_scene(1234, B, plane_res=256, scene_range=ds['scene_range'], device='cuda')
cm = synthetic.make_cameras(1234, B, radius=ds['radius'], device='cuda')
nt, nu = synthetic.make_noise(1234, B, H, W, S, device='cuda')
cfg = fused.RenderConfig(scene_range=sc['scene_range'])
def step():
    with torch.no_grad():
        return fused.fused_render(sc['planes'], sc['w1'], sc['b1'], sc['w2'], sc['b2'], sc['palette'],
                                  sc['beta'], sc['alpha'], cm['c2w'], cm['focal'], None, None, cfg,
                                  H, W, S, nt, nu)
for _ in range(3):
    out = step()
torch.cuda.synchronize()
ev = []
fused.KERNEL_EVENTS = ev
for _ in range(K):
    out = step()
fused.KERNEL_EVENTS = None
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)
digest = hashlib.sha256(out[0].cpu().numpy().tobytes()).hexdigest()[:12]
print('%-60s B=%d S=%d %dx%d kernel ms: median %.3f min %.3f max %.3f  (%.1f M rays/s)  rgb sha %s'
      % (_lib.LIB_PATH.split('/')[-1], B, S, H, W, ms[len(ms) // 2], ms[0], ms[-1],
         B * H * W / ms[len(ms) // 2] / 1e3, digest))
