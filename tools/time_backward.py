"""Times the fused render at config-2 geometry: forward (no grad), forward in grad mode (saves
the fine depths), and forward+backward in the inversion setting (grads to planes and palette,
decoder frozen; with `cam` also to tform_cam2world), the backward on its own events.
Usage: python tools/time_backward.py [batch] [cam]"""
import sys, torch
sys.path.insert(0, '.')
from nerf_from_image_b200 import fused, _lib
from fixtures import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
CAM = len(sys.argv) > 2 and sys.argv[2] == 'cam'
H, W, S = 128, 128, 64
ds = synthetic.DATASET_CONFIGS['p3d_car']
sc = synthetic.make_scene(1, B, plane_res=256, scene_range=ds['scene_range'], device='cuda')
cm = synthetic.make_cameras(1, B, radius=ds['radius'], device='cuda')
nt, nu = synthetic.make_noise(1, B, H, W, S, device='cuda')
cfg = fused.RenderConfig(scene_range=sc['scene_range'])
planes = sc['planes'].clone().requires_grad_()
pal = sc['palette'].clone().requires_grad_()
c2w = cm['c2w'].clone().requires_grad_() if CAM else cm['c2w']
ev = lambda: torch.cuda.Event(enable_timing=True)
bwd_ms = []
def run(mode):
    with torch.set_grad_enabled(mode != 'nograd'):
        rgb, depth, mask, _ = fused.fused_render(planes, sc['w1'], sc['b1'], sc['w2'], sc['b2'], pal, sc['beta'],
                                                 sc['alpha'], c2w, cm['focal'], None, None, cfg, H, W, S, nt, nu)
        if mode == 'bwd':
            loss = rgb.square().mean() + mask.mean()
            a, b = ev(), ev()
            a.record()
            loss.backward()
            b.record()
            bwd_ms.append((a, b))
            planes.grad = None; pal.grad = None
            if CAM: c2w.grad = None
print('library:', _lib.LIB_PATH)
for mode, label in (('nograd', 'fwd only'), ('grad', 'fwd, grad mode (z_fine saved)'),
                    ('bwd', 'fwd+bwd (planes, palette%s grads)' % (', c2w' if CAM else ''))):
    for _ in range(2): run(mode)
    torch.cuda.synchronize()
    bwd_ms.clear()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(3): run(mode)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    extra = ''
    if mode == 'bwd':
        extra = '; backward() alone %.2f ms' % (sum(a.elapsed_time(b) for a, b in bwd_ms) / len(bwd_ms))
    print('%s: %.2f ms per step at B=%d (%.1f M rays/s)%s' % (label, ms, B, B * H * W / ms / 1e3, extra))
