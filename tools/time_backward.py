"""Times forward and forward+backward (inversion setting: grads to planes, palette, cameras;
decoder frozen) of the fused render at config-2 geometry.  Usage: python tools/time_backward.py [batch]"""
import sys, torch
sys.path.insert(0, '.')
from nerf_from_image_b200 import fused, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W, S = 128, 128, 64
ds = synthetic.DATASET_CONFIGS['p3d_car']
sc = synthetic.make_scene(1, B, plane_res=256, scene_range=ds['scene_range'], device='cuda')
cm = synthetic.make_cameras(1, B, radius=ds['radius'], device='cuda')
nt, nu = synthetic.make_noise(1, B, H, W, S, device='cuda')
cfg = fused.RenderConfig(scene_range=sc['scene_range'])
planes = sc['planes'].clone().requires_grad_()
pal = sc['palette'].clone().requires_grad_()
def run(grad):
    with torch.set_grad_enabled(grad):
        rgb, depth, mask, _ = fused.fused_render(planes, sc['w1'], sc['b1'], sc['w2'], sc['b2'], pal, sc['beta'],
                                                 sc['alpha'], cm['c2w'], cm['focal'], None, None, cfg, H, W, S, nt, nu)
        if grad:
            (rgb.square().mean() + mask.mean()).backward()
            planes.grad = None; pal.grad = None
for grad in (False, True):
    for _ in range(2): run(grad)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run(grad)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('%s: %.2f ms per step at B=%d (%.1f M rays/s)' % ('fwd+bwd (planes, palette grads)' if grad else 'fwd only', ms, B, B * H * W / ms / 1e3))
