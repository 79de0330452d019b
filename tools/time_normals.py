"""compute_normals render (evaluation / visualisation calls) at config-2 geometry: the pipelined
pair render_forward_pipe + render_normals_pipe (NFI_MLP_AUTO) against the fp32 SIMT kernel.
Usage: python tools/time_normals.py [batch]"""
import sys, torch
sys.path.insert(0, '.')
from nerf_from_image_b200 import fused
from fixtures import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H, W, S = 128, 128, 64
ds = synthetic.DATASET_CONFIGS['p3d_car']
sc = synthetic.make_scene(1, B, plane_res=256, scene_range=ds['scene_range'], device='cuda')
cm = synthetic.make_cameras(1, B, radius=ds['radius'], device='cuda')
nt, nu = synthetic.make_noise(1, B, H, W, S, device='cuda')
outs = {}
for mode, label in ((0, 'render_forward_pipe + render_normals_pipe'), (1, 'render_forward_simt<NORM>')):
    cfg = fused.RenderConfig(scene_range=sc['scene_range'], mlp_mode=mode)
    def step():
        with torch.no_grad():
            return fused.fused_render(sc['planes'], sc['w1'], sc['b1'], sc['w2'], sc['b2'], sc['palette'], sc['beta'],
                                      sc['alpha'], cm['c2w'], cm['focal'], None, None, cfg, H, W, S, nt, nu,
                                      compute_normals=True)
    for _ in range(2): out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out = step()
    e1.record(); torch.cuda.synchronize()
    outs[mode] = out[4]
    print('%-44s %.2f ms per render with normals at B=%d' % (label, e0.elapsed_time(e1) / 5, B))
print('normals: pipelined vs SIMT rel-L2 %.2e' % ((outs[0] - outs[1]).norm() / outs[1].norm()).item())
