"""Finds the rays on which the tensor-core backward disagrees with the SIMT backward (losses
restricted to one image row, then one pixel) and prints what is special about their samples."""
import sys
import torch
sys.path.insert(0, '.')
from fixtures import synthetic
from tests import helpers as Hh
from oracle import render_oracle as O
B, H, W, S = 1, 128, 128, 64
case = sys.argv[1] if len(sys.argv) > 1 else 'p3d_plain'
scene, cams = Hh.make_case(case, batch=B, plane_res=256, device='cuda')
nt, nu = synthetic.make_noise(51, B, H, W, S, device='cuda')
g = torch.Generator().manual_seed(0)
wr = torch.randn(B, H, W, 3, generator=g).cuda()
wm = torch.randn(B, H, W, generator=g).cuda()
rel = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-20)).item()
outs = {}
for tag, mode in (('simt', 1), ('tc', 4)):
    sc = dict(scene, planes=scene['planes'].clone().requires_grad_())
    rgb, _, mask, _ = Hh.run_cuda(sc, cams, H, W, S, nt, nu, mlp_mode=mode)
    outs[tag] = (sc['planes'], rgb, mask)
def grads(sel):
    res = []
    for tag in ('simt', 'tc'):
        pl, rgb, mask = outs[tag]
        loss = (rgb * wr * sel[..., None]).sum() + (mask * wm * sel).sum()
        res.append(torch.autograd.grad(loss, pl, retain_graph=True)[0])
    return res
bad_rows = []
for r in range(H):
    sel = torch.zeros(B, H, W, device='cuda'); sel[:, r] = 1
    a, b = grads(sel)
    e = rel(b, a)
    if e > 5e-4: bad_rows.append((r, e))
print(case, 'rows with tc-vs-simt plane-gradient error > 5e-4:', [(r, '%.1e' % e) for r, e in bad_rows])
bad = []
for r, _ in bad_rows[:6]:
    for c in range(W):
        sel = torch.zeros(B, H, W, device='cuda'); sel[:, r, c] = 1
        a, b = grads(sel)
        if a.norm() > 0 and rel(b, a) > 5e-4: bad.append((r, c, rel(b, a), a.norm().item()))
print('bad pixels (row, col, rel err, |g|):', [(r, c, '%.1e' % e, '%.2e' % n) for r, c, e, n in bad])
# what is special about them
fn = outs['tc'][1].grad_fn
saved = dict(zip(fn.saved_names, fn.saved_tensors))
zf = saved['z_fine'].view(B, H, W, S)
o, d = O.ray_bundle(H, W, cams['focal'], cams['c2w'], cams['bbox'], cams['center'])
d = torch.nn.functional.normalize(d, dim=-1)
near, far, hit = O.near_far_planes(o, d, scene['scene_range'])
tc_ = O.coarse_depths(near, far, S, nt)
for r, c, e, n in bad[:8]:
    z = torch.cat((tc_[0, r, c], zf[0, r, c])).sort().values
    dz = z[1:] - z[:-1]
    pts = (o[0, r, c][None] + d[0, r, c][None] * z[:, None]) / scene['scene_range']
    print(' pixel (%d,%d): near %.4f far %.4f hit %d  min dz %.3e  #dz==0 %d  #|x|>1 %d  mask %.4f  fine z range [%.4f, %.4f]  dup fine %d'
          % (r, c, near[0, r, c], far[0, r, c], int(hit[0, r, c]), dz.min().item(), int((dz == 0).sum()),
             int((pts.abs() > 1).any(-1).sum()), outs['tc'][2][0, r, c].item(), zf[0, r, c].min().item(),
             zf[0, r, c].max().item(), int((zf[0, r, c][1:] == zf[0, r, c][:-1]).sum())))
    ix = (pts + 1) * 0.5 * 255
    fr = ix - ix.floor()
    print('    samples with a fractional texel coordinate < 1e-4 or > 1-1e-4:', int(((fr < 1e-4) | (fr > 1 - 1e-4)).any(-1).sum()),
          ' ties coarse==fine:', int((tc_[0, r, c][:, None] == zf[0, r, c][None, :]).sum()))
