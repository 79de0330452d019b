"""For the rays found by grad_diag4: which SAMPLES of the ray carry the tensor-core backward's
plane-gradient error (position in the merged depth order, coarse / fine, weight)."""
import sys
import torch
sys.path.insert(0, '.')
from fixtures import synthetic
from tests import helpers as Hh
from oracle import render_oracle as O
B, H, W, S = 1, 128, 128, 64
case = 'p3d_plain'
pixels = [(74, 2), (107, 117), (111, 32)]
scene, cams = Hh.make_case(case, batch=B, plane_res=256, device='cuda')
nt, nu = synthetic.make_noise(51, B, H, W, S, device='cuda')
g = torch.Generator().manual_seed(0)
wr = torch.randn(B, H, W, 3, generator=g).cuda()
wm = torch.randn(B, H, W, generator=g).cuda()
outs = {}
for tag, mode in (('simt', 1), ('tc', 4)):
    sc = dict(scene, planes=scene['planes'].clone().requires_grad_())
    rgb, _, mask, _ = Hh.run_cuda(sc, cams, H, W, S, nt, nu, mlp_mode=mode)
    outs[tag] = (sc['planes'], rgb, mask)
fn = outs['tc'][1].grad_fn
saved = dict(zip(fn.saved_names, fn.saved_tensors))
zf = saved['z_fine'].view(B, H, W, S)
o, d = O.ray_bundle(H, W, cams['focal'], cams['c2w'], cams['bbox'], cams['center'])
d = torch.nn.functional.normalize(d, dim=-1)
near, far, hit = O.near_far_planes(o, d, scene['scene_range'])
tcz = O.coarse_depths(near, far, S, nt)
R = 256
for (r, c) in pixels:
    sel = torch.zeros(B, H, W, device='cuda'); sel[:, r, c] = 1
    gs = []
    for tag in ('simt', 'tc'):
        pl, rgb, mask = outs[tag]
        loss = (rgb * wr * sel[..., None]).sum() + (mask * wm * sel).sum()
        gs.append(torch.autograd.grad(loss, pl, retain_graph=True)[0][0])   # [3,32,R,R]
    err = (gs[1] - gs[0]).square().sum(1)       # [3,R,R]
    ref = gs[0].square().sum(1)
    zall = torch.cat((tcz[0, r, c], zf[0, r, c]))
    kind = torch.cat((torch.zeros(S), torch.ones(S))).cuda()
    order = zall.argsort(stable=True)
    z, kind = zall[order], kind[order]
    pts = (o[0, r, c][None] + d[0, r, c][None] * z[:, None]) / scene['scene_range']
    ix = ((pts + 1) * 0.5 * (R - 1)).clamp(0, R - 1)
    i0 = ix.floor().long().clamp(max=R - 2)
    pairs = ((0, 1), (0, 2), (1, 2))
    e_s, r_s = torch.zeros(2 * S), torch.zeros(2 * S)
    for s in range(2 * S):
        for pl, (a, b) in enumerate(pairs):
            xx, yy = i0[s, a].item(), i0[s, b].item()
            e_s[s] += err[pl, yy:yy + 2, xx:xx + 2].sum().item()
            r_s[s] += ref[pl, yy:yy + 2, xx:xx + 2].sum().item()
    top = e_s.topk(6)
    print('pixel (%d,%d): total err^2 %.3e, ref^2 %.3e' % (r, c, err.sum().item(), ref.sum().item()))
    for v, s in zip(top.values.tolist(), top.indices.tolist()):
        print('   merged index %3d (%s #%d)  z %.5f  err^2 near its taps %.3e  (ref^2 %.3e)  x/range (%.4f %.4f %.4f)'
              % (s, 'fine' if kind[s] > 0 else 'coarse', int((kind[:s] == kind[s]).sum()), z[s].item(), v, r_s[s],
                 pts[s, 0].item(), pts[s, 1].item(), pts[s, 2].item()))
