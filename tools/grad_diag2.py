"""Localises the plane-gradient error of the tensor-core backward: worst texels, and the mixed
run (tensor-core forward outputs into the SIMT backward).  Usage: python tools/grad_diag2.py [case]"""
import sys
import torch
sys.path.insert(0, '.')
from fixtures import synthetic
from tests import helpers as Hh
B, H, W, S = 1, 128, 128, 64
case = sys.argv[1] if len(sys.argv) > 1 else 'p3d_plain'
scene, cams = Hh.make_case(case, batch=B, plane_res=256, device='cuda')
nt, nu = synthetic.make_noise(51, B, H, W, S, device='cuda')
g = torch.Generator().manual_seed(0)
wr = torch.randn(B, H, W, 3, generator=g).cuda()
wm = torch.randn(B, H, W, generator=g).cuda()
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
res = {}
for tag in ('simt', 'tc', 'tc2', 'mix'):
    sc = dict(scene, planes=scene['planes'].clone().requires_grad_())
    if tag == 'mix':  # tensor-core forward, weights need grads -> SIMT backward
        for k in ('w1', 'b1', 'w2', 'b2'):
            sc[k] = scene[k].clone().requires_grad_()
    rgb, _, mask, _ = Hh.run_cuda(sc, cams, H, W, S, nt, nu, mlp_mode=1 if tag == 'simt' else 4)
    if tag == 'tc2':
        torch.cuda.synchronize()
    loss = (rgb * wr).sum() + (mask * wm).sum()
    res[tag] = torch.autograd.grad(loss, sc['planes'])[0]
    res[tag + '_rgb'] = rgb.detach()
print('tc run 1 vs run 2 (same inputs): rel %.2e, max abs %.3e' % (rel(res['tc2'], res['tc']), (res['tc2'] - res['tc']).abs().max().item()))
print(case, 'tc vs simt %.2e   mix vs simt %.2e   rgb tc vs simt %.2e' % (
    rel(res['tc'], res['simt']), rel(res['mix'], res['simt']), rel(res['tc_rgb'], res['simt_rgb'])))
for tag in ('tc', 'mix'):
    err = (res[tag] - res['simt'])[0]                 # [3,32,R,R]
    e = err.square().sum(1)                           # [3,R,R]
    top = e.flatten().topk(12)
    print(tag, 'worst texels (plane, y, x): sum|g| simt, sum|g| %s, |err|' % tag)
    for v, idx in zip(top.values.tolist(), top.indices.tolist()):
        pl, y, x = idx // (256 * 256), (idx // 256) % 256, idx % 256
        print('   (%d,%3d,%3d)  %.4e  %.4e  %.4e' % (pl, y, x, res['simt'][0, pl, :, y, x].abs().sum().item(),
                                                      res[tag][0, pl, :, y, x].abs().sum().item(), v ** 0.5))
# per-pixel image difference, largest
d = (res['tc_rgb'] - res['simt_rgb']).abs().sum(-1)[0]
top = d.flatten().topk(8)
print('largest |rgb tc - rgb simt| pixels (y, x, diff, mask):')
for v, idx in zip(top.values.tolist(), top.indices.tolist()):
    print('   (%3d,%3d) %.3e' % (idx // W, idx % W, v))
