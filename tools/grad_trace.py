"""Per-sample backward trace of one ray from both backward kernels (debug bit 0x4000 of mlp_mode):
z, sigma, w, T, dL/dsigma, dOut[0] and the 32-vector dL/d(feature) of every sample in merged
depth order -- where do the tensor-core and the SIMT kernel part?  Usage: grad_trace.py y x"""
import sys
import torch
sys.path.insert(0, '.')
from fixtures import synthetic
from tests import helpers as Hh
from nerf_from_image_b200 import fused
B, H, W, S = 1, 128, 128, 64
py, px = int(sys.argv[1]), int(sys.argv[2])
scene, cams = Hh.make_case('p3d_plain', batch=B, plane_res=256, device='cuda')
nt, nu = synthetic.make_noise(51, B, H, W, S, device='cuda')
g = torch.Generator().manual_seed(0)
wr = torch.randn(B, H, W, 3, generator=g).cuda()
wm = torch.randn(B, H, W, generator=g).cuda()
tr = {}
for tag, mode in (('simt', 1), ('tc', 4)):
    sc = dict(scene, planes=scene['planes'].clone().requires_grad_())
    rgb, _, mask, _ = Hh.run_cuda(sc, cams, H, W, S, nt, nu, mlp_mode=mode)
    buf = torch.zeros(2 * S * 40, device='cuda')
    fused.DEBUG_BUF, fused.DEBUG_RAY = buf, py * W + px
    ((rgb * wr).sum() + (mask * wm).sum()).backward()
    fused.DEBUG_BUF = fused.DEBUG_RAY = None
    torch.cuda.synchronize()
    tr[tag] = (buf[:2 * S * 8].view(2 * S, 8).cpu(), buf[2 * S * 8:].view(2 * S, 32).cpu())
a, b = tr['simt'], tr['tc']
print('pixel (%d,%d)   columns: z sigma w T dsig dOut0 s_i delta | |dF| simt, |dF| tc, |dF diff|' % (py, px))
for i in range(2 * S):
    dd = (a[1][i] - b[1][i]).norm().item()
    flag = ' <<<' if dd > 0.05 * max(a[1][i].norm().item(), 1e-12) and a[1][i].norm() > 1e-6 else ''
    if flag or i % 16 == 0:
        print('%3d simt %s | %.3e' % (i, ' '.join('% .4e' % v for v in a[0][i].tolist()), a[1][i].norm().item()))
        print('    tc   %s | %.3e  diff %.3e%s' % (' '.join('% .4e' % v for v in b[0][i].tolist()), b[1][i].norm().item(), dd, flag))
        if flag:
            print('    dF simt', ' '.join('% .2e' % v for v in a[1][i][:16].tolist()))
            print('    dF tc  ', ' '.join('% .2e' % v for v in b[1][i][:16].tolist()))
