import sys, torch
sys.path.insert(0, '.')
from nerf_from_image_b200 import fused, synthetic
B,H,W,S=8,128,128,64
ds=synthetic.DATASET_CONFIGS['p3d_car']
sc=synthetic.make_scene(1,B,plane_res=256,scene_range=ds['scene_range'],device='cuda')
cm=synthetic.make_cameras(1,B,radius=ds['radius'],device='cuda')
nt,nu=synthetic.make_noise(1,B,H,W,S,device='cuda')
for mode,name in ((0x1002,'normal'),(0x1102,'skip_gather'),(0x1202,'skip_consumer')):
    cfg=fused.RenderConfig(scene_range=sc['scene_range'],mlp_mode=mode)
    buf=torch.zeros(32,device='cuda'); fused.DEBUG_BUF=buf
    with torch.no_grad():
        for _ in range(2):
            fused.fused_render(sc['planes'],sc['w1'],sc['b1'],sc['w2'],sc['b2'],sc['palette'],sc['beta'],sc['alpha'],cm['c2w'],cm['focal'],None,None,cfg,H,W,S,nt,nu)
    torch.cuda.synchronize()
    b=buf.cpu()
    steps_p=b[6].item()/2; steps_c=b[14].item()
    print(name)
    print('  producer set0 (per step it handled, cycles): wait_free %.0f taps %.0f gather %.0f fence+bar %.0f issue1 %.0f' % tuple((b[i]/max(steps_p,1)).item() for i in range(5)))
    print('  consumer (per step, cycles): wait_d1 %.0f epi1 %.0f waitst+fence+bar %.0f issue2 %.0f wait_d2 %.0f rest(head,composite) %.0f' % tuple((b[8+i]/max(steps_c,1)).item() for i in range(6)))
