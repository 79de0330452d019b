"""oracle/synthesis_oracle.py (the restatement of the reference's plane producer, SURVEY.md
section 8f N1) against the golden fixtures the reference produced, and -- where the reference
can be imported -- against the UNMODIFIED models.stylegan.SynthesisNetwork block by block."""
import pytest
import torch

from oracle import reference_lift as RL
from oracle import synthesis_oracle as SO
from tests import helpers_synth as HS


@pytest.mark.parametrize('case', HS.CASES)
def test_oracle_reproduces_the_reference_fixtures(case):
    p, ws, img, mode = HS.load_case(case)
    noises = HS.const_noises(p) if mode == 'const' else None
    with torch.no_grad():
        got = SO.synthesis_forward(p, ws, noises)
    assert got.shape == img.shape
    assert (got - img).abs().max().item() < 2e-5 * max(1.0, img.abs().max().item())


@pytest.mark.skipif(not RL.available(), reason='reference not importable')
def test_oracle_matches_the_reference_module_block_by_block():
    RL._import_reference()
    from models import stylegan
    torch.manual_seed(3)
    net = stylegan.SynthesisNetwork(512, 32, 96, channel_base=4096, channel_max=128).eval()
    with torch.no_grad():
        for n, q in net.named_parameters():
            if n.endswith('noise_strength'):
                q.fill_(0.05)
    ws = torch.randn(2, net.num_ws, 512)
    p = SO.extract_params(net)
    with torch.no_grad():
        img, blocks = SO.synthesis_forward(p, ws, HS.const_noises(p), return_blocks=True)
        # the reference, one block at a time (SynthesisNetwork.forward, stylegan.py:475-490)
        x = ref_img = None
        w_idx = 0
        for r in net.block_resolutions:
            blk = getattr(net, 'b%d' % r)
            cur = ws.narrow(1, w_idx, blk.num_conv + blk.num_torgb)
            w_idx += blk.num_conv
            x, ref_img = blk(x, ref_img, cur, noise_mode='const')
            ox, oimg = blocks['b%d' % r]
            assert (ox - x).abs().max().item() < 1e-4 * max(1.0, x.abs().max().item()), r
            assert (oimg - ref_img).abs().max().item() < 1e-4 * max(1.0, ref_img.abs().max().item()), r
        assert torch.equal(img, blocks['b32'][1])
    assert net.num_ws == 2 * len(net.block_resolutions)


def test_up_layer_is_the_composition_the_kernels_implement():
    """The stride-2 transposed conv as four parity-phase convolutions over the input grid
    (4 + 2 + 2 + 1 taps) followed by the 4x4 FIR -- the decomposition of csrc/nfi_synth.cu --
    equals conv_transpose2d + filter (stylegan.py:98-102)."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    x = torch.randn(2, 8, 5, 7)
    w = torch.randn(6, 8, 3, 3)
    f = SO.fir_kernel()
    ref = F.conv_transpose2d(x, w.transpose(0, 1), stride=2)
    H, W = x.shape[2:]
    raw = torch.zeros(2, 6, 2 * H + 1, 2 * W + 1)
    xp = F.pad(x, (1, 1, 1, 1))
    for py in range(2):
        for px in range(2):
            DH, DW = (H if py else H + 1), (W if px else W + 1)
            acc = torch.zeros(2, 6, DH, DW)
            for ky in range(py, 3, 2):
                for kx in range(px, 3, 2):
                    dy, dx = -(ky // 2), -(kx // 2)
                    patch = xp[:, :, 1 + dy:1 + dy + DH, 1 + dx:1 + dx + DW]
                    acc += torch.einsum('bchw,oc->bohw', patch, w[:, :, ky, kx])
            raw[:, :, py::2, px::2] = acc
    assert (raw - ref).abs().max().item() < 1e-4
    # image upsampling weights of the ToRGB epilogue: out[2i] = 3/4 in[i] + 1/4 in[i-1], ...
    img = torch.randn(1, 3, 4, 4)
    up = SO.upsample_img(img, f)
    k = torch.tensor([0.25, 0.75, 0.75, 0.25])
    man = torch.zeros(1, 3, 8, 8)
    for u in range(8):
        for v in range(8):
            iy, ix = u // 2, v // 2
            jy, jx = (iy + 1 if u % 2 else iy - 1), (ix + 1 if v % 2 else ix - 1)
            val = 0.75 * 0.75 * img[:, :, iy, ix]
            if 0 <= jx < 4:
                val = val + 0.75 * 0.25 * img[:, :, iy, jx]
            if 0 <= jy < 4:
                val = val + 0.25 * 0.75 * img[:, :, jy, ix]
            if 0 <= jx < 4 and 0 <= jy < 4:
                val = val + 0.25 * 0.25 * img[:, :, jy, jx]
            man[:, :, u, v] = val
    assert (man - up).abs().max().item() < 1e-5
