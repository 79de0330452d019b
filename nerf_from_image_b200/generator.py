"""Generator front-end with the plane producer on sm_100a.

``FusedGeneratorFront(G)`` is the part of the reference's ``Generator.forward``
(/root/reference/models/generator.py:407-477,500-502) that turns the model input into the
radiance field -- latent selection / mapping network, palette (texture mapper), synthesis
network -- with the synthesis network run by ``synthesis.FusedSynthesis`` (tcgen05 kernels,
planes emitted channel-last).  ``G`` stays the reference's own module and the owner of every
parameter; its tiny MLPs (mapping network, texture mapper, encoder) run as they are.

It returns the reference's ``model_outputs`` dict with a ``'triplane'`` entry instead of the
``'sampler'`` closure: ``render.render`` hands that entry to the fused render kernels, which is
what it would have done with the closure's captured planes anyway (render.extract_field).

Envelope: requests within {'sampler', 'attention_values'} under ``torch.no_grad()`` (evaluation
renders, encoder-training targets, visualisation).  The regulariser heads and every call that
differentiates through the synthesis network are the reference module's; ``render`` picks this
front-end only when a call is inside the envelope (``render.enable_fused_synthesis``).
"""
import torch

from . import _lib
from .synthesis import FusedSynthesis

SUPPORTED_OUTPUTS = ('sampler', 'attention_values')


class FusedGeneratorFront:
    def __init__(self, generator):
        self.g = generator
        self.synthesis = FusedSynthesis(generator.synthesis_network)

    @staticmethod
    def supports(request_model_outputs, model_inputs):
        return (not torch.is_grad_enabled()
                and all(o in SUPPORTED_OUTPUTS for o in request_model_outputs)
                and all(k in ('freeze_noise', 'attention_values', 'attention_values_bias')
                        for k in model_inputs))

    def __call__(self, viewdir, c, request_model_outputs=['sampler'], model_inputs={}):
        g = self.g
        if not self.supports(request_model_outputs, model_inputs):
            raise _lib.NfiError('FusedGeneratorFront: outside its envelope (no_grad, outputs '
                                'within %r)' % (SUPPORTED_OUTPUTS,))
        if g.use_viewdir and viewdir is not None:
            raise NotImplementedError('--use_viewdir is outside the fused path')
        # ---- model input -> ws (generator.py:423-446)
        if g.use_encoder:
            z, image = c
            batch = z.shape[0]
            ws = g.mapping_network(z, g.emb(image))
        else:
            label = None
            if g.num_classes:
                if isinstance(c, (list, tuple)):
                    c, label = c
                    assert len(c.shape) == 2
                    label = g.class_embedding(label)
                else:
                    assert len(c.shape) == 3
            batch = c.shape[0]
            if len(c.shape) == 3:
                ws = (c.expand(-1, g.mapping_network.backbone.num_ws, -1).contiguous()
                      if c.shape[1] == 1 else c)
            else:
                ws = g.mapping_network(c, label)
        # ---- palette (generator.py:452-466)
        attention_values = None
        if g.attention_values > 0:
            assert ws.shape[1] == 15
            w_tex, w_synthesis = ws[:, 14], ws[:, :14]
            if 'attention_values' in model_inputs:
                attention_values = model_inputs['attention_values']
            else:  # 'sampler' is always among the requests of render()
                attention_values = g.texture_mapper(w_tex)
                if 'attention_values_bias' in model_inputs:
                    attention_values = attention_values + model_inputs['attention_values_bias']
        else:
            w_synthesis = ws
        # ---- planes (generator.py:471-477), channel-last
        noise_mode = 'const' if model_inputs.get('freeze_noise') else 'random'
        planes_cl = self.synthesis(w_synthesis, noise_mode=noise_mode)
        assert planes_cl.shape[0] == batch
        dec = g.decoder.net
        l1, l2 = dec[0], dec[2]
        out = {}
        if 'attention_values' in request_model_outputs:
            assert g.attention_values > 0
            out['attention_values'] = attention_values
        out['triplane'] = dict(
            planes=planes_cl, planes_layout='channel_last', palette=attention_values,
            w1=l1.weight * l1.weight_gain, b1=l1.bias * l1.bias_gain,
            w2=l2.weight * l2.weight_gain, b2=l2.bias * l2.bias_gain,
            beta=getattr(g, 'beta', None), alpha=getattr(g, 'alpha', None))
        return out
