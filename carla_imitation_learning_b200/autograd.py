"""autograd glue: the two differentiable entry points the LightningModule contract needs.

`net(x)` -> logits must be differentiable w.r.t. every parameter (any loss may follow), and
`training_step` must return a 0-dim tensor whose `.backward()` fills `.grad` of all 14
parameters (/root/reference/src/models/imitation.py:38-45 relies on torch autograd for that).
Both Functions run the CUDA kernels in forward and backward; nothing here is numeric.
"""
from __future__ import annotations

import torch


def _param_grads(net, flat: torch.Tensor):
    """Per-parameter views of one arena-shaped gradient buffer (fresh, so autograd may adopt them)."""
    out = []
    for p in net._ordered_params:
        g = flat[p._bc_offset:p._bc_offset + p.numel()].view(p.shape)
        g._bc_flat = flat
        out.append(g)
    return out


class NetFunction(torch.autograd.Function):
    """logits = ConvNet1(x); backward takes d/dlogits."""

    @staticmethod
    def forward(ctx, x, net, *params):
        eng = net.engine()
        bufs = eng.forward(x, None, backward=False)
        ctx.net, ctx.bufs = net, bufs
        return bufs.logits

    @staticmethod
    def backward(ctx, dlogits):
        net, bufs = ctx.net, ctx.bufs
        eng = net.engine()
        bufs.dlogits.copy_(dlogits)
        flat = eng.backward(bufs).clone()
        return (None, None, *_param_grads(net, flat))


class LossFunction(torch.autograd.Function):
    """loss = CrossEntropy(ConvNet1(x), y) with mean reduction, fused (one head launch computes
    logits, loss and d loss/d logits)."""

    @staticmethod
    def forward(ctx, x, y, net, *params):
        eng = net.engine()
        bufs = eng.forward(x, y, backward=False)
        ctx.net, ctx.bufs = net, bufs
        return bufs.loss

    @staticmethod
    def backward(ctx, gloss):
        net, bufs = ctx.net, ctx.bufs
        eng = net.engine()
        saved = bufs.dlogits
        bufs.dlogits = saved * gloss      # device-side scale: no host sync; keeps `saved` for a second backward
        flat = eng.backward(bufs).clone()
        bufs.dlogits = saved
        return (None, None, None, *_param_grads(net, flat))
