"""autograd glue: the differentiable entry points the LightningModule contract needs.

`net(x)` -> logits must be differentiable w.r.t. every parameter (any loss may follow), and
`training_step` must return a 0-dim tensor whose `.backward()` fills `.grad` of all 14
parameters (/root/reference/src/models/imitation.py:38-45 relies on torch autograd for that).
All Functions run the CUDA kernels in forward and backward; nothing here is numeric.

Three entry points:
  NetFunction        logits = net(x); backward from d/dlogits (any loss may follow).
  LossFunction       loss = CE(net(x), y) fused into the head launch; backward runs the CUDA backward.
  FusedStepFunction  the static-shape fast path (configs/model `cuda_graph: true`): forward, CE and the WHOLE backward are
                     enqueued (or replayed as one CUDA graph) when the loss is computed; `.backward()` only hands the
                     already computed gradients to autograd. No per-step allocation, no weight re-pack, no gradient copy.
"""
from __future__ import annotations

import torch

from . import _lib
from .engine import on_device, StagedBatch, _stream_ptr


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
        eng.backward(bufs)
        flat = eng._last_flat = eng.grad_view().clone()
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
        eng.backward(bufs)
        flat = eng._last_flat = eng.grad_view().clone()
        bufs.dlogits = saved
        return (None, None, None, *_param_grads(net, flat))


class _StepGraphs:
    """CUDA graphs of the fused forward+backward, keyed by (input pointer, label pointer, batch, gradient slot): a loader
    that rotates over a few staging slots (data.SequentialFrames) gets one graph per slot. A key is captured the second
    time it is seen (the first pass runs eagerly and doubles as the warm-up the capture needs)."""

    def __init__(self):
        self.seen, self.graphs = set(), {}

    def run(self, key, enqueue) -> None:
        g = self.graphs.get(key)
        if g is not None:
            g.replay()
            return
        if key in self.seen and not torch.cuda.is_current_stream_capturing():
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                enqueue()
            self.graphs[key] = g
            g.replay()
            return
        self.seen.add(key)
        enqueue()


class FusedStepFunction(torch.autograd.Function):
    """loss = CE(ConvNet1(x), y) with the gradients of all 14 parameters computed in the same enqueue.

    backward() does what autograd's AccumulateGrad would do with the 14 gradients -- adopt them as .grad when there is none,
    add them otherwise -- but with per-slot CACHED arena views, so a step costs 14 attribute stores instead of 14 tensor
    constructions, and FusedAdam recognises the arena by identity (no gather, no copy)."""

    @staticmethod
    def forward(ctx, x, y, net, *params):
        eng = net.engine()
        x = eng.check_input(x)
        B = x.shape[0]
        flat, views = eng.next_grad_slot(net._ordered_params)   # a gradient arena no live .grad aliases
        bufs = eng.static_buffers(B, x, y)
        if net.fast_step == "graph":
            xp = x.tp.data_ptr() if isinstance(x, StagedBatch) else x.data_ptr()
            graphs = net.__dict__.get("_step_graphs")
            if graphs is None:
                graphs = net.__dict__["_step_graphs"] = _StepGraphs()
            graphs.run((xp, y.data_ptr(), B, flat.data_ptr()), lambda: eng.enqueue_train(bufs))
        else:
            eng.enqueue_train(bufs)
        ctx.net, ctx.flat, ctx.views = net, flat, views
        return bufs.loss.clone()         # the static loss cell is rewritten by the next step; epoch-end hooks keep these

    @staticmethod
    def backward(ctx, gloss):
        net, flat, views = ctx.net, ctx.flat, ctx.views
        eng = net.engine()
        g = gloss.reshape(1)
        if g.dtype != torch.float32:
            g = g.float()
        with on_device(eng.device):
            _lib.check(eng.lib.bc_scale_inplace(flat.data_ptr(), flat.numel(), g.data_ptr(), _stream_ptr()), "bc_scale_inplace")
        for p, v in zip(net._ordered_params, views):
            if p.grad is None:
                p.grad = v
            else:
                p.grad.add_(v)
        return (None,) * (3 + len(views))
