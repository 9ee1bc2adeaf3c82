"""ConvNet1 -- the behaviour-cloning policy CNN, B200-native.

Same constructor, attributes and state_dict as the reference's
/root/reference/src/architectures/nets.py:6-39 (`ConvNet1(hparams)`, `cnn_base`, `fc`,
`example_input_array`, keys `cnn_base.{0,3,6,9}.*`, `fc.{0,2,4}.*`, OIHW / (out,in) f32),
but there are no nn.Conv2d / nn.Linear modules inside: all 14 tensors are views of ONE flat
f32 arena (also the Adam arena and the DDP bucket), and `forward` is a call into the
hand-written sm_100a kernels. No cuDNN/cuBLAS, no CPU fallback.
"""
from __future__ import annotations

import math

import torch
from torch import nn

try:  # the reference derives from pl.LightningModule (nets.py:6); keep that when Lightning exists
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # pragma: no cover - Lightning is not installed in the build image
    _Base = nn.Module

from carla_imitation_learning_b200 import _lib
from carla_imitation_learning_b200.autograd import FusedStepFunction, LossFunction, NetFunction
from carla_imitation_learning_b200.engine import BCEngine

# layer table in constructor order: (container, slot, weight shape builder)
_CONV = ((0, 16, 7), (3, 32, 5), (6, 64, 4), (9, 128, 3))      # slot in cnn_base, out channels, kernel
_FC = ((0, 128, 64), (2, 64, 32), (4, 32, None))              # slot in fc, in, out (None = n_actions)


class _Slot(nn.Module):
    """Holds one layer's weight/bias Parameters (arena views) under the reference's key names."""

    def __init__(self, kind: str):
        super().__init__()
        self.kind = kind

    def extra_repr(self) -> str:
        return f"{self.kind}, weight={tuple(self.weight.shape)} (arena view)"

    def forward(self, *_a, **_k):
        raise RuntimeError("ConvNet1 layers are not separately callable: the network runs as fused CUDA kernels")


class _Stack(nn.Module):
    """Index-addressable container so that state_dict keys read cnn_base.0.weight, fc.4.bias, ..."""

    def __getitem__(self, i: int) -> _Slot:
        return self._modules[str(i)]

    def forward(self, *_a, **_k):
        raise RuntimeError("call ConvNet1.forward; its stacks are parameter holders only")


class ConvNet1(_Base):
    def __init__(self, hparams):
        super().__init__()
        obs_size = int(hparams['obs_size'])
        n_actions = int(hparams['n_actions'])
        self.obs_size, self.n_actions = obs_size, n_actions
        # B200 addition (configs/model/imitation.yaml `precision`): 'fp32' = exact FFMA kernels,
        # 'bf16' = tcgen05 kernels. The reference has no such key; absent means fp32.
        try:
            self.precision = str(hparams['precision']) if 'precision' in hparams else 'fp32'
        except TypeError:
            self.precision = 'fp32'
        if self.precision not in ('fp32', 'bf16'):
            raise ValueError(f"precision must be 'fp32' or 'bf16', got {self.precision!r}")
        # B200 additions (configs/model/imitation.yaml), all absent from the reference and all optional:
        #   cuda_graph: true  -> training_step runs the fused forward+backward as ONE CUDA graph per loader slot
        #   fused_step: true  -> the same fused enqueue without graph capture
        #   overlap_backward  -> weight-gradient kernels of conv4..conv2 on a side stream (bc_backward_overlap)
        def _opt(key, default=False):
            try:
                return hparams[key] if key in hparams else default
            except TypeError:
                return default
        self.fast_step = "graph" if _opt('cuda_graph') else ("eager" if _opt('fused_step') else None)
        self.overlap_backward = bool(_opt('overlap_backward', False))

        # same RNG consumption order as the reference: example input first (nets.py:14) ...
        self.example_input_array = torch.randn((1, obs_size, 256, 256))

        # ... then every layer's default init, conv stack before fc (nets.py:17-33)
        inits = []
        cin = obs_size
        for _slot, cout, k in _CONV:
            inits.append(self._default_init((cout, cin, k, k)))
            cin = cout
        for _slot, fin, fout in _FC:
            inits.append(self._default_init((n_actions if fout is None else fout, fin)))

        total, offsets, sizes = self._layout(obs_size, n_actions)
        arena = torch.zeros(total, dtype=torch.float32)
        self.cnn_base, self.fc = _Stack(), _Stack()
        self._ordered_params = []
        for li, (w, b) in enumerate(inits):
            stack, slot, kind = (self.cnn_base, _CONV[li][0], "conv") if li < 4 else (self.fc, _FC[li - 4][0], "linear")
            holder = _Slot(kind)
            for j, (name, t) in enumerate((("weight", w), ("bias", b))):
                off = offsets[2 * li + j]
                arena[off:off + t.numel()] = t.reshape(-1)
                p = nn.Parameter(arena[off:off + t.numel()].view(t.shape))
                p._bc_offset, p._bc_owner = off, self
                holder.register_parameter(name, p)
                self._ordered_params.append(p)
            stack.add_module(str(slot), holder)
        self._engine = None
        self._rebind(arena)
        if torch.cuda.is_available():
            self.to(torch.device("cuda", torch.cuda.current_device()))

    # ------------------------------------------------------------------ construction helpers
    @staticmethod
    def _default_init(shape):
        """nn.Conv2d / nn.Linear reset_parameters: kaiming_uniform(a=sqrt 5) then U(+-1/sqrt(fan_in))."""
        w = torch.empty(shape)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        fan_in = w[0].numel()
        bound = 1 / math.sqrt(fan_in)
        b = torch.empty(shape[0])
        nn.init.uniform_(b, -bound, bound)
        return w, b

    @staticmethod
    def _layout(obs_size, n_actions):
        """Arena offsets from the C ABI (bc_arena_layout) -- one source of truth for host and device."""
        return _lib.arena_layout(obs_size, n_actions)

    def _rebind(self, arena: torch.Tensor) -> None:
        """Point every Parameter at its slice of `arena` (after construction or a device move)."""
        self._arena = arena
        for p in self._ordered_params:
            p.data = arena[p._bc_offset:p._bc_offset + p.numel()].view(p.shape)
            p._bc_arena = arena
            p.grad = None
        self._engine = None

    def _apply(self, fn, recurse=True):
        new = fn(self._arena)
        if new.dtype != torch.float32:
            raise TypeError("ConvNet1 keeps f32 master weights; bf16 is a compute mode of the kernels, not a storage dtype")
        if new is not self._arena:
            self._rebind(new.contiguous())
        return self

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Copy INTO the arena views (reference checkpoints load both ways, train.py:198-201)."""
        own = dict(self.named_parameters())
        missing = [k for k in own if k not in state_dict]
        unexpected = [k for k in state_dict if k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError(f"state_dict mismatch: missing {missing}, unexpected {unexpected}")
        with torch.no_grad():
            for k, p in own.items():
                if k in state_dict:
                    p.copy_(state_dict[k])
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    @torch.no_grad()
    def fold_batchnorm(self, layer: int, gamma, beta, running_mean, running_var, eps: float = 1e-5) -> None:
        """EXTENSION (the reference has no BatchNorm layer): fold an inference-time BatchNorm2d that follows conv `layer`
        (0..3) into that conv -- W' = W * s, b' = (b - mean) * s + beta, s = gamma / sqrt(var + eps) -- so that the
        conv+BN+ReLU block the north star names runs as the same fused conv+bias+ReLU+pool kernel (the BN scale/shift IS the
        kernel's weight/bias epilogue). Specification: oracle/ext_oracle.py::fold_batchnorm."""
        slot = self.cnn_base[_CONV[layer][0]]
        dev = slot.weight.device
        s = gamma.to(dev, torch.float32) / torch.sqrt(running_var.to(dev, torch.float32) + eps)
        slot.bias.copy_((slot.bias - running_mean.to(dev, torch.float32)) * s + beta.to(dev, torch.float32))
        slot.weight.mul_(s.view(-1, 1, 1, 1))

    # ------------------------------------------------------------------ compute
    def engine(self) -> BCEngine:
        if self._engine is None:
            if not self._arena.is_cuda:
                raise RuntimeError(
                    "ConvNet1 runs only on a CUDA sm_100 (B200) device: move it with .to('cuda'). "
                    "There is deliberately no CPU / PyTorch fallback for the hot path.")
            self._engine = BCEngine(self._arena, self.obs_size, self.n_actions)
            if self.precision == 'bf16':         # obs_size 4, and 12 = three 4-frame camera streams (BASELINE configs[3])
                self._engine.set_mode('bf16')
            self._engine.overlap = self.overlap_backward
            params, arena = self._ordered_params, self._arena
            # a Parameter re-pointed with `p.data = view` keeps its OWN version counter: p.copy_() (load_state_dict) bumps
            # that one, not the arena's -- the operand images follow both
            self._engine.weights_version = lambda: (arena._version, sum(p._version for p in params))
        self._engine.ensure_packed()        # bf16 operand images: re-derived only when the master weights changed outside FusedAdam
        return self._engine

    def _to_device(self, t: torch.Tensor) -> torch.Tensor:
        return t if t.device == self._arena.device else t.to(self._arena.device, non_blocking=True)

    def forward(self, x):
        """(B, obs_size, 256, 256) -> (B, n_actions) logits (nets.py:35-39)."""
        return NetFunction.apply(self._to_device(x), self, *self._ordered_params)

    def loss(self, x, y):
        """Fused forward + CrossEntropyLoss() (mean) -- what Imitation.training_step needs."""
        x, y = self._to_device(x), self._to_device(y)
        if self.fast_step and torch.is_grad_enabled():
            return FusedStepFunction.apply(x, y, self, *self._ordered_params)
        return LossFunction.apply(x, y, self, *self._ordered_params)

    @torch.no_grad()
    def act(self, x):
        """Greedy action ids = argmax over logits (src/data/stat.py:41, imitation.py:177)."""
        return self.engine().forward_act(self._to_device(x))[0]


class ConvNetRawSegment(_Base):
    """The two-stream variant of /root/reference/src/architectures/nets.py:42-78. In the reference its constructor starts with
    `super(ConvNet1, self).__init__()` (nets.py:44) -- `self` is not a ConvNet1, so EVERY construction raises TypeError before
    a single layer exists; no caller, config or checkpoint in the reference uses the class. The drop-in keeps that behaviour
    (same exception type at the same point) instead of inventing semantics the reference never had; the kernels of the hot
    path are built for ConvNet1's geometry (out of scope)."""

    def __init__(self, hparams):
        raise TypeError("super(type, obj): obj must be an instance or subtype of type "
                        "[/root/reference/src/architectures/nets.py:44 calls super(ConvNet1, self).__init__() inside ConvNetRawSegment: "
                        "the class cannot be constructed in the reference either]")
