"""Fused multi-tensor Adam over the flat parameter arena (one kernel per step).

Drop-in for the `Adam(self.parameters(), lr=1e-3)` the reference builds in
Imitation.configure_optimizers (/root/reference/src/models/imitation.py:82-87): a
torch.optim.Optimizer subclass, `step(closure)` invokes the closure, `param_groups[0]['lr']`
stays live for MultiStepLR, state_dict()/load_state_dict() use torch.optim.Adam's keys.
Bias corrections are derived on the device (bc_adam_tick) so a captured CUDA graph replays
correctly step after step; the learning rate lives in a device scalar for the same reason.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .engine import on_device, _stream_ptr

# state vector layout (csrc/abi.cu), f64: lr beta1 beta2 eps step grad_scale step_size bc2_sqrt
_LR, _B1, _B2, _EPS, _STEP, _GS, _SS, _BC2 = range(8)


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, amsgrad: bool = False):
        if weight_decay != 0.0 or amsgrad:
            raise ValueError("the reference uses Adam defaults (imitation.py:83): weight_decay=0, amsgrad=False")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0.0, amsgrad=False))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam drives one flat arena: a single param group")
        self._params = list(self.param_groups[0]["params"])
        arena = getattr(self._params[0], "_bc_arena", None)
        if arena is None or any(getattr(p, "_bc_arena", None) is not arena for p in self._params):
            raise ValueError("FusedAdam needs parameters that are views of one ConvNet1 arena")
        self._arena_of = self._params[0]._bc_owner  # the ConvNet1 shell; its arena may move (.to())
        self._bound = None
        self._dev_lr: Optional[float] = None
        self.grad_scale = 1.0
        self.exchange = None          # parallel.ModuleExchange when the module path trains data-parallel
        self.stats = {"zero_copy": 0, "gathered": 0}   # how step() found the gradients: as one arena (no copy) / gathered per tensor

    # ------------------------------------------------------------------ state
    def _bind(self):
        net = self._arena_of
        arena = net._arena
        if self._bound is not None and self._bound[0] is arena:
            return self._bound
        if not arena.is_cuda:
            raise RuntimeError("FusedAdam.step needs the parameters on a CUDA (sm_100) device; there is no CPU fallback")
        if torch.cuda.is_current_stream_capturing():
            # the zero-fills below would be baked into the graph and reset the moments, the step count and the LR at
            # every replay
            raise RuntimeError("FusedAdam state must exist before CUDA-graph capture: call optimizer.prepare() (or run one eager step) first")
        old = self._bound
        m = torch.zeros_like(arena)
        v = torch.zeros_like(arena)
        st = torch.zeros(9, dtype=torch.float64, device=arena.device)     # 8 scalars (csrc/abi.cu) + the CTA counter of the fused tick
        if old is not None:  # the arena moved: carry the moments over
            m.copy_(old[1]); v.copy_(old[2]); st.copy_(old[3])
        g = self.param_groups[0]
        st[_B1], st[_B2], st[_EPS], st[_GS] = g["betas"][0], g["betas"][1], g["eps"], self.grad_scale
        self._dev_lr = None
        flat_g = torch.zeros_like(arena)
        self._bound = (arena, m, v, st, flat_g)
        for p in self._params:   # torch.optim.Adam-compatible per-parameter views
            off, n = p._bc_offset, p.numel()
            self.state[p] = dict(step=st[_STEP:_STEP + 1].view(()), exp_avg=m[off:off + n].view(p.shape),
                                 exp_avg_sq=v[off:off + n].view(p.shape))
        return self._bound

    def _sync_scalars(self, st) -> None:
        g = self.param_groups[0]
        if self._dev_lr != g["lr"]:
            st[_LR:_LR + 1].fill_(float(g["lr"]))   # async fill, no host sync; graph-safe (value lives on device)
            self._dev_lr = g["lr"]

    def prepare(self) -> None:
        """Allocate the state (first call) and push a changed learning rate to its device scalar. Call before capturing a
        CUDA graph that contains the step and before every replay: MultiStepLR milestones (imitation.py:84-86) change
        param_groups[0]['lr'] on the host, the captured kernels read the device scalar."""
        _arena, _m, _v, st, _ = self._bind()
        self._sync_scalars(st)

    def _packed(self):
        """(w_packed pointer or None, obs_size, n_actions): bf16 mode lets the Adam kernel refresh the MMA operand images."""
        net = self._arena_of
        eng = getattr(net, "_engine", None)
        return (eng.packed_ptr() if eng is not None else None), int(net.obs_size), int(net.n_actions)

    def set_grad_scale(self, scale: float) -> None:
        """Fold the data-parallel 1/world_size mean into the update's gradient read."""
        self.grad_scale = float(scale)
        if self._bound is not None:
            self._bound[3][_GS:_GS + 1].fill_(self.grad_scale)

    # ------------------------------------------------------------------ step
    def _flat_grads(self, arena, flat_g) -> torch.Tensor:
        """The gradients as one arena-shaped buffer; zero-copy when the .grad tensors are (still) the views the CUDA
        backward handed to autograd, i.e. they sit at their arena offsets inside one gradient arena of the engine."""
        eng = getattr(self._arena_of, "_engine", None)
        if eng is not None:
            owner = eng.grads_as_arena(self._params)           # the fused step's cached views, recognised by identity
            if owner is not None:
                self.stats["zero_copy"] += 1
                return owner
        first = self._params[0].grad
        if first is not None:
            base = first.data_ptr() - self._params[0]._bc_offset * 4
            if all(p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous()
                   and p.grad.data_ptr() == base + p._bc_offset * 4 for p in self._params):
                eng = getattr(self._arena_of, "_engine", None)
                owner = eng.grad_arena_at(base) if eng is not None else None
                if owner is not None:
                    self.stats["zero_copy"] += 1
                    return owner
        self.stats["gathered"] += 1
        for p in self._params:
            off, n = p._bc_offset, p.numel()
            if p.grad is None:
                flat_g[off:off + n].zero_()
            else:
                flat_g[off:off + n].copy_(p.grad.reshape(-1))
        return flat_g

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        arena, m, v, st, flat_g = self._bind()
        self._sync_scalars(st)
        g = self._flat_grads(arena, flat_g)
        if self.exchange is not None:          # data parallel through the module path: this rank's gradients -> the peer-visible arena
            self.exchange.step_from(g)
        else:
            self.step_flat(g)
        return loss

    def step_flat(self, flat_grads: torch.Tensor) -> None:
        """Tick and fused update in one launch, gradients given as an arena-shaped tensor (the engine's grads)."""
        arena, m, v, st, _ = self._bind()
        if not torch.cuda.is_current_stream_capturing():
            self._sync_scalars(st)
        wp, obs, na = self._packed()
        lib, s = _lib.lib(), _stream_ptr()
        with on_device(arena.device):
            _lib.check(lib.bc_adam_tick_step(arena.data_ptr(), flat_grads.data_ptr(), m.data_ptr(), v.data_ptr(),
                                             st.data_ptr(), arena.numel(), wp, obs, na, s), "bc_adam_tick_step")

    def step_exchange(self, peer, lo: int = 0, hi: Optional[int] = None, bucket: int = 1, publish: bool = True, stream=None) -> None:
        """Data-parallel step of the arena floats [lo, hi): ONE kernel that sums every rank's gradients straight from NVLink
        peer memory (rank order), folds the 1/world mean and applies Adam (bc_adam_step_exchange; `peer` = parallel.PeerGrads).
        The launch with publish=True also ticks the step counter and completes the exchange epoch."""
        arena, m, v, st, _ = self._bind()
        if not torch.cuda.is_current_stream_capturing():
            self._sync_scalars(st)
        wp, obs, na = self._packed()
        lib = _lib.lib()
        s = _stream_ptr() if stream is None else stream
        with on_device(arena.device):
            _lib.check(lib.bc_adam_step_exchange(arena.data_ptr(), m.data_ptr(), v.data_ptr(), st.data_ptr(), arena.numel(),
                                                 peer.c_struct, lo, arena.numel() if hi is None else hi, bucket, int(publish),
                                                 wp, obs, na, s), "bc_adam_step_exchange")

    def zero_grad(self, set_to_none: bool = True) -> None:
        if not set_to_none:
            return super().zero_grad(set_to_none=False)
        for p in self._params:
            p.grad = None

    # ------------------------------------------------------------------ checkpoints
    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.Adam state_dict (reference checkpoints, train.py:198-201)."""
        arena, m, v, st, _ = self._bind()
        packed = state_dict["state"]
        ids = state_dict["param_groups"][0]["params"]
        for pid, p in zip(ids, self._params):
            if pid in packed:
                s = packed[pid]
                self.state[p]["exp_avg"].copy_(s["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(s["exp_avg_sq"])
                st[_STEP] = float(s["step"])
        g = state_dict["param_groups"][0]
        self.param_groups[0]["lr"] = g["lr"]
        self._dev_lr = None
