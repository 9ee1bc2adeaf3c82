"""CPU oracle for the north-star EXTENSIONS that have NO counterpart in the reference.  TEST INFRASTRUCTURE ONLY.

    EXTENSION -- NO REFERENCE PARITY.

BASELINE.json's north_star asks for command-conditioned per-branch heads with a branch-select mask, an L1/MSE
steer/throttle/brake loss, normalise / crop / colour-jitter inside the staging kernel and a folded-BatchNorm conv epilogue.
The reference has none of them: its ConvNet1 has one un-branched head (/root/reference/src/architectures/nets.py:31-33),
its loss is CrossEntropyLoss (/root/reference/src/models/imitation.py:43-44), its only transform pipeline is
ToTensor/Normalize for MNIST (/root/reference/src/transforms/mnist_transforms.py:7-13) and no BatchNorm layer exists
anywhere in the tree. So there is nothing to pin these functions against; they are written in the reference's own style
(plain torch / numpy) as the specification the CUDA kernels are tested against, and the tests say "extension".

Only tests/ may import this module (same rule as oracle/bc_oracle.py).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import bc_oracle as O

LOSS_KINDS = ("ce", "l1", "mse")


# --------------------------------------------------------------------------- branched heads
def branch_param_names(n_branches: int):
    return tuple(f"branches.{g}.{i}.{s}" for g in range(n_branches) for i in (0, 2, 4) for s in ("weight", "bias"))


def init_branched_params(seed: int, obs_size: int, n_out: int, n_branches: int) -> "OrderedDict[str, torch.Tensor]":
    """ConvNet1's own initialisation first (bc_oracle.init_params: example input, conv stack, the reference's single head --
    which the branched net keeps in its arena but never uses), then G heads 128->64->32->n_out with torch's default Linear
    init, branch after branch, from the same RNG stream."""
    base = O.init_params(seed, obs_size, n_out)
    out = OrderedDict((k, v) for k, v in base.items() if k.startswith("cnn_base."))
    for g in range(n_branches):
        for i, (fi, fo) in zip((0, 2, 4), ((128, 64), (64, 32), (32, n_out))):
            m = torch.nn.Linear(fi, fo)
            out[f"branches.{g}.{i}.weight"], out[f"branches.{g}.{i}.bias"] = m.weight.detach().clone(), m.bias.detach().clone()
    return out


def trunk_features(params, x):
    h = x
    for name, k, s, p in O.CONV_SPECS:
        h = F.max_pool2d(F.relu(F.conv2d(h, params[f"{name}.weight"], params[f"{name}.bias"], stride=s)), kernel_size=p)
    return torch.flatten(h, 1)


def branched_forward(params, x, command, n_branches: int):
    """out[b] = head[command[b]](trunk(x[b])): every sample goes through the branch its high-level command selects."""
    feat = trunk_features(params, x)
    outs = []
    for g in range(n_branches):
        h = F.relu(F.linear(feat, params[f"branches.{g}.0.weight"], params[f"branches.{g}.0.bias"]))
        h = F.relu(F.linear(h, params[f"branches.{g}.2.weight"], params[f"branches.{g}.2.bias"]))
        outs.append(F.linear(h, params[f"branches.{g}.4.weight"], params[f"branches.{g}.4.bias"]))
    allb = torch.stack(outs, 1)                                          # (B, G, n_out)
    return allb.gather(1, command.view(-1, 1, 1).expand(-1, 1, allb.shape[-1])).squeeze(1)


def branched_loss(out, target, kind: str):
    """'ce': CrossEntropyLoss() on class ids; 'l1' / 'mse': nn.L1Loss() / nn.MSELoss() (mean over B x n_out) on the
    (steer, throttle, brake) regression targets."""
    if kind == "ce":
        return F.cross_entropy(out, target)
    return F.l1_loss(out, target) if kind == "l1" else F.mse_loss(out, target)


def branched_loss_and_grads(params, x, command, target, n_branches: int, kind: str, dtype=torch.float64):
    leaf = OrderedDict((k, v.detach().to(dtype).clone().requires_grad_(True)) for k, v in params.items())
    out = branched_forward(leaf, x.to(dtype), command, n_branches)
    loss = branched_loss(out, target if kind == "ce" else target.to(dtype), kind)
    grads = torch.autograd.grad(loss, list(leaf.values()), allow_unused=True)
    grads = [torch.zeros_like(p) if g is None else g for g, p in zip(grads, leaf.values())]
    return loss.detach(), out.detach(), OrderedDict(zip(leaf.keys(), grads))


# --------------------------------------------------------------------------- staging augmentation
AUG_FIELDS = ("crop_y", "crop_x", "brightness", "contrast", "saturation", "mean", "inv_std")


def augment_params(seed: int, n_frames: int, src_hw, out_hw=(256, 256), brightness=0.2, contrast=0.2, saturation=0.2,
                   mean: float = 0.0, std: float = 1.0) -> np.ndarray:
    """Per-frame augmentation table, reproducible from `seed` on the host (numpy PCG64): (n_frames, 8) f32 rows
    [crop_y, crop_x, brightness, contrast, saturation, mean, 1/std, 0]. Crop offsets are uniform over the valid range,
    the jitter factors uniform in [1 - a, 1 + a] (torchvision ColorJitter's parameterisation)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.zeros((n_frames, 8), np.float32)
    t[:, 0] = rng.integers(0, src_hw[0] - out_hw[0] + 1, size=n_frames)
    t[:, 1] = rng.integers(0, src_hw[1] - out_hw[1] + 1, size=n_frames)
    t[:, 2] = rng.uniform(1 - brightness, 1 + brightness, size=n_frames)
    t[:, 3] = rng.uniform(1 - contrast, 1 + contrast, size=n_frames)
    t[:, 4] = rng.uniform(1 - saturation, 1 + saturation, size=n_frames)
    t[:, 5] = mean
    t[:, 6] = 1.0 / std
    return t


def stage_augmented(frames_u8: np.ndarray, table: np.ndarray, out_hw=(256, 256)) -> np.ndarray:
    """(n,Hs,Ws,3) u8 + per-frame table -> (n,256,256) f32 gray planes. Per pixel, all in f32, in this order:
        crop;  c = rgb * brightness, clamped to [0,255];
        c = (c - 127.5) * contrast + 127.5, clamped   (contrast about mid-grey: no per-frame mean pass);
        y = 0.299 r + 0.587 g + 0.114 b;  c = y + (c - y) * saturation, clamped;
        gray = (0.299 r + 0.587 g + 0.114 b) / 255;  out = (gray - mean) * inv_std.
    With brightness = contrast = saturation = 1, mean = 0, inv_std = 1 and no crop this is the reference's gray conversion
    to f32 rounding (the un-augmented staging kernel stays the bit-exact one)."""
    n = frames_u8.shape[0]
    H, W = out_hw
    out = np.empty((n, H, W), np.float32)
    f32 = np.float32
    for i in range(n):
        cy, cx = int(table[i, 0]), int(table[i, 1])
        br, ct, sa, mean, istd = (f32(v) for v in table[i, 2:7])
        c = frames_u8[i, cy:cy + H, cx:cx + W].astype(np.float32)
        c = np.clip(c * br, f32(0), f32(255))
        c = np.clip((c - f32(127.5)) * ct + f32(127.5), f32(0), f32(255))
        y = (f32(0.299) * c[..., 0] + f32(0.587) * c[..., 1]) + f32(0.114) * c[..., 2]
        c = np.clip(y[..., None] + (c - y[..., None]) * sa, f32(0), f32(255))
        g = ((f32(0.299) * c[..., 0] + f32(0.587) * c[..., 1]) + f32(0.114) * c[..., 2]) * f32(1.0 / 255.0)
        out[i] = (g - mean) * istd
    return out


# --------------------------------------------------------------------------- folded BatchNorm
def fold_batchnorm(weight, bias, gamma, beta, running_mean, running_var, eps: float = 1e-5):
    """conv -> BatchNorm2d(eval) == conv with W' = W * s[:,None,None,None], b' = (b - mean) * s + beta, s = gamma / sqrt(var + eps):
    the BN of an inference-time conv+BN+ReLU block disappears into the weights the conv kernels already read."""
    s = gamma / torch.sqrt(running_var + eps)
    return weight * s.view(-1, 1, 1, 1), (bias - running_mean) * s + beta
