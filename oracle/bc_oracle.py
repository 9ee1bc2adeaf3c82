"""CPU oracle for the behaviour-cloning hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (carla_imitation_learning_b200, src/)
never does and raises when its CUDA library is missing.

It restates, with plain torch-CPU / numpy arithmetic, what the reference computes
on this path. Each function cites the reference lines it follows
(paths relative to /root/reference):

  gray_stack / sequential_samples  src/dataset/imitation_dataset.py:115-133
  discretise_actions               src/dataset/imitation_dataset.py:317-339
  init_params                      src/architectures/nets.py:7-33 (torch default init)
  forward                          src/architectures/nets.py:17-39
  cross_entropy / loss_and_grads   src/models/imitation.py:38-45
  explicit_backward                the autograd graph of the above, written out
  adam_update                      src/models/imitation.py:82-87 (torch.optim.Adam defaults)
  lr_at_epoch                      src/models/imitation.py:84-86 (MultiStepLR [20,30] x0.1)

Parity pin: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF:
oracle/make_golden.py imports the unmodified reference classes in the build container
(via the two import stubs in oracle/_stubs) and writes tests/golden/*.npz;
tests/test_oracle_golden.py checks every function here against those files, and
tests/test_reference_live.py re-checks against the live reference whenever
/root/reference is present.

All arithmetic lives in third-party code absent from /root/reference (torch ATen /
oneDNN); the reference pins no versions. This oracle runs on the image's
torch 2.11.0 CPU kernels, i.e. the same library the reference would dispatch to here.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# (state_dict prefix, kernel, stride, pool) -- nets.py:18-29
CONV_SPECS = (
    ("cnn_base.0", 7, 3, 3),
    ("cnn_base.3", 5, 1, 2),
    ("cnn_base.6", 4, 1, 2),
    ("cnn_base.9", 3, 1, 2),
)
CONV_CHANNELS = (16, 32, 64, 128)
FC_SPECS = (("fc.0", 128, 64), ("fc.2", 64, 32), ("fc.4", 32, None))  # nets.py:31-33
GRAY_WEIGHTS = (0.299, 0.587, 0.114)  # imitation_dataset.py:121

PARAM_ORDER = tuple(
    f"{p}.{s}" for p in [c[0] for c in CONV_SPECS] + [f[0] for f in FC_SPECS] for s in ("weight", "bias")
)


# --------------------------------------------------------------------------- data
def gray_stack(frames_u8: np.ndarray) -> np.ndarray:
    """(n,H,W,3) u8 -> (n,H,W) f32: f64 dot with the luma weights, /255.0, cast.

    imitation_dataset.py:120-121 (np.dot in f64, division by 255.0) and :130 (cast to float32).
    """
    g = np.dot(frames_u8[..., :], list(GRAY_WEIGHTS)) / 255.0
    return g.astype(np.float32)


def sequential_samples(frames_u8: np.ndarray, labels: np.ndarray, frame_skip: int = 4):
    """Sliding-window samples of a frame sequence, as SequentialTorchDataset yields them.

    Sample i stacks frames [i, i+frame_skip) and takes the label of frame i+frame_skip
    (imitation_dataset.py:117 with index = i + 4 from :125, label at :131).
    Returns x (N-frame_skip, frame_skip, H, W) f32 and y (N-frame_skip,) int64.
    """
    g = gray_stack(frames_u8)
    n = g.shape[0] - frame_skip
    x = np.stack([g[i:i + frame_skip] for i in range(n)], axis=0)
    y = np.asarray(labels[frame_skip:frame_skip + n]).astype(np.int64)
    return x, y


def discretise_actions(steer, throttle, brake, steer_threshold: float = 0.05) -> np.ndarray:
    """Continuous (steer, throttle, brake) -> class id acc*3+steer in 0..8.

    imitation_dataset.py:317-339. steer: >thr -> 2, <-thr -> 0, else 1.
    acc starts as a copy of brake and is overwritten for the three listed
    (brake, throttle) combinations; everything else keeps the brake value.
    """
    steer = np.asarray(steer, dtype=np.float64)
    throttle = np.asarray(throttle, dtype=np.float64)
    brake = np.asarray(brake, dtype=np.float64)
    s = np.ones_like(steer)
    s[steer > steer_threshold] = 2.0
    s[steer < -steer_threshold] = 0.0
    # NB the reference tests `steer == 0.0 / 2.0` after overwriting, so a raw steer value of
    # exactly 0.0 or 2.0 inside the dead band keeps its raw value (0.0 or 2.0).
    inside = ~((steer > steer_threshold) | (steer < -steer_threshold))
    s[inside & (steer == 0.0)] = 0.0
    s[inside & (steer == 2.0)] = 2.0
    acc = brake.copy()
    acc[(brake == 0.0) & (throttle == 1.0)] = 2.0
    acc[(brake == 0.0) & (throttle == 0.5)] = 1.0
    acc[(brake == 1.0) & (throttle == 0.0)] = 0.0
    return acc * 3 + s


# --------------------------------------------------------------------------- parameters
def param_shapes(obs_size: int = 4, n_actions: int = 9) -> "OrderedDict[str, tuple]":
    shapes = OrderedDict()
    cin = obs_size
    for (name, k, _s, _p), cout in zip(CONV_SPECS, CONV_CHANNELS):
        shapes[f"{name}.weight"] = (cout, cin, k, k)
        shapes[f"{name}.bias"] = (cout,)
        cin = cout
    for name, fin, fout in FC_SPECS:
        fout = n_actions if fout is None else fout
        shapes[f"{name}.weight"] = (fout, fin)
        shapes[f"{name}.bias"] = (fout,)
    return shapes


def init_params(seed: int = 12345, obs_size: int = 4, n_actions: int = 9) -> "OrderedDict[str, torch.Tensor]":
    """Default torch init in the order the reference constructor consumes the RNG.

    train.py:103 seeds, nets.py:14 draws example_input_array FIRST, then nets.py:17-33
    builds conv/linear layers in order; each layer draws weight = kaiming_uniform(a=sqrt5)
    then bias = U(+-1/sqrt(fan_in)).
    """
    torch.manual_seed(seed)
    torch.randn((1, obs_size, 256, 256))  # example_input_array consumes the generator first
    out = OrderedDict()
    for name, shape in param_shapes(obs_size, n_actions).items():
        t = torch.empty(shape, dtype=torch.float32)
        if name.endswith("weight"):
            torch.nn.init.kaiming_uniform_(t, a=math.sqrt(5))
            fan_in = int(np.prod(shape[1:]))
        else:
            bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
            torch.nn.init.uniform_(t, -bound, bound)
        out[name] = t
    return out


# --------------------------------------------------------------------------- forward / loss
def forward(params, x: torch.Tensor, keep: bool = False):
    """nets.py:35-39: 4 x (conv, ReLU, floor-mode max-pool), flatten, 3-layer MLP."""
    acts = []
    h = x
    for name, k, s, p in CONV_SPECS:
        h = F.conv2d(h, params[f"{name}.weight"], params[f"{name}.bias"], stride=s)
        h = F.max_pool2d(F.relu(h), kernel_size=p)
        acts.append(h)
    h = torch.flatten(h, start_dim=1)
    for i, (name, _fi, _fo) in enumerate(FC_SPECS):
        h = F.linear(h, params[f"{name}.weight"], params[f"{name}.bias"])
        if i < len(FC_SPECS) - 1:
            h = F.relu(h)
        acts.append(h)
    return (h, acts) if keep else h


def cross_entropy(logits: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """imitation.py:43-44: nn.CrossEntropyLoss() = mean_b(logsumexp(z_b) - z_b[y_b])."""
    lse = torch.logsumexp(logits, dim=1)
    return (lse - logits.gather(1, y[:, None]).squeeze(1)).mean()


def aux_criterion(params, x, y2, dtype=torch.float32):
    """ImitationAux's lossCriterion (/root/reference/src/models/imitation.py:11-24 with 109-117): of the three terms only
    l3 = cross_entropy(output[2], y[:, 1]) is live -- the action logits against column 1 of the (B, 2) labels."""
    return loss_and_grads(params, x, y2[:, 1], dtype=dtype)


def loss_and_grads(params, x, y, dtype=torch.float32):
    """Loss and d loss / d param for every tensor, by autograd on the restated forward."""
    leaf = OrderedDict((k, v.detach().to(dtype).clone().requires_grad_(True)) for k, v in params.items())
    logits = forward(leaf, x.to(dtype))
    loss = cross_entropy(logits, y)
    grads = torch.autograd.grad(loss, list(leaf.values()))
    return loss.detach(), logits.detach(), OrderedDict(zip(leaf.keys(), grads))


def explicit_backward(params, x, y, dtype=torch.float64, argmax_override=None, relu_override=None):
    """The same gradients written out operation by operation (no autograd).

    `argmax_override` (list of 4 int64 tensors (B,C,Hp,Wp), window-local row-major index)
    replaces the pool routing decision. Max-pool routing is discontinuous: when two window
    entries differ by less than the rounding error of the conv that produced them, f32 CPU,
    f64 and GPU arithmetic may legitimately pick different winners, and one flipped winner
    moves conv1/conv2 weight gradients by ~3e-3 relative (measured: f32 vs f64 oracle on
    synth seed 0). Parity tests therefore check (i) every routing decision of the device is
    a maximum of its window to within rounding, and (ii) gradients GIVEN that routing.

    `relu_override` = {"pooled": 4 bool tensors (B,C,Hp,Wp), "fc": 2 bool tensors (B,64), (B,32)} replaces the ReLU
    derivative by the device's own decisions (pooled activation > 0 / hidden activation > 0). ReLU'(z) is the same kind of
    discontinuity as the pool routing: in bf16 mode a pre-activation within the operand rounding of zero may fall on the
    other side, and one flipped unit changes a weight gradient by that unit's whole contribution (measured on the B200:
    3-17 % of max|grad| with the oracle's own masks, see tests/test_gpu_step.py). The returned aux["relu_flips"] lists, per
    layer, the oracle pre-activations of the units whose decision differs, so a test can bound them by the rounding error.

    This is the specification the CUDA backward kernels implement:
      dlogits = (softmax - onehot)/B; linear: dW = dY^T X, db = sum dY, dX = dY W;
      ReLU'(0) = 0; max-pool routes the gradient to the FIRST maximum of each window in
      row-major order (floor mode: trailing rows/cols that fit no window get zero);
      conv: dW = sum patches^T dY, db = sum dY, dX = transposed conv (not needed for conv1).
    """
    P = {k: v.detach().to(dtype) for k, v in params.items()}
    x = x.to(dtype)
    B = x.shape[0]
    # forward, keeping what backward needs
    conv_in, conv_out, pooled, amax = [], [], [], []
    h = x
    for name, k, s, p in CONV_SPECS:
        conv_in.append(h)
        z = F.conv2d(h, P[f"{name}.weight"], P[f"{name}.bias"], stride=s)
        conv_out.append(z)
        a = torch.clamp_min(z, 0)
        Hc, Wc = a.shape[-2:]
        Hp, Wp = Hc // p, Wc // p
        win = a[..., :Hp * p, :Wp * p].reshape(B, -1, Hp, p, Wp, p).permute(0, 1, 2, 4, 3, 5).reshape(B, -1, Hp, Wp, p * p)
        # first maximum in row-major window order
        m = win.max(dim=-1, keepdim=True).values
        first = (win == m).to(torch.int64).argmax(dim=-1)
        if argmax_override is not None:
            first = argmax_override[len(amax)].to(torch.int64)
            m = win.gather(-1, first[..., None])
        amax.append(first)
        h = m.squeeze(-1)
        pooled.append(h)
    feats = [torch.flatten(h, 1)]
    pre = []
    for i, (name, _fi, _fo) in enumerate(FC_SPECS):
        z = feats[-1] @ P[f"{name}.weight"].t() + P[f"{name}.bias"]
        pre.append(z)
        feats.append(torch.clamp_min(z, 0) if i < 2 else z)
    logits = feats[-1]
    sm = torch.softmax(logits, dim=1)
    loss = (torch.logsumexp(logits, 1) - logits.gather(1, y[:, None]).squeeze(1)).mean()
    g = sm.clone()
    g[torch.arange(B), y] -= 1.0
    g /= B
    grads = OrderedDict()
    relu_flips = {}
    for i in (2, 1, 0):
        name = FC_SPECS[i][0]
        if i < 2:
            mask = pre[i] > 0
            if relu_override is not None:
                dev_mask = relu_override["fc"][i].to(torch.bool)
                relu_flips[name] = pre[i][mask != dev_mask]
                mask = dev_mask
            g = g * mask.to(dtype)
        grads[f"{name}.weight"] = g.t() @ feats[i]
        grads[f"{name}.bias"] = g.sum(0)
        g = g @ P[f"{name}.weight"]
    g = g.reshape(pooled[-1].shape)
    for li in (3, 2, 1, 0):
        name, k, s, p = CONV_SPECS[li]
        z = conv_out[li]
        Hc, Wc = z.shape[-2:]
        Hp, Wp = Hc // p, Wc // p
        # un-pool: scatter to the first-max position, then ReLU mask
        onehot = F.one_hot(amax[li], p * p).to(dtype) * g[..., None]
        dz = torch.zeros_like(z)
        dz[..., :Hp * p, :Wp * p] = onehot.reshape(B, -1, Hp, Wp, p, p).permute(0, 1, 2, 4, 3, 5).reshape(B, -1, Hp * p, Wp * p)
        if relu_override is not None:
            # the gradient sits only at the routed position of each window, whose activation is the pooled value
            dev_mask = relu_override["pooled"][li].to(torch.bool)
            relu_flips[name] = pooled[li][(pooled[li] > 0) != dev_mask] if argmax_override is None else \
                z[..., :Hp * p, :Wp * p].reshape(B, -1, Hp, p, Wp, p).permute(0, 1, 2, 4, 3, 5).reshape(B, -1, Hp, Wp, p * p) \
                .gather(-1, amax[li][..., None]).squeeze(-1)[(pooled[li] > 0) != dev_mask]
            m_full = torch.zeros_like(z)
            m_full[..., :Hp * p, :Wp * p] = dev_mask.to(dtype).repeat_interleave(p, dim=-2).repeat_interleave(p, dim=-1)
            dz = dz * m_full
        else:
            dz = dz * (z > 0).to(dtype)
        xin = conv_in[li]
        cols = F.unfold(xin, kernel_size=k, stride=s)                    # (B, Cin*k*k, L)
        dzf = dz.reshape(B, dz.shape[1], -1)                             # (B, Cout, L)
        grads[f"{name}.weight"] = torch.einsum("bol,bkl->ok", dzf, cols).reshape(P[f"{name}.weight"].shape)
        grads[f"{name}.bias"] = dz.sum(dim=(0, 2, 3))
        if li > 0:
            w = P[f"{name}.weight"].reshape(dz.shape[1], -1)             # (Cout, Cin*k*k)
            dcols = torch.einsum("ok,bol->bkl", w, dzf)
            g = F.fold(dcols, output_size=xin.shape[-2:], kernel_size=k, stride=s)
    ordered = OrderedDict((k_, grads[k_]) for k_ in PARAM_ORDER)
    return loss, logits, ordered, {"pooled": pooled, "argmax": amax, "conv_out": conv_out, "relu_flips": relu_flips}


# --------------------------------------------------------------------------- optimiser
ADAM_DEFAULTS = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8)  # imitation.py:83 + torch defaults


def adam_update(p, g, m, v, step: int, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """One torch.optim.Adam step (no weight decay, no amsgrad), in place on p, m, v.

    m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
    p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)      (torch/optim/adam.py single-tensor path)
    """
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))
    return p, m, v


def lr_at_epoch(epoch: int, base_lr: float = 1e-3, milestones=(20, 30), gamma: float = 0.1) -> float:
    """MultiStepLR(milestones=[20,30], gamma=0.1) -- imitation.py:84-86."""
    return base_lr * (gamma ** sum(1 for m_ in milestones if epoch >= m_))


class OracleTrainer:
    """training_step -> zero_grad -> backward -> Adam.step, on CPU, in the oracle's arithmetic."""

    def __init__(self, params, dtype=torch.float32, **adam):
        self.p = OrderedDict((k, v.detach().to(dtype).clone()) for k, v in params.items())
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in self.p.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in self.p.items())
        self.t = 0
        self.dtype = dtype
        self.adam = {**ADAM_DEFAULTS, **adam}

    def step(self, x, y) -> float:
        loss, _logits, grads = loss_and_grads(self.p, x, y, self.dtype)
        self.t += 1
        for k in self.p:
            adam_update(self.p[k], grads[k], self.m[k], self.v[k], self.t, **self.adam)
        return float(loss)


# --------------------------------------------------------------------------- synthetic data
def synth_frames(seed: int, n_frames: int, h: int = 256, w: int = 256, n_actions: int = 9,
                 label_noise: float = 0.25):
    """Deterministic CARLA-shaped synthetic sequence (numpy PCG64, platform independent).

    Frames are uniform-noise RGB in [0,128) with a +100 band whose row position announces the
    NEXT frame's label (wrong with probability `label_noise`), so a 1k-step run has a loss
    curve that falls from ln 9 and then plateaus near 1.0 instead of collapsing to 0;
    labels are uniform in [0, n_actions). The reference has no data in-tree (data/ holds only
    .gitkeep); BASELINE.json asks for synthetic frames at the 256x256 nets.py:14 hard-codes.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    labels = rng.integers(0, n_actions, size=n_frames + 1, dtype=np.int64)
    shown = labels.copy()
    flip = rng.random(n_frames + 1) < label_noise
    shown[flip] = rng.integers(0, n_actions, size=int(flip.sum()))
    frames = rng.integers(0, 128, size=(n_frames, h, w, 3), dtype=np.uint8)
    band = h // n_actions
    for i in range(n_frames):
        # frame i announces the label of frame i+1: a sample's last stacked frame then
        # predicts the sample's label (label index = window end, imitation_dataset.py:125,131)
        r0 = int(shown[i + 1]) * band
        frames[i, r0:r0 + band] += 100
    return frames, labels[:n_frames]
