"""Run the UNMODIFIED reference in a separate process and dump what it computes (build container only).

    python oracle/ref_live.py OUT.npz [PRODUCT_STATE_DICT.pt]

tests/test_reference_live.py calls this (the reference's `src` package and the product's `src` package cannot share one
interpreter) and compares the oracle and the product's host-side functions with the LIVE reference, not only with the
committed golden files. Imports /root/reference through the stubs in oracle/_stubs like oracle/make_golden.py.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("BC_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "_stubs"), REF, ROOT]

from src.architectures.nets import ConvNet1 as RefNet  # noqa: E402  (reference)
from src.models.imitation import Imitation as RefImitation  # noqa: E402  (reference)
from src.dataset.imitation_dataset import continous_to_discreet  # noqa: E402 (reference)

from oracle import bc_oracle as O  # noqa: E402


def main(out, product_sd=None):
    hp = dict(obs_size=4, n_actions=9)
    torch.manual_seed(12345)
    net = RefNet(hp)
    model = RefImitation(hp, net, {})
    res = {"init": np.concatenate([p.detach().reshape(-1).numpy() for p in net.parameters()]),
           "keys": np.array(list(model.state_dict().keys()))}
    frames, labels = O.synth_frames(77, 3 + 4)
    x_np, y_np = O.sequential_samples(frames, labels)
    x, y = torch.from_numpy(x_np), torch.from_numpy(y_np)
    opts, schs = model.configure_optimizers()
    opt, sch = opts[0], schs[0]
    model._stub_schedulers = sch
    loss = model.training_step((x, y), 0)
    res["logits"] = model(x).detach().numpy()
    res["loss"] = float(loss.detach())
    opt.zero_grad()
    loss.backward()
    res["grads"] = np.concatenate([p.grad.reshape(-1).numpy() for p in net.parameters()])
    opt.step()
    l2 = model.training_step((x, y), 1)
    opt.zero_grad(); l2.backward(); opt.step()
    res["after2"] = np.concatenate([p.detach().reshape(-1).numpy() for p in net.parameters()])
    res["opt_state_keys"] = np.array(sorted(opt.state_dict()["state"][0].keys()))
    # epoch-end hook: one scheduler step per epoch (imitation.py:57-60); LR after 21 and 31 epochs
    lrs = []
    for ep in range(31):
        model.training_epoch_end([{"loss": torch.tensor(1.0)}])
        lrs.append(opt.param_groups[0]["lr"])
    res["lr_after_epochs"] = np.array(lrs)
    # label discretisation on a grid that hits every branch (imitation_dataset.py:317-339)
    rng = np.random.default_rng(0)
    steer = np.concatenate([rng.uniform(-1, 1, 200), [0.0, 0.05, -0.05, 0.049, 2.0 * 0 + 0.0]])
    throttle = rng.choice([0.0, 0.5, 1.0], size=steer.size)
    brake = rng.choice([0.0, 1.0], size=steer.size)
    # pandas 3 returns read-only `.values` (copy-on-write), which the reference's in-place writes predate: feed it a column
    # holder with the 2021-era semantics (writable arrays), as oracle/make_golden.py does
    class _Col:
        def __init__(self, a):
            self.values = a
    df = dict(steer=_Col(steer.copy()), throttle=_Col(throttle.copy()), brake=_Col(brake.copy()))
    res["labels"] = np.asarray(continous_to_discreet(df), dtype=np.float64)
    res["label_inputs"] = np.stack([steer, throttle, brake])
    if product_sd:
        sd = torch.load(product_sd, map_location="cpu", weights_only=False)
        missing, unexpected = model.load_state_dict(sd["state_dict"], strict=True)
        res["product_ckpt_loaded"] = np.array(1)
        if sd.get("optimizer_states"):
            opt.load_state_dict(sd["optimizer_states"][0])      # torch.optim.Adam accepts the fused optimiser's state
            res["product_opt_loaded"] = np.array(1)
        res["logits_after_load"] = model(x).detach().numpy()
    np.savez(out, **res)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
