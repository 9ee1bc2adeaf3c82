"""The oracle and the product's host-side functions against the LIVE reference (not only the committed golden files).

Runs only where /root/reference exists (the build container): oracle/ref_live.py imports the UNMODIFIED reference classes in
a separate interpreter and dumps what they compute; on the GPU box the test skips (the golden fixtures under tests/golden
carry the same facts there). Reference lines: src/architectures/nets.py:6-39, src/models/imitation.py:27-91,
src/dataset/imitation_dataset.py:317-339.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import bc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BC_REFERENCE", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="the reference tree is not present on this box")


@pytest.fixture(scope="module")
def live(tmp_path_factory):
    d = tmp_path_factory.mktemp("ref_live")
    out = str(d / "ref.npz")
    # a checkpoint in the Lightning layout written from the PRODUCT's module (CPU: state_dict only, no optimiser state)
    from src.architectures.nets import ConvNet1
    from src.models.imitation import Imitation
    torch.manual_seed(4321)
    hp = {"obs_size": 4, "n_actions": 9}
    model = Imitation(hp, ConvNet1(hp), {})
    ck = str(d / "product.ckpt")
    torch.save({"state_dict": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, "optimizer_states": [],
                "epoch": 0, "global_step": 0}, ck)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_live.py"), out, ck], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return dict(np.load(out, allow_pickle=True)), {k: v.detach().cpu() for k, v in model.state_dict().items()}


def _flat(named):
    return np.concatenate([np.asarray(named[k].detach().double()).reshape(-1) for k in O.PARAM_ORDER])


def test_oracle_equals_live_reference(live):
    ref, _ = live
    P = O.init_params(12345)
    assert np.array_equal(_flat(P).astype(np.float32), ref["init"])                 # same RNG consumption order, bit for bit
    frames, labels = O.synth_frames(77, 7)
    x, y = O.sequential_samples(frames, labels)
    x, y = torch.from_numpy(x), torch.from_numpy(y)
    loss, logits, grads = O.loss_and_grads(P, x, y)
    assert np.allclose(logits.numpy(), ref["logits"], rtol=0, atol=1e-6)
    assert abs(float(loss) - float(ref["loss"])) <= 1e-6
    g = _flat(grads)
    assert np.abs(g - ref["grads"]).max() <= 1e-6 * np.abs(ref["grads"]).max()
    tr = O.OracleTrainer(P)
    tr.step(x, y); tr.step(x, y)
    assert np.abs(_flat(tr.p) - ref["after2"]).max() <= 2e-5      # |update| = 2e-3 after two sign-like Adam steps (same bound as test_oracle_golden)
    assert [O.lr_at_epoch(e + 1) for e in range(31)] == pytest.approx(list(ref["lr_after_epochs"]), rel=1e-12)


def test_product_host_contract_equals_live_reference(live):
    ref, product_sd = live
    from carla_imitation_learning_b200.data import continous_to_discreet
    assert list(ref["keys"]) == ["net." + k for k in O.PARAM_ORDER] == list(product_sd.keys())
    s, t, b = ref["label_inputs"]
    assert np.array_equal(continous_to_discreet(s, t, b), ref["labels"])
    assert np.array_equal(O.discretise_actions(s, t, b), ref["labels"])
    assert sorted(ref["opt_state_keys"]) == ["exp_avg", "exp_avg_sq", "step"]      # what FusedAdam.state_dict() carries per parameter


def test_product_checkpoint_loads_into_the_live_reference(live):
    """A Lightning-layout state_dict written by the product loads strictly into the reference's Imitation and gives the
    oracle's logits for those weights (train.py:198-201 load_from_checkpoint, the other direction)."""
    ref, product_sd = live
    assert int(ref["product_ckpt_loaded"]) == 1
    P = {k[len("net."):]: v for k, v in product_sd.items()}
    frames, labels = O.synth_frames(77, 7)
    x, _ = O.sequential_samples(frames, labels)
    logits = O.forward(P, torch.from_numpy(x))
    assert np.allclose(logits.numpy(), ref["logits_after_load"], rtol=0, atol=1e-6)
