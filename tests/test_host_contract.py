"""Host-side logic and the C-ABI surface, on CPU (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import bc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from carla_imitation_learning_b200 import _lib
    _lib.build()
    header = open(os.path.join(ROOT, "include", "bc_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(bc_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.lib().bc_abi_version() == 2


def test_no_device_is_an_error_not_a_fallback():
    from carla_imitation_learning_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("this is the CPU-box check")
    assert _lib.lib().bc_device_check() != 0
    assert b"no CPU fallback" in _lib.lib().bc_last_error_string()


def test_arena_layout_matches_reference_shapes():
    from carla_imitation_learning_b200 import _lib
    total, offsets, sizes = _lib.arena_layout(4, 9)
    shapes = O.param_shapes(4, 9)
    assert [int(np.prod(s)) for s in shapes.values()] == sizes
    assert sum(sizes) == 133305
    spans = sorted(zip(offsets, sizes))
    for (o0, n0), (o1, _n1) in zip(spans, spans[1:]):
        assert o0 + n0 <= o1 and o1 % 32 == 0        # disjoint, 128 B aligned
    assert total % 32 == 0 and spans[-1][0] + spans[-1][1] <= total
    total12, _, sizes12 = _lib.arena_layout(12, 9)   # BASELINE config 4: 3 cameras x 4 frames
    assert sizes12[0] == 16 * 12 * 49 and total12 > total


def test_convnet1_shell_matches_reference_init_and_keys(golden_dir):
    from src.architectures.nets import ConvNet1
    g = np.load(os.path.join(golden_dir, "ref_step_b4.npz"))
    torch.manual_seed(12345)
    net = ConvNet1({"obs_size": 4, "n_actions": 9}).cpu()
    sd = net.state_dict()
    assert list(sd.keys()) == list(O.PARAM_ORDER)
    flat = np.concatenate([sd[k].reshape(-1).numpy() for k in O.PARAM_ORDER])
    assert np.array_equal(flat, g["init"])           # bit-identical to the reference under seed 12345
    assert tuple(net.example_input_array.shape) == (1, 4, 256, 256)
    assert float(np.abs(net.example_input_array.numpy()).sum()) == float(g["example_checksum"])
    assert net.cnn_base[0].weight.shape == (16, 4, 7, 7) and net.fc[4].weight.shape == (9, 32)
    # every parameter is a view of one arena
    base = net._arena.data_ptr()
    for p in net.parameters():
        assert p.data_ptr() == base + 4 * p._bc_offset
    # checkpoints load both ways
    ref_sd = {k: torch.from_numpy(np.random.default_rng(0).standard_normal(tuple(v.shape)).astype(np.float32)) for k, v in sd.items()}
    net.load_state_dict(ref_sd)
    assert all(torch.equal(net.state_dict()[k], ref_sd[k]) for k in ref_sd)
    with pytest.raises(RuntimeError):
        net.load_state_dict({"bogus": torch.zeros(1)})
    with pytest.raises(TypeError):
        net.half()


def test_imitation_contract_without_a_device():
    from src.architectures.nets import ConvNet1
    from src.models.imitation import Imitation
    net = ConvNet1({"obs_size": 4, "n_actions": 9}).cpu()
    loaders = {"train_dataloader": "a", "val_dataloader": "b", "test_dataloader": "c"}
    m = Imitation({"obs_size": 4}, net, loaders)
    assert m.h_params == {"obs_size": 4} and m.net is net and m.data_loader is loaders
    assert (m.train_dataloader(), m.val_dataloader(), m.test_dataloader()) == ("a", "b", "c")
    opts, scheds = m.configure_optimizers()
    assert isinstance(opts[0], torch.optim.Optimizer) and opts[0].param_groups[0]["lr"] == 1e-3
    assert opts[0].defaults["betas"] == (0.9, 0.999) and opts[0].defaults["eps"] == 1e-8
    sch = scheds[0]
    assert list(sch.milestones.keys()) == [20, 30] and sch.gamma == 0.1
    assert torch.equal(m.scale_image(torch.tensor([-1.0, 1.0])), torch.tensor([0.0, 1.0]))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU"):
            m.training_step((torch.zeros(1, 4, 256, 256), torch.zeros(1, dtype=torch.int64)), 0)
        with pytest.raises(RuntimeError):
            opts[0].step()
    m.validation_epoch_end([torch.tensor(1.0), torch.tensor(3.0)])   # logger absent: must not raise


def test_hydra_compose_restatement_keeps_reference_keys():
    from carla_imitation_learning_b200.config import compose
    hp = compose("config", overrides=["model=imitation"])
    ref_keys = ["logs", "camera", "NUM_EPOCHS", "BATCH_SIZE", "LEARNING_RATE", "DROP_OUT", "TEST_SIZE", "VALID_SIZE",
                "obs_size", "n_actions", "frame_skip", "train_logs", "test_logs", "image_size", "alpha", "beta",
                "log_dir", "data_dir", "pytorch_seed"]
    assert all(k in hp for k in ref_keys)
    assert (hp["obs_size"], hp["n_actions"], hp["frame_skip"], hp["BATCH_SIZE"], hp.pytorch_seed) == (4, 9, 4, 64, 12345)
    assert hp["image_size"] == [1, 224, 224] and hp["NUM_EPOCHS"] == 50 and hp["LEARNING_RATE"] == 0.001
    for k in ("_target_", "data_dir", "batch_size", "train_val_test_split", "num_workers", "pin_memory"):
        assert k in hp.datamodule
    assert hp.datamodule.data_dir == hp.data_dir == "data/"
    hp["camera"] = "semantic"                        # train.py:99 mutates it
    assert hp.camera == "semantic"
    assert compose("config", overrides=["model=imitation", "BATCH_SIZE=32"])["BATCH_SIZE"] == 32


def test_product_synthetic_data_equals_oracle_generator():
    from carla_imitation_learning_b200.data import continous_to_discreet, synthetic_sequence
    f1, l1 = synthetic_sequence(5, 6, 64, 64)
    f2, l2 = O.synth_frames(5, 6, 64, 64)
    assert np.array_equal(f1, f2) and np.array_equal(l1, l2)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_labels.npz"))
    assert np.array_equal(continous_to_discreet(g["steer"], g["throttle"], g["brake"]), g["action"])


def test_product_code_never_imports_the_oracle():
    bad = []
    for top in ("carla_imitation_learning_b200", "src"):
        for dp, _dn, fn in os.walk(os.path.join(ROOT, top)):
            for f in fn:
                if f.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(dp, f)).read(), re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_imitation_aux_and_raw_segment_mirror_the_reference_contract():
    """ImitationAux (imitation.py:94-159): criterion attribute = lossCriterion, Imitation's hooks; ConvNetRawSegment (nets.py:42-78)
    raises TypeError on construction exactly like the reference's class does (recorded in ref_aux_step_b4.npz)."""
    from src.architectures.nets import ConvNet1, ConvNetRawSegment
    from src.models import imitation as M
    hp = {"obs_size": 4, "n_actions": 9}
    model = M.ImitationAux(hp, ConvNet1(hp), {"train_dataloader": 1, "val_dataloader": 2, "test_dataloader": 3})
    assert model.criterion is M.lossCriterion
    assert (model.train_dataloader(), model.val_dataloader(), model.test_dataloader()) == (1, 2, 3)
    for hook in ("forward", "training_step", "validation_step", "training_epoch_end", "validation_epoch_end", "configure_optimizers", "scale_image"):
        assert callable(getattr(model, hook))
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_aux_step_b4.npz"))
    with pytest.raises(TypeError):
        ConvNetRawSegment(hp)
    assert str(g["raw_segment_error"]) == "TypeError"
    with pytest.raises(IndexError):       # one-column labels: the reference indexes y[:, 1]
        M.lossCriterion(model, None, [torch.zeros(2, 4, 256, 256), torch.zeros(2, dtype=torch.int64)])


def test_stacked_camera_window_is_a_view_with_the_channel_order_frame_major():
    """BASELINE configs[3] on the host side: planes of 3 cameras interleaved frame by frame; sliding_window(frame_skip=12, step=3)
    gives sample i = planes [3i, 3i+12) without a copy, channel 3*f + cam = frame i+f of camera cam; StagedBatch reports the same shape."""
    from carla_imitation_learning_b200 import StagedBatch, sliding_window
    B, h = 5, 4
    planes = torch.arange(3 * (B + 4) * h * h, dtype=torch.float32).reshape(3 * (B + 4), h, h)
    x = sliding_window(planes, frame_skip=12, step=3)
    assert tuple(x.shape) == (B, 12, h, h) and x.data_ptr() == planes.data_ptr()
    assert x.stride(0) == 3 * h * h and x.stride(1) == h * h
    for i in (0, 2, B - 1):
        for f in range(4):
            for cam in range(3):
                assert torch.equal(x[i, 3 * f + cam], planes[3 * (i + f) + cam])
    sb = StagedBatch(torch.empty((3 * (B + 4), 8)), None, 12, 3)
    assert sb.shape == (B, 12, 256, 256) and StagedBatch(torch.empty((B + 4, 8))).shape == (B, 4, 256, 256)
