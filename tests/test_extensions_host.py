"""EXTENSIONS -- NO REFERENCE PARITY (host side, no GPU): the branched net's parameter layout and initialisation against
oracle/ext_oracle.py, the augmentation table, the folded-BN algebra."""
import numpy as np
import torch

from oracle import bc_oracle as O
from oracle import ext_oracle as X


def test_branched_net_keys_init_and_arena_views():
    from src.architectures.branched import ConvNet1Branched
    hp = {"obs_size": 4, "n_actions": 3, "n_branches": 4, "branch_loss": "mse"}
    torch.manual_seed(12345)
    net = ConvNet1Branched(hp)
    P = X.init_branched_params(12345, 4, 3, 4)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    assert list(sd) == list(P) == [f"cnn_base.{i}.{s}" for i in (0, 3, 6, 9) for s in ("weight", "bias")] + list(X.branch_param_names(4))
    assert all(torch.equal(sd[k], P[k]) for k in P)
    # every branch parameter is a view of ONE flat arena in the head-segment order fc.4, fc.2, fc.0
    base = net._barena.data_ptr()
    for p in net.branch_parameters():
        assert p.data_ptr() == base + 4 * p._bc_offset
    assert net._barena.numel() == 4 * net._head_len
    net.load_state_dict({k: v + 1 for k, v in sd.items()})
    assert torch.equal(net.state_dict()["branches.3.4.bias"], sd["branches.3.4.bias"] + 1)


def test_augment_table_is_the_specifications_table():
    from carla_imitation_learning_b200.data import augment_table
    a = augment_table(3, 9, (300, 310), brightness=0.3, contrast=0.1, saturation=0.25, mean=0.5, std=0.25)
    assert np.array_equal(a, X.augment_params(3, 9, (300, 310), brightness=0.3, contrast=0.1, saturation=0.25, mean=0.5, std=0.25))
    assert a[:, 0].max() <= 44 and a[:, 1].max() <= 54 and (a[:, 6] == 4.0).all()


def test_identity_augmentation_is_the_reference_gray_conversion():
    frames, _ = O.synth_frames(2, 3, 256, 256)
    ident = np.zeros((3, 8), np.float32)
    ident[:, 2:5] = 1.0
    ident[:, 6] = 1.0
    assert np.abs(X.stage_augmented(frames, ident) - O.gray_stack(frames)).max() <= 2.5e-7


def test_fold_batchnorm_algebra():
    gen = torch.Generator().manual_seed(0)
    w, b = torch.randn(8, 3, 3, 3, generator=gen, dtype=torch.float64), torch.randn(8, generator=gen, dtype=torch.float64)
    gamma, beta = torch.randn(8, generator=gen, dtype=torch.float64), torch.randn(8, generator=gen, dtype=torch.float64)
    mean, var = torch.randn(8, generator=gen, dtype=torch.float64), torch.rand(8, generator=gen, dtype=torch.float64) + 0.1
    x = torch.randn(2, 3, 9, 9, generator=gen, dtype=torch.float64)
    F = torch.nn.functional
    ref = F.batch_norm(F.conv2d(x, w, b), mean, var, gamma, beta, training=False, eps=1e-5)
    w2, b2 = X.fold_batchnorm(w, b, gamma, beta, mean, var)
    assert float((F.conv2d(x, w2, b2) - ref).abs().max()) <= 1e-12
