"""Host model of the staging kernel's gray arithmetic (csrc/stage.cu::gray_px) against the reference formula
`np.dot(img, [0.299, 0.587, 0.114]) / 255.0 -> float32` (/root/reference/src/dataset/imitation_dataset.py:121,130)
for ALL 2^24 (R,G,B) triples. The device runs the same three f32 operations; the GPU test
tests/test_gpu_parity.py::test_stage_gray_bit_exact_all_rgb checks the kernel itself."""
from fractions import Fraction

import numpy as np


def _all_rgb():
    r, g, b = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
    return np.stack([r, g, b], -1).reshape(-1, 3).astype(np.uint8)


def gray_px_model(s: np.ndarray) -> np.ndarray:
    """q0 = s*r ; rem = fma(-q0, 255000, s) ; q = fma(rem, r, q0), r = rn_f32(1/255000) -- in f32.
    The FMAs are emulated in f64: -q0*255000 + s is exact there (24 + 18 bits), rem*r + q0 carries 48 + 24 bits."""
    r = np.float32(1.0) / np.float32(255000.0)
    sf = s.astype(np.float32)
    q0 = (sf * r).astype(np.float32)
    rem = sf.astype(np.float64) - q0.astype(np.float64) * 255000.0
    rem32 = rem.astype(np.float32)
    assert (rem32.astype(np.float64) == rem).all(), "the remainder must be exact in f32"
    return (q0.astype(np.float64) + rem32.astype(np.float64) * np.float64(r)).astype(np.float32)


def test_gray_depends_only_on_the_integer_dot_and_is_its_correctly_rounded_quotient():
    rgb = _all_rgb()
    ref = (np.dot(rgb, [0.299, 0.587, 0.114]) / 255.0).astype(np.float32)       # the reference's formula, verbatim
    c = rgb.astype(np.int64)
    s = 299 * c[:, 0] + 587 * c[:, 1] + 114 * c[:, 2]
    assert int(s.max()) == 255000
    got = gray_px_model(s)
    assert (got.view(np.uint32) == ref.view(np.uint32)).all()
    # and both are the correctly rounded quotient s / 255000 (exact rational arithmetic on a sample of distinct s)
    us = np.unique(s)
    rng = np.random.default_rng(0)
    for v in np.concatenate([us[:64], us[-64:], rng.choice(us, 4096, replace=False)]):
        exact = np.float32(float(Fraction(int(v), 255000)))
        assert gray_px_model(np.array([v]))[0] == exact
