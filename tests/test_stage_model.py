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


def gray_px_bf16_model(rgb: np.ndarray) -> np.ndarray:
    """csrc/stage.cu::gray_px_bf16_exact (the Toeplitz-ready bf16 staging kernel): s from two dp4a over the bytes (R,G,B,G)
    with the byte weights (255,255,114,255) + (44,77,0,0) on top of the bit pattern of 2^23, one FADD, then
    q = fma(s, r_lo, s * r) in f32 (the FMA emulated in f64: the product carries 18 + 24 bits), rounded to bf16 (nearest even)."""
    c = rgb.astype(np.int64)
    px = np.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1]], -1)
    w = np.array([[255, 255, 114, 255], [44, 77, 0, 0]], dtype=np.int64)
    acc = np.int64(0x4B000000) + (px @ w.T).sum(-1)
    sf = acc.astype(np.uint32).view(np.float32) - np.float32(8388608.0)
    r = np.float32(1.0) / np.float32(255000.0)
    r_lo = np.float32(1.0 / 255000.0 - float(r))
    q0 = (sf * r).astype(np.float32)
    q = (q0.astype(np.float64) + sf.astype(np.float64) * np.float64(r_lo)).astype(np.float32)
    return _bf16_bits(q)


def _bf16_bits(x32: np.ndarray) -> np.ndarray:
    u = x32.view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)           # round to nearest even (no NaN / inf here)


def test_bf16_staging_arithmetic_equals_the_rounded_reference_for_all_triples():
    """The bf16 planes are bf16_rn(reference f32 gray) for ALL 2^24 triples although the kernel never forms the correctly
    rounded f32 quotient; a single multiplication would NOT do (4 values of s sit within 2^-25 of a bf16 midpoint)."""
    rgb = _all_rgb()
    ref = (np.dot(rgb, [0.299, 0.587, 0.114]) / 255.0).astype(np.float32)
    want = _bf16_bits(ref)
    assert (gray_px_bf16_model(rgb) == want).all()
    c = rgb.astype(np.int64)
    s = 299 * c[:, 0] + 587 * c[:, 1] + 114 * c[:, 2]
    r = np.float32(1.0) / np.float32(255000.0)
    single = _bf16_bits((s.astype(np.float32) * r).astype(np.float32))
    assert sorted(set(s[single != want].tolist())) == [106333, 180791, 212666, 244541]
