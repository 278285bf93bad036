"""The arithmetic claim behind the tensor path, checked on the CPU with NumPy: a three-term hi/lo split of the operands
(hi.hi + hi.lo + lo.hi, accumulated in higher precision) reproduces FP32-class dot products, both with TF32 planes
(KMB_PATH_TENSOR_3XTF32) and with FP16 planes of power-of-two scaled data (KMB_PATH_TENSOR_3XF16: the scale the device
prepass picks puts the largest operand in [2^13, 2^14)).  DESIGN.md section 7 states the error model tested here."""
import numpy as np
import pytest


def split_f16(w):
    """What split_points_f16_kernel does: p with 2^p max|w| in [2^13, 2^14); hi = fp16(2^p w), lo = fp16(2^p w - hi)."""
    p = 13 - int(np.floor(np.log2(np.abs(w).max())))
    s = np.float32(w) * np.float32(2.0 ** p)
    hi = s.astype(np.float16)
    lo = (s - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64), 2.0 ** (-2 * p)


def to_tf32(x):
    """cvt.rna.tf32.f32: round to nearest (ties away) on the 13 dropped mantissa bits."""
    u = np.float32(x).view(np.uint32).astype(np.uint64)
    u = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return u.view(np.float32)


def split_tf32(w):
    hi = to_tf32(w)
    lo = to_tf32(np.float32(w) - hi)
    return hi.astype(np.float64), lo.astype(np.float64), 1.0


def three_term(split, u, v):
    uh, ul, su = split(u)
    vh, vl, sv = split(v)
    assert su == sv
    return (uh @ vh.T + uh @ vl.T + ul @ vh.T) * su


@pytest.mark.parametrize("split", [split_f16, split_tf32])
@pytest.mark.parametrize("D", [64, 784])
def test_three_term_split_is_fp32_class(split, D):
    rng = np.random.RandomState(D)
    r = (3.0 / D) ** 0.5
    pts = np.float32(r * rng.rand(400, D))
    pts = (pts - pts.mean(0)).astype(np.float32)      # the prepass centres on the column means
    u, v = pts[:150], pts[150:]
    if split is split_f16:   # one common scale for both operand sets, as the prepass uses
        both = lambda w: split_f16(np.concatenate([w, pts]))   # noqa: E731
        uh, ul, s = both(u); vh, vl, _ = both(v)
        got = (uh[:150] @ vh[:250].T + uh[:150] @ vl[:250].T + ul[:150] @ vh[:250].T) * s
    else:
        got = three_term(split, u, v)
    exact = u.astype(np.float64) @ v.astype(np.float64).T
    scale = np.linalg.norm(u, axis=1)[:, None] * np.linalg.norm(v, axis=1)[None, :]
    err = np.abs(got - exact) / scale
    assert err.max() <= 2.0 ** -20, err.max()         # dropped lo.lo term + plane rounding: ~2^-22 per product
    # the squared distances the kernel forms from it keep 1e-6 of the typical distance
    d2 = (u.astype(np.float64) ** 2).sum(1)[:, None] + (v.astype(np.float64) ** 2).sum(1)[None, :] - 2 * got
    d2_exact = ((u[:, None, :].astype(np.float64) - v[None, :, :]) ** 2).sum(-1)
    assert np.abs(d2 - d2_exact).max() <= 2e-6 * d2_exact.mean()


def test_f16_planes_with_an_outlier_and_tiny_coordinates():
    """One far point stretches the power-of-two scale; coordinates far below the largest fall into FP16 subnormals for
    lo.  The absolute error stays at 2^-24 of the LARGEST operand, i.e. the accuracy of the bulk degrades gracefully
    (what tests/test_product_gpu.py::test_tensor_f16_scaling_range sees on the device)."""
    rng = np.random.RandomState(1)
    D = 48
    pts = np.float32(0.25 * rng.rand(300, D))
    pts[0] += 25.0                                     # outlier, 100x the spread
    pts[:, 5] *= 1e-4                                  # one tiny column
    hi, lo, s = split_f16(pts)
    rec = (hi + lo) * np.sqrt(s)
    big = np.abs(pts).max()
    assert np.abs(rec - pts).max() <= big * 2.0 ** -23
    got = (hi @ hi.T + hi @ lo.T + lo @ hi.T) * s
    exact = pts.astype(np.float64) @ pts.astype(np.float64).T
    bulk = slice(1, None)
    assert np.abs(got - exact)[bulk, bulk].max() <= big * np.abs(pts[1:]).max() * D * 2.0 ** -22
