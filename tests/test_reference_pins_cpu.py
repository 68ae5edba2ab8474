"""CPU tests (-m "not gpu"): the oracle against outputs of the REFERENCE's own compiled code for the rows whose sources
compile here against the container-only shims (oracle/shim): vk::halfSample (a1), createPatchFromPatchWithBorder (c2),
align2D / align1D (c3, c4), ZMSSD<4> (c5), TukeyWeightFunction (b6), RadialTangentialDistortion (s1), seed.h helpers (d1),
OccupandyGrid2D::getCellIndex (a5). tests/golden/direct_ref_golden.npz holds the reference's outputs (made by
tests/golden/make_golden.py); the live library is used as well wherever oracle/_ref travelled."""
import os

import numpy as np
import pytest

import helpers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALIGN_TOL_PX = 1e-3  # north_star: align2D/1D sub-pixel results within 1e-3 px


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "direct_ref_golden.npz"))


@pytest.fixture(scope="module")
def mine(orc):
    return helpers.direct_outputs(orc, "orc")


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


def test_pyramid_bytes_equal_reference_halfsample(mine, gold):
    for w, h, _ in helpers.PYR_SHAPES:
        assert list(mine[f"pyr_sha_{w}x{h}"]) == list(gold[f"pyr_sha_{w}x{h}"]), (w, h)


def test_integer_rows_bit_exact(mine, gold):
    assert same(mine["zmssd"], gold["zmssd"])
    assert same(mine["patch_from_border"], gold["patch_from_border"])
    assert same(mine["grid_cells"], gold["grid_cells"])


def test_float_helpers_bit_exact(mine, gold):
    """Same operations in the same order, no FMA contraction on either side -> identical bits."""
    for k in ("tukey", "radtan_distort", "radtan_undistort", "radtan_jacobian", "seed_helpers"):
        assert same(mine[k], gold[k]), k


def test_align1d_bit_exact_and_align2d_within_tolerance(mine, gold):
    a1, g1 = mine["align1d"], gold["align1d"]
    assert same(a1, g1), "align1D (3x3 float normal equations) must reproduce the reference bit for bit"
    a2, g2 = mine["align2d"], gold["align2d"]
    assert same(a2[:, 0], g2[:, 0]), "converged flags differ"
    assert 20 < g2[:, 0].sum() < len(g2), "cases must cover converged and non-converged outcomes"
    d = np.abs(a2[:, 1:] - g2[:, 1:])
    d = d[np.isfinite(d)]
    # the only restated third-party arithmetic on this row is Eigen's 4x4 inverse (cofactor form, summation order unpinned)
    assert d.max() <= ALIGN_TOL_PX, d.max()
    assert (d == 0).mean() > 0.9


def test_live_compiled_reference_random_cases(orc):
    if orc.ref_direct_lib() is None:
        pytest.skip("oracle/_ref/libdirect_ref.so not built on this box")
    rng = np.random.default_rng(3)
    from svo_pro_universal_b200 import synth
    for trial in range(6):  # halfSample: SSE2 branch, unaligned start, non-continuous rows, odd sizes
        w, h = (int(rng.integers(2, 12)) * 16, int(rng.integers(4, 60))) if trial % 2 == 0 else (int(rng.integers(9, 200)), int(rng.integers(5, 90)))
        img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        assert same(orc.create_img_pyramid(img, 2)[1], orc.ref_half_sample(img))
        # off the 16-byte alignment or with padded rows the reference falls back to the truncating mean (vision.cpp:80-98)
        trunc = orc.create_img_pyramid(img, 2, 0)[1]
        assert same(trunc, orc.ref_half_sample(img, align_offset=4))
        assert same(trunc, orc.ref_half_sample(img, extra_stride=16))
    img = synth.make_image(21)
    for i in range(120):
        x, y = int(rng.integers(12, 740)), int(rng.integers(12, 468))
        pwb = img[y - 5:y + 5, x - 5:x + 5]
        px0 = np.array([x + rng.uniform(-2, 2), y + rng.uniform(-2, 2)])
        th = rng.uniform(0, 2 * np.pi)
        o1, p1, h1 = orc.align1d(img, (np.cos(th), np.sin(th)), pwb, px0, which="orc")
        o2, p2, h2 = orc.align1d(img, (np.cos(th), np.sin(th)), pwb, px0, which="ref")
        assert o1 == o2 and same(p1, p2) and h1 == h2
        o1, p1 = orc.align2d(img, pwb, px0, which="orc")
        o2, p2 = orc.align2d(img, pwb, px0, which="ref")
        assert o1 == o2 and (same(p1, p2) or np.abs(p1 - p2).max() <= ALIGN_TOL_PX)
