"""GPU parity tests against outputs of the REFERENCE's own compiled code (tests/golden/direct_ref_golden.npz, made by
tests/golden/make_golden.py from oracle/_ref/libdirect_ref.so): pyramid bytes (a1), align2D / align1D (c3, c4)."""
import hashlib
import os

import numpy as np
import pytest

import helpers
from svo_pro_universal_b200 import capi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALIGN_TOL_PX = 1e-3  # north_star: align2D/1D sub-pixel results within 1e-3 px


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "direct_ref_golden.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_pyramid_bytes_equal_reference_halfsample(ctx, gold):
    for w, h, nl in helpers.PYR_SHAPES:
        img = synth.make_image(100 + w, w, h, n_rect=max(8, w * h // 400))
        p = capi.Pyramid(ctx, 1, w, h, nl)
        p.upload(img[None])
        p.build()
        assert [sha(p.download(0, l)) for l in range(nl)] == list(gold[f"pyr_sha_{w}x{h}"]), (w, h)


def test_align2d_align1d_equal_reference(ctx, gold):
    img, cases = helpers.align_cases()
    pyr = capi.Pyramid(ctx, 1, img.shape[1], img.shape[0], 1)
    pyr.upload(img[None])
    g2, g1 = gold["align2d"], gold["align1d"]
    groups = {}
    for i, c in enumerate(cases):
        groups.setdefault((c["n_iter"], c["est_offset"], c["est_gain"]), []).append(i)
    n_exact2 = 0
    for (n_iter, eo, eg), idx in groups.items():
        z = np.zeros(len(idx), np.int32)
        pwb = np.stack([cases[i]["pwb"].reshape(100) for i in idx])
        px0 = np.stack([cases[i]["px0"] for i in idx])
        dirs = np.stack([cases[i]["dir"] for i in idx])
        px2, conv2 = capi.align2d(ctx, pyr, z, z, pwb, px0, n_iter, eo, eg)
        px1, conv1, hinv = capi.align1d(ctx, pyr, z, z, dirs, pwb, px0, n_iter, eo, eg)
        assert np.array_equal(conv2.astype(bool), g2[idx, 0].astype(bool))
        assert np.array_equal(conv1.astype(bool), g1[idx, 0].astype(bool))
        # align1D: the 3x3 float path reproduces the reference bit for bit (pixel and h_inv)
        assert np.array_equal(px1, g1[idx, 1:3], equal_nan=True)
        assert np.array_equal(hinv, g1[idx, 3], equal_nan=True)
        d = np.abs(px2 - g2[idx, 1:3])
        d = d[np.isfinite(d)]
        assert d.max() <= ALIGN_TOL_PX
        n_exact2 += int((d == 0).sum())
    assert n_exact2 > 0.9 * 2 * len(cases)
