"""CPU tests (-m "not gpu"): the oracle against the committed golden vectors of the reference's own FAST code, against the
compiled reference when oracle/_ref is present, and against independent numpy restatements."""
import hashlib
import os

import numpy as np
import pytest

from svo_pro_universal_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def fast_gold():
    return np.load(os.path.join(GOLD, "fast_ref_golden.npz"))


@pytest.fixture(scope="module")
def orc_gold():
    return np.load(os.path.join(GOLD, "oracle_golden.npz"))


def _case_pyr(orc, seed, w, h):
    img0 = synth.make_image(seed, w, h, n_rect=max(8, w * h // 400))
    return img0, orc.create_img_pyramid(img0, 3 if min(w, h) >= 28 else 1)


def test_fast_oracle_matches_reference_golden(orc, fast_gold):
    """Rows a2-a4: restated closed-form FAST == the reference's generated trees (golden vectors made from oracle/_ref)."""
    for seed, w, h, thr in fast_gold["cases"]:
        img0, pyr = _case_pyr(orc, int(seed), int(w), int(h))
        assert sha(img0) == str(fast_gold[f"img_sha_{seed}"]), "synthetic image generator drifted"
        for l, im in enumerate(pyr):
            xy = orc.fast_detect(im, int(thr), 10)
            assert np.array_equal(xy, fast_gold[f"xy_{seed}_{l}"])
            sc = orc.fast_score10(im, xy, int(thr))
            assert np.array_equal(sc, fast_gold[f"score_{seed}_{l}"])
            nm = orc.fast_nonmax3x3(xy, sc)
            assert np.array_equal(nm, fast_gold[f"nonmax_{seed}_{l}"])
            assert np.array_equal(orc.fast_detect(im, int(thr), 9), fast_gold[f"xy9_{seed}_{l}"])


def test_fast_oracle_matches_compiled_reference(orc):
    """Same comparison against the live compiled reference (only where oracle/_ref travelled)."""
    if orc.ref_lib() is None:
        pytest.skip("oracle/_ref/libfast_ref.so not built on this box")
    rng = np.random.default_rng(5)
    for trial in range(4):
        w, h = int(rng.integers(24, 200)), int(rng.integers(8, 120))
        img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        img = np.kron(rng.integers(0, 256, (h // 4 + 1, w // 4 + 1)), np.ones((4, 4)))[:h, :w].astype(np.uint8) if trial % 2 else img
        for thr in (3, 10, 40):
            a = orc.fast_detect(img, thr, 10, "orc")
            assert np.array_equal(a, orc.fast_detect(img, thr, 10, "ref_sse2"))
            assert np.array_equal(a, orc.fast_detect(img, thr, 10, "ref_plain10"))
            s = orc.fast_score10(img, a, thr, "orc")
            assert np.array_equal(s, orc.fast_score10(img, a, thr, "ref"))
            assert np.array_equal(orc.fast_nonmax3x3(a, s, "orc"), orc.fast_nonmax3x3(a, s, "ref"))


def test_fast_empty_and_tiny_images(orc):
    assert len(orc.fast_detect(np.zeros((6, 40), np.uint8), 10)) == 0      # h < 7: no rows to test
    assert len(orc.fast_detect(np.full((30, 30), 77, np.uint8), 1)) == 0   # flat image
    assert len(orc.fast_nonmax3x3(np.zeros((0, 2), np.int16), np.zeros(0, np.int32))) == 0


def _half_sse2_numpy(img):
    """Independent numpy statement of the SSE2 sequence: avg_epu8 vertically, then avg_epu16 of even/odd bytes."""
    a = img.astype(np.int32)
    h, w = a.shape
    sw = w >> 4
    v = (a[0:h - (h % 2):2] + a[1:h:2] + 1) >> 1
    v = v[:, :16 * sw]
    return ((v[:, 0::2] + v[:, 1::2] + 1) >> 1).astype(np.uint8)


def _half_scalar_numpy(img):
    a = img.astype(np.int32)
    h, w = a.shape
    oh, ow = h // 2, w // 2
    a = a[:2 * oh, :2 * ow]
    return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2]) // 4).astype(np.uint8)


def test_halfsample_formulas(orc):
    """Row a1: both halfSample branches against independent numpy restatements; the per-level predicate for 752x480."""
    img = synth.make_image(3)
    pyr = orc.create_img_pyramid(img, 5)
    assert [p.shape for p in pyr] == [(480, 752), (240, 376), (120, 188), (60, 94), (30, 47)]
    assert np.array_equal(pyr[1], _half_sse2_numpy(pyr[0]))       # 752 % 16 == 0 -> SSE2 rounding formula
    for l in (2, 3, 4):                                           # 376, 188, 94 are not multiples of 16 -> truncating
        assert np.array_equal(pyr[l], _half_scalar_numpy(pyr[l - 1]))
    assert (pyr[1] != _half_scalar_numpy(pyr[0])).mean() > 0.5    # the two formulas really differ
    pyr0 = orc.create_img_pyramid(img, 5, 0)
    assert np.array_equal(pyr0[1], _half_scalar_numpy(pyr0[0]))
    odd = synth.make_image(4, 47, 31)
    p = orc.create_img_pyramid(odd, 2)
    assert p[1].shape == (15, 23) and np.array_equal(p[1], _half_scalar_numpy(odd))


def test_oracle_regression_pins(orc, orc_gold):
    """Oracle outputs pinned by committed vectors (guards the restatement against accidental edits)."""
    img = synth.make_image(11)
    for mode in (-1, 0):
        pyr = orc.create_img_pyramid(img, 5, mode)
        assert [sha(p) for p in pyr] == list(orc_gold[f"pyr_sha_mode{mode}"])
    for seed in (0, 5):
        c = orc.fast_detector(synth.make_image(seed))
        g = orc_gold[f"corners_{seed}"]
        for k in ("x", "y", "level", "score"):
            assert np.array_equal(c[k], g[k])


def test_oracle_sparse_align_pins_and_convergence(orc, orc_gold):
    from helpers import oracle_align, pose_diff
    for seed in (1, 2, 3):
        d = synth.make_align_pair(seed)
        for name, kw in (("default", {}), ("illum_robust", dict(estimate_illumination_gain=1, estimate_illumination_offset=1,
                                                                 robustification=1))):
            r = oracle_align(orc, d, orc.default_align_options(**kw))
            assert r.n_tracked == int(orc_gold[f"align_{name}_{seed}_n"])
            assert list(r.iters) == list(orc_gold[f"align_{name}_{seed}_iters"])
            np.testing.assert_allclose(np.array(r.T_icur_iref), orc_gold[f"align_{name}_{seed}_T"], rtol=0, atol=1e-12)
            dq, dt = pose_diff(r.T_icur_iref, d["T_icur_iref_gt"])
            assert dq < 2e-3 and dt < 5e-3, "alignment should land near the synthetic ground truth"


def test_oracle_align2d_recovers_known_shift(orc):
    """align2D on a shifted copy of the same smooth patch converges to the known shift (self-consistency, row c3)."""
    img = synth.make_image(21, blur=2)
    pyr = [img]
    cam = synth.EUROC_CAM
    yy, xx = np.mgrid[0:480, 0:752]
    shifted = np.clip(np.rint(synth.bilinear(img, xx - 0.6, yy + 0.35)), 0, 255).astype(np.uint8)  # content moves +0.6, -0.35
    errs = []
    for px0 in synth.pick_features(img, 40, 3):
        x, y = int(px0[0]), int(px0[1])
        pwb = np.ascontiguousarray(img[y - 5:y + 5, x - 5:x + 5]).reshape(-1)
        px = (px0 + np.array([0.3, -0.2])).copy()
        ok = orc.lib().orc_align2d(shifted.ctypes.data_as(orc.u8p), 752, 480, 752, pwb.ctypes.data_as(orc.u8p), 10, 1, 0,
                                   px.ctypes.data_as(orc.f64p))
        if ok == 1:
            errs.append(np.abs(px - (px0 + np.array([0.6, -0.35]))).max())
    # the reference stops once an update is < 0.03 px, so the residual error is a few hundredths of a pixel
    assert len(errs) >= 30 and np.median(errs) < 0.12, (len(errs), np.median(errs))
    assert pyr and cam


def test_oracle_vogiatzis_known_answers(orc, orc_gold):
    st = np.array([0.25, 0.0123, 10.0, 10.0])
    for row in orc_gold["vogiatzis_rows"]:
        ok = orc.lib().orc_update_filter_vogiatzis(row[0], row[1], 1.0 / 1.5, st.ctypes.data_as(orc.f64p))
        assert ok == int(row[2])
        np.testing.assert_allclose(st, row[3:], rtol=1e-13)
    # hand-checkable limit: an uninformative measurement (huge tau2) barely moves mu and keeps sigma2 finite/positive
    st = np.array([0.25, 0.01, 10.0, 10.0])
    orc.lib().orc_update_filter_vogiatzis(0.5, 1e6, 1.0, st.ctypes.data_as(orc.f64p))
    assert abs(st[0] - 0.25) < 1e-3 and st[1] > 0
    # negative mean -> reference resets mu to 1 and reports failure
    st = np.array([1e-3, 100.0, 10.0, 1.0])
    ok = orc.lib().orc_update_filter_vogiatzis(-5.0, 1e-6, 1.0, st.ctypes.data_as(orc.f64p))
    assert ok == 0 and st[0] == 1.0
