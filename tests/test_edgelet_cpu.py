"""CPU tests of SURVEY §8 row f2 (edgelet detector) and rows a5 / a6 (fastDetector, fillFeatures, the detector classes): the oracle
restatement against (i) the committed outputs of the REAL OpenCV for the two imgproc functions the detector executes, (ii) the
committed outputs of the reference's own compiled detectors, and (iii) that compiled reference itself where it travelled."""
import os

import numpy as np

import helpers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_blur_and_scharr_match_real_opencv(orc):
    g = np.load(os.path.join(GOLD, "cv_imgproc_golden.npz"))
    for i in range(len(g["shapes"])):
        img = g[f"img_{i}"]
        blur = orc.gaussian_blur3x3(img)
        assert np.array_equal(blur, g[f"blur_{i}"]), f"GaussianBlur differs from cv2 {g['cv2_version']} on case {i}"
        dx, dy = orc.scharr3x3(blur)
        assert np.array_equal(dx, g[f"dx_{i}"]) and np.array_equal(dy, g[f"dy_{i}"]), f"Scharr differs from cv2 on case {i}"


def test_detectors_match_reference_golden(orc):
    g = np.load(os.path.join(GOLD, "detect_ref_golden.npz"))
    o = helpers.detect_outputs(orc, "orc")
    assert set(o) == set(g.files)
    n_edgelets = 0
    for k in o:
        if k.startswith("det_"):
            continue
        assert np.array_equal(o[k], g[k]), k
        if k.startswith("edgelet_") and k.endswith("_score"):
            n_edgelets += int((o[k] > o[k].min()).sum())
    assert n_edgelets > 1500  # the cases really produce edgelets
    for i in range(len(helpers.DETECT_CASES)):
        for t, max_n in ((0, None), (2, None), (5, None), (2, 60)):
            a = {f: o[f"det_{i}_{t}_{max_n}_{f}"] for f in ("px", "score", "level", "grad", "type")}
            b = {f: g[f"det_{i}_{t}_{max_n}_{f}"] for f in ("px", "score", "level", "grad", "type")}
            helpers.assert_features_equal(a, b, f"case {i} detector {t} max_n {max_n}")
    assert np.array_equal(o["hist_angle"], g["hist_angle"])


def test_detectors_match_compiled_reference(orc):
    if orc.ref_detect_lib() is None:
        return  # the compiled reference did not travel; the golden test above covers the same cases
    a, b = helpers.detect_outputs(orc, "orc"), helpers.detect_outputs(orc, "ref")
    for k in a:
        if not k.startswith("det_"):
            assert np.array_equal(a[k], b[k]), k


def test_fastgrad_semantics(orc):
    """FastGrad = FAST corners first, edgelets only in the cells left empty, capped at max_n; edgelets report level 0."""
    case = helpers.DETECT_CASES[0]
    _, pyr, _ = helpers.detect_case_inputs(orc, case)
    fast = orc.detect_features(orc.DETECTOR_FAST, pyr)
    both = orc.detect_features(orc.DETECTOR_FAST_GRAD, pyr)
    n = len(fast["score"])
    assert n > 100 and len(both["score"]) > n
    assert np.array_equal(both["px"][:n], fast["px"]) and (both["type"][:n] == 7).all() and (both["type"][n:] == 6).all()
    assert (both["level"][n:] == 0).all() and (both["px"][n:] % 2 == 0).all()
    cells = (both["px"][:, 1] // 30) * 26 + both["px"][:, 0] // 30
    assert len(np.unique(cells)) == len(cells)  # one feature per cell
    assert np.allclose(np.hypot(both["grad"][:, 0], both["grad"][:, 1]), 1.0, atol=1e-6)
    capped = orc.detect_features(orc.DETECTOR_FAST_GRAD, pyr, max_n=60)
    assert len(capped["score"]) == 60 and (capped["type"] == 7).all()


def test_float_sqrt_claim():
    """The CUDA kernel computes float(std::sqrt(double(n))) as a correctly rounded float sqrt for n < 2^24 (edgelet.cu)."""
    n = np.arange(0, 1 << 24, dtype=np.int64)
    f = np.sqrt(n.astype(np.float64)).astype(np.float32)
    assert np.array_equal(f, np.sqrt(n.astype(np.float32)))
    # ... and keeps the SQUARED magnitude in its score tile: n -> float(sqrt(n)) is strictly increasing below 2^22, so integer
    # comparisons decide the reference's float comparisons there (scoreGE / scoreGT in edgelet.cu), and non-decreasing above
    assert (np.diff(f[:1 << 22]) > 0).all() and (np.diff(f) >= 0).all()
    t = np.arange(0, 2048, dtype=np.int64)
    assert np.array_equal(f[t * t], t.astype(np.float32))  # mag(thr^2) == thr exactly


def test_angle_bins_on_the_diagonals(orc):
    """Diagonal gradients sit exactly on a histogram bin boundary; the CUDA kernel hard-codes the correctly rounded angles there."""
    for s in (1, 7, 255):
        for gx, gy, const in ((s, s, 0.7853981633974483), (-s, s, 2.356194490192345), (-s, -s, -2.356194490192345), (s, -s, -0.7853981633974483)):
            assert np.arctan2(float(gy), float(gx)) == const
            t = 36 * (const + np.pi) / (2.0 * np.pi)
            b = int(np.floor(t + 0.5))
            assert orc.angle_histogram_bin(gx, gy) == (b if b < 36 else 0)
