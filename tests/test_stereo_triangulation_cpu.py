"""CPU tests of SURVEY §8 row f3 (StereoTriangulation::compute): the oracle restatement of the matching loop against the committed
outputs of the reference's own compiled compute() (detector + std::random_shuffle + epipolar matching + frame bookkeeping)."""
import os

import numpy as np

import helpers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_matches_reference_golden(orc):
    g = np.load(os.path.join(GOLD, "stereo_tri_ref_golden.npz"))
    n_failed_total = 0
    for i, case in enumerate(helpers.STEREO_TRI_CASES):
        keep = []
        d, s1, p0, p1, f0, f1 = helpers.stereo_tri_frames(orc, case, keep)
        order = g[f"order_{i}"]
        det, f = helpers.stereo_tri_entries(orc, case, d, p0, order)
        assert np.array_equal(det["px"], g[f"px0_{i}"]) and np.array_equal(det["type"], g[f"type0_{i}"])  # frame0's new columns
        ft = orc.make_features(det["px"][order], f, det["grad"][order], det["type"][order], det["level"][order])
        res, ns, nf = orc.stereo_triangulate(f0, f1, ft, case[3], 0, case[4], case[5], case[6])
        helpers.assert_stereo_matches_reference(res, order, g, i, tol=1e-9)
        assert ns == min(case[3], int((res["status"] == 2).sum())) and nf == int((res["status"] == 1).sum())
        if ns == case[3]:  # stopped early: nothing behind the last success was touched
            last = np.flatnonzero(res["status"] == 2)[-1]
            assert (res["status"][last + 1:] == 0).all()
        n_failed_total += nf
    assert n_failed_total > 50


def test_compiled_reference_agrees_with_golden(orc):
    if orc.ref_frontend_lib() is None or not hasattr(orc.ref_frontend_lib(), "ref_stereo_triangulation_compute"):
        return
    g = np.load(os.path.join(GOLD, "stereo_tri_ref_golden.npz"))
    r = helpers.stereo_tri_reference(orc)
    for k in r:
        assert np.array_equal(r[k], g[k]), k
