"""CPU tests (-m "not gpu"): the oracle against outputs of the REFERENCE's own compiled front-end —
SparseImgAlign::run (rows b1-b9), patch_warp + Matcher::findMatchDirect / findEpipolarMatchDirect (c1, c6, c7) and
depth_filter_utils::updateSeed / updateFilterVogiatzis / computeTau (d1-d3) — built from /root/reference against the
container-only stand-ins of oracle/shim (oracle/_ref/libfrontend_ref.so). tests/golden/frontend_ref_golden.npz holds the
reference's outputs (tests/golden/make_golden.py); where oracle/_ref travelled the live library is exercised too."""
import os

import numpy as np
import pytest

import helpers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROT_TOL, TRANS_TOL = 1e-4, 1e-4   # north_star: poses within 1e-4 rad / 1e-4 m
PX_TOL = 1e-3                     # north_star: align2D/1D within 1e-3 px
REL_TOL = 1e-4                    # north_star: seed mean / variance within 1e-4 relative


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "frontend_ref_golden.npz"))


@pytest.fixture(scope="module")
def mine(orc):
    return helpers.frontend_outputs(orc, "orc")


def test_sparse_img_align_equals_reference(mine, gold):
    a, g = mine["align_rows"], gold["align_rows"]
    assert a.shape == g.shape and len(a) == 3 * (len(helpers.ALIGN_OPTION_SETS) + 1) + 2
    assert np.array_equal(a[:, 10], g[:, 10]), "number of tracked features"
    for i in range(len(a)):
        dq, dt = helpers.pose_diff(a[i, :7], g[i, :7])
        assert dq < ROT_TOL and dt < TRANS_TOL, (i, dq, dt)
        # far inside the tolerance in fact: same operations, only the depth of a feature enters through |landmark - camera|
        assert dq < 1e-9 and dt < 1e-9, (i, dq, dt)
    np.testing.assert_allclose(a[:, 9], g[:, 9], rtol=1e-6)                         # chi2 (float accumulator)
    np.testing.assert_allclose(mine["align_H"], gold["align_H"], rtol=1e-9, atol=1e-6)  # last evaluated Hessian
    dq, dt = helpers.pose_diff(mine["stereo_row"][:7], gold["stereo_row"][:7])
    assert dq < 1e-9 and dt < 1e-9 and mine["stereo_row"][10] == gold["stereo_row"][10]
    for c in range(2):
        dq, dt = helpers.pose_diff(mine["stereo_T_f_w"][c], gold["stereo_T_f_w"][c])
        assert dq < 1e-9 and dt < 1e-9


def test_matcher_equals_reference(mine, gold):
    n_ok = 0
    for name in ("fmd_default", "fmd_gain", "epi_sphere", "epi_plane", "epi_a1d", "epi_nosub"):
        res = gold[f"{name}_result"]
        assert np.array_equal(mine[f"{name}_result"], res), name
        ok = res == 0
        n_ok += int(ok.sum())
        assert ok.sum() > 0.5 * len(res) and (~ok).sum() > 8, "cases must cover successes and failures"
        assert np.array_equal(mine[f"{name}_search_level"][ok], gold[f"{name}_search_level"][ok])
        assert np.array_equal(mine[f"{name}_patch_with_border"][ok], gold[f"{name}_patch_with_border"][ok]), "warped patch bytes"
        assert np.abs(mine[f"{name}_px_cur"][ok] - gold[f"{name}_px_cur"][ok]).max() < PX_TOL
        np.testing.assert_allclose(mine[f"{name}_f_cur"][ok], gold[f"{name}_f_cur"][ok], atol=1e-6)
        np.testing.assert_allclose(mine[f"{name}_A_cur_ref"][ok], gold[f"{name}_A_cur_ref"][ok], rtol=1e-12, atol=1e-14)
        if name.startswith("epi"):
            np.testing.assert_allclose(mine[f"{name}_depth"][ok], gold[f"{name}_depth"][ok], rtol=1e-4)
            assert np.array_equal(mine[f"{name}_epi_length_pyramid"][ok], gold[f"{name}_epi_length_pyramid"][ok])
            assert np.array_equal(mine[f"{name}_reject"], gold[f"{name}_reject"])
            assert np.array_equal(mine[f"{name}_epi_image"], gold[f"{name}_epi_image"]), "Matcher::epi_image_ (set before every return)"
    # Matcher::scanEpipolarLine on its own: best ZMSSD bit-equal, best pixel up to the rounding of the rotated bearing
    for name in ("sphere", "plane", "capped", "low_start"):
        assert np.array_equal(mine[f"scan_{name}_zmssd"], gold[f"scan_{name}_zmssd"]), name
        assert np.abs(mine[f"scan_{name}_px"] - gold[f"scan_{name}_px"]).max() < 1e-9, name
    assert (gold["scan_sphere_zmssd"] < 2000 * 64).sum() > 100 and (gold["scan_capped_zmssd"] != gold["scan_sphere_zmssd"]).any()
    assert ((gold["scan_low_start_zmssd"] == 9000) & (gold["scan_sphere_zmssd"] > 9000)).any(), "a scan that never beats the caller's start"
    # wherever the sub-pixel position is bit-identical (everything except a few align2D refinements, whose 4x4 inverse is
    # Eigen arithmetic restated on both sides) the triangulated depth is bit-identical too
    for name in ("epi_sphere", "epi_plane", "epi_a1d", "epi_nosub"):
        same = (gold[f"{name}_result"] == 0) & (mine[f"{name}_px_cur"] == gold[f"{name}_px_cur"]).all(axis=1)
        assert same.sum() > 0.9 * (gold[f"{name}_result"] == 0).sum()
        assert np.array_equal(mine[f"{name}_depth"][same], gold[f"{name}_depth"][same])


def test_depth_filter_equals_reference(mine, gold):
    for name in ("vog", "gauss", "conv"):
        assert int(mine[f"seeds_{name}_n"]) == int(gold[f"seeds_{name}_n"]) > 300
        assert np.array_equal(mine[f"seeds_{name}_types"], gold[f"seeds_{name}_types"])
        assert np.array_equal(mine[f"seeds_{name}_ok"], gold[f"seeds_{name}_ok"])
        np.testing.assert_allclose(mine[f"seeds_{name}_state"], gold[f"seeds_{name}_state"], rtol=REL_TOL)
        same = (mine[f"seeds_{name}_state"] == gold[f"seeds_{name}_state"]).all(axis=1)
        assert same.mean() > 0.9, "all but the seeds refined through align2D's restated 4x4 inverse are bit-identical"


def test_live_compiled_reference_filter_and_tau(orc):
    L = orc.ref_frontend_lib()
    if L is None:
        pytest.skip("oracle/_ref/libfrontend_ref.so not built on this box")
    rng = np.random.default_rng(2)
    for _ in range(3000):  # updateFilterVogiatzis incl. its guards (NaN normalisation, sigma2 < 0, mu < 0)
        st = np.array([rng.uniform(-0.1, 1), rng.uniform(1e-6, 0.05), rng.uniform(1, 30), rng.uniform(1, 30)])
        st2 = st.copy()
        z, tau2 = rng.uniform(-0.2, 1.5), 10.0 ** rng.uniform(-8, -1)
        assert orc.lib().orc_update_filter_vogiatzis(z, tau2, 0.66, orc._f64(st)) == L.ref_update_filter_vogiatzis(z, tau2, 0.66, orc._f64(st2))
        assert np.array_equal(st, st2, equal_nan=True)
        if z > 0 and st[0] > 0 and st[1] > 0:  # the reference CHECK-aborts on a negative mean / variance (depth_filter.cpp:575-576)
            assert orc.lib().orc_update_filter_gaussian(z, tau2, orc._f64(st)) == L.ref_update_filter_gaussian(z, tau2, orc._f64(st2))
            assert np.array_equal(st, st2, equal_nan=True)
    for _ in range(500):   # computeTau
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        T = np.concatenate([q, rng.normal(size=3) * 0.3])
        f = rng.normal(size=3); f /= np.linalg.norm(f)
        z = rng.uniform(0.3, 12)
        assert orc.lib().orc_compute_tau(orc._f64(T), orc._f64(f), z, 0.00218) == L.ref_compute_tau(orc._f64(T), orc._f64(f), z, 0.00218)


def test_bench_cpu_step_reference_equals_port(orc):
    """bench.py's CPU legs: the frame-pair step on the compiled reference (vk::halfSample pyramid + SparseImgAlign::run) and on
    the oracle port produce the same poses — the timed baseline is the computation the parity tests check."""
    if orc.ref_frontend_lib() is None:
        pytest.skip("oracle/_ref/libfrontend_ref.so not built on this box")
    import sys
    sys.path.insert(0, os.path.dirname(GOLD.rstrip("/")).rsplit("/tests", 1)[0])
    import bench
    uniq = bench.make_unique_pairs(3, 1000)
    keep = []
    refs, l0 = bench.orc_frames(orc, uniq, keep)
    T0 = bench.perturbed_initial_poses(np.stack([d["T_imu_world_ref"] for d in uniq]), 7)   # per-pair initial guess, as the bench legs
    curs = [orc.make_frame([d["cur_img"]], d["cam"], d["T_cam_imu"], T0[k], keep=keep) for k, d in enumerate(uniq)]
    opt = orc.default_align_options()
    a = orc.pyramid_align_batch(l0, refs, curs, opt, bench.N_LEVELS, 2)
    b = orc.ref_pyramid_align_batch(l0, refs, curs, opt, bench.N_LEVELS, 2)
    for x, y in zip(a, b):
        assert x.n_tracked == y.n_tracked > 100
        dq, dt = helpers.pose_diff(np.array(x.T_icur_iref[:]), np.array(y.T_icur_iref[:]))
        assert dq < 1e-9 and dt < 1e-9
