"""GPU parity tests against outputs of the REFERENCE's own compiled front-end (tests/golden/frontend_ref_golden.npz, made by
tests/golden/make_golden.py from oracle/_ref/libfrontend_ref.so): SparseImgAlign::run (b), Matcher::findMatchDirect /
findEpipolarMatchDirect (c) and the updateSeed chain (d) on the seeded cases of tests/helpers.py:frontend_outputs."""
import os

import numpy as np
import pytest

import helpers
from svo_pro_universal_b200 import capi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROT_TOL, TRANS_TOL, PX_TOL, REL_TOL = 1e-4, 1e-4, 1e-3, 1e-4  # north_star tolerances


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "frontend_ref_golden.npz"))


def test_sparse_align_equals_reference(ctx, gold):
    rows = gold["align_rows"]
    k = 0
    for seed in (1, 2, 3):
        d = synth.make_align_pair(seed)
        for kw in helpers.ALIGN_OPTION_SETS:
            res, _, _ = helpers.gpu_align(ctx, [d], capi.sparse_align_options(**kw))
            dq, dt = helpers.pose_diff(res[0]["T_icur_iref"], rows[k, :7])
            assert dq < ROT_TOL and dt < TRANS_TOL and dq < 1e-8 and dt < 1e-8, (seed, kw, dq, dt)
            assert res[0]["n_tracked"] == rows[k, 10]
            np.testing.assert_allclose(res[0]["chi2"], rows[k, 9], rtol=1e-5)
            k += 1
        priors = np.zeros(1, capi.ALIGN_PRIOR_DTYPE)
        priors[0]["T"] = synth.se3_mul(d["T_icur_iref_gt"], synth.se3_exp_small(np.array([1e-3, -2e-3, 1e-3]), np.array([2e-3, 0, -1e-3])))
        res, _, _ = helpers.gpu_align(ctx, [d], capi.sparse_align_options(lambda_rot=0.5, lambda_trans=0.1), priors=priors)
        dq, dt = helpers.pose_diff(res[0]["T_icur_iref"], rows[k, :7])
        assert dq < 1e-8 and dt < 1e-8, ("prior", seed, dq, dt)
        k += 1
    d = synth.make_align_pair(5, cam=synth.EUROC_CAM_RADTAN)
    for kw in (dict(), dict(use_distortion_jacobian=1)):
        res, _, _ = helpers.gpu_align(ctx, [d], capi.sparse_align_options(**kw))
        dq, dt = helpers.pose_diff(res[0]["T_icur_iref"], rows[k, :7])
        assert dq < 1e-8 and dt < 1e-8, ("radtan", kw, dq, dt)
        k += 1
    assert k == len(rows)


def _match_set():
    ms = synth.make_match_set(7, n_features=240)
    ms["px"][:8] = np.array([[3.0, 3.0]]) + np.arange(8)[:, None] * 0.25
    ms["depth"][8:16] = 0.05
    return ms


def test_matcher_equals_reference(ctx, gold):
    ms = _match_set()
    ref = capi.Pyramid(ctx, 1, 752, 480, 5); cur = capi.Pyramid(ctx, 1, 752, 480, 5)
    ref.upload(ms["ref_img"]); cur.upload(ms["cur_img"]); ref.build(); cur.build()
    cam = capi.Camera.from_dict(ms["cam"])
    ft = capi.make_features(ms["px"], ms["f"], ms["grad"], ms["type"], ms["level"])
    for name, kw in (("default", dict()), ("gain", dict(affine_est_gain=1))):
        got = capi.find_match_direct(ctx, ref, cur, cam, cam, ms["T_cur_ref"], ft, ms["depth"], np.ascontiguousarray(ms["px_guess"]), capi.matcher_options(**kw))
        res = gold[f"fmd_{name}_result"]
        assert np.array_equal(got["result"], res)
        ok = res == 0
        assert np.abs(got["px_cur"][ok] - gold[f"fmd_{name}_px_cur"][ok]).max() < PX_TOL
        assert np.array_equal(got["search_level"][ok], gold[f"fmd_{name}_search_level"][ok])
        np.testing.assert_allclose(got["A_cur_ref"][ok], gold[f"fmd_{name}_A_cur_ref"][ok], rtol=1e-9, atol=1e-12)
    d_inv = 1.0 / ms["depth"]
    d3 = np.ascontiguousarray(np.stack([d_inv * np.random.default_rng(1).uniform(0.9, 1.1, len(d_inv)), d_inv * 1.5, d_inv * 0.6], 1))
    for name, kw in (("sphere", dict()), ("plane", dict(scan_on_unit_sphere=0)), ("a1d", dict(align_1d=1)), ("nosub", dict(subpix_refinement=0))):
        got = capi.find_epipolar_match_direct(ctx, ref, cur, cam, cam, ms["T_cur_ref"], ft, d3, capi.matcher_options(**kw))
        res = gold[f"epi_{name}_result"]
        assert np.array_equal(got["result"], res), name
        ok = res == 0
        assert np.abs(got["px_cur"][ok] - gold[f"epi_{name}_px_cur"][ok]).max() < PX_TOL
        np.testing.assert_allclose(got["depth"][ok], gold[f"epi_{name}_depth"][ok], rtol=1e-4)
        np.testing.assert_allclose(got["epi_length_pyramid"][ok], gold[f"epi_{name}_epi_length_pyramid"][ok], rtol=1e-9)
        assert np.array_equal(got["reject"], gold[f"epi_{name}_reject"])
        np.testing.assert_allclose(got["epi_image"], gold[f"epi_{name}_epi_image"], rtol=1e-12, atol=1e-12)  # Matcher::epi_image_


def test_scan_epipolar_line_equals_reference(ctx, orc, gold):
    """Matcher::scanEpipolarLine on its own (svo_cuda_scan_epipolar_line, all scans in one launch) against the reference's own
    compiled scan on the same segments / patches: best score bit-equal, best pixel to rounding."""
    ms = _match_set()
    cur = capi.Pyramid(ctx, 1, 752, 480, 5)
    cur.upload(ms["cur_img"]); cur.build()
    cam = capi.Camera.from_dict(ms["cam"])
    keep = []
    rf = orc.make_frame(orc.create_img_pyramid(ms["ref_img"], 5), ms["cam"], keep=keep)
    cf = orc.make_frame(orc.create_img_pyramid(ms["cur_img"], 5), ms["cam"], keep=keep)
    d_inv = 1.0 / ms["depth"]
    d3 = np.stack([d_inv * np.random.default_rng(1).uniform(0.9, 1.1, len(d_inv)), d_inv * 1.5, d_inv * 0.6], 1)
    sc = helpers.scan_cases(orc, ms, rf, cf, orc.make_features(ms["px"], ms["f"], ms["grad"], ms["type"], ms["level"]), d3)
    for name, kw, z0 in (("sphere", dict(), None), ("plane", dict(scan_on_unit_sphere=0), None), ("capped", dict(max_epi_search_steps=4), None),
                         ("low_start", dict(), np.full(len(sc["A"]), 9000, np.int32))):
        px, z = capi.scan_epipolar_line(ctx, cur, cam, sc["A"], sc["B"], sc["C"], sc["patch"], sc["level"], sc["epi_length"],
                                        capi.matcher_options(**kw), zmssd_best=z0)
        assert np.array_equal(z, gold[f"scan_{name}_zmssd"]), name
        assert np.abs(px - gold[f"scan_{name}_px"]).max() < 1e-9, name


def test_update_seeds_equals_reference(ctx, gold):
    sq = synth.make_seed_sequence(17, n_seeds=160, n_obs=6)
    S, O = len(sq["px"]), len(sq["cur_imgs"])
    ref = capi.Pyramid(ctx, 1, 752, 480, 5); cur = capi.Pyramid(ctx, O, 752, 480, 5)
    ref.upload(sq["ref_img"]); cur.upload(np.stack(sq["cur_imgs"])); ref.build(); cur.build()
    cam = capi.Camera.from_dict(sq["cam"])
    ft = capi.make_features(sq["px"], sq["f"], sq["grad"], sq["type"].astype(np.int32), sq["level"])
    obs = np.ascontiguousarray(np.tile(np.arange(O, dtype=np.int32)[:, None], (1, S)))
    for name, dkw in (("vog", dict()), ("gauss", dict(use_vogiatzis_update=0)), ("conv", dict(check_convergence=1, seed_convergence_sigma2_thresh=50.0))):
        types, state = sq["type"].copy(), sq["state"].copy()
        n, _ = capi.update_seeds(ctx, ref, cur, cam, cam, ft, types, state, np.full(S, sq["mu_range"]), obs, obs,
                                 np.ascontiguousarray(sq["T_cur_ref"]), capi.matcher_options(), capi.depth_filter_options(**dkw))
        assert int(n[0]) == int(gold[f"seeds_{name}_n"])
        assert np.array_equal(types, gold[f"seeds_{name}_types"])
        np.testing.assert_allclose(state, gold[f"seeds_{name}_state"], rtol=REL_TOL)
