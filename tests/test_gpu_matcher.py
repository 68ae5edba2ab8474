"""GPU parity tests, rows c1-c7: affine warp (bit-exact patch bytes), ZMSSD-driven epipolar search, align1D / align2D and the
two Matcher entry points against the oracle. Tolerance: sub-pixel results within 1e-3 px (north_star)."""
import ctypes as C

import numpy as np
import pytest

from svo_pro_universal_b200 import capi, synth

pytestmark = pytest.mark.gpu
PX_TOL = 1e-3


def _setup(ctx, orc, ms, n_levels=5):
    ref = capi.Pyramid(ctx, 1, 752, 480, n_levels)
    cur = capi.Pyramid(ctx, 1, 752, 480, n_levels)
    ref.upload(ms["ref_img"]); cur.upload(ms["cur_img"])
    ref.build(); cur.build()
    keep = []
    rf = orc.make_frame(orc.create_img_pyramid(ms["ref_img"], n_levels), ms["cam"], keep=keep)
    cf = orc.make_frame(orc.create_img_pyramid(ms["cur_img"], n_levels), ms["cam"], keep=keep)
    return ref, cur, rf, cf, keep


def _mopts(orc, **kw):
    return capi.matcher_options(**kw), orc.default_matcher_options(**kw)


def test_warp_affine_patch_bytes_bit_exact(ctx, orc):
    ms = synth.make_match_set(3, n_features=600, max_rot_deg=8.0, max_trans=0.3)
    ref, cur, rf, cf, keep = _setup(ctx, orc, ms)
    cam = capi.Camera.from_dict(ms["cam"])
    ft = capi.make_features(ms["px"], ms["f"], ms["grad"], ms["type"], ms["level"])
    A, sl, pwb, ok = capi.warp_affine(ctx, ref, cam, cam, ms["T_cur_ref"], ft, ms["depth"])
    T = np.ascontiguousarray(ms["T_cur_ref"], np.float64)
    n_ok = 0
    for i in range(len(ft)):
        Ao = np.zeros(4)
        px = np.ascontiguousarray(ms["px"][i]); f = np.ascontiguousarray(ms["f"][i])
        orc.lib().orc_get_warp_matrix_affine(C.byref(rf), C.byref(cf), px.ctypes.data_as(orc.f64p), f.ctypes.data_as(orc.f64p),
                                             float(ms["depth"][i]), T.ctypes.data_as(orc.f64p), int(ms["level"][i]),
                                             Ao.ctypes.data_as(orc.f64p))
        np.testing.assert_allclose(A[i], Ao, rtol=1e-9, atol=1e-12)
        slo = orc.lib().orc_get_best_search_level(Ao.ctypes.data_as(orc.f64p), 4)
        assert sl[i] == slo
        lv = int(ms["level"][i])
        img = rf._keep[0][lv] if hasattr(rf, "_keep") else keep[0][0][lv]
        patch = np.zeros(100, np.uint8)
        oko = orc.lib().orc_warp_affine(Ao.ctypes.data_as(orc.f64p), img.ctypes.data_as(orc.u8p), img.shape[1], img.shape[0],
                                        img.strides[0], px.ctypes.data_as(orc.f64p), lv, slo, 5, patch.ctypes.data_as(orc.u8p))
        assert ok[i] == oko
        if oko:
            assert np.array_equal(pwb[i], patch), f"feature {i}: warped patch bytes differ"
            n_ok += 1
    assert n_ok > 400


@pytest.mark.parametrize("kw", [dict(), dict(affine_est_gain=1), dict(affine_est_offset=0), dict(align_max_iter=3)])
def test_find_match_direct(ctx, orc, kw):
    """BASELINE config 3 shape on one pair: 2000 features, corners -> align2D, 25 % edgelets -> align1D."""
    ms = synth.make_match_set(7, n_features=2000)
    ref, cur, rf, cf, keep = _setup(ctx, orc, ms)
    cam = capi.Camera.from_dict(ms["cam"])
    gopt, oopt = _mopts(orc, **kw)
    ft = capi.make_features(ms["px"], ms["f"], ms["grad"], ms["type"], ms["level"])
    got = capi.find_match_direct(ctx, ref, cur, cam, cam, ms["T_cur_ref"], ft, ms["depth"], np.ascontiguousarray(ms["px_guess"]), gopt)
    oft = orc.make_features(ms["px"], ms["f"], ms["grad"], ms["type"], ms["level"])
    exp = orc.find_match_direct_batch(rf, cf, ms["T_cur_ref"], oft, ms["depth"], ms["px_guess"], oopt, n_threads=8)
    assert np.array_equal(got["result"], exp["result"]), np.flatnonzero(got["result"] != exp["result"])[:10]
    okm = exp["result"] == 0
    assert okm.sum() > 0.5 * len(ft) and (ms["type"][okm] == synth.K_EDGELET).sum() > 50
    assert np.abs(got["px_cur"][okm] - exp["px_cur"][okm]).max() < PX_TOL
    np.testing.assert_allclose(got["f_cur"][okm], exp["f_cur"][okm], atol=1e-5)
    assert np.array_equal(got["search_level"], exp["search_level"])
    np.testing.assert_allclose(got["A_cur_ref"], exp["A_cur_ref"], rtol=1e-9, atol=1e-12)
    ed = okm & (ms["type"] == synth.K_EDGELET)
    np.testing.assert_allclose(got["h_inv"][ed], exp["h_inv"][ed], rtol=1e-5)
    # matches land close to the ground-truth reprojection
    err = np.linalg.norm(got["px_cur"][okm & (ms["type"] == synth.K_CORNER)] - ms["px_true"][okm & (ms["type"] == synth.K_CORNER)], axis=1)
    assert np.median(err) < 0.5


def test_find_match_direct_visibility_and_warp_failures(ctx, orc):
    ms = synth.make_match_set(9, n_features=300)
    # push some features against the border / give absurd depth so kFailVisibility / kFailWarp / kFailAlignment all occur
    ms["px"][:20] = np.array([[3.0, 3.0]]) + np.arange(20)[:, None] * 0.25
    ms["px"][20:40, 0] = 748.0
    ms["depth"][40:60] = 0.05
    ms["px_guess"][60:80] += 40.0
    ref, cur, rf, cf, keep = _setup(ctx, orc, ms)
    cam = capi.Camera.from_dict(ms["cam"])
    gopt, oopt = _mopts(orc)
    ft = capi.make_features(ms["px"], ms["f"], ms["grad"], ms["type"], ms["level"])
    got = capi.find_match_direct(ctx, ref, cur, cam, cam, ms["T_cur_ref"], ft, ms["depth"], np.ascontiguousarray(ms["px_guess"]), gopt)
    exp = orc.find_match_direct_batch(rf, cf, ms["T_cur_ref"], orc.make_features(ms["px"], ms["f"], ms["grad"], ms["type"], ms["level"]),
                                      ms["depth"], ms["px_guess"], oopt)
    assert np.array_equal(got["result"], exp["result"])
    assert {3, 5} <= set(exp["result"].tolist())  # kFailVisibility and kFailAlignment both exercised
    okm = exp["result"] == 0
    assert np.abs(got["px_cur"][okm] - exp["px_cur"][okm]).max() < PX_TOL


def test_align2d_align1d_standalone(ctx, orc):
    """The free functions of feature_alignment.h on caller-provided 10x10 patches."""
    ms = synth.make_match_set(5, n_features=500)
    ref, cur, rf, cf, keep = _setup(ctx, orc, ms)
    cam = capi.Camera.from_dict(ms["cam"])
    ft = capi.make_features(ms["px"], ms["f"], ms["grad"], ms["type"], ms["level"])
    A, sl, pwb, ok = capi.warp_affine(ctx, ref, cam, cam, ms["T_cur_ref"], ft, ms["depth"])
    sel = np.flatnonzero(ok)
    px0 = ms["px_guess"][sel] / (2.0 ** sl[sel])[:, None]
    cur_pyr = orc.create_img_pyramid(ms["cur_img"], 5)
    frame_idx = np.zeros(len(sel), np.int32)
    for est_off, est_gain in ((1, 0), (1, 1), (0, 0)):
        px2, conv2 = capi.align2d(ctx, cur, frame_idx, sl[sel], pwb[sel], px0, 10, est_off, est_gain)
        dirs = np.ascontiguousarray(ms["grad"][sel])
        px1, conv1, hinv = capi.align1d(ctx, cur, frame_idx, sl[sel], dirs, pwb[sel], px0, 10, est_off, est_gain)
        n_conv = 0
        for k, i in enumerate(sel):
            img = cur_pyr[sl[i]]
            p = px0[k].copy()
            c = orc.lib().orc_align2d(img.ctypes.data_as(orc.u8p), img.shape[1], img.shape[0], img.strides[0],
                                      pwb[i].ctypes.data_as(orc.u8p), 10, est_off, est_gain, p.ctypes.data_as(orc.f64p))
            assert c == conv2[k]
            if c:
                assert np.abs(p - px2[k]).max() < PX_TOL
                n_conv += 1
            p = px0[k].copy()
            h = C.c_double()
            c = orc.lib().orc_align1d(img.ctypes.data_as(orc.u8p), img.shape[1], img.shape[0], img.strides[0],
                                      dirs[k].ctypes.data_as(orc.f64p), pwb[i].ctypes.data_as(orc.u8p), 10, est_off, est_gain,
                                      p.ctypes.data_as(orc.f64p), C.byref(h))
            assert c == conv1[k]
            assert abs(h.value - hinv[k]) <= 1e-5 * abs(h.value)
            if c:
                assert np.abs(p - px1[k]).max() < PX_TOL
        assert n_conv > 0.5 * len(sel)


@pytest.mark.parametrize("kw", [dict(), dict(scan_on_unit_sphere=0), dict(max_epi_search_steps=20), dict(subpix_refinement=0),
                                dict(align_1d=1), dict(epi_search_edgelet_filtering=0)])
def test_find_epipolar_match_direct(ctx, orc, kw):
    """Epipolar search with depth priors of varying quality: long lines (ZMSSD scan, both scan variants), short lines
    (direct alignment), edgelet angle rejection, triangulated depth."""
    ms = synth.make_match_set(13, n_features=1200, max_rot_deg=1.0, max_trans=0.25)
    ref, cur, rf, cf, keep = _setup(ctx, orc, ms)
    cam = capi.Camera.from_dict(ms["cam"])
    gopt, oopt = _mopts(orc, **kw)
    rng = np.random.default_rng(5)
    n = len(ms["px"])
    inv = 1.0 / ms["depth"]
    est = inv * rng.uniform(0.7, 1.4, n)
    spread = np.where(rng.uniform(size=n) < 0.2, 0.01, rng.uniform(0.1, 0.8, n)) * inv
    d_inv = np.ascontiguousarray(np.stack([est, est + spread, np.maximum(est - spread, 1e-8)], 1))
    ft = capi.make_features(ms["px"], ms["f"], ms["grad"], ms["type"], ms["level"])
    got = capi.find_epipolar_match_direct(ctx, ref, cur, cam, cam, ms["T_cur_ref"], ft, d_inv, gopt)
    exp = orc.find_epipolar_match_direct_batch(rf, cf, ms["T_cur_ref"], orc.make_features(ms["px"], ms["f"], ms["grad"], ms["type"], ms["level"]),
                                               d_inv, oopt, n_threads=8)
    assert np.array_equal(got["result"], exp["result"]), np.flatnonzero(got["result"] != exp["result"])[:10]
    assert np.array_equal(got["reject"], exp["reject"]) and np.array_equal(got["search_level"], exp["search_level"])
    np.testing.assert_allclose(got["epi_length_pyramid"], exp["epi_length_pyramid"], rtol=1e-9)
    okm = exp["result"] == 0
    assert okm.sum() > 0.3 * n
    assert np.abs(got["px_cur"][okm] - exp["px_cur"][okm]).max() < PX_TOL
    np.testing.assert_allclose(got["depth"][okm], exp["depth"][okm], rtol=1e-4)
    long_lines = okm & (exp["epi_length_pyramid"] >= 2.0)
    assert long_lines.sum() > 100 and (okm & (exp["epi_length_pyramid"] < 2.0)).sum() > 20
    rel = np.abs(got["depth"][long_lines] - ms["depth"][long_lines]) / ms["depth"][long_lines]
    assert np.median(rel) < 0.1  # the search finds the true surface


def test_large_call_work_order(ctx, orc):
    """Calls with >= 16384 features are worked on grouped by (type, level) — a stable counting sort of the feature indices on the device —
    and written back by feature index: 17 shuffled copies of a 1200-feature set (20400 features, types and levels interleaved) give, copy
    by copy, exactly the records of the small (unordered) call, for both Matcher entry points."""
    ms = synth.make_match_set(13, n_features=1200, max_rot_deg=1.0, max_trans=0.25)
    ref, cur, rf, cf, keep = _setup(ctx, orc, ms)
    cam = capi.Camera.from_dict(ms["cam"])
    gopt = capi.matcher_options()
    n = len(ms["px"])
    rng = np.random.default_rng(8)
    inv = 1.0 / ms["depth"]
    est = inv * rng.uniform(0.7, 1.4, n)
    spread = rng.uniform(0.1, 0.8, n) * inv
    d_inv = np.ascontiguousarray(np.stack([est, est + spread, np.maximum(est - spread, 1e-8)], 1))
    ft = capi.make_features(ms["px"], ms["f"], ms["grad"], ms["type"], ms["level"])
    guess = np.ascontiguousarray(ms["px_guess"])
    small0 = capi.find_match_direct(ctx, ref, cur, cam, cam, ms["T_cur_ref"], ft, ms["depth"], guess, gopt)
    small1 = capi.find_epipolar_match_direct(ctx, ref, cur, cam, cam, ms["T_cur_ref"], ft, d_inv, gopt)
    idx = np.concatenate([rng.permutation(n) for _ in range(17)])
    assert len(idx) >= 16384 and len(np.unique(ms["type"])) >= 2 and len(np.unique(ms["level"])) >= 2
    big0 = capi.find_match_direct(ctx, ref, cur, cam, cam, ms["T_cur_ref"], np.ascontiguousarray(ft[idx]), np.ascontiguousarray(ms["depth"][idx]),
                                  np.ascontiguousarray(guess[idx]), gopt)
    big1 = capi.find_epipolar_match_direct(ctx, ref, cur, cam, cam, ms["T_cur_ref"], np.ascontiguousarray(ft[idx]), np.ascontiguousarray(d_inv[idx]), gopt)
    assert big0.tobytes() == small0[idx].tobytes()
    assert big1.tobytes() == small1[idx].tobytes()
    assert (small0["result"] == 0).sum() > 0.5 * n and (small1["result"] == 0).sum() > 0.3 * n


def test_per_feature_frame_and_transform_indices(ctx, orc):
    """Features of several frame pairs in one launch (ref/cur frame index + T index per feature)."""
    sets = [synth.make_match_set(s, n_features=200) for s in (21, 22, 23)]
    B = len(sets)
    ref = capi.Pyramid(ctx, B, 752, 480, 5)
    cur = capi.Pyramid(ctx, B, 752, 480, 5)
    ref.upload(np.stack([m["ref_img"] for m in sets])); cur.upload(np.stack([m["cur_img"] for m in sets]))
    ref.build(); cur.build()
    cam = capi.Camera.from_dict(sets[0]["cam"])
    cat = lambda k: np.ascontiguousarray(np.concatenate([m[k] for m in sets]))
    idx = np.concatenate([np.full(len(m["px"]), i, np.int32) for i, m in enumerate(sets)])
    ft = capi.make_features(cat("px"), cat("f"), cat("grad"), cat("type"), cat("level"))
    T = np.stack([m["T_cur_ref"] for m in sets])
    gopt, oopt = _mopts(orc)
    got = capi.find_match_direct(ctx, ref, cur, cam, cam, T, ft, cat("depth"), cat("px_guess"), gopt, ref_frame_idx=idx,
                                 cur_frame_idx=idx, T_idx=idx)
    off = 0
    for i, m in enumerate(sets):
        keep = []
        rf = orc.make_frame(orc.create_img_pyramid(m["ref_img"], 5), m["cam"], keep=keep)
        cf = orc.make_frame(orc.create_img_pyramid(m["cur_img"], 5), m["cam"], keep=keep)
        exp = orc.find_match_direct_batch(rf, cf, m["T_cur_ref"], orc.make_features(m["px"], m["f"], m["grad"], m["type"], m["level"]),
                                          m["depth"], m["px_guess"], oopt)
        g = got[off:off + len(m["px"])]
        assert np.array_equal(g["result"], exp["result"])
        okm = exp["result"] == 0
        assert np.abs(g["px_cur"][okm] - exp["px_cur"][okm]).max() < PX_TOL
        off += len(m["px"])
