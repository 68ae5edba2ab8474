"""GPU test of the C++ host facades (svo_pro_universal_b200/host/svo_b200.h): the reference-shaped classes
svo::SparseImgAlign, svo::Matcher, svo::DepthFilter, svo::FastDetector and frame_utils::createImgPyramid, driven by
tests/cpp/facade_driver.cpp, give the oracle's results on a synthetic frame pair."""
import os
import subprocess

import numpy as np
import pytest

from helpers import oracle_align, pose_diff
from svo_pro_universal_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _build_drivers():
    """The facade library and the C++ drivers are built once per module (no-ops when they travelled with the snapshot)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "svo_pro_universal_b200", "host")], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], check=True)


def _T_f_w(T_cam_imu, T_imu_world):
    return synth.se3_mul(T_cam_imu, T_imu_world)


def test_cpp_facades_match_oracle(orc, tmp_path):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "svo_pro_universal_b200", "host")], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], check=True)
    d = synth.make_align_pair(3, n_features=120)
    N = len(d["px"])
    rng = np.random.default_rng(9)
    level = rng.integers(0, 3, N).astype(np.int32)
    ftype = np.where(rng.uniform(size=N) < 0.25, synth.K_EDGELET, synth.K_CORNER).astype(np.int32)
    grad = rng.normal(size=(N, 2)); grad /= np.linalg.norm(grad, axis=1)[:, None]
    X = d["f"] * d["depth"][:, None]
    R, t = synth.se3_to_Rt(d["T_cur_ref_gt"])
    guess = synth.cam_project(d["cam"], X @ R.T + t) + rng.uniform(-1.5, 1.5, (N, 2))
    T_ref = _T_f_w(d["T_cam_imu"], d["T_imu_world_ref"])
    T_cur_init = _T_f_w(d["T_cam_imu"], d["T_imu_world_cur_init"])
    T_cur_true = synth.se3_mul(d["T_cur_ref_gt"], T_ref)
    cam = d["cam"]
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        np.array([752, 480, 5, N], np.int32).tofile(f)
        d["ref_img"].tofile(f); d["cur_img"].tofile(f)
        np.array([cam[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")], np.float64).tofile(f)
        np.array([cam["width"], cam["height"], cam["distortion"]], np.int32).tofile(f)
        for T in (d["T_cam_imu"], T_ref, T_cur_init, T_cur_true):
            np.asarray(T, np.float64).tofile(f)
        for a in (d["px"], d["f"], d["depth"], grad, guess):
            np.ascontiguousarray(a, np.float64).tofile(f)
        ftype.tofile(f); level.tofile(f)
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "facade_driver"), str(fin), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(fout, np.float64)
    p = 0

    # (a) FastDetector::detect == oracle FastDetector::detect (sets compared: std::sort is unstable on ties)
    n_det = int(out[p]); p += 1
    det = out[p:p + 4 * n_det].reshape(n_det, 4); p += 4 * n_det
    cs = out[p]; p += 1
    px_o, sc_o, lv_o = np.zeros((416, 2)), np.zeros(416), np.zeros(416, np.int32)
    n_o = orc.lib().orc_fast_detect_features(d["ref_img"].ctypes.data_as(orc.u8p), 752, 480, 5, -1, 10.0, 8, 0, 2, 30, None, 416,
                                             px_o.ctypes.data_as(orc.f64p), sc_o.ctypes.data_as(orc.f64p), lv_o.ctypes.data_as(orc.i32p))
    assert n_det == n_o > 100
    exp = sorted(zip(px_o[:n_o, 0], px_o[:n_o, 1], sc_o[:n_o], lv_o[:n_o]))
    assert sorted(map(tuple, det)) == [tuple(map(float, e)) for e in exp]
    assert (np.diff(det[:, 2]) <= 0).all()  # sorted by score, best first
    pyr = orc.create_img_pyramid(d["ref_img"], 5)
    cs_o = sum(float((im.astype(np.float64) * ((np.arange(im.shape[1])[None, :] + 3 * np.arange(im.shape[0])[:, None]) % 7 + 1)).sum()) for im in pyr)
    assert cs == cs_o  # host mirror of the GPU pyramid is bit-exact

    # (a') FastGradDetector / GradientDetectorGrid ::detect == the oracle's (identical to the reference's own compiled detectors)
    for det_type in (orc.DETECTOR_FAST_GRAD, orc.DETECTOR_GRID_GRAD):
        n_det = int(out[p]); p += 1
        det = out[p:p + 8 * n_det].reshape(n_det, 8); p += 8 * n_det
        o = orc.detect_features(det_type, pyr)
        assert n_det == len(o["score"]) > 300
        got = {"px": det[:, 0:2], "score": det[:, 2], "level": det[:, 3].astype(np.int32), "type": det[:, 4].astype(np.int32), "grad": det[:, 5:7]}
        from helpers import assert_features_equal
        assert_features_equal(got, o, f"facade detector {det_type}")
        assert (det[:, 7] > 0.5).all()  # unit bearing vectors computed for the new features
        if det_type == orc.DETECTOR_FAST_GRAD:
            assert (got["type"] == 6).sum() > 5 and (got["type"] == 7).sum() > 100

    # (d4) DepthFilter::addKeyframe: the 200 best FastGrad features become seeds with the reference's initial state
    n_seed = int(out[p]); p += 1
    mu_range = out[p]; p += 1
    seeds = out[p:p + 9 * n_seed].reshape(n_seed, 9); p += 9 * n_seed
    n_after = int(out[p]); p += 1
    o = orc.detect_features(orc.DETECTOR_FAST_GRAD, pyr, max_n=200)
    assert n_seed == len(o["score"]) == 200 == n_after and mu_range == 1.0
    assert np.array_equal(seeds[:, 2], o["score"])
    cut = o["score"].min()
    keep = seeds[:, 2] > cut  # std::sort is unstable: entries at the cut score may differ
    assert sorted(map(tuple, seeds[keep][:, :4])) == sorted(zip(o["px"][o["score"] > cut, 0], o["px"][o["score"] > cut, 1], o["score"][o["score"] > cut],
                                                            o["level"][o["score"] > cut].astype(float)))
    assert set(seeds[:, 4]) <= {0.0, 1.0} and (seeds[:, 4] == 1.0).sum() > 100      # kCornerSeed / kEdgeletSeed
    assert np.array_equal(seeds[:, 5:9], np.tile([1.0 / 3.0, 1.0 / 36.0, 10.0, 10.0], (n_seed, 1)))

    # (b) SparseImgAlign::run writes cur->T_f_w_
    n_tracked = int(out[p]); p += 1
    T_f_w = out[p:p + 7]; p += 7
    chi2 = out[p]; p += 1
    o = oracle_align(orc, d, orc.default_align_options())
    assert n_tracked == o.n_tracked
    dq, dt = pose_diff(T_f_w, o.T_f_w[0])
    assert dq < 1e-4 and dt < 1e-4
    np.testing.assert_allclose(chi2, o.chi2, rtol=1e-4)

    # (c) Matcher::findMatchDirect / findEpipolarMatchDirect per feature
    m = out[p:p + 7 * N].reshape(N, 7); p += 7 * N
    keep = []
    rf = orc.make_frame(pyr, cam, keep=keep)
    cf = orc.make_frame(orc.create_img_pyramid(d["cur_img"], 5), cam, keep=keep)
    oft = orc.make_features(d["px"], d["f"], grad, ftype, level)
    e1 = orc.find_match_direct_batch(rf, cf, d["T_cur_ref_gt"], oft, d["depth"], guess, orc.default_matcher_options())
    inv = 1.0 / d["depth"]
    e2 = orc.find_epipolar_match_direct_batch(rf, cf, d["T_cur_ref_gt"], oft, np.stack([inv, 1.3 * inv, 0.7 * inv], 1), orc.default_matcher_options())
    assert np.array_equal(m[:, 0].astype(int), e1["result"]) and np.array_equal(m[:, 3].astype(int), e2["result"])
    ok1, ok2 = e1["result"] == 0, e2["result"] == 0
    assert ok1.sum() > N // 2 and ok2.sum() > N // 4
    assert np.abs(m[ok1, 1:3] - e1["px_cur"][ok1]).max() < 1e-3
    assert np.abs(m[ok2, 5:7] - e2["px_cur"][ok2]).max() < 1e-3
    np.testing.assert_allclose(m[ok2, 4], e2["depth"][ok2], rtol=1e-4)

    # (d) DepthFilter::updateSeeds
    n_succ = int(out[p]); p += 1
    sd = out[p:p + 5 * N].reshape(N, 5); p += 5 * N
    # (e) two host threads (own contexts / streams) racing into the lazy device upload of shared frames: same results as (c)
    assert out[p] == 0.0, f"{int(out[p])} results differ between the two-thread and the sequential run"
    p += 1
    # (f) fast:: leaves with list-shaped results == the oracle's (== the reference's own, tests/test_oracle_cpu.py)
    nc = int(out[p]); p += 1
    lst = out[p:p + 3 * nc].reshape(nc, 3); p += 3 * nc
    n_nm = int(out[p]); p += 1
    nm = out[p:p + n_nm].astype(int); p += n_nm
    n9 = int(out[p]); p += 1
    appended = out[p]; p += 1
    oxy = orc.fast_detect(d["ref_img"], 10, 10)
    osc = orc.fast_score10(d["ref_img"], oxy, 10)
    assert nc == len(oxy) > 1000 and np.array_equal(lst[:, :2].astype(np.int16), oxy) and np.array_equal(lst[:, 2].astype(np.int32), osc)
    assert np.array_equal(nm, orc.fast_nonmax3x3(oxy, osc))
    assert n9 == len(orc.fast_detect(d["ref_img"], 10, 9)) > nc and appended == 1.0
    # Matcher members (patch_, patch_with_border_, epi_image_) and scanEpipolarLine on its own
    n_chk = int(out[p]); p += 1
    rec = out[p:p + n_chk * 181].reshape(n_chk, 181); p += n_chk * 181
    inv = 1.0 / d["depth"][:n_chk]
    oft_c = orc.make_features(d["px"][:n_chk], d["f"][:n_chk], grad[:n_chk], ftype[:n_chk], level[:n_chk])
    e = orc.find_epipolar_match_direct_batch(rf, cf, d["T_cur_ref_gt"], oft_c, np.stack([inv, 1.3 * inv, 0.7 * inv], 1), orc.default_matcher_options())
    assert np.array_equal(rec[:, 0].astype(int), e["result"])
    np.testing.assert_allclose(rec[:, 1:3], e["epi_image"], rtol=1e-9, atol=1e-9)
    warped = (e["result"] != 4) & (e["result"] != 7)   # kFailWarp / kFailAngle return before the patch exists
    assert warped.sum() > n_chk // 2
    assert np.array_equal(rec[warped, 5:105].astype(np.uint8), e["patch_with_border"][warped])
    assert np.array_equal(rec[warped, 105:169].astype(np.uint8).reshape(-1, 8, 8), e["patch_with_border"][warped].reshape(-1, 10, 10)[:, 1:9, 1:9])
    n_scanned = 0
    for i in np.flatnonzero(warped):
        A, B, Cc = rec[i, 169:178:3], rec[i, 170:178:3], rec[i, 171:178:3]
        if rec[i, 4] < 1e-9:
            continue  # a zero-length segment has no scan steps
        best, z = orc.scan_epipolar_line(cf, A, B, Cc, rec[i, 105:169].astype(np.uint8), int(rec[i, 3]), rec[i, 4], orc.default_matcher_options())
        assert z == int(rec[i, 180]) and np.abs(best - rec[i, 178:180]).max() < 1e-9, i
        n_scanned += 1
    assert n_scanned > n_chk // 2
    okg = out[p]; stg = out[p + 1:p + 5]; p += 5
    sto = np.array([0.31, 0.004, 10.0, 10.0])
    assert okg == orc.lib().orc_update_filter_gaussian(0.33, 0.0007, orc._f64(sto))
    np.testing.assert_allclose(stg, sto, rtol=1e-12)
    assert out[p] == 1.0, "setPatchSize(8) must be refused"
    p += 1
    assert list(out[p:p + 4]) == [26.0 * 16.0, 51.0 * 32.0, 1.0, 0.0]   # grid 30 px cells, closeness grid 15 px cells; fill; resetGrid
    p += 4
    assert list(out[p:p + 2]) == [1.0, 1.0]
    p += 2
    # (g) DepthFilter with its parallel thread: updateSeeds returned 0 at once, the worker's results equal the synchronous run's bit for
    # bit, and the keyframe job initialised the reference's 200 seeds
    assert list(out[p:p + 3]) == [0.0, 0.0, 200.0], list(out[p:p + 3])
    p += 3
    assert p == len(out)
    types = np.where(ftype == synth.K_EDGELET, synth.K_EDGELET_SEED, synth.K_CORNER_SEED).astype(np.uint8)
    st = np.tile(np.array([0.25, (1 / 1.5) ** 2 / 36.0, 10.0, 10.0]), (N, 1))
    n_o, _, _ = orc.update_seeds(rf, [cf], d["T_cur_ref_gt"][None], oft, types, st, 1 / 1.5, orc.default_matcher_options())
    assert n_succ == n_o > N // 4
    np.testing.assert_allclose(sd[:, :4], st, rtol=1e-4)
    assert np.array_equal(sd[:, 4].astype(np.uint8), types)


def test_cpp_reprojector_facade_matches_reference(tmp_path):
    """svo::Reprojector::reprojectFrames of the C++ facade (three svo_cuda_reproject_match passes + the host bookkeeping of
    reprojector.cpp:28-310) against the outputs of the REFERENCE's own compiled Reprojector::reprojectFrames
    (tests/golden/reproject_ref_golden.npz): same features in the same slots, same grid, statistics, trashed points, landmark
    counters and seed states."""
    import helpers
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "svo_pro_universal_b200", "host")], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], check=True)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reproject_ref_golden.npz"))
    for ci, (seed, max_n, unconv, ratio, min_req, rm) in enumerate(helpers.REPROJECT_FRAMES_CASES):
        sc = synth.make_reproject_scene(seed)
        t, cam = sc["tables"], sc["cam"]
        K, NF, NP, NO = t["n_kfs"], t["n_feat"], t["n_points"], t["n_obs"]
        fin, fout = tmp_path / f"rin{ci}.bin", tmp_path / f"rout{ci}.bin"
        with open(fin, "wb") as f:
            np.array([752, 480, 5, K, K, NF, NP, NO, max_n, unconv, min_req, rm], np.int32).tofile(f)
            np.array([ratio], np.float64).tofile(f)
            np.array([cam[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")], np.float64).tofile(f)
            np.array([cam["width"], cam["height"], cam["distortion"]], np.int32).tofile(f)
            for im in sc["kf_imgs"]:
                im.tofile(f)
            sc["cur_img"].tofile(f)
            np.ascontiguousarray(t["kf_T_f_w"], np.float64).tofile(f)
            np.ascontiguousarray(sc["cur_T_f_w"], np.float64).tofile(f)
            np.ascontiguousarray(t["kf_seed_mu_range"], np.float64).tofile(f)
            np.ascontiguousarray(t["kf_feat_begin"], np.int32).tofile(f)
            for k in ("px", "f", "grad"):
                np.ascontiguousarray(t["feat"][k], np.float64).tofile(f)
            for k in ("type", "level"):
                np.ascontiguousarray(t["feat"][k], np.int32).tofile(f)
            np.ascontiguousarray(t["feat_score"], np.float64).tofile(f)
            np.ascontiguousarray(t["feat_seed_state"], np.float64).tofile(f)
            np.ascontiguousarray(t["feat_point"], np.int32).tofile(f)
            np.ascontiguousarray(t["pt_pos"], np.float64).tofile(f)
            for k in ("pt_n_failed", "pt_n_succeeded", "pt_obs_begin", "obs_feat"):
                np.ascontiguousarray(t[k], np.int32).tofile(f)
        r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "reproject_driver"), str(fin), str(fout)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out = np.fromfile(fout, np.float64)
        n = int(out[0]); p = 1
        rows = out[p:p + 16 * n].reshape(n, 16); p += 16 * n
        n_cells = int(out[p]); p += 1
        occ = out[p:p + n_cells]; p += n_cells
        stats = out[p:p + 3]; p += 3
        ptc = out[p:p + 2 * NP].reshape(NP, 2); p += 2 * NP
        fs = out[p:p + 5 * NF].reshape(NF, 5); p += 5 * NF
        g = {k: gold[f"rf{ci}_{k}"] for k in helpers.REPROJ_FRAMES_KEYS}
        assert n == len(g["type"]) > 100, (ci, n, len(g["type"]))
        assert np.array_equal(rows[:, 0].astype(int), g["type"]), ci
        assert np.abs(rows[:, 1:3] - g["px"]).max() < 1e-3, ci                     # north_star: 1e-3 px
        assert np.array_equal(rows[:, 3].astype(int), g["level"]), ci
        assert np.array_equal(rows[:, 4].astype(int), g["point"]), ci
        assert np.array_equal(rows[:, 5].astype(int), g["seed_feat"]), ci
        np.testing.assert_allclose(rows[:, 6:10], g["state"], rtol=1e-4, atol=1e-12)   # north_star: 1e-4 relative
        assert np.abs(rows[:, 10:13] - g["f"]).max() < 1e-5, ci
        assert np.abs(rows[:, 13:15] - g["grad"]).max() < 1e-9, ci
        assert np.array_equal(rows[:, 15], g["score"]), ci
        assert np.array_equal(occ.astype(np.uint8), g["occupancy"][:n_cells]), ci
        assert np.array_equal(stats.astype(int), g["stats"]), (ci, stats, g["stats"])
        assert np.array_equal(ptc.astype(int), g["pt_counters"]), ci
        np.testing.assert_allclose(fs[:, :4], g["feat_state"], rtol=1e-4, atol=1e-12)
        assert np.array_equal(fs[:, 4].astype(int), g["feat_type"]), ci


def test_cpp_pose_optimizer_facade_matches_reference(tmp_path):
    """svo::PoseOptimizer::run of the C++ facade against the outputs of the REFERENCE's own compiled PoseOptimizer::run
    (tests/golden/pose_opt_ref_golden.npz): pose, remaining measurements, outlier marks, MAD sigma, iterations."""
    import helpers
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "svo_pro_universal_b200", "host")], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], check=True)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "pose_opt_ref_golden.npz"))
    for ci, spec in enumerate(helpers.POSE_OPT_CASES):
        c, prior = helpers.pose_opt_case(spec)
        cam, N = c["cam"], len(c["px"])
        fin, fout = tmp_path / f"pin{ci}.bin", tmp_path / f"pout{ci}.bin"
        with open(fin, "wb") as f:
            np.array([spec[1], N, spec[2], int(prior is not None)], np.int32).tofile(f)
            np.array([cam[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")], np.float64).tofile(f)
            np.array([cam["width"], cam["height"], cam["distortion"]], np.int32).tofile(f)
            np.ascontiguousarray(np.stack(c["T_cam_imu"]), np.float64).tofile(f)
            np.ascontiguousarray(c["T_imu_world_init"], np.float64).tofile(f)
            np.ascontiguousarray(prior if prior is not None else [1.0, 0, 0, 0], np.float64).tofile(f)
            for k in ("px", "f", "grad", "xyz_world"):
                np.ascontiguousarray(c[k], np.float64).tofile(f)
            for k in ("level", "type", "feat_cam"):
                np.ascontiguousarray(c[k], np.int32).tofile(f)
            np.ascontiguousarray(c["has_xyz"], np.uint8).tofile(f)
        r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "pose_opt_driver"), str(fin), str(fout)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out = np.fromfile(fout, np.float64)
        assert int(out[0]) == int(gold[f"p{ci}_n"]), ci
        dq, dt = pose_diff(out[1:8], gold[f"p{ci}_T"])
        assert dq < 1e-4 and dt < 1e-4 and dq < 1e-9 and dt < 1e-9, (ci, dq, dt)
        np.testing.assert_allclose(out[8:11], gold[f"p{ci}_stats"][:3], rtol=1e-6)
        assert int(out[11]) == int(gold[f"p{ci}_stats"][3])
        assert np.array_equal(out[12:12 + N].astype(np.uint8), gold[f"p{ci}_outlier"]), ci


def test_stereo_triangulation_facade_matches_reference(orc, tmp_path):
    """svo::StereoTriangulation::compute through makeDetector + the facade (C++ driver) leaves in both frames what the REFERENCE's
    own compiled compute() left there for the same srand seed (tests/golden/stereo_tri_ref_golden.npz)."""
    import helpers
    g = np.load(os.path.join(ROOT, "tests", "golden", "stereo_tri_ref_golden.npz"))
    for i, case in enumerate(helpers.STEREO_TRI_CASES):
        d, s1 = helpers.stereo_case(case[0])
        cam = d["cam"]
        fin, fout = tmp_path / f"st_in_{i}.bin", tmp_path / f"st_out_{i}.bin"
        with open(fin, "wb") as f:
            np.array([5, case[2], case[3], case[1], 752, 480], np.int32).tofile(f)
            np.array([cam[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")], np.float64).tofile(f)
            np.array([cam["width"], cam["height"], cam["distortion"]], np.int32).tofile(f)
            for T in (d["T_cam_imu"], s1["T_cam_imu"], d["T_imu_world_ref"]):
                np.asarray(T, np.float64).tofile(f)
            np.array(case[4:7], np.float64).tofile(f)
            d["ref_img"].tofile(f); s1["ref_img"].tofile(f)
        r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "stereo_tri_driver"), str(fin), str(fout)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out = np.fromfile(fout, np.float64)
        n0, n1 = int(out[0]), int(out[1])
        assert n0 == int(g[f"n0_{i}"]) and n1 == int(g[f"n1_{i}"]), (i, n0, n1)
        a = out[2:2 + 4 * n0].reshape(n0, 4)
        b = out[2 + 4 * n0:2 + 4 * n0 + 15 * n1].reshape(n1, 15)
        opt = out[2 + 4 * n0 + 15 * n1:].reshape(n1, 4)
        assert np.array_equal(a[:, :2], g[f"px0_{i}"]) and np.array_equal(a[:, 2], g[f"type0_{i}"])
        assert np.array_equal(b[:, 13].astype(np.int32), g[f"ref_index1_{i}"])            # the same frame0 features in the same slots
        assert np.array_equal(np.flatnonzero(a[:, 3]), np.sort(g[f"ref_index1_{i}"]))     # ... and they own the new landmarks
        assert (b[:, 14] == 2).all()                                                     # observed by both frames
        assert np.array_equal(b[:, 7], g[f"level1_{i}"]) and np.array_equal(b[:, 8], g[f"type1_{i}"]) and np.array_equal(b[:, 9], g[f"score1_{i}"])
        np.testing.assert_allclose(b[:, 0:2], g[f"px1_{i}"], rtol=0, atol=1e-3)
        np.testing.assert_allclose(b[:, 2:5], g[f"f1_{i}"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(b[:, 5:7], g[f"grad1_{i}"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(b[:, 10:13], g[f"xyz1_{i}"], rtol=1e-4, atol=1e-6)
        # optimizeStructure: every landmark refined from its two observations == the oracle's Point::optimize (bit-equal to the
        # reference's compiled point.cpp, tests/test_point_optimizer_cpu.py); frame1 is visited last
        T0, T1 = synth.se3_mul(d["T_cam_imu"], d["T_imu_world_ref"]), synth.se3_mul(s1["T_cam_imu"], d["T_imu_world_ref"])
        f0 = synth.cam_backproject(cam, a[b[:, 13].astype(int), :2])
        f0 /= np.linalg.norm(f0, axis=1, keepdims=True)
        for k in range(0, n1, 7):
            # frame0's pass optimises the point first, frame1's pass (non-edgelets again) a second time from that result
            p1, _ = orc.point_optimize(np.stack([T0, T1]), np.stack([f0[k], b[k, 2:5]]), b[k, 10:13], 5)
            edgelet = int(b[k, 8]) == 6
            p2 = p1 if edgelet else orc.point_optimize(np.stack([T0, T1]), np.stack([f0[k], b[k, 2:5]]), p1, 5)[0]
            want = b[k, 10:13] if edgelet else p2
            np.testing.assert_allclose(opt[k, :3], want, rtol=0, atol=1e-9)
        assert (opt[b[:, 8] != 6, 3] == 2).all()


def test_feature_tracker_facade_matches_reference(tmp_path):
    """svo::FeatureTracker::trackAndDetect through the facade (C++ driver; every active track of a frame in one svo_cuda_align_pyr2d
    call) leaves in every frame of three sequences what the REFERENCE's own compiled tracker left there
    (tests/golden/tracker_ref_golden.npz): plain tracking, reset + re-detection, detection on top of live tracks."""
    import helpers
    g = np.load(os.path.join(ROOT, "tests", "golden", "tracker_ref_golden.npz"))
    for i, case in enumerate(helpers.TRACKER_CASES):
        d, imgs = helpers.tracker_sequence(case)
        cam = d["cam"]
        fin, fout = tmp_path / f"tr_in_{i}.bin", tmp_path / f"tr_out_{i}.bin"
        with open(fin, "wb") as f:
            np.array([case[1], case[2], case[3], int(case[4]), int(case[5]), 752, 480, 5], np.int32).tofile(f)
            np.array([cam[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")], np.float64).tofile(f)
            np.array([cam["width"], cam["height"], cam["distortion"]], np.int32).tofile(f)
            for im in imgs:
                np.ascontiguousarray(im, np.uint8).tofile(f)
        r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "tracker_driver"), str(fin), str(fout)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out = np.fromfile(fout, np.float64)
        p = 0
        for k in range(case[1]):
            n = int(out[p]); p += 1
            cols = out[p:p + 4 * n].reshape(n, 4); p += 4 * n
            n_active, n_term, disp = out[p:p + 3]; p += 3
            assert n == len(g[f"px_{i}_{k}"]), (i, k, n)
            assert np.array_equal(cols[:, 2].astype(np.int64), g[f"track_id_{i}_{k}"]) and np.array_equal(cols[:, 3], g[f"score_{i}_{k}"])
            np.testing.assert_allclose(cols[:, :2], g[f"px_{i}_{k}"], rtol=0, atol=1e-3)   # KLT positions (bit-equal in practice)
            assert int(n_active) == int(g[f"n_active_{i}_{k}"]) and int(n_term) == int(g[f"n_terminated_{i}_{k}"])
            assert abs(disp - float(g[f"disparity_{i}_{k}"])) < 1e-3
