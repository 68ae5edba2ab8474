"""Shared helpers of the parity tests: build oracle frames and GPU pyramids from the same synthetic bytes."""
import numpy as np

from svo_pro_universal_b200 import capi, synth, batch  # noqa: F401


def pose_diff(Ta, Tb):
    """(rotation angle [rad], translation distance [m]) between two 7-vector transformations."""
    Ta, Tb = np.asarray(Ta, float), np.asarray(Tb, float)
    dq = 2 * np.arccos(min(1.0, abs(float(np.dot(Ta[:4], Tb[:4])) / (np.linalg.norm(Ta[:4]) * np.linalg.norm(Tb[:4])))))
    return dq, float(np.linalg.norm(Ta[4:] - Tb[4:]))


def oracle_align(orc, d, opt, n_levels=5, keep=None):
    keep = [] if keep is None else keep
    rp = orc.create_img_pyramid(d["ref_img"], n_levels)
    cp = orc.create_img_pyramid(d["cur_img"], n_levels)
    rf = orc.make_frame(rp, d["cam"], d["T_cam_imu"], d["T_imu_world_ref"], d["px"], d["f"], d["depth"], d["eligible"], keep=keep)
    cf = orc.make_frame(cp, d["cam"], d["T_cam_imu"], d["T_imu_world_cur_init"], keep=keep)
    return orc.sparse_align([rf], [cf], opt)


def to_orc_options(orc, gopt, prior=None):
    """Mirror a capi.SparseAlignOptions (+ optional prior row) into the oracle's option struct."""
    o = orc.default_align_options(
        max_level=gopt.max_level, min_level=gopt.min_level,
        estimate_illumination_gain=gopt.estimate_illumination_gain,
        estimate_illumination_offset=gopt.estimate_illumination_offset,
        use_distortion_jacobian=gopt.use_distortion_jacobian, robustification=gopt.robustification,
        weight_scale=gopt.weight_scale, max_iter=gopt.max_iter, eps=gopt.eps, alpha_init=gopt.alpha_init,
        beta_init=gopt.beta_init, lambda_rot=gopt.lambda_rot, lambda_trans=gopt.lambda_trans,
        lambda_alpha=gopt.lambda_alpha, lambda_beta=gopt.lambda_beta)
    if prior is not None:
        o.have_prior = 1
        o.prior_T[:] = list(prior["T"])
        o.prior_alpha, o.prior_beta = float(prior["alpha"]), float(prior["beta"])
    return o


def gpu_align(ctx, pairs, gopt, priors=None, n_levels=5):
    pk = batch.pack_align_batch(pairs)
    B = len(pairs)
    h, w = pairs[0]["ref_img"].shape
    ref = capi.Pyramid(ctx, B, w, h, n_levels)
    cur = capi.Pyramid(ctx, B, w, h, n_levels)
    ref.upload(pk["ref_imgs"]); cur.upload(pk["cur_imgs"])
    ref.build(); cur.build()
    res = capi.sparse_align(ctx, [ref], [cur], [capi.Camera.from_dict(pairs[0]["cam"])], pk["T_cam_imu"], pk["T_imu_world_ref"],
                            pk["T_imu_world_cur"], pk["n_features"], pk["px"], pk["f"], pk["depth"], pk["eligible"], gopt,
                            priors=priors)
    return res, ref, cur


# ---- seeded cases for the rows pinned by the compiled reference (oracle/_ref/libdirect_ref.so) ------------------------------
PYR_SHAPES = ((752, 480, 5), (640, 480, 4), (94, 60, 3), (47, 30, 2), (100, 75, 3), (33, 17, 2))
N_ALIGN_CASES = 160


def align_cases(seed=7, n=N_ALIGN_CASES):
    """n seeded align2D / align1D problems on one synthetic 752x480 image: the 10x10 reference patch is cut at an integer
    position of the same image and the start is displaced by up to 1.5 px (some cases sit at the image border so the
    `break` paths run); affine flags and iteration counts vary."""
    rng = np.random.default_rng(seed)
    img = synth.make_image(seed)
    cases = []
    for i in range(n):
        if i % 16 == 15:   # start closer than 4 px to the border -> immediate break, not converged
            x, y = int(rng.integers(6, 740)), 6
            px0 = (x + 0.3, 3.5)
        else:
            x, y = int(rng.integers(12, 740)), int(rng.integers(12, 468))
            px0 = (x + rng.uniform(-1.5, 1.5), y + rng.uniform(-1.5, 1.5))
        th = rng.uniform(0, 2 * np.pi)
        cases.append(dict(pwb=img[y - 5:y + 5, x - 5:x + 5].copy(), px0=np.array(px0), dir=np.array([np.cos(th), np.sin(th)]),
                          n_iter=(10, 10, 3, 30)[i % 4], est_offset=bool((i // 2) % 2 == 0), est_gain=bool(i % 8 == 5)))
    return img, cases


def direct_outputs(orc, which):
    """Every pinned function evaluated through `which` ("orc" = the restatement, "ref" = the compiled reference)."""
    import hashlib
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    out = {}
    for w, h, nl in PYR_SHAPES:  # a1: pyramid bytes (752 -> SSE2 formula, 94/47/100/33 -> truncating mean, mixed chains)
        img = synth.make_image(100 + w, w, h, n_rect=max(8, w * h // 400))
        pyr = orc.create_img_pyramid(img, nl) if which == "orc" else orc.ref_create_img_pyramid(img, nl)
        out[f"pyr_sha_{w}x{h}"] = np.array([sha(p) for p in pyr])
    img, cases = align_cases()
    a2 = np.zeros((len(cases), 3)); a1 = np.zeros((len(cases), 4))
    for i, c in enumerate(cases):  # c3 / c4
        ok, p = orc.align2d(img, c["pwb"], c["px0"], c["n_iter"], c["est_offset"], c["est_gain"], which=which)
        a2[i] = (ok, p[0], p[1])
        ok, p, hinv = orc.align1d(img, c["dir"], c["pwb"], c["px0"], c["n_iter"], c["est_offset"], c["est_gain"], which=which)
        a1[i] = (ok, p[0], p[1], hinv)
    out["align2d"], out["align1d"] = a2, a1
    rng = np.random.default_rng(11)  # c2 / c5
    xy = np.stack([rng.integers(0, 744, 400), rng.integers(0, 472, 400)], 1)
    out["zmssd"] = np.stack([orc.zmssd(cases[k]["pwb"][1:9, 1:9], img, xy, which) for k in range(4)])
    out["patch_from_border"] = np.stack([orc.patch_from_patch_with_border(cases[k]["pwb"], which) for k in range(4)])
    err = np.concatenate([rng.normal(0, 3, 500), [0.0, 4.6851, -4.6851, 4.68509, 100.0]]).astype(np.float32)  # b6
    out["tukey"] = orc.tukey_weight(err, which=which)
    k = (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)  # s1: EuRoC cam0 radtan
    pts = rng.uniform(-0.9, 0.9, (300, 2))
    out["radtan_distort"] = orc.radtan(k, pts, "distort", which)
    out["radtan_undistort"] = orc.radtan(k, pts, "undistort", which)
    out["radtan_jacobian"] = orc.radtan(k, pts, "jacobian", which)
    states = np.abs(rng.normal(0.5, 0.3, (50, 4))) + 1e-3  # d1: seed helpers
    out["seed_helpers"] = np.stack([orc.seed_helpers(s, 1.0 / 1.5, (200.0, 500.0)[j % 2], 1.0 / s[0], 0.01 * (1 + j % 5), which)
                                    for j, s in enumerate(states)])
    gxy = np.stack([rng.integers(0, 752, 600), rng.integers(0, 480, 600)], 1).astype(np.int32)  # a5: cell index
    lvl = rng.integers(0, 3, 600)
    gxy = (gxy >> lvl[:, None]).astype(np.int32)
    out["grid_cells"] = orc.grid_cell_index(30, 26, 16, gxy, (1 << lvl).astype(np.int32), which)
    return out


# ---- rows b / c1 / c6-c7 / d1-d3 through the restatement ("orc") or the compiled reference front-end ("ref") ---------------
ALIGN_OPTION_SETS = (dict(), dict(estimate_illumination_gain=1, estimate_illumination_offset=1), dict(robustification=1),
                     dict(estimate_illumination_gain=1, estimate_illumination_offset=1, robustification=1),
                     dict(max_level=4, min_level=2, max_iter=5), dict(alpha_init=0.02, beta_init=-1.5))


def stereo_case(seed=81):
    """Two-camera bundle (11 cm baseline) around synth.make_align_pair(seed): per-camera images, features and T_cam_imu."""
    d = synth.make_align_pair(seed)
    T_c1_c0 = synth.se3_exp_small(np.zeros(3), np.array([-0.11, 0.0, 0.0]))
    scene = d["scene"]
    ref1 = scene.render(T_c1_c0)
    cur1 = scene.render(synth.se3_mul(T_c1_c0, d["T_cur_ref_gt"]))
    px1 = synth.pick_features(ref1, 150, 5)
    f1 = synth.cam_backproject(d["cam"], px1)
    R01, t01 = synth.se3_to_Rt(synth.se3_inv(T_c1_c0))
    lam = (scene.d - scene.n @ t01) / ((f1 @ R01.T) @ scene.n)
    X1 = f1 * lam[:, None]
    depth1 = np.linalg.norm(X1, axis=1)
    return d, dict(ref_img=ref1, cur_img=cur1, px=px1, f=X1 / depth1[:, None], depth=depth1, T_cam_imu=synth.se3_mul(T_c1_c0, d["T_cam_imu"]))


def _align_arrays(r):
    return np.concatenate([np.array(r.T_icur_iref), [r.alpha if hasattr(r, "alpha") else 0.0, r.beta if hasattr(r, "beta") else 0.0, r.chi2,
                                                    r.n_tracked]]), np.array(r.H), np.array([list(t) for t in r.T_f_w])


def scan_cases(orc, ms, rf, cf, oft, d3):
    """Inputs of stand-alone Matcher::scanEpipolarLine calls: the segment end points A, B and the estimate C of every feature of the
    match set whose epipolar segment is long enough to be scanned (matcher.cpp:170-176, 222-224), with the search level, the warped
    8x8 patch and epi_length_pyramid_ the restatement's findEpipolarMatchDirect leaves (always the restatement, so that the inputs do
    not depend on the implementation under test)."""
    r = orc.find_epipolar_match_direct_batch(rf, cf, ms["T_cur_ref"], oft, d3, orc.default_matcher_options(), n_threads=8)
    R, t = synth.se3_to_Rt(np.asarray(ms["T_cur_ref"], np.float64))
    Rf = ms["f"] @ R.T
    A, B, Cc = Rf + t[None] * d3[:, 1:2], Rf + t[None] * d3[:, 2:3], Rf + t[None] * d3[:, 0:1]
    sel = np.flatnonzero((r["epi_length_pyramid"] >= 2.0) & (r["result"] != 4) & (r["result"] != 7))
    pwb = r["patch_with_border"][sel].reshape(-1, 10, 10)
    return {"A": A[sel], "B": B[sel], "C": Cc[sel], "patch": np.ascontiguousarray(pwb[:, 1:9, 1:9]).reshape(-1, 64),
            "level": r["search_level"][sel].astype(np.int32), "epi_length": r["epi_length_pyramid"][sel], "sel": sel}


def frontend_outputs(orc, which):
    align = orc.sparse_align if which == "orc" else orc.ref_sparse_align
    fmd = (lambda *a: orc.find_match_direct_batch(*a, n_threads=8)) if which == "orc" else orc.ref_find_match_direct_batch
    epi = (lambda *a: orc.find_epipolar_match_direct_batch(*a, n_threads=8)) if which == "orc" else orc.ref_find_epipolar_match_direct_batch
    out, keep = {}, []
    # b: SparseImgAlign::run, mono, every option set; radtan camera with the distortion Jacobian; weighted prior; stereo bundle
    rows, Hs = [], []
    for seed in (1, 2, 3):
        d = synth.make_align_pair(seed)
        rp, cp = orc.create_img_pyramid(d["ref_img"], 5), orc.create_img_pyramid(d["cur_img"], 5)
        rf = orc.make_frame(rp, d["cam"], d["T_cam_imu"], d["T_imu_world_ref"], d["px"], d["f"], d["depth"], d["eligible"], keep=keep)
        cf = orc.make_frame(cp, d["cam"], d["T_cam_imu"], d["T_imu_world_cur_init"], keep=keep)
        for kw in ALIGN_OPTION_SETS:
            a, H, _ = _align_arrays(align([rf], [cf], orc.default_align_options(**kw)))
            rows.append(a); Hs.append(H)
        o = orc.default_align_options(lambda_rot=0.5, lambda_trans=0.1)
        o.have_prior = 1
        o.prior_T[:] = list(synth.se3_mul(d["T_icur_iref_gt"], synth.se3_exp_small(np.array([1e-3, -2e-3, 1e-3]), np.array([2e-3, 0, -1e-3]))))
        a, H, _ = _align_arrays(align([rf], [cf], o))
        rows.append(a); Hs.append(H)
    d = synth.make_align_pair(5, cam=synth.EUROC_CAM_RADTAN)
    rf = orc.make_frame(orc.create_img_pyramid(d["ref_img"], 5), d["cam"], d["T_cam_imu"], d["T_imu_world_ref"], d["px"], d["f"], d["depth"], d["eligible"], keep=keep)
    cf = orc.make_frame(orc.create_img_pyramid(d["cur_img"], 5), d["cam"], d["T_cam_imu"], d["T_imu_world_cur_init"], keep=keep)
    for kw in (dict(), dict(use_distortion_jacobian=1)):
        a, H, _ = _align_arrays(align([rf], [cf], orc.default_align_options(**kw)))
        rows.append(a); Hs.append(H)
    out["align_rows"], out["align_H"] = np.array(rows), np.array(Hs)
    d0, c1 = stereo_case()
    rfs = [orc.make_frame(orc.create_img_pyramid(d0["ref_img"], 5), d0["cam"], d0["T_cam_imu"], d0["T_imu_world_ref"], d0["px"], d0["f"], d0["depth"], keep=keep),
           orc.make_frame(orc.create_img_pyramid(c1["ref_img"], 5), d0["cam"], c1["T_cam_imu"], d0["T_imu_world_ref"], c1["px"], c1["f"], c1["depth"], keep=keep)]
    cfs = [orc.make_frame(orc.create_img_pyramid(d0["cur_img"], 5), d0["cam"], d0["T_cam_imu"], d0["T_imu_world_cur_init"], keep=keep),
           orc.make_frame(orc.create_img_pyramid(c1["cur_img"], 5), d0["cam"], c1["T_cam_imu"], d0["T_imu_world_cur_init"], keep=keep)]
    a, H, Tfw = _align_arrays(align(rfs, cfs, orc.default_align_options(estimate_illumination_gain=1, estimate_illumination_offset=1)))
    out["stereo_row"], out["stereo_T_f_w"] = a, Tfw[:2]
    # c: findMatchDirect / findEpipolarMatchDirect
    ms = synth.make_match_set(7, n_features=240)
    ms["px"][:8] = np.array([[3.0, 3.0]]) + np.arange(8)[:, None] * 0.25   # kFailVisibility
    ms["depth"][8:16] = 0.05                                              # warp / alignment failures
    rf = orc.make_frame(orc.create_img_pyramid(ms["ref_img"], 5), ms["cam"], keep=keep)
    cf = orc.make_frame(orc.create_img_pyramid(ms["cur_img"], 5), ms["cam"], keep=keep)
    oft = orc.make_features(ms["px"], ms["f"], ms["grad"], ms["type"], ms["level"])
    fields = ("result", "px_cur", "f_cur", "search_level", "A_cur_ref", "h_inv", "epi_length_pyramid", "depth", "patch_with_border")
    for name, kw in (("default", dict()), ("gain", dict(affine_est_gain=1))):
        r = fmd(rf, cf, ms["T_cur_ref"], oft, ms["depth"], ms["px_guess"], orc.default_matcher_options(**kw))
        for k in fields:
            out[f"fmd_{name}_{k}"] = r[k]
    d_inv = 1.0 / ms["depth"]
    d3 = np.stack([d_inv * np.random.default_rng(1).uniform(0.9, 1.1, len(d_inv)), d_inv * 1.5, d_inv * 0.6], 1)
    for name, kw in (("sphere", dict()), ("plane", dict(scan_on_unit_sphere=0)), ("a1d", dict(align_1d=1)), ("nosub", dict(subpix_refinement=0))):
        r = epi(rf, cf, ms["T_cur_ref"], oft, d3, orc.default_matcher_options(**kw))
        for k in fields + ("epi_image", "reject"):
            out[f"epi_{name}_{k}"] = r[k]
    # c7: Matcher::scanEpipolarLine on its own (both scan variants, a capped scan, a caller-supplied starting score)
    sc = scan_cases(orc, ms, rf, cf, oft, d3)
    for name, kw, z0 in (("sphere", dict(), 2000 * 64), ("plane", dict(scan_on_unit_sphere=0), 2000 * 64),
                         ("capped", dict(max_epi_search_steps=4), 2000 * 64), ("low_start", dict(), 9000)):
        o = orc.default_matcher_options(**kw)
        res = [orc.scan_epipolar_line(cf, sc["A"][i], sc["B"][i], sc["C"][i], sc["patch"][i], sc["level"][i], sc["epi_length"][i], o, z0,
                                      which=("orc" if which == "orc" else "ref")) for i in range(len(sc["A"]))]
        out[f"scan_{name}_px"] = np.array([r[0] for r in res])
        out[f"scan_{name}_zmssd"] = np.array([r[1] for r in res], np.int32)
    # d: updateSeed chain over ordered observations
    sq = synth.make_seed_sequence(17, n_seeds=160, n_obs=6)
    rf = orc.make_frame(orc.create_img_pyramid(sq["ref_img"], 5), sq["cam"], keep=keep)
    cfs = [orc.make_frame(orc.create_img_pyramid(im, 5), sq["cam"], keep=keep) for im in sq["cur_imgs"]]
    oft = orc.make_features(sq["px"], sq["f"], sq["grad"], sq["type"].astype(np.int32), sq["level"])
    for name, kw in (("vog", dict()), ("gauss", dict(use_vogiatzis=0)), ("conv", dict(check_convergence=1, sigma2_thresh=50.0))):
        t, s = sq["type"].copy(), sq["state"].copy()
        if which == "orc":
            n, _, ok = orc.update_seeds(rf, cfs, sq["T_cur_ref"], oft, t, s, sq["mu_range"], orc.default_matcher_options(), **kw)
        else:
            n, ok = orc.ref_update_seeds(rf, cfs, sq["T_cur_ref"], oft, t, s, sq["mu_range"], orc.default_matcher_options(), **kw)
        out[f"seeds_{name}_types"], out[f"seeds_{name}_state"], out[f"seeds_{name}_ok"], out[f"seeds_{name}_n"] = t, s, ok, np.array(n)
    return out


# ---- f1: Reprojector candidate matching ------------------------------------------------------------------------------------------
IDENTITY7 = np.array([1.0, 0, 0, 0, 0, 0, 0])
# (scene seed, max_n_features, n_features_in, occupied-cell fraction, sort_by_num_obs, scene kwargs)
REPROJECT_CASES = [
    (3, 120, 0, 0.10, 0, {}),
    (3, 1000, 0, 0.10, 0, {}),                     # never reaches the cap: landmarks, converged and unconverged seeds are all tried
    (4, 60, 10, 0.30, 0, {}),
    (4, 0, 0, 0.0, 0, {}),                         # unlimited: occupancy ignored, every candidate tried
    (5, 150, 0, 0.05, 1, {}),                      # sortCandidatesByNumObs
    (6, 400, 200, 0.0, 0, dict(max_rot_deg=9.0, max_trans=0.5)),  # large motion: visibility / warp / alignment failures
    (7, 30, 40, 0.0, 0, {}),                       # frame already holds more than max_n: exactly one more match is taken
]
REPROJ_INT_FIELDS = ("status", "order", "slot", "level", "type_out", "d_failed", "d_succeeded")
REPROJ_FLOAT_FIELDS = ("cur_px", "px", "f", "grad", "seed_state")


def reproject_px_error_angle(cam):
    return float(np.arctan(1.0 / (2.0 * cam["fx"])) + np.arctan(1.0 / (2.0 * cam["fy"])))


def reproject_case_inputs(case):
    seed, max_n, n_in, occ_frac, by_obs, kw = case
    sc = synth.make_reproject_scene(seed, **kw)
    n_cells = ((sc["cam"]["width"] + 29) // 30) * ((sc["cam"]["height"] + 29) // 30)
    occ = (np.random.default_rng(seed + 5).uniform(size=n_cells) < occ_frac).astype(np.uint8)
    return sc, occ


def reproject_outputs(orc, which):
    """Every REPROJECT_CASES case through the oracle ("orc") or the reference's own compiled reprojector.cpp ("ref")."""
    out = {}
    for ci, case in enumerate(REPROJECT_CASES):
        seed, max_n, n_in, occ_frac, by_obs, kw = case
        sc, occ = reproject_case_inputs(case)
        keep = []
        kfs = [orc.make_frame(orc.create_img_pyramid(im, 5), sc["cam"], IDENTITY7, T, keep=keep)
               for im, T in zip(sc["kf_imgs"], sc["tables"]["kf_T_f_w"])]
        cur = orc.make_frame(orc.create_img_pyramid(sc["cur_img"], 5), sc["cam"], IDENTITY7, sc["cur_T_f_w"], keep=keep)
        opt = orc.ReprojOptions(30, max_n, 1, 0, by_obs, 200.0, reproject_px_error_angle(sc["cam"]))
        res, st = orc.reproject_match(kfs, sc["tables"], cur, sc["entry_feat"], n_in, occ, opt, which=which)
        for k in REPROJ_INT_FIELDS + REPROJ_FLOAT_FIELDS:
            out[f"c{ci}_{k}"] = res[k]
        out[f"c{ci}_stats"] = np.array([st["n_candidates"], st["n_trials"], st["n_matches"], st["n_consumed"]])
        out[f"c{ci}_occ"] = occ
    return out


def assert_reproject_equal(a, b, px_tol, rel_tol, tag=""):
    """Two reproject_outputs dicts: integer fields identical, sub-pixel results within px_tol, seed states within rel_tol."""
    for ci in range(len(REPROJECT_CASES)):
        for k in REPROJ_INT_FIELDS + ("stats", "occ"):
            assert np.array_equal(a[f"c{ci}_{k}"], b[f"c{ci}_{k}"]), (tag, ci, k)
        assert np.abs(a[f"c{ci}_cur_px"] - b[f"c{ci}_cur_px"]).max() < 1e-9, (tag, ci)
        assert np.abs(a[f"c{ci}_px"] - b[f"c{ci}_px"]).max() < px_tol, (tag, ci)
        assert np.abs(a[f"c{ci}_f"] - b[f"c{ci}_f"]).max() < 1e-5, (tag, ci)
        assert np.abs(a[f"c{ci}_grad"] - b[f"c{ci}_grad"]).max() < 1e-9, (tag, ci)
        np.testing.assert_allclose(a[f"c{ci}_seed_state"], b[f"c{ci}_seed_state"], rtol=rel_tol, atol=1e-12)


# (scene seed, max_n_features_per_frame, reproject_unconverged_seeds, max_unconverged_seeds_ratio, min_required_features, remove_unconstrained)
REPROJECT_FRAMES_CASES = [
    (3, 120, 1, -1.0, 0, 1),     # stops inside the landmark pass
    (3, 400, 1, -1.0, 0, 0),     # all three passes: landmarks, converged seeds, unconverged seeds (updateSeed)
    (4, 260, 1, 0.3, 0, 1),      # unconverged seeds capped by the ratio
    (5, 300, 0, -1.0, 0, 1),     # unconverged seeds not reprojected
    (6, 150, 1, -1.0, 0, 0),     # enough features after the landmark pass: converged seeds only occupy cells
]
REPROJ_FRAMES_KEYS = ("type", "px", "level", "point", "seed_feat", "state", "f", "grad", "score", "occupancy", "stats", "pt_counters",
                      "feat_state", "feat_type")


def reproject_frames_reference(orc):
    """The reference's whole Reprojector::reprojectFrames (compiled reprojector.cpp) on every REPROJECT_FRAMES_CASES case."""
    out = {}
    for ci, (seed, max_n, unconv, ratio, min_req, rm) in enumerate(REPROJECT_FRAMES_CASES):
        sc = synth.make_reproject_scene(seed)
        keep = []
        kfs = [orc.make_frame(orc.create_img_pyramid(im, 5), sc["cam"], IDENTITY7, T, keep=keep)
               for im, T in zip(sc["kf_imgs"], sc["tables"]["kf_T_f_w"])]
        cur = orc.make_frame(orc.create_img_pyramid(sc["cur_img"], 5), sc["cam"], IDENTITY7, sc["cur_T_f_w"], keep=keep)
        o = orc.ref_reproject_frames(kfs, sc["tables"], len(kfs), cur, max_n, unconv, ratio, min_req, rm)
        for k in REPROJ_FRAMES_KEYS:
            out[f"rf{ci}_{k}"] = o[k][:416] if k == "occupancy" else o[k]
    return out


# ---- f4: PoseOptimizer -----------------------------------------------------------------------------------------------------------
# (case seed, cameras, error type, radtan camera, rotation prior)
POSE_OPT_CASES = [(1, 1, 0, False, False), (2, 2, 0, False, True), (3, 1, 2, False, False), (4, 1, 1, False, True), (5, 2, 1, True, False),
                  (6, 1, 2, True, True), (7, 2, 2, False, False), (8, 1, 0, True, False)]


def pose_opt_case(spec):
    seed, n_cams, err_type, radtan, prior = spec
    c = synth.make_pose_opt_case(seed, n_cams=n_cams, cam=synth.EUROC_CAM_RADTAN if radtan else None)
    return c, (c["T_imu_world_true"][:4].copy() if prior else None)


def pose_opt_outputs(orc, which):
    """Every POSE_OPT_CASES case through the oracle ("orc") or the reference's own compiled pose_optimizer.cpp ("ref")."""
    out = {}
    for ci, spec in enumerate(POSE_OPT_CASES):
        c, prior = pose_opt_case(spec)
        n, T, outl, stats = orc.pose_optimize(c, orc.pose_opt_options(err_type=spec[2], prior_q=prior, prior_lambda=0.5), which)
        out[f"p{ci}_n"], out[f"p{ci}_T"], out[f"p{ci}_outlier"], out[f"p{ci}_stats"] = np.array(n), T, outl, stats[:4]
    return out


# ---- f2: edgelet detector / detector classes (oracle/_ref/libdetect_ref.so) ----------------------------------------------------------
# (seed, width, height, n_levels, kind, threshold_secondary, border, with_occupancy)
DETECT_CASES = [(11, 752, 480, 5, "rect", 100, 8, False), (12, 752, 480, 5, "rect", 30, 8, True), (13, 640, 480, 4, "rect", 100, 4, False),
                (14, 500, 300, 3, "stripes", 60, 8, False), (15, 752, 480, 5, "noise", 100, 8, True), (16, 376, 240, 2, "rect", 250, 10, False),
                (17, 100, 75, 2, "rect", 40, 8, False), (18, 44, 40, 2, "noise", 20, 8, False),
                # 0 / 255 blobs: squared magnitudes beyond 2^22, where float(sqrt(n)) stops being injective (edgelet.cu's slow paths),
                # the second one with a threshold above 2048 as well
                (19, 376, 240, 2, "binary", 100, 8, False), (20, 376, 240, 2, "binary", 2500, 8, True)]
CORNER_FIELDS = ("x", "y", "level", "score", "angle")


def detect_image(seed, w, h, kind):
    """Level-0 test image: piece-wise constant rectangles, 8-periodic stripes with many equal gradient scores, or noise."""
    rng = np.random.default_rng(seed)
    if kind == "rect":
        return synth.make_image(seed, w, h, n_rect=max(8, w * h // 400))
    if kind == "binary":
        yy, xx = np.mgrid[0:h, 0:w]
        blobs = np.sin(xx / 7.0 + seed) * np.cos(yy / 5.0) + 0.3 * np.sin((xx + 2 * yy) / 11.0) + 0.15 * rng.standard_normal((h, w))
        return np.where(blobs > 0, 255, 0).astype(np.uint8)
    if kind == "stripes":
        yy, xx = np.mgrid[0:h, 0:w]
        return (((xx // 8 + yy // 16) % 2) * 150 + 40 + rng.integers(0, 2, (h, w))).astype(np.uint8)
    return rng.integers(0, 256, (h, w)).astype(np.uint8)


def detect_case_inputs(orc, case):
    seed, w, h, n_levels, kind, thr2, border, with_occ = case
    img = detect_image(seed, w, h, kind)
    pyr = orc.create_img_pyramid(img, n_levels)
    n_cells = (-(-w // 30)) * (-(-h // 30))
    occ = (np.random.default_rng(seed + 100).random(n_cells) < 0.3).astype(np.uint8) if with_occ else None
    return img, pyr, occ


def detect_outputs(orc, which):
    """Per-cell edgelets / FAST corners and the three detectors' feature lists for DETECT_CASES, from the oracle restatement
    (which="orc") or the reference's own compiled detectors (which="ref")."""
    out = {}
    for i, case in enumerate(DETECT_CASES):
        seed, w, h, n_levels, kind, thr2, border, with_occ = case
        img, pyr, occ = detect_case_inputs(orc, case)
        e = orc.edgelet_detector_v2(pyr, thr2, border, 30, occ, which=which)
        f = orc.fast_detector_pyr(pyr, 10, border, 0, min(2, n_levels - 1), 30, occ, which=which)
        for k in CORNER_FIELDS:
            out[f"edgelet_{i}_{k}"] = e[k]
            out[f"fast_{i}_{k}"] = f[k]
        for t, max_n in ((orc.DETECTOR_FAST, None), (orc.DETECTOR_FAST_GRAD, None), (orc.DETECTOR_GRID_GRAD, None),
                         (orc.DETECTOR_FAST_GRAD, 60)):
            d = orc.detect_features(t, pyr, 10.0, float(thr2), border, 0, min(2, n_levels - 1), 30, occ, max_n, which=which)
            for k, v in d.items():
                out[f"det_{i}_{t}_{max_n}_{k}"] = v
    g = np.random.default_rng(5)
    img = detect_image(21, 376, 240, "rect")
    pts = np.stack([g.integers(0, 376, 200), g.integers(0, 240, 200)], 1)
    pts[:8] = [[0, 0], [375, 239], [1, 1], [3, 236], [374, 5], [4, 4], [0, 120], [200, 239]]
    out["hist_pts"] = pts
    out["hist_angle"] = np.array([orc.angle_at_pixel_histogram(img, x, y, 4, which=which) for x, y in pts])
    return out


def assert_features_equal(a, b, tag=""):
    """Feature lists of AbstractDetector::detect: fillFeatures sorts by score with std::sort (unstable), so features with equal
    scores are compared as sets."""
    assert len(a["score"]) == len(b["score"]), tag
    assert np.array_equal(a["score"], b["score"]) and np.array_equal(a["type"], b["type"]), tag
    ka = sorted(zip(a["score"], a["px"][:, 0], a["px"][:, 1], a["level"], a["grad"][:, 0], a["grad"][:, 1]))
    kb = sorted(zip(b["score"], b["px"][:, 0], b["px"][:, 1], b["level"], b["grad"][:, 0], b["grad"][:, 1]))
    assert ka == kb, tag


# ---- f3: StereoTriangulation::compute ------------------------------------------------------------------------------------------------
# (scene seed, srand seed, detector type, triangulate_n_features, mean / min / max inverse depth)
STEREO_TRI_CASES = [(81, 5, 2, 120, 1 / 3.0, 1.0, 1 / 50.0),    # the defaults on the FastGrad detector: stops after 120 successes
                    (82, 9, 0, 1000, 1 / 3.0, 1.0, 1 / 50.0),   # FAST only, more wanted than exist: every feature is tried
                    (83, 3, 5, 40, 1 / 3.0, 1.0, 1 / 50.0),     # edgelets only (align_1d for every feature)
                    (84, 7, 2, 200, 1 / 2.0, 1 / 1.5, 1 / 4.2)]  # a depth range that cuts part of the scene off: many failures


def stereo_tri_frames(orc, case, keep):
    d, s1 = stereo_case(case[0])
    p0, p1 = orc.create_img_pyramid(d["ref_img"], 5), orc.create_img_pyramid(s1["ref_img"], 5)
    f0 = orc.make_frame(p0, d["cam"], d["T_cam_imu"], d["T_imu_world_ref"], keep=keep)
    f1 = orc.make_frame(p1, d["cam"], s1["T_cam_imu"], d["T_imu_world_ref"], keep=keep)
    return d, s1, p0, p1, f0, f1


def stereo_tri_entries(orc, case, d, p0, order):
    """frame0's detected features (the oracle's detector, identical to the reference's) in the visiting order `order`."""
    det = orc.detect_features(case[2], p0)
    f = synth.cam_backproject(d["cam"], det["px"][order])
    f = f / np.linalg.norm(f, axis=1, keepdims=True)
    return det, f


def stereo_tri_reference(orc):
    """The reference's own StereoTriangulation::compute on STEREO_TRI_CASES (oracle/_ref/libfrontend_ref.so) + the visiting orders."""
    out = {}
    for i, case in enumerate(STEREO_TRI_CASES):
        keep = []
        d, s1, p0, p1, f0, f1 = stereo_tri_frames(orc, case, keep)
        r = orc.ref_stereo_triangulation_compute(f0, f1, case[2], 10.0, 100.0, case[3], case[4], case[5], case[6], seed=case[1])
        n_c = int((r["type0"] == 7).sum())
        out[f"order_{i}"] = orc.ref_stereo_shuffle_order(case[1], 0, n_c, r["n0"])
        for k, v in r.items():
            out[f"{k}_{i}"] = np.asarray(v)
    return out


def assert_stereo_matches_reference(res, order, g, i, tol=1e-9):
    """Per-entry results (oracle or CUDA) against what the reference's compute() left in frame1."""
    ok = res[res["status"] == 2]
    assert len(ok) == int(g[f"n1_{i}"]), (i, len(ok), int(g[f"n1_{i}"]))
    assert np.array_equal(order[res["status"] == 2], g[f"ref_index1_{i}"]), i   # the same frame0 features, in the same order
    assert np.array_equal(ok["slot"], np.arange(len(ok))) and np.array_equal(ok["level"], g[f"level1_{i}"]) and np.array_equal(ok["type"], g[f"type1_{i}"])
    np.testing.assert_allclose(ok["px_cur"], g[f"px1_{i}"], rtol=0, atol=1e-3)   # sub-pixel results within 1e-3 px (north_star)
    np.testing.assert_allclose(ok["f_cur"], g[f"f1_{i}"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(ok["grad_cur"], g[f"grad1_{i}"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(ok["xyz_world"], g[f"xyz1_{i}"], rtol=1e-4, atol=1e-6)


# ---- f4 (second half): Point::optimize -------------------------------------------------------------------------------------------------
def point_opt_cases(seed=17, n_points=400, n_frames=12):
    """Seeded structure-refinement problems: n_frames camera poses around the origin, n_points landmarks 2-8 m in front of them, each
    observed by 1-10 of the frames with bearing noise, starting from a perturbed position (a few hard cases: 1 observation, a start
    far off, observations from nearly the same place)."""
    rng = np.random.default_rng(seed)
    T_f_w = np.stack([synth.se3_exp_small(rng.normal(0, 0.05, 3), rng.normal(0, 0.25, 3)) for _ in range(n_frames)])
    pos_true = np.stack([rng.uniform(-2, 2, n_points), rng.uniform(-1.5, 1.5, n_points), rng.uniform(2, 8, n_points)], 1)
    begin, obs_frame, obs_f, pos0 = [0], [], [], []
    for i in range(n_points):
        k = 1 if i % 57 == 0 else int(rng.integers(2, 11))
        fr = rng.choice(n_frames, k, replace=False)
        for j in fr:
            R, t = synth.se3_to_Rt(T_f_w[j])
            p = R @ pos_true[i] + t
            b = p / np.linalg.norm(p) + rng.normal(0, 2e-3, 3)
            obs_f.append(b / np.linalg.norm(b))
            obs_frame.append(j)
        begin.append(len(obs_frame))
        pos0.append(pos_true[i] + rng.normal(0, 0.15 if i % 13 else 1.5, 3))
    return dict(T_f_w=T_f_w, pos_true=pos_true, pos0=np.array(pos0), obs_begin=np.array(begin, np.int32),
                obs_frame=np.array(obs_frame, np.int32), obs_f=np.array(obs_f))


POINT_OPT_INPUT_KEYS = ("T_f_w", "pos_true", "pos0", "obs_begin", "obs_frame", "obs_f")


def point_opt_outputs(orc, which, n_iter=5, c=None):
    """Point::optimize on the cases `c` (default: freshly generated). The golden file stores the INPUTS next to the reference's
    outputs: numpy's SIMD sin / cos differ in the last bit between hosts, and a 1-ulp change of an input moves the result of the
    non-converged cases by up to 4e-8 m, so tests on another box must start from the stored bytes."""
    c = point_opt_cases() if c is None else c
    out = {k: np.asarray(c[k]) for k in POINT_OPT_INPUT_KEYS}
    for sphere in (0, 1):
        res = []
        for i in range(len(c["pos0"])):
            lo, hi = c["obs_begin"][i], c["obs_begin"][i + 1]
            p, _ = orc.point_optimize(c["T_f_w"][c["obs_frame"][lo:hi]], c["obs_f"][lo:hi], c["pos0"][i], n_iter, bool(sphere), which=which)
            res.append(p)
        out[f"pos_{sphere}"] = np.array(res)
    return out


# ---- f3 (tracker part): FeatureTracker::trackAndDetect ---------------------------------------------------------------------------------
# (scene seed, number of frames, detector type, min_tracks_to_detect_new_features, reset_before_detection, template = first observation)
TRACKER_CASES = [(91, 5, 0, 50, True, True), (92, 5, 2, 381, True, True), (93, 4, 0, 366, False, False)]


def tracker_sequence(case):
    """A mono sequence: the textured plane of synth.make_align_pair(seed) seen from a camera that moves a little more every frame."""
    seed, n = case[0], case[1]
    d = synth.make_align_pair(seed)
    rng = np.random.default_rng(seed)
    imgs, T = [d["ref_img"]], synth.IDENTITY.copy()
    for _ in range(n - 1):
        T = synth.se3_mul(synth.se3_exp_small(rng.normal(0, 0.004, 3), rng.normal(0, 0.03, 3)), T)
        imgs.append(d["scene"].render(T))
    return d, imgs


def tracker_outputs(orc, which):
    out = {}
    for i, case in enumerate(TRACKER_CASES):
        d, imgs = tracker_sequence(case)
        pyrs = [orc.create_img_pyramid(im, 5) for im in imgs]
        if which == "ref":
            keep = []
            frames = [orc.make_frame(p, d["cam"], keep=keep) for p in pyrs]
            seq = orc.ref_feature_tracker_sequence(frames, case[2], 10.0, 100.0, case[3], case[4], case[5])
        else:
            seq = orc.feature_tracker_sequence(pyrs, case[2], 10.0, 100.0, case[3], case[4], case[5])
        for k, fr in enumerate(seq):
            for key, v in fr.items():
                out[f"{key}_{i}_{k}"] = np.asarray(v)
    return out
