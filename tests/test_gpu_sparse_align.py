"""GPU parity tests, rows b1-b9: svo_cuda_sparse_align against the oracle's SparseImgAlign::run on identical synthetic
752x480 EuRoC-shaped inputs. Tolerance (BASELINE.json north_star): poses within 1e-4 rad / 1e-4 m."""
import numpy as np
import pytest

from helpers import gpu_align, oracle_align, pose_diff, to_orc_options
from svo_pro_universal_b200 import capi, synth

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-4   # rad
TRANS_TOL = 1e-4  # m


def _compare(orc, pairs, res, gopt, priors=None, check_iters=True, h_rtol=1e-6):
    for i, d in enumerate(pairs):
        o = oracle_align(orc, d, to_orc_options(orc, gopt, None if priors is None else priors[i]))
        r = res[i]
        assert r["n_tracked"] == o.n_tracked
        dq, dt = pose_diff(r["T_icur_iref"], o.T_icur_iref)
        assert dq < ROT_TOL and dt < TRANS_TOL, f"pair {i}: dR={dq:.3e} rad dt={dt:.3e} m"
        if o.n_tracked:  # without features run() returns before touching the frames' poses
            dq, dt = pose_diff(r["T_f_w"][0], o.T_f_w[0])
            assert dq < ROT_TOL and dt < TRANS_TOL
        assert abs(r["alpha"] - o.alpha) < 1e-5 and abs(r["beta"] - o.beta) < 1e-3
        if check_iters:
            assert list(r["iters"]) == list(o.iters), f"pair {i}: GN iterations per level differ"
        if o.n_tracked:
            np.testing.assert_allclose(r["chi2"], o.chi2, rtol=1e-4)
            np.testing.assert_allclose(r["H"].reshape(8, 8), np.array(o.H).reshape(8, 8), rtol=h_rtol, atol=1e-6)
        assert r["stop"] == o.stop


def test_default_options_batch(ctx, orc):
    """BASELINE config 1 (4-level, ~180 features, 4x4 patches), a batch of independent pairs in one launch."""
    pairs = [synth.make_align_pair(s) for s in range(1, 9)]
    gopt = capi.sparse_align_options()
    res, _, _ = gpu_align(ctx, pairs, gopt)
    _compare(orc, pairs, res, gopt)
    for d, r in zip(pairs, res):
        dq, dt = pose_diff(r["T_icur_iref"], d["T_icur_iref_gt"])
        assert dq < 2e-3 and dt < 5e-3  # and it lands near the synthetic ground truth


@pytest.mark.parametrize("split", ["0", "1"])
def test_both_thread_mappings(ctx, orc, monkeypatch, split):
    """The plain 6-DoF kernel exists with one thread per patch (throughput: large batches) and with two threads per patch (latency:
    batches of at most one pair per SM, which is what the small parity cases are). SVO_ALIGN_SPLIT forces either: both must pass the
    same parity checks on the same batch, and agree with each other to rounding."""
    pairs = [synth.make_align_pair(s) for s in range(1, 9)]
    gopt = capi.sparse_align_options()
    monkeypatch.setenv("SVO_ALIGN_SPLIT", split)
    res, _, _ = gpu_align(ctx, pairs, gopt)
    _compare(orc, pairs, res, gopt)
    monkeypatch.setenv("SVO_ALIGN_SPLIT", "1" if split == "0" else "0")
    other, _, _ = gpu_align(ctx, pairs, gopt)
    for a, b in zip(res, other):
        assert list(a["iters"]) == list(b["iters"]) and a["n_tracked"] == b["n_tracked"]
        dq, dt = pose_diff(a["T_icur_iref"], b["T_icur_iref"])
        assert dq < 1e-7 and dt < 1e-11


def test_subpixel_feature_positions(ctx, orc):
    """Features at sub-pixel positions (what the Reprojector's refined matches are): the interpolated reference patch values are then
    not representable in the kernel's FP32 patch cache, whose rounding (relative 2^-24 of a grey level) must stay far inside the
    tolerance: pose within 1e-7 rad / 1e-8 m of the FP64 oracle, same iteration counts, for several option sets."""
    pairs = []
    for s in range(31, 39):
        d = synth.make_align_pair(s)
        rng = np.random.default_rng(s)
        d["px"] = d["px"] + rng.uniform(-0.5, 0.5, d["px"].shape)
        X = d["scene"].ref_points(d["px"])
        d["depth"] = np.linalg.norm(X, axis=1)
        d["f"] = X / d["depth"][:, None]
        pairs.append(d)
    for kw in (dict(), dict(estimate_illumination_gain=1, estimate_illumination_offset=1), dict(robustification=1, weight_scale=10.0),
               dict(max_level=2, min_level=0)):
        gopt = capi.sparse_align_options(**kw)
        res, _, _ = gpu_align(ctx, pairs, gopt)
        # H: the illumination cross terms sum dx * ref over patches with cancelling signs, so the 2^-24 rounding of ref shows up as a few 1e-6
        # relative in those (small) entries; poses, chi2 and iteration counts are held to the usual bounds
        _compare(orc, pairs, res, gopt, h_rtol=2e-5)
        for d, r in zip(pairs, res):
            o = oracle_align(orc, d, to_orc_options(orc, gopt))
            dq, dt = pose_diff(r["T_icur_iref"], o.T_icur_iref)
            assert dq < 1e-7 and dt < 1e-8, (kw, dq, dt)


@pytest.mark.parametrize("kw", [
    dict(estimate_illumination_gain=1, estimate_illumination_offset=1),
    dict(robustification=1, weight_scale=10.0),
    dict(estimate_illumination_gain=1, estimate_illumination_offset=1, robustification=1),
    dict(estimate_illumination_offset=1),
    dict(max_level=4, min_level=2),           # the shipped YAMLs (examples/param/pinhole.yaml:76-77)
    dict(max_level=3, min_level=0, max_iter=6, eps=1e-5),
    dict(use_distortion_jacobian=1),
    dict(alpha_init=0.02, beta_init=1.5, estimate_illumination_gain=1, estimate_illumination_offset=1),
])
def test_option_variants(ctx, orc, kw):
    pairs = [synth.make_align_pair(s) for s in (11, 12, 13)]
    gopt = capi.sparse_align_options(**kw)
    res, _, _ = gpu_align(ctx, pairs, gopt)
    _compare(orc, pairs, res, gopt)


def test_radtan_camera_with_distortion_jacobian(ctx, orc):
    pairs = [synth.make_align_pair(s, cam=synth.EUROC_CAM_RADTAN) for s in (21, 22)]
    for kw in (dict(), dict(use_distortion_jacobian=1)):
        gopt = capi.sparse_align_options(**kw)
        res, _, _ = gpu_align(ctx, pairs, gopt)
        _compare(orc, pairs, res, gopt)


def test_weighted_prior(ctx, orc):
    """setWeightedPrior (sparse_img_align_base.cpp:44-62, applyPrior :77-107), one prior per pair."""
    pairs = [synth.make_align_pair(s) for s in (31, 32, 33)]
    gopt = capi.sparse_align_options(lambda_rot=0.5, lambda_trans=0.1, lambda_alpha=0.0, lambda_beta=0.0)
    priors = np.zeros(len(pairs), capi.ALIGN_PRIOR_DTYPE)
    for i, d in enumerate(pairs):
        priors[i]["T"] = synth.se3_mul(d["T_icur_iref_gt"], synth.se3_exp_small(np.array([1e-3, -2e-3, 1e-3]), np.array([2e-3, 0, -1e-3])))
    res, _, _ = gpu_align(ctx, pairs, gopt, priors=priors)
    _compare(orc, pairs, res, gopt, priors=priors)


def test_edge_cases_no_features_ineligible_and_out_of_bounds(ctx, orc):
    d0 = synth.make_align_pair(41)
    none = dict(d0, eligible=np.zeros(len(d0["px"]), np.uint8))                      # nothing eligible -> run() returns 0
    border = dict(d0, px=np.concatenate([d0["px"][:50], [[5.0, 5.0], [750.0, 470.0], [39.9, 200.0], [700.0, 100.0]]]),
                  f=np.concatenate([d0["f"][:50], d0["f"][:4]]), depth=np.concatenate([d0["depth"][:50], d0["depth"][:4]]),
                  eligible=np.ones(54, np.uint8))                                    # 4 features fail the level-4 bounds test
    few = dict(d0, px=d0["px"][:7], f=d0["f"][:7], depth=d0["depth"][:7], eligible=np.ones(7, np.uint8))
    mixed = dict(d0, eligible=(np.arange(len(d0["px"])) % 3 != 0).astype(np.uint8))
    pairs = [none, border, few, mixed]
    gopt = capi.sparse_align_options()
    res, _, _ = gpu_align(ctx, pairs, gopt)
    assert res[0]["n_tracked"] == 0 and res[1]["n_tracked"] == 50 and res[2]["n_tracked"] == 7
    np.testing.assert_allclose(res[0]["T_icur_iref"], synth.IDENTITY, atol=1e-12)   # pose untouched without features
    _compare(orc, pairs, res, gopt)


def test_far_initial_guess_and_points_behind_camera(ctx, orc):
    """Visibility handling (sparse_img_align.cpp:432-460): patches leaving the image / z < 0 drop out per iteration."""
    d = synth.make_align_pair(51, max_rot_deg=6.0, max_trans=0.3)
    behind = dict(d, depth=d["depth"].copy())
    behind["T_imu_world_cur_init"] = synth.se3_mul(synth.se3_exp_small(np.zeros(3), np.array([0.0, 0.0, 30.0])), d["T_imu_world_ref"])
    pairs = [d, behind]
    gopt = capi.sparse_align_options()
    res, _, _ = gpu_align(ctx, pairs, gopt)
    _compare(orc, pairs, res, gopt)


def test_many_features_and_max_capacity(ctx, orc):
    d = synth.make_align_pair(61, n_features=400)
    assert len(d["px"]) > 200
    gopt = capi.sparse_align_options()
    res, _, _ = gpu_align(ctx, [d], gopt)
    _compare(orc, [d], res, gopt)
    # more features than one CTA's shared memory can hold (about 1380 with the FP32 patch cache) -> explicit error, not a silent truncation
    n_big = 1600
    with pytest.raises(capi.SvoCudaError):
        p = capi.Pyramid(ctx, 1, 752, 480, 5)
        capi.sparse_align(ctx, [p], [p], [capi.Camera.from_dict(d["cam"])], np.array([synth.IDENTITY]), np.array([synth.IDENTITY]),
                          np.array([synth.IDENTITY]), np.zeros((1, 1), np.int32), np.zeros((1, 1, n_big, 2)), np.zeros((1, 1, n_big, 3)),
                          np.ones((1, 1, n_big)), np.zeros((1, 1, n_big), np.uint8), gopt)
    # a bundle that fills most of that capacity: the same 300-odd features four times over (two passes of the 384-thread variant)
    rep = 4
    d4 = dict(d)
    for k in ("px", "f", "depth", "eligible"):
        d4[k] = np.concatenate([d[k]] * rep)
    res4, _, _ = gpu_align(ctx, [d4], gopt)
    _compare(orc, [d4], res4, gopt)
    assert res4[0]["n_tracked"] == rep * res[0]["n_tracked"] > 1000


def test_frame_index_indirection_and_device_resident_io(ctx, orc):
    """ref/cur frame index arrays + SVO_MEM_DEVICE arrays (torch tensors) give the same results as host staging."""
    import torch
    pairs = [synth.make_align_pair(s) for s in (71, 72, 73)]
    gopt = capi.sparse_align_options()
    res_host, ref, cur = gpu_align(ctx, pairs, gopt)
    from svo_pro_universal_b200 import batch
    pk = batch.pack_align_batch(pairs)
    perm = np.array([2, 0, 1], np.int32)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = torch.zeros(3 * capi.ALIGN_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    a = {k: t(pk[k][perm]) for k in ("T_imu_world_ref", "T_imu_world_cur", "n_features", "px", "f", "depth", "eligible")}
    idx = t(perm.reshape(3, 1))
    torch.cuda.synchronize()  # the context runs on its own non-blocking stream: make torch's uploads visible first
    capi.sparse_align(ctx, [ref], [cur], [capi.Camera.from_dict(pairs[0]["cam"])], pk["T_cam_imu"], a["T_imu_world_ref"],
                      a["T_imu_world_cur"], a["n_features"], a["px"], a["f"], a["depth"], a["eligible"], gopt, ref_frame_idx=idx,
                      cur_frame_idx=idx, results=out)
    ctx.synchronize()
    res_dev = out.cpu().numpy().view(capi.ALIGN_RESULT_DTYPE)
    for k, i in enumerate(perm):
        np.testing.assert_array_equal(res_dev[k]["T_icur_iref"], res_host[i]["T_icur_iref"])  # same kernel, same bits


def test_stereo_bundle_two_cameras(ctx, orc):
    """FrameBundle with 2 cameras sharing one IMU pose (stereo): features of both ref frames drive one 6-DoF state."""
    d = synth.make_align_pair(81)
    # second camera: same scene seen from a camera displaced 11 cm along x (EuRoC-like baseline)
    T_c1_c0 = synth.se3_exp_small(np.zeros(3), np.array([-0.11, 0.0, 0.0]))
    T_cam1_imu = synth.se3_mul(T_c1_c0, d["T_cam_imu"])
    scene = d["scene"]
    ref1 = scene.render(T_c1_c0)
    T_cur1_ref0 = synth.se3_mul(T_c1_c0, d["T_cur_ref_gt"])
    cur1 = scene.render(T_cur1_ref0)
    px1 = synth.pick_features(ref1, 150, 5)
    # depth of cam-1 features: intersect cam-1 rays with the plane (expressed in cam 0)
    f1 = synth.cam_backproject(d["cam"], px1)
    R01, t01 = synth.se3_to_Rt(synth.se3_inv(T_c1_c0))
    dirs0 = f1 @ R01.T
    lam = (scene.d - scene.n @ t01) / (dirs0 @ scene.n)
    X1 = f1 * lam[:, None]
    depth1 = np.linalg.norm(X1, axis=1)
    fb1 = X1 / depth1[:, None]

    B, F = 1, 180
    pyr = {}
    for name, img in (("r0", d["ref_img"]), ("c0", d["cur_img"]), ("r1", ref1), ("c1", cur1)):
        p = capi.Pyramid(ctx, 1, 752, 480, 5)
        p.upload(img)
        p.build()
        pyr[name] = p
    px = np.zeros((B, 2, F, 2)); f = np.zeros((B, 2, F, 3)); dep = np.ones((B, 2, F)); el = np.zeros((B, 2, F), np.uint8)
    n0, n1 = len(d["px"]), len(px1)
    px[0, 0, :n0], f[0, 0, :n0], dep[0, 0, :n0], el[0, 0, :n0] = d["px"], d["f"], d["depth"], 1
    px[0, 1, :n1], f[0, 1, :n1], dep[0, 1, :n1], el[0, 1, :n1] = px1, fb1, depth1, 1
    gopt = capi.sparse_align_options(estimate_illumination_gain=1, estimate_illumination_offset=1)
    cams = [capi.Camera.from_dict(d["cam"])] * 2
    res = capi.sparse_align(ctx, [pyr["r0"], pyr["r1"]], [pyr["c0"], pyr["c1"]], cams, np.stack([d["T_cam_imu"], T_cam1_imu]),
                            d["T_imu_world_ref"][None], d["T_imu_world_cur_init"][None], np.array([[n0, n1]], np.int32), px, f, dep, el,
                            gopt)[0]
    keep = []
    rp0, cp0 = orc.create_img_pyramid(d["ref_img"], 5), orc.create_img_pyramid(d["cur_img"], 5)
    rp1, cp1 = orc.create_img_pyramid(ref1, 5), orc.create_img_pyramid(cur1, 5)
    rf = [orc.make_frame(rp0, d["cam"], d["T_cam_imu"], d["T_imu_world_ref"], d["px"], d["f"], d["depth"], keep=keep),
          orc.make_frame(rp1, d["cam"], T_cam1_imu, d["T_imu_world_ref"], px1, fb1, depth1, keep=keep)]
    cf = [orc.make_frame(cp0, d["cam"], d["T_cam_imu"], d["T_imu_world_cur_init"], keep=keep),
          orc.make_frame(cp1, d["cam"], T_cam1_imu, d["T_imu_world_cur_init"], keep=keep)]
    o = orc.sparse_align(rf, cf, to_orc_options(orc, gopt))
    assert res["n_tracked"] == o.n_tracked == n0 + n1
    dq, dt = pose_diff(res["T_icur_iref"], o.T_icur_iref)
    assert dq < ROT_TOL and dt < TRANS_TOL
    for c in range(2):
        dq, dt = pose_diff(res["T_f_w"][c], o.T_f_w[c])
        assert dq < ROT_TOL and dt < TRANS_TOL
    dq, dt = pose_diff(res["T_icur_iref"], d["T_icur_iref_gt"])
    assert dq < 2e-3 and dt < 5e-3


def test_full_size_batch_properties(ctx):
    """Size-independent properties on a larger tiled batch: identical inputs -> identical outputs across the batch
    (per-CTA determinism), and a second call reproduces the first bit for bit."""
    from svo_pro_universal_b200 import batch
    uniq = [synth.make_align_pair(s) for s in (91, 92, 93, 94)]
    pk = batch.tile_batch(batch.pack_align_batch(uniq), 256)
    B = 256
    ref = capi.Pyramid(ctx, B, 752, 480, 5)
    cur = capi.Pyramid(ctx, B, 752, 480, 5)
    ref.upload(pk["ref_imgs"]); cur.upload(pk["cur_imgs"])
    ref.build(); cur.build()
    args = ([ref], [cur], [capi.Camera.from_dict(uniq[0]["cam"])], pk["T_cam_imu"], pk["T_imu_world_ref"], pk["T_imu_world_cur"],
            pk["n_features"], pk["px"], pk["f"], pk["depth"], pk["eligible"], capi.sparse_align_options())
    a = capi.sparse_align(ctx, *args)
    b = capi.sparse_align(ctx, *args)
    assert a.tobytes() == b.tobytes()
    for i in range(4, B):
        assert a[i].tobytes() == a[i % 4].tobytes()
    for i in range(4):
        dq, dt = pose_diff(a[i]["T_icur_iref"], uniq[i]["T_icur_iref_gt"])
        assert dq < 2e-3 and dt < 5e-3
