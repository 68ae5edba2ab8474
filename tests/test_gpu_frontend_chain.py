"""GPU parity test of the whole batched stereo front-end (BASELINE.json configs[4]: detect + sparse align + feature align +
depth filter): svo_pro_universal_b200.frontend.StereoFrontendBatch chains pyramid -> 2-camera SparseImgAlign -> Reprojector ->
DepthFilter::updateSeeds -> FastGrad detector (FAST + edgelets) on device-resident arrays; every stage is compared with the oracle on the same inputs (the
oracle's reprojection stage is fed the poses the aligner produced, so that each stage is checked on identical inputs).
Tolerances: poses 1e-4 rad / 1e-4 m (measured ~1e-12), sub-pixel 1e-3 px, seed states 1e-4 relative, corners bit-exact."""
import numpy as np
import pytest

import helpers
from svo_pro_universal_b200 import capi, frontend, synth

pytestmark = pytest.mark.gpu


def test_stereo_frontend_chain_matches_oracle(ctx, orc):
    import torch
    scenes = [frontend.make_stereo_scene(s) for s in (81, 82)]
    B = 3   # pair 2 tiles scene 0 again
    fb = frontend.StereoFrontendBatch(ctx, scenes, B, torch.device("cuda", 0))
    fb.step()
    fb.step()   # a second pass over the same batch gives the same answer (state is reset per step)
    out = fb.results()
    fb.release()
    ang = helpers.reproject_px_error_angle(scenes[0]["cam"])
    n_match_total = 0
    for i in range(B):
        sc = scenes[i % 2]
        keep = []
        cam = sc["cam"]
        pyr = {k: orc.create_img_pyramid(v, 5) for k, v in sc["imgs"].items()}
        # ---- SparseImgAlign::run on the two-camera bundle (illumination gain + offset)
        rfs = [orc.make_frame(pyr[f"r{c}"], cam, sc["T_cam_imu"][c], sc["T_imu_world_ref"], sc["px"][c], sc["f"][c], sc["depth"][c], keep=keep)
               for c in range(2)]
        cfs = [orc.make_frame(pyr[f"c{c}"], cam, sc["T_cam_imu"][c], sc["T_imu_world_ref"], keep=keep) for c in range(2)]
        r = orc.sparse_align(rfs, cfs, orc.default_align_options(estimate_illumination_gain=1, estimate_illumination_offset=1))
        g = out["align"][i]
        dq, dt = helpers.pose_diff(g["T_icur_iref"], np.array(r.T_icur_iref[:]))
        assert dq < 1e-4 and dt < 1e-4 and dq < 1e-9 and dt < 1e-9, (i, dq, dt)
        assert g["n_tracked"] == r.n_tracked > 250 and list(g["iters"][:4]) == list(r.iters[:4])
        for c in range(2):
            dq, dt = helpers.pose_diff(g["T_f_w"][c], np.array(r.T_f_w[c][:]))
            assert dq < 1e-9 and dt < 1e-9
        # the aligner recovered the true motion
        T_true = synth.se3_mul(sc["T_cur_ref_gt"], sc["T_f_w_ref"][0])
        dq, dt = helpers.pose_diff(g["T_f_w"][0], T_true)
        assert dq < 2e-3 and dt < 5e-3, (dq, dt)
        # ---- Reprojector per camera, on the pose the aligner produced
        for c in range(2):
            j = 2 * i + c
            lo, hi = out["entry_begin"][j], out["entry_begin"][j + 1]
            n = hi - lo
            kf = orc.make_frame(pyr[f"r{c}"], cam, helpers.IDENTITY7, sc["T_f_w_ref"][c], keep=keep)
            cf = orc.make_frame(pyr[f"c{c}"], cam, helpers.IDENTITY7, np.ascontiguousarray(g["T_f_w"][c]), keep=keep)
            ftype = synth.K_CORNER if c == 0 else synth.K_CORNER_SEED_CONV
            st = np.tile([1.0, 1e-6, 10.0, 10.0], (n, 1)); st[:, 0] = 1.0 / sc["depth"][c]
            R, tt = synth.se3_to_Rt(synth.se3_inv(sc["T_f_w_ref"][0]))
            feat = np.zeros(n, capi.FEATURE_DTYPE)
            feat["px"], feat["f"], feat["grad"], feat["type"] = sc["px"][c], sc["f"][c], [1.0, 0.0], ftype
            tables = dict(n_kfs=1, n_feat=n, n_points=n if c == 0 else 0, n_obs=n if c == 0 else 0, kf_seed_mu_range=np.array([1.0 / 1.5]),
                          kf_feat_begin=np.array([0, n], np.int32), feat=feat, feat_score=np.linspace(60.0, 11.0, n), feat_seed_state=st,
                          feat_point=(np.arange(n) if c == 0 else np.full(n, -1)).astype(np.int32), feat_kf=np.zeros(n, np.int32),
                          pt_pos=((sc["f"][0] * sc["depth"][0][:, None]) @ R.T + tt) if c == 0 else np.zeros((1, 3)),
                          pt_n_failed=np.zeros(max(n, 1), np.int32), pt_n_succeeded=np.zeros(max(n, 1), np.int32),
                          pt_obs_begin=(np.arange(n + 1) if c == 0 else np.zeros(1)).astype(np.int32),
                          obs_feat=(np.arange(n) if c == 0 else np.zeros(1)).astype(np.int32))
            occ = np.zeros(416, np.uint8)
            ro, so = orc.reproject_match([kf], tables, cf, np.arange(n, dtype=np.int32), 0, occ, orc.ReprojOptions(30, 120, 1, 0, 0, 200.0, ang))
            rg = out["reproj"][lo:hi]
            for k in helpers.REPROJ_INT_FIELDS:
                assert np.array_equal(rg[k], ro[k]), (i, c, k)
            assert np.abs(rg["px"] - ro["px"]).max() < 1e-3 and np.abs(rg["cur_px"] - ro["cur_px"]).max() < 1e-9
            sg = out["reproj_stats"][j]
            assert (sg["n_candidates"], sg["n_trials"], sg["n_matches"], sg["n_consumed"]) == \
                   (so["n_candidates"], so["n_trials"], so["n_matches"], so["n_consumed"])
            assert np.array_equal(out["occupancy"][j], occ)
            n_match_total += int(so["n_matches"])
        # ---- PoseOptimizer::run on the matched entries of both cameras (same start pose, same measurements)
        lo, hi = out["entry_begin"][2 * i], out["entry_begin"][2 * i + 2]
        rg = out["reproj"][lo:hi]
        n0 = out["entry_begin"][2 * i + 1] - lo
        case = dict(cam=cam, T_cam_imu=sc["T_cam_imu"], T_imu_world_init=out["pose_opt_T0"][i], px=rg["px"], f=rg["f"], grad=rg["grad"],
                    level=rg["level"], type=np.where(np.arange(hi - lo) < n0, synth.K_CORNER, synth.K_CORNER_SEED_CONV).astype(np.int32),
                    xyz_world=out["pose_opt_xyz"][lo:hi], has_xyz=(rg["status"] == 4).astype(np.uint8),
                    feat_cam=(np.arange(hi - lo) >= n0).astype(np.int32))
        n_po, T_po, outl_po, st_po = orc.pose_optimize(case, orc.pose_opt_options())
        pg = out["pose_opt"][i]
        dq, dt = helpers.pose_diff(pg["T_imu_world"], T_po)
        if not (dq < 1e-9 and dt < 1e-9):
            import os
            os.makedirs("gpurun_out", exist_ok=True)
            np.savez("gpurun_out/chain_po_debug.npz", T_gpu=pg["T_imu_world"], T_orc=T_po, st_orc=st_po, sigma_gpu=pg["measurement_sigma"],
                     iters_gpu=pg["iters"], n_gpu=pg["n_meas"], outl_gpu=out["pose_opt_outlier"][lo:hi], outl_orc=outl_po,
                     **{"case_" + k: np.asarray(v) for k, v in case.items() if k != "cam"})
        assert dq < 1e-9 and dt < 1e-9 and pg["n_meas_final"] == n_po and pg["iters"] == int(st_po[3]), (i, dq, dt)
        assert np.array_equal(out["pose_opt_outlier"][lo:hi], outl_po) and pg["n_meas"] > 200
        dq, dt = helpers.pose_diff(synth.se3_mul(sc["T_cam_imu"][0], pg["T_imu_world"]), T_true)   # still at the true pose
        assert dq < 2e-3 and dt < 5e-3
        # ---- FastGrad detector on the new left frame: FAST corners, then edgelets in the cells without a corner
        co = orc.fast_detector(sc["imgs"]["c0"])
        cg = out["corners"][i]
        for k in ("x", "y", "level", "score"):
            assert np.array_equal(cg[k], co[k]), (i, k)
        eo = orc.edgelet_detector_v2(orc.create_img_pyramid(sc["imgs"]["c0"], 5), 100, 8, 30, (co["score"] > 10).astype(np.uint8))
        eg = out["edgelets"][i]
        for k in ("x", "y", "level", "score", "angle"):
            assert np.array_equal(eg[k], eo[k]), (i, k)
    assert n_match_total > 100 * B
    # ---- DepthFilter::updateSeeds: the seeds of every left reference frame against the new left frame
    p = 0
    n_ok = 0
    for i in range(B):
        sc = scenes[i % 2]
        keep = []
        rf = orc.make_frame(orc.create_img_pyramid(sc["imgs"]["r0"], 5), sc["cam"], keep=keep)
        cf = orc.make_frame(orc.create_img_pyramid(sc["imgs"]["c0"], 5), sc["cam"], keep=keep)
        n = len(sc["seed_px"])
        lv = np.random.default_rng(1).integers(0, 3, sum(len(scenes[k % 2]["seed_px"]) for k in range(B))).astype(np.int32)[p:p + n]
        oft = orc.make_features(sc["seed_px"], sc["seed_f"], np.tile([1.0, 0.0], (n, 1)), np.full(n, synth.K_CORNER_SEED, np.int32), lv)
        ty, st = np.full(n, synth.K_CORNER_SEED, np.uint8), sc["seed_state"].copy()
        k_ok, _, _ = orc.update_seeds(rf, [cf], sc["T_cur_ref_gt"].reshape(1, 7), oft, ty, st, sc["seed_mu_range"], orc.default_matcher_options())
        assert np.array_equal(out["seed_types"][p:p + n], ty)
        np.testing.assert_allclose(out["seed_state"][p:p + n], st, rtol=1e-4, atol=1e-12)
        n_ok += k_ok
        p += n
    assert out["n_seed_ok"] == n_ok > 0


def test_chain_stereo_triangulation_stage(ctx, orc):
    """The optional keyframe stage of the chain: the corners / edgelets FastGrad just found in the new left frame are triangulated
    against the new right frame on the device (fixed-shape entry lists with holes, pose from the aligner). Checked against the
    oracle's StereoTriangulation loop on the very features and pose the device stage used."""
    import torch
    scenes = [frontend.make_stereo_scene(s) for s in (81, 82)]
    B = 3
    fb = frontend.StereoFrontendBatch(ctx, scenes, B, torch.device("cuda", 0), stereo_triangulation=True)
    fb.step()
    out = fb.results()
    fb.release()
    n_tri = 0
    for i in range(B):
        sc = scenes[i % 2]
        keep = []
        ft_all, res_all = out["stereo_ftrs"][i], out["stereo"][i]
        sel = np.flatnonzero(ft_all["type"] >= 0)
        assert len(sel) > 300 and (res_all["status"][ft_all["type"] < 0] == capi.STEREO_NOT_REACHED).all()
        # the entry list = corners then edgelets in cell order, exactly what the detector stage reported
        cg, eg = out["corners"][i], out["edgelets"][i]
        exp_px = np.concatenate([np.stack([cg["x"], cg["y"]], 1)[cg["score"] > 10], np.stack([eg["x"], eg["y"]], 1)[eg["score"] > 100]])
        assert np.array_equal(ft_all["px"][sel], exp_px.astype(np.float64))
        # oracle frames: the new left / right images, left pose = what the aligner produced (T_f_w of camera 0)
        T_f_w0 = out["align_T_f_w"][i][0]
        T_imu_world = synth.se3_mul(synth.se3_inv(sc["T_cam_imu"][0]), T_f_w0)
        dq, dt = helpers.pose_diff(synth.se3_inv(T_f_w0), out["stereo_Twc"][i])
        assert dq < 1e-12 and dt < 1e-12
        f0 = orc.make_frame(orc.create_img_pyramid(sc["imgs"]["c0"], 5), sc["cam"], sc["T_cam_imu"][0], T_imu_world, keep=keep)
        f1 = orc.make_frame(orc.create_img_pyramid(sc["imgs"]["c1"], 5), sc["cam"], sc["T_cam_imu"][1], T_imu_world, keep=keep)
        ft = ft_all[sel]
        oft = orc.make_features(ft["px"], ft["f"], ft["grad"], ft["type"], ft["level"])
        o, ns, nf = orc.stereo_triangulate(f0, f1, oft, 120)
        r = res_all[sel]
        for k in ("status", "slot", "match_result", "level", "type"):
            assert np.array_equal(r[k], o[k]), (i, k)
        ok = r["status"] == 2
        np.testing.assert_allclose(r["px_cur"][ok], o["px_cur"][ok], rtol=0, atol=1e-3)
        np.testing.assert_allclose(r["xyz_world"][ok], o["xyz_world"][ok], rtol=1e-4, atol=1e-6)
        assert out["stereo_stats"]["n_succeeded"][i] == ns and out["stereo_stats"]["n_failed"][i] == nf
        n_tri += ns
    assert n_tri > 200
