"""GPU parity tests, row f4: svo_cuda_pose_optimize (one CTA per frame bundle runs the whole PoseOptimizer::run) against the
oracle and the outputs of the REFERENCE's own compiled pose_optimizer.cpp (tests/golden/pose_opt_ref_golden.npz). Poses within
1e-4 rad / 1e-4 m (measured ~1e-12: only the summation order of H and g differs), outlier flags, iteration counts and the
measurement count identical, MAD sigma and medians to float precision."""
import os

import numpy as np
import pytest

import helpers
from svo_pro_universal_b200 import capi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pose_opt_ref_golden.npz")


def _gpu_case(ctx, spec):
    c, prior = helpers.pose_opt_case(spec)
    ft = capi.make_features(c["px"], c["f"], c["grad"], c["type"], c["level"])
    cams = [capi.Camera.from_dict(c["cam"])] * spec[1]
    opt = capi.pose_optimizer_options(err_type=spec[2], prior_lambda=0.5)
    res, outl = capi.pose_optimize(ctx, cams, np.stack(c["T_cam_imu"]), c["T_imu_world_init"].reshape(1, 7).copy(),
                                   np.array([0, len(ft)], np.int32), ft, c["feat_cam"], np.ascontiguousarray(c["xyz_world"]), c["has_xyz"], opt,
                                   prior_q=None if prior is None else prior.reshape(1, 4))
    return c, res[0], outl


def test_pose_optimize_equals_oracle_and_reference(ctx, orc):
    gold = np.load(GOLD)
    mine = helpers.pose_opt_outputs(orc, "orc")
    for ci, spec in enumerate(helpers.POSE_OPT_CASES):
        c, r, outl = _gpu_case(ctx, spec)
        for name, ref in (("oracle", mine), ("reference golden", gold)):
            dq, dt = helpers.pose_diff(r["T_imu_world"], ref[f"p{ci}_T"])
            assert dq < 1e-4 and dt < 1e-4 and dq < 1e-9 and dt < 1e-9, (name, ci, dq, dt)
            assert r["n_meas_final"] == int(ref[f"p{ci}_n"]) and np.array_equal(outl, ref[f"p{ci}_outlier"]), (name, ci)
            st = ref[f"p{ci}_stats"]
            np.testing.assert_allclose([r["measurement_sigma"], r["reproj_error_before"], r["reproj_error_after"]], st[:3], rtol=1e-6)
            assert r["iters"] == int(st[3]), (name, ci)
        for cam_i in range(spec[1]):   # frame->T_f_w_ = T_cam_imu * T_imu_world
            dq, dt = helpers.pose_diff(r["T_f_w"][cam_i], synth.se3_mul(c["T_cam_imu"][cam_i], r["T_imu_world"]))
            assert dq < 1e-12 and dt < 1e-12


def test_pose_optimize_batch_device_arrays(ctx, orc):
    """B = 64 bundles (8 unique, tiled) in ONE launch with device-resident arrays; every bundle equals its single-bundle result."""
    import torch
    dev = torch.device("cuda", 0)
    specs = [(s, 1, 0, False, False) for s in range(20, 28)]
    cases = [helpers.pose_opt_case(s)[0] for s in specs]
    B = 64
    idx = np.arange(B) % 8
    fts = [capi.make_features(c["px"], c["f"], c["grad"], c["type"], c["level"]) for c in cases]
    begin = np.concatenate([[0], np.cumsum([len(fts[i]) for i in idx])]).astype(np.int32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    res, outl = capi.pose_optimize(ctx, [capi.Camera.from_dict(cases[0]["cam"])], np.stack(cases[0]["T_cam_imu"]),
                                   t(np.stack([cases[i]["T_imu_world_init"] for i in idx])), t(begin),
                                   t(np.concatenate([fts[i] for i in idx]).view(np.uint8)), None,
                                   t(np.concatenate([cases[i]["xyz_world"] for i in idx])), t(np.concatenate([cases[i]["has_xyz"] for i in idx])),
                                   capi.pose_optimizer_options())
    ctx.synchronize()
    res = res.cpu().numpy().view(capi.POSE_OPT_RESULT_DTYPE)
    outl = outl.cpu().numpy()
    for b in range(B):
        n, T, o, st = orc.pose_optimize(cases[idx[b]], orc.pose_opt_options())
        dq, dt = helpers.pose_diff(res[b]["T_imu_world"], T)
        assert dq < 1e-9 and dt < 1e-9 and res[b]["n_meas_final"] == n
        assert np.array_equal(outl[begin[b]:begin[b + 1]], o)


def test_pose_optimize_rejects_bad_arguments(ctx):
    c, _ = helpers.pose_opt_case(helpers.POSE_OPT_CASES[0])
    ft = capi.make_features(c["px"], c["f"], c["grad"], c["type"], c["level"])
    cam = [capi.Camera.from_dict(c["cam"])]
    with pytest.raises(capi.SvoCudaError):   # n_features != feat_begin[B]
        capi.pose_optimize(ctx, cam, np.stack(c["T_cam_imu"]), c["T_imu_world_init"].reshape(1, 7).copy(), np.array([0, len(ft) - 1], np.int32),
                           ft, c["feat_cam"], np.ascontiguousarray(c["xyz_world"]), c["has_xyz"], capi.pose_optimizer_options())
    with pytest.raises(capi.SvoCudaError):   # unknown error type
        capi.pose_optimize(ctx, cam, np.stack(c["T_cam_imu"]), c["T_imu_world_init"].reshape(1, 7).copy(), np.array([0, len(ft)], np.int32),
                           ft, c["feat_cam"], np.ascontiguousarray(c["xyz_world"]), c["has_xyz"], capi.pose_optimizer_options(err_type=5))
