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
