"""Small driver for ncu captures: every hot-path kernel launched 3 times at one-wave-or-more sizes (no CPU oracle work).
    ncu --set full -k regex:<kernel> -s 1 -c 1 ... python tools/prof_paths.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svo_pro_universal_b200 import capi, synth, batch  # noqa: E402

# PROF_SMALL=1: every path at a few CTAs (tools/sanitize.sh runs this under compute-sanitizer, which slows kernels 10-100x)
SMALL = os.environ.get("PROF_SMALL", "0") == "1"
REPS = 1 if SMALL else 3
ctx = capi.Context(0)

# (a)+(b): B pairs, pyramid + FAST + sparse alignment
B = int(os.environ.get("PROF_PAIRS", "12" if SMALL else "1184"))
uniq = [synth.make_align_pair(5000 + s) for s in range(8)]
pk = batch.tile_batch(batch.pack_align_batch(uniq, max_features=180), B)
ref = capi.Pyramid(ctx, B, 752, 480, 5); cur = capi.Pyramid(ctx, B, 752, 480, 5)
ref.upload(pk["ref_imgs"]); cur.upload(pk["cur_imgs"]); ref.build()
cam = capi.Camera.from_dict(uniq[0]["cam"])
for _ in range(REPS):
    cur.build()
    res = capi.sparse_align(ctx, [ref], [cur], [cam], pk["T_cam_imu"], pk["T_imu_world_ref"], pk["T_imu_world_cur"], pk["n_features"],
                            pk["px"], pk["f"], pk["depth"], pk["eligible"], capi.sparse_align_options())
print("align iters", res["iters"][:2].tolist(), "n", res["n_tracked"][:2])
del ref
# (a) detector on BASELINE configs[1]-shaped frames (the images bench_kernels.py times)
uimg = np.stack([synth.make_image(200 + s) for s in range(16)])
cur.upload(uimg[np.arange(B) % 16])
for _ in range(REPS):
    capi.fast_detect(ctx, cur, capi.detector_options(), fused_pyramid=True)
for _ in range(REPS):
    capi.fast_detect(ctx, cur, capi.detector_options())
for _ in range(REPS):  # (f2): edgelet detector alone, then FastGrad (FAST + merge + edgelets)
    capi.edgelet_detect(ctx, cur, 100, 8, 30)
for _ in range(REPS):
    capi.fastgrad_detect(ctx, cur, capi.detector_options(), 100)
del cur
if os.environ.get("PROF_UNTIL") == "detect":  # pyramid + alignment + detectors only (a quick sanitizer pass after a detector change)
    sys.exit(0)

# (c): matcher paths
NP, NF, NU = (2, 300, 2) if SMALL else (64, 2000, 4)
sets = [synth.make_match_set(300 + s, n_features=NF) for s in range(NU)]
ref = capi.Pyramid(ctx, NU, 752, 480, 5); cur = capi.Pyramid(ctx, NU, 752, 480, 5)
ref.upload(np.stack([m["ref_img"] for m in sets])); cur.upload(np.stack([m["cur_img"] for m in sets]))
ref.build(); cur.build()
pid = np.arange(NP) % NU
cat = lambda k: np.concatenate([sets[i][k] for i in pid])
ft = capi.make_features(cat("px"), cat("f"), cat("grad"), cat("type"), cat("level"))
fidx = np.concatenate([np.full(len(sets[i]["px"]), i, np.int32) for i in pid])
T = np.stack([m["T_cur_ref"] for m in sets])
mopt = capi.matcher_options()
for _ in range(REPS):
    r = capi.find_match_direct(ctx, ref, cur, cam, cam, T, ft, cat("depth"), cat("px_guess"), mopt, ref_frame_idx=fidx, cur_frame_idx=fidx, T_idx=fidx)
print("findMatchDirect success", float((r["result"] == 0).mean()))
rng = np.random.default_rng(1)
inv = 1.0 / cat("depth"); est = inv * rng.uniform(0.7, 1.4, len(ft)); spread = rng.uniform(0.1, 0.8, len(ft)) * inv
dinv = np.stack([est, est + spread, np.maximum(est - spread, 1e-8)], 1)
for _ in range(REPS):
    r = capi.find_epipolar_match_direct(ctx, ref, cur, cam, cam, T, ft, dinv, mopt, ref_frame_idx=fidx, cur_frame_idx=fidx, T_idx=fidx)
print("epipolar success", float((r["result"] == 0).mean()))
if SMALL:  # the grouped work order of large calls (>= 16384 features: counting sort of the feature indices) under the sanitizer as well
    rep_o = -(-16400 // len(ft))
    big = lambda a: np.ascontiguousarray(np.concatenate([a] * rep_o)[:16400])
    rb = capi.find_match_direct(ctx, ref, cur, cam, cam, T, big(ft), big(cat("depth")), big(cat("px_guess")), mopt, ref_frame_idx=big(fidx),
                                cur_frame_idx=big(fidx), T_idx=big(fidx))
    re_ = capi.find_epipolar_match_direct(ctx, ref, cur, cam, cam, T, big(ft), big(dinv), mopt, ref_frame_idx=big(fidx), cur_frame_idx=big(fidx), T_idx=big(fidx))
    print("ordered large call: results equal the small call's", bool(np.array_equal(rb["result"][:len(ft)], capi.find_match_direct(
        ctx, ref, cur, cam, cam, T, ft, cat("depth"), cat("px_guess"), mopt, ref_frame_idx=fidx, cur_frame_idx=fidx, T_idx=fidx)["result"])),
        bool(np.array_equal(re_["result"][:len(ft)], r["result"])))
# Matcher::scanEpipolarLine on its own: the long segments of the call above with their warped patches (svo_cuda_warp_affine)
long_ = np.flatnonzero(r["epi_length_pyramid"] >= 2.0)[: (64 if SMALL else 20000)]
if len(long_):
    Aw, slw, pwb, okw = capi.warp_affine(ctx, ref, cam, cam, T, ft[long_], 1.0 / np.maximum(dinv[long_, 0], 1e-6), ref_frame_idx=fidx[long_], T_idx=fidx[long_])
    Rs = np.stack([synth.se3_to_Rt(T[i])[0] for i in range(NU)]); ts = np.stack([synth.se3_to_Rt(T[i])[1] for i in range(NU)])
    fr = np.stack([ft["f"][i] for i in long_]); Rf = np.einsum("nij,nj->ni", Rs[fidx[long_]], fr); tt = ts[fidx[long_]]
    for _ in range(REPS):
        best, zb = capi.scan_epipolar_line(ctx, cur, cam, Rf + tt * dinv[long_, 1:2], Rf + tt * dinv[long_, 2:3], Rf + tt * dinv[long_, 0:1],
                                           np.ascontiguousarray(pwb.reshape(-1, 10, 10)[:, 1:9, 1:9]), r["search_level"][long_],
                                           r["epi_length_pyramid"][long_], mopt, cur_frame_idx=fidx[long_])
    print("stand-alone scans", len(long_), "below threshold", float((zb < 2000 * 64).mean()))
# fast:: leaves with list-shaped results on one level (device-side ordered compaction)
for _ in range(REPS):
    xy_l, sc_l, nm_l = capi.fast_corner_list(ctx, ref, 0, 0, 10, 10)
    sc2 = capi.fast_corner_score(ctx, ref, 0, 0, xy_l, 10, 10)
    nm2 = capi.fast_nonmax_3x3(ctx, xy_l, sc_l)
print("corner list", len(xy_l), "scores equal", bool(np.array_equal(sc_l, sc2)), "non-max equal", bool(np.array_equal(np.flatnonzero(nm_l), nm2)))
del ref, cur

# (d): depth filter
S = 2000 if SMALL else 50000
state = np.tile(np.array([0.25, (1 / 1.5) ** 2 / 36.0, 10.0, 10.0]), (S, 1))
z = 0.25 + rng.normal(size=S) * 0.01
for _ in range(REPS):
    capi.update_filter_vogiatzis(ctx, z, np.full(S, 1e-4), np.full(S, 1 / 1.5), state)
if hasattr(capi.lib(), "svo_cuda_update_filter_seq"):
    for _ in range(REPS):
        capi.update_filter_seq(ctx, np.ascontiguousarray(np.broadcast_to(z, (16, S))), np.full((16, S), 1e-4), np.full(S, 1 / 1.5), state)
        capi.update_filter_seq(ctx, np.ascontiguousarray(np.broadcast_to(z, (4, S))), np.full((4, S), 1e-4), None, state, gaussian=True)
q = synth.make_seed_sequence(400, n_seeds=400, n_obs=8)
n = len(q["px"])
ref = capi.Pyramid(ctx, 1, 752, 480, 5); cur = capi.Pyramid(ctx, 8, 752, 480, 5)
ref.upload(q["ref_img"][None]); cur.upload(np.stack(q["cur_imgs"])); ref.build(); cur.build()
rep = 1 if SMALL else 32
ftq = capi.make_features(*(np.concatenate([q[k]] * rep) for k in ("px", "f", "grad")), np.concatenate([q["type"].astype(np.int32)] * rep),
                         np.concatenate([q["level"]] * rep))
obs = np.tile(np.arange(8, dtype=np.int32)[:, None], (1, n * rep))
for _ in range(REPS):
    ty = np.concatenate([q["type"].astype(np.uint8)] * rep); st = np.concatenate([q["state"]] * rep)
    ns, _ = capi.update_seeds(ctx, ref, cur, cam, cam, ftq, ty, st, np.full(n * rep, q["mu_range"]), obs, obs, q["T_cur_ref"], mopt,
                              capi.depth_filter_options(), want_match_results=False)
print("update_seeds successes", ns)
del ref, cur

# (f1): reprojector, 296 current frames sharing one map
sc = synth.make_reproject_scene(21, n_cur=8)
K, F = len(sc["kf_imgs"]), (3 if SMALL else 296)
ref = capi.Pyramid(ctx, K, 752, 480, 5); cur = capi.Pyramid(ctx, 8, 752, 480, 5)
ref.upload(np.stack(sc["kf_imgs"])); cur.upload(np.stack(sc["cur_imgs"])); ref.build(); cur.build()
tb = dict(sc["tables"])
tb["feat"] = capi.make_features(tb["feat"]["px"], tb["feat"]["f"], tb["feat"]["grad"], tb["feat"]["type"], tb["feat"]["level"])
ef = np.ascontiguousarray(sc["entry_feat"], np.int32)
for _ in range(REPS):
    res, st = capi.reproject_match(ctx, ref, cur, cam, cam, tb, np.ascontiguousarray(sc["cur_Ts"][np.arange(F) % 8]), np.zeros(F, np.int32),
                                   (np.arange(F + 1) * len(ef)).astype(np.int32), np.tile(ef, F), np.zeros((F, 416), np.uint8),
                                   capi.reprojector_options(max_n_features=120), cur_frame_idx=(np.arange(F) % 8).astype(np.int32))
print("reproject matches/frame", float(st["n_matches"].mean()), "trials/frame", float(st["n_trials"].mean()))
del ref, cur

# (f4): pose optimizer, 4736 bundles
pcs = [synth.make_pose_opt_case(40 + s) for s in range(8)]
BP = 24 if SMALL else 4736
pidx = np.arange(BP) % 8
pft = [capi.make_features(c["px"], c["f"], c["grad"], c["type"], c["level"]) for c in pcs]
pbeg = np.concatenate([[0], np.cumsum([len(pft[i]) for i in pidx])]).astype(np.int32)
for _ in range(REPS):
    pres, _ = capi.pose_optimize(ctx, [capi.Camera.from_dict(pcs[0]["cam"])], np.stack(pcs[0]["T_cam_imu"]),
                                 np.stack([pcs[i]["T_imu_world_init"] for i in pidx]), pbeg, np.concatenate([pft[i] for i in pidx]), None,
                                 np.concatenate([pcs[i]["xyz_world"] for i in pidx]), np.concatenate([pcs[i]["has_xyz"] for i in pidx]),
                                 capi.pose_optimizer_options())
print("pose optimizer iterations", float(pres["iters"].mean()), "measurements", float(pres["n_meas"].mean()))

# (f3): StereoTriangulation — 256 stereo pairs x ~390 detected features (FastGrad on the device), progressive epipolar matching + commit
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
prs = [helpers.stereo_case(81), helpers.stereo_case(82)]
BS = 4 if SMALL else 256
sidS = np.arange(BS) % 2
q0 = capi.Pyramid(ctx, 2, 752, 480, 5); q1 = capi.Pyramid(ctx, 2, 752, 480, 5)
q0.upload(np.stack([p[0]["ref_img"] for p in prs])); q1.upload(np.stack([p[1]["ref_img"] for p in prs])); q0.build(); q1.build()
cS, eS = capi.fastgrad_detect(ctx, q0, capi.detector_options(), 100)
ftU = []
for k in range(2):
    sel = [(cS[k], 10.0, 7), (eS[k], 100.0, 6)]
    px = np.concatenate([np.stack([a["x"], a["y"]], 1)[a["score"] > t] for a, t, _ in sel]).astype(np.float64)
    ang = np.concatenate([a["angle"][a["score"] > t] for a, t, _ in sel])
    lv = np.concatenate([a["level"][a["score"] > t] for a, t, _ in sel])
    ty = np.concatenate([np.full(int((a["score"] > t).sum()), v, np.int32) for a, t, v in sel])
    fS = synth.cam_backproject(prs[k][0]["cam"], px); fS /= np.linalg.norm(fS, axis=1, keepdims=True)
    ftU.append(capi.make_features(px, fS, np.stack([np.cos(ang), np.sin(ang)], 1).astype(np.float64), ty, lv))
begS = np.concatenate([[0], np.cumsum([len(ftU[i]) for i in sidS])]).astype(np.int32)
ftS = np.concatenate([ftU[i] for i in sidS])
TwcS = np.stack([synth.se3_inv(synth.se3_mul(prs[i][0]["T_cam_imu"], prs[i][0]["T_imu_world_ref"])) for i in sidS])
T_f1f0 = synth.se3_mul(prs[0][1]["T_cam_imu"], synth.se3_inv(prs[0][0]["T_cam_imu"]))
camS = capi.Camera.from_dict(prs[0][0]["cam"])
for _ in range(REPS):
    rS, sS = capi.stereo_triangulate(ctx, q0, q1, camS, camS, T_f1f0, TwcS, begS, ftS, np.full(BS, 120, np.int32), np.zeros(BS, np.int32),
                                     capi.matcher_options(max_epi_search_steps=500, subpix_refinement=1),
                                     frame0_idx=sidS.astype(np.int32), frame1_idx=sidS.astype(np.int32))
print("stereo triangulated per pair", float(sS["n_succeeded"].mean()))

# (f4, second half): Point::optimize — 400 points x 128 copies
cP = helpers.point_opt_cases()
KP = 1 if SMALL else 128
nP, nO = len(cP["pos0"]), len(cP["obs_frame"])
posP = np.tile(cP["pos0"], (KP, 1))
begP = np.concatenate([[0], (cP["obs_begin"][1:][None, :] + nO * np.arange(KP)[:, None]).ravel()]).astype(np.int32)
for _ in range(REPS):
    pp = posP.copy()
    itP = capi.optimize_points(ctx, pp, begP, np.tile(cP["obs_frame"], KP), np.tile(cP["obs_f"], (KP, 1)), cP["T_f_w"], 5, False)
print("point optimizer iterations", float(itP.mean()))
