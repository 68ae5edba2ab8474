"""GPU parity tests, row f1: svo_cuda_reproject_match (getCandidate -> candidate sort -> per-cell matchCandidate queues -> stop
position / slots / occupancy / statistics) against the oracle and against the outputs of the REFERENCE's own compiled
reprojector.cpp (tests/golden/reproject_ref_golden.npz). Integer outputs (statuses, candidate order, slots, levels, types,
landmark counters, statistics, occupancy) are bit-exact; sub-pixel positions within 1e-3 px, seed states within 1e-4 relative."""
import os

import numpy as np
import pytest

import helpers
from svo_pro_universal_b200 import capi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reproject_ref_golden.npz")
PX_TOL, REL_TOL = 1e-3, 1e-4


def _gpu_tables(sc):
    t = dict(sc["tables"])
    t["feat"] = capi.make_features(t["feat"]["px"], t["feat"]["f"], t["feat"]["grad"], t["feat"]["type"], t["feat"]["level"])
    for k in ("kf_T_f_w", "kf_seed_mu_range", "feat_score", "feat_seed_state", "pt_pos"):
        t[k] = np.ascontiguousarray(t[k], np.float64)
    for k in ("feat_point", "feat_kf", "pt_n_failed", "pt_n_succeeded", "pt_obs_begin", "obs_feat"):
        t[k] = np.ascontiguousarray(t[k], np.int32)
    return t


def _run_case(ctx, case, device=False):
    seed, max_n, n_in, occ_frac, by_obs, kw = case
    sc, occ = helpers.reproject_case_inputs(case)
    K = len(sc["kf_imgs"])
    ref = capi.Pyramid(ctx, K, 752, 480, 5)
    cur = capi.Pyramid(ctx, 1, 752, 480, 5)
    ref.upload(np.stack(sc["kf_imgs"])); cur.upload(sc["cur_img"])
    ref.build(); cur.build()
    cam = capi.Camera.from_dict(sc["cam"])
    opt = capi.reprojector_options(max_n_features=max_n, sort_by_num_obs=by_obs, px_error_angle=helpers.reproject_px_error_angle(sc["cam"]))
    t = _gpu_tables(sc)
    ef = np.ascontiguousarray(sc["entry_feat"], np.int32)
    args = dict(cur_T_f_w=np.ascontiguousarray(sc["cur_T_f_w"], np.float64).reshape(1, 7), n_features_in=np.array([n_in], np.int32),
                entry_begin=np.array([0, len(ef)], np.int32), entry_feat=ef, occupancy=occ.reshape(1, -1).copy())
    if device:
        import torch
        dev = torch.device("cuda", 0)
        tt = {k: (torch.from_numpy(v.view(np.uint8) if v.dtype.fields else v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in t.items()}
        args = {k: torch.from_numpy(v).to(dev) for k, v in args.items()}
        res, st = capi.reproject_match(ctx, ref, cur, cam, cam, tt, args["cur_T_f_w"], args["n_features_in"], args["entry_begin"],
                                       args["entry_feat"], args["occupancy"], opt)
        ctx.synchronize()
        res = res.cpu().numpy().view(capi.REPROJ_RESULT_DTYPE)[:len(ef)]
        st = st.cpu().numpy().view(capi.REPROJ_STATS_DTYPE)
        occ_out = args["occupancy"].cpu().numpy()[0]
    else:
        res, st = capi.reproject_match(ctx, ref, cur, cam, cam, t, args["cur_T_f_w"], args["n_features_in"], args["entry_begin"],
                                       args["entry_feat"], args["occupancy"], opt)
        occ_out = args["occupancy"][0]
    return res, st[0], occ_out


def _as_outputs(per_case):
    out = {}
    for ci, (res, st, occ) in enumerate(per_case):
        for k in helpers.REPROJ_INT_FIELDS + helpers.REPROJ_FLOAT_FIELDS:
            out[f"c{ci}_{k}"] = res[k]
        out[f"c{ci}_stats"] = np.array([st["n_candidates"], st["n_trials"], st["n_matches"], st["n_consumed"]])
        out[f"c{ci}_occ"] = occ
    return out


@pytest.fixture(scope="module")
def gpu_outputs(ctx):
    return _as_outputs([_run_case(ctx, c) for c in helpers.REPROJECT_CASES])


def test_reproject_match_equals_oracle(gpu_outputs, orc):
    mine = helpers.reproject_outputs(orc, "orc")
    helpers.assert_reproject_equal(gpu_outputs, mine, PX_TOL, REL_TOL, "oracle")
    for ci in range(len(helpers.REPROJECT_CASES)):   # the Matcher::MatchResult of every attempt (the reference cannot report it)
        assert np.array_equal(gpu_outputs[f"c{ci}_status"] >= 3, mine[f"c{ci}_status"] >= 3)


def test_reproject_match_equals_reference_golden(gpu_outputs):
    helpers.assert_reproject_equal(gpu_outputs, np.load(GOLD), PX_TOL, REL_TOL, "reference golden")


def test_reproject_match_device_arrays(ctx, gpu_outputs):
    """SVO_MEM_DEVICE: tables, entries, occupancy and results resident in HBM — same bytes as the host-array path."""
    for ci in (0, 3):
        res, st, occ = _run_case(ctx, helpers.REPROJECT_CASES[ci], device=True)
        for k in helpers.REPROJ_INT_FIELDS + helpers.REPROJ_FLOAT_FIELDS:
            assert np.array_equal(res[k], gpu_outputs[f"c{ci}_{k}"]), (ci, k)
        assert np.array_equal(occ, gpu_outputs[f"c{ci}_occ"])


def test_reproject_match_batch_of_frames_and_ties(ctx, orc):
    """F = 3 current frames sharing one map in ONE call (different poses, feature counts and grids), with integer scores so
    that many candidates compare equal: ties keep the visiting order (stable), exactly like the oracle's std::stable_sort."""
    sc = synth.make_reproject_scene(11, n_cur=3, integer_scores=True)
    K = len(sc["kf_imgs"])
    ref = capi.Pyramid(ctx, K, 752, 480, 5)
    cur = capi.Pyramid(ctx, 3, 752, 480, 5)
    ref.upload(np.stack(sc["kf_imgs"])); cur.upload(np.stack(sc["cur_imgs"]))
    ref.build(); cur.build()
    cam = capi.Camera.from_dict(sc["cam"])
    ang = helpers.reproject_px_error_angle(sc["cam"])
    opt = capi.reprojector_options(max_n_features=90, px_error_angle=ang)
    ef = np.ascontiguousarray(sc["entry_feat"], np.int32)
    E = len(ef)
    rng = np.random.default_rng(2)
    occ = (rng.uniform(size=(3, 416)) < 0.15).astype(np.uint8)
    n_in = np.array([0, 25, 80], np.int32)
    subsets = [ef, ef[::2].copy(), ef[: E // 3].copy()]
    entry_begin = np.concatenate([[0], np.cumsum([len(s) for s in subsets])]).astype(np.int32)
    occ_gpu = occ.copy()
    res, st = capi.reproject_match(ctx, ref, cur, cam, cam, _gpu_tables(sc), np.ascontiguousarray(sc["cur_Ts"], np.float64), n_in,
                                   entry_begin, np.concatenate(subsets), occ_gpu, opt)
    keep = []
    kfs = [orc.make_frame(orc.create_img_pyramid(im, 5), sc["cam"], helpers.IDENTITY7, T, keep=keep)
           for im, T in zip(sc["kf_imgs"], sc["tables"]["kf_T_f_w"])]
    oopt = orc.ReprojOptions(30, 90, 1, 0, 0, 200.0, ang)
    for j in range(3):
        cf = orc.make_frame(orc.create_img_pyramid(sc["cur_imgs"][j], 5), sc["cam"], helpers.IDENTITY7, sc["cur_Ts"][j], keep=keep)
        o = occ[j].copy()
        r, s = orc.reproject_match(kfs, sc["tables"], cf, subsets[j], int(n_in[j]), o, oopt)
        g = res[entry_begin[j]:entry_begin[j + 1]]
        for k in helpers.REPROJ_INT_FIELDS:
            assert np.array_equal(g[k], r[k]), (j, k)
        assert np.abs(g["px"] - r["px"]).max() < PX_TOL
        assert (st[j]["n_candidates"], st[j]["n_trials"], st[j]["n_matches"], st[j]["n_consumed"]) == \
               (s["n_candidates"], s["n_trials"], s["n_matches"], s["n_consumed"])
        assert np.array_equal(occ_gpu[j], o)
        assert s["n_matches"] > 5


def test_progressive_matching_quotas_and_second_pass(ctx, orc):
    """The matching is progressive (pass 1: quota + 25 % + 8 cells per frame in list order; a progress kernel lists the unvisited cells
    that can still lie before the stop; pass 2). One call with eight frames whose quotas run from 1 to 120 over a large-motion scene
    (many failed attempts, so pass 1 alone does not fill the larger quotas) and with occupied cells in front of the list: every entry's
    status, order, slot, counters and the statistics equal the oracle's sequential walk."""
    sc = synth.make_reproject_scene(6, n_cur=2, max_rot_deg=9.0, max_trans=0.5)
    K = len(sc["kf_imgs"])
    ref = capi.Pyramid(ctx, K, 752, 480, 5)
    cur = capi.Pyramid(ctx, 2, 752, 480, 5)
    ref.upload(np.stack(sc["kf_imgs"])); cur.upload(np.stack(sc["cur_imgs"]))
    ref.build(); cur.build()
    cam = capi.Camera.from_dict(sc["cam"])
    ang = helpers.reproject_px_error_angle(sc["cam"])
    max_n = 120
    opt = capi.reprojector_options(max_n_features=max_n, px_error_angle=ang)
    ef = np.ascontiguousarray(sc["entry_feat"], np.int32)
    n_in = np.array([0, 0, 60, 60, 110, 119, 150, 30], np.int32)   # quotas 120, 120, 60, 60, 10, 1, 1 (already full), 90
    F = len(n_in)
    fidx = (np.arange(F) % 2).astype(np.int32)
    rng = np.random.default_rng(4)
    occ = (rng.uniform(size=(F, 416)) < np.array([0.0, 0.5, 0.0, 0.3, 0.0, 0.0, 0.2, 0.7])[:, None]).astype(np.uint8)
    entry_begin = (np.arange(F + 1) * len(ef)).astype(np.int32)
    occ_gpu = occ.copy()
    res, st = capi.reproject_match(ctx, ref, cur, cam, cam, _gpu_tables(sc), np.ascontiguousarray(sc["cur_Ts"][fidx], np.float64), n_in,
                                   entry_begin, np.tile(ef, F), occ_gpu, opt, cur_frame_idx=fidx)
    keep = []
    kfs = [orc.make_frame(orc.create_img_pyramid(im, 5), sc["cam"], helpers.IDENTITY7, T, keep=keep)
           for im, T in zip(sc["kf_imgs"], sc["tables"]["kf_T_f_w"])]
    oopt = orc.ReprojOptions(30, max_n, 1, 0, 0, 200.0, ang)
    second_pass_needed = 0
    for j in range(F):
        cf = orc.make_frame(orc.create_img_pyramid(sc["cur_imgs"][fidx[j]], 5), sc["cam"], helpers.IDENTITY7, sc["cur_Ts"][fidx[j]], keep=keep)
        o = occ[j].copy()
        r, s = orc.reproject_match(kfs, sc["tables"], cf, ef, int(n_in[j]), o, oopt)
        g = res[entry_begin[j]:entry_begin[j + 1]]
        for k in helpers.REPROJ_INT_FIELDS:
            assert np.array_equal(g[k], r[k]), (j, k, np.flatnonzero(g[k] != r[k])[:8])
        assert np.abs(g["px"] - r["px"]).max() < PX_TOL
        assert (st[j]["n_candidates"], st[j]["n_trials"], st[j]["n_matches"], st[j]["n_consumed"]) == \
               (s["n_candidates"], s["n_trials"], s["n_matches"], s["n_consumed"]), j
        assert np.array_equal(occ_gpu[j], o)
        quota = max(1, max_n - int(n_in[j]))
        second_pass_needed += int(s["n_trials"] + int((r["status"] == capi.REPROJ_SKIPPED).sum()) > quota + quota // 4 + 8)
    assert second_pass_needed >= 2, "the case must exercise the second pass"


def test_reproject_match_rejects_bad_arguments(ctx):
    sc = synth.make_reproject_scene(3, n_kfs=1, n_per_kf=20)
    ref = capi.Pyramid(ctx, 1, 752, 480, 5); cur = capi.Pyramid(ctx, 1, 752, 480, 5)
    cam = capi.Camera.from_dict(sc["cam"])
    ef = np.ascontiguousarray(sc["entry_feat"], np.int32)
    with pytest.raises(capi.SvoCudaError):   # n_entries != entry_begin[F]
        capi.reproject_match(ctx, ref, cur, cam, cam, _gpu_tables(sc), np.zeros((1, 7)), np.zeros(1, np.int32),
                             np.array([0, len(ef) - 1], np.int32), ef, np.zeros((1, 416), np.uint8), capi.reprojector_options())
    with pytest.raises(capi.SvoCudaError):   # cell size that yields more than 4096 cells
        capi.reproject_match(ctx, ref, cur, cam, cam, _gpu_tables(sc), np.zeros((1, 7)), np.zeros(1, np.int32),
                             np.array([0, len(ef)], np.int32), ef, np.zeros((1, 416), np.uint8), capi.reprojector_options(cell_size=4))
