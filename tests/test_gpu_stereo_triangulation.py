"""GPU parity tests, row f3 (StereoTriangulation::compute): svo_cuda_stereo_triangulate against the oracle entry by entry and against
what the reference's own compiled compute() left in frame1 (tests/golden/stereo_tri_ref_golden.npz), through the C ABI."""
import os

import numpy as np
import pytest

import helpers
from svo_pro_universal_b200 import capi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _setup(ctx, orc, cases, g, case_ids):
    """Pyramids of all cases' frames + concatenated entries in the reference's visiting order."""
    B = len(cases)
    pyr0, pyr1 = capi.Pyramid(ctx, B, 752, 480, 5), capi.Pyramid(ctx, B, 752, 480, 5)
    imgs0, imgs1, ftrs, begin, Twc, orders, oracle = [], [], [], [0], [], [], []
    for case, i in zip(cases, case_ids):
        keep = []
        d, s1, p0, p1, f0, f1 = helpers.stereo_tri_frames(orc, case, keep)
        order = g[f"order_{i}"]
        det, f = helpers.stereo_tri_entries(orc, case, d, p0, order)
        ftrs.append(capi.make_features(det["px"][order], f, det["grad"][order], det["type"][order], det["level"][order]))
        oft = orc.make_features(det["px"][order], f, det["grad"][order], det["type"][order], det["level"][order])
        oracle.append(orc.stereo_triangulate(f0, f1, oft, case[3], 0, case[4], case[5], case[6]))
        begin.append(begin[-1] + len(order))
        T_c0_w = synth.se3_mul(d["T_cam_imu"], d["T_imu_world_ref"])
        Twc.append(synth.se3_inv(T_c0_w))
        orders.append(order)
        imgs0.append(d["ref_img"]); imgs1.append(s1["ref_img"])
        T_f1f0 = synth.se3_mul(s1["T_cam_imu"], synth.se3_inv(d["T_cam_imu"]))
        cam = capi.Camera.from_dict(d["cam"])
    pyr0.upload(np.stack(imgs0)); pyr1.upload(np.stack(imgs1))
    pyr0.build(); pyr1.build()
    return pyr0, pyr1, cam, T_f1f0, np.stack(Twc), np.array(begin, np.int32), np.concatenate(ftrs), orders, oracle


def _check(res, stats, begin, orders, oracle, cases, case_ids, g):
    for b, (case, i) in enumerate(zip(cases, case_ids)):
        r = res[begin[b]:begin[b + 1]]
        o, ns, nf = oracle[b]
        for k in ("status", "slot", "match_result", "level", "type"):
            assert np.array_equal(r[k], o[k]), (i, k)
        ok = r["status"] == 2
        np.testing.assert_allclose(r["px_cur"][ok], o["px_cur"][ok], rtol=0, atol=1e-3)   # align1D / align2D within 1e-3 px (bit-equal in practice)
        np.testing.assert_allclose(r["depth"][ok], o["depth"][ok], rtol=1e-4)
        np.testing.assert_allclose(r["xyz_world"][ok], o["xyz_world"][ok], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(r["grad_cur"][ok], o["grad_cur"][ok], rtol=0, atol=1e-6)
        np.testing.assert_allclose(r["f_cur"][ok], o["f_cur"][ok], rtol=0, atol=1e-5)
        assert stats["n_succeeded"][b] == ns and stats["n_failed"][b] == nf
        helpers.assert_stereo_matches_reference(r, orders[b], g, i)


def test_stereo_triangulation_each_case(ctx, orc):
    g = np.load(os.path.join(GOLD, "stereo_tri_ref_golden.npz"))
    for i, case in enumerate(helpers.STEREO_TRI_CASES):
        pyr0, pyr1, cam, T_f1f0, Twc, begin, ftrs, orders, oracle = _setup(ctx, orc, [case], g, [i])
        mopt = capi.matcher_options(max_epi_search_steps=500, subpix_refinement=1)
        res, stats = capi.stereo_triangulate(ctx, pyr0, pyr1, cam, cam, T_f1f0, Twc, begin, ftrs, np.array([case[3]], np.int32),
                                             np.zeros(1, np.int32), mopt, case[4], case[5], case[6])
        _check(res, stats, begin, orders, oracle, [case], [i], g)


def test_stereo_triangulation_batch_on_device_arrays(ctx, orc):
    """Three pairs in one call, every array device-resident, non-zero first slots."""
    import torch
    g = np.load(os.path.join(GOLD, "stereo_tri_ref_golden.npz"))
    ids = [0, 1, 2]
    cases = [helpers.STEREO_TRI_CASES[i] for i in ids]
    pyr0, pyr1, cam, T_f1f0, Twc, begin, ftrs, orders, oracle = _setup(ctx, orc, cases, g, ids)
    mopt = capi.matcher_options(max_epi_search_steps=500, subpix_refinement=1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    want = np.array([c[3] for c in cases], np.int32)
    res, stats = capi.stereo_triangulate(ctx, pyr0, pyr1, cam, cam, T_f1f0, t(Twc), t(begin), t(ftrs.view(np.uint8)), t(want),
                                         t(np.array([7, 0, 3], np.int32)), mopt)
    ctx.synchronize()
    res = res.cpu().numpy().view(capi.STEREO_RESULT_DTYPE)
    stats = stats.cpu().numpy().view(capi.STEREO_STATS_DTYPE)
    for b, first in enumerate((7, 0, 3)):  # slots start at frame1's feature count
        r = res[begin[b]:begin[b + 1]]
        ok = r["status"] == 2
        assert np.array_equal(r["slot"][ok], first + np.arange(ok.sum()))
        r["slot"][ok] -= first
    _check(res, stats, begin, orders, oracle, cases, ids, g)


def test_stereo_triangulation_arguments(ctx):
    import ctypes as C
    pyr = capi.Pyramid(ctx, 1, 752, 480, 5)
    L = capi.lib()
    assert L.svo_cuda_stereo_triangulate(ctx._h, pyr._h, pyr._h, None, None, None, None, None, None, 1, None, 0, None, None, None,
                                         C.c_double(0.3), C.c_double(1.0), C.c_double(0.02), None, None, None, 0) == -1


def test_stereo_triangulation_ragged_and_empty_pairs(ctx, orc):
    """Pairs without entries and pairs that want nothing (n_desired = 0) sit between ordinary pairs in one call."""
    g = np.load(os.path.join(GOLD, "stereo_tri_ref_golden.npz"))
    case = helpers.STEREO_TRI_CASES[0]
    pyr0, pyr1, cam, T_f1f0, Twc, begin, ftrs, orders, oracle = _setup(ctx, orc, [case], g, [0])
    n = len(ftrs)
    # pair 0: ordinary; pair 1: no entries; pair 2: same entries, wants 0; pair 3: wants 5
    begin4 = np.array([0, n, n, 2 * n, 3 * n], np.int32)
    ft4 = np.concatenate([ftrs, ftrs, ftrs])
    idx = np.zeros(4, np.int32)
    mopt = capi.matcher_options(max_epi_search_steps=500, subpix_refinement=1)
    res, stats = capi.stereo_triangulate(ctx, pyr0, pyr1, cam, cam, T_f1f0, np.repeat(Twc, 4, 0), begin4, ft4, np.array([case[3], 7, 0, 5], np.int32),
                                         np.zeros(4, np.int32), mopt, case[4], case[5], case[6], frame0_idx=idx, frame1_idx=idx)
    o, ns, nf = oracle[0]
    assert np.array_equal(res["status"][:n], o["status"]) and stats["n_succeeded"][0] == ns and stats["n_failed"][0] == nf
    assert stats["n_succeeded"][1] == 0 and stats["n_failed"][1] == 0
    assert (res["status"][n:2 * n] == capi.STEREO_NOT_REACHED).all() and stats["n_succeeded"][2] == 0 and stats["n_failed"][2] == 0
    r3 = res[2 * n:]
    first5 = np.flatnonzero(o["match_result"] == 0)[:5]
    assert stats["n_succeeded"][3] == 5 and np.array_equal(np.flatnonzero(r3["status"] == 2), first5)
    assert (r3["status"][first5[-1] + 1:] == capi.STEREO_NOT_REACHED).all()
    # an empty batch is fine
    res0, stats0 = capi.stereo_triangulate(ctx, pyr0, pyr1, cam, cam, T_f1f0, Twc, np.zeros(2, np.int32), ftrs[:0], np.array([3], np.int32),
                                           np.zeros(1, np.int32), mopt)
    assert stats0["n_succeeded"][0] == 0 and len(res0) == 0
