"""Two svo_cuda contexts in ONE process — on devices 0 and 1 when the box has two GPUs, else both on device 0 — without the caller
ever touching cudaSetDevice: every entry point binds its own context's device, and per-kernel attributes (dynamic shared memory of the
pyramid and Reprojector sort kernels, the edgelet angle table) are set per context. Both contexts must give the oracle's answers."""
import numpy as np
import pytest

import helpers
from svo_pro_universal_b200 import capi, synth

pytestmark = pytest.mark.gpu


def test_two_contexts_one_process(orc):
    n_dev = capi.lib().svo_cuda_device_count()
    devices = [0, 1] if n_dev >= 2 else [0, 0]
    ctxs = [capi.Context(d) for d in devices]
    img = synth.make_image(321)
    opyr = orc.create_img_pyramid(img, 5)
    ocorners = orc.fast_detector(img)
    oedge = orc.edgelet_detector_v2(opyr, 100, 8, 30, (ocorners["score"] > 10).astype(np.uint8))
    d = synth.make_align_pair(9)
    # the Reprojector path (its sort kernel needs the large dynamic shared-memory attribute on every device)
    case = helpers.REPROJECT_CASES[0]
    sc, occ0 = helpers.reproject_case_inputs(case)
    outs = []
    for ctx in reversed(ctxs):   # the second context first: nothing may depend on which device was used first
        pyr = capi.Pyramid(ctx, 1, 752, 480, 5)
        pyr.upload(img); pyr.build()
        for l in range(5):
            assert np.array_equal(pyr.download(0, l), opyr[l]), (ctx.device, l)
        fast, edge = capi.fastgrad_detect(ctx, pyr, capi.detector_options(), 100)
        for k in ("x", "y", "level", "score"):
            assert np.array_equal(fast[0][k], ocorners[k]), (ctx.device, k)
        for k in ("x", "y", "level", "score", "angle"):
            assert np.array_equal(edge[0][k], oedge[k]), (ctx.device, k)
        res, _, _ = helpers.gpu_align(ctx, [d], capi.sparse_align_options())
        K = len(sc["kf_imgs"])
        ref = capi.Pyramid(ctx, K, 752, 480, 5); cur = capi.Pyramid(ctx, 1, 752, 480, 5)
        ref.upload(np.stack(sc["kf_imgs"])); cur.upload(sc["cur_img"]); ref.build(); cur.build()
        tb = dict(sc["tables"])
        tb["feat"] = capi.make_features(tb["feat"]["px"], tb["feat"]["f"], tb["feat"]["grad"], tb["feat"]["type"], tb["feat"]["level"])
        ef = np.ascontiguousarray(sc["entry_feat"], np.int32)
        cam = capi.Camera.from_dict(sc["cam"])
        occ = occ0.copy().reshape(1, -1)
        r, st = capi.reproject_match(ctx, ref, cur, cam, cam, tb, np.ascontiguousarray(sc["cur_T_f_w"]).reshape(1, 7), np.array([case[2]], np.int32),
                                     np.array([0, len(ef)], np.int32), ef, occ, capi.reprojector_options(max_n_features=case[1], sort_by_num_obs=case[4]))
        outs.append((res[0]["T_icur_iref"].copy(), r["status"].copy(), r["slot"].copy(), int(st["n_matches"][0])))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2]) and outs[0][3] == outs[1][3] > 20
