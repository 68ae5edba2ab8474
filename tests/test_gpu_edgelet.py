"""GPU parity tests, row f2: the edgelet detector (Gaussian 3x3 -> Scharr -> score -> neighbour test -> cell arg-max -> angle
histogram) and the FastGrad combination, BIT-EXACT against the oracle and against the committed outputs of the reference's own
compiled detectors (tests/golden/detect_ref_golden.npz), through the C ABI."""
import os

import numpy as np
import pytest

import helpers
from svo_pro_universal_b200 import capi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gpu_pyr(ctx, imgs, n_levels):
    imgs = np.ascontiguousarray(imgs)
    if imgs.ndim == 2:
        imgs = imgs[None]
    B, h, w = imgs.shape
    p = capi.Pyramid(ctx, B, w, h, n_levels)
    p.upload(imgs)
    p.build()
    return p


def _assert_corners_equal(g, o, tag):
    for k in helpers.CORNER_FIELDS:
        assert np.array_equal(g[k], o[k]), f"{tag}: field {k} differs at cells {np.flatnonzero(g[k] != o[k])[:8]}"


def test_angle_histogram_bins_exhaustive(ctx, orc):
    """Every central-difference gradient in [-255, 255]^2 falls in the same histogram bin as with the host libm."""
    bins = capi.angle_histogram_bins(ctx)
    ref = np.array([[orc.angle_histogram_bin(gx, gy) for gx in range(-255, 256)] for gy in range(-255, 256)], np.int8)
    assert np.array_equal(bins, ref), np.argwhere(bins != ref)[:8]


@pytest.mark.parametrize("ci", range(len(helpers.DETECT_CASES)))
def test_edgelets_bit_exact(ctx, orc, ci):
    case = helpers.DETECT_CASES[ci]
    seed, w, h, n_levels, kind, thr2, border, with_occ = case
    img, pyr, occ = helpers.detect_case_inputs(orc, case)
    p = _gpu_pyr(ctx, img, n_levels)
    g = capi.edgelet_detect(ctx, p, thr2, border, 30, occupancy=None if occ is None else occ[None])[0]
    _assert_corners_equal(g, orc.edgelet_detector_v2(pyr, thr2, border, 30, occ), f"case {ci} vs oracle")
    gold = np.load(os.path.join(GOLD, "detect_ref_golden.npz"))
    for k in helpers.CORNER_FIELDS:
        assert np.array_equal(g[k], gold[f"edgelet_{ci}_{k}"]), f"case {ci} vs the compiled reference's golden output, field {k}"


def test_fastgrad_bit_exact_and_feature_lists(ctx, orc):
    """svo_cuda_fastgrad_detect + the facade's fillFeatures ordering against FastGradDetector::detect of the reference (golden)."""
    gold = np.load(os.path.join(GOLD, "detect_ref_golden.npz"))
    for ci, case in enumerate(helpers.DETECT_CASES):
        seed, w, h, n_levels, kind, thr2, border, with_occ = case
        img, pyr, occ = helpers.detect_case_inputs(orc, case)
        p = _gpu_pyr(ctx, img, n_levels)
        opt = capi.detector_options(threshold=10, border=border, max_level=min(2, n_levels - 1))
        for max_n in (None, 60):
            fast, edge = capi.fastgrad_detect(ctx, p, opt, thr2, max_n, occupancy=None if occ is None else occ[None])
            fast, edge = fast[0], edge[0]
            for k in helpers.CORNER_FIELDS:
                assert np.array_equal(fast[k], gold[f"fast_{ci}_{k}"]), f"case {ci}: FAST stage field {k}"
            # fillFeatures (feature_detection_utils.cpp:72-142): score > threshold, sort by score, cap
            n_cells = len(fast)
            cap = n_cells if max_n is None else max_n
            feats = {"px": [], "score": [], "level": [], "grad": [], "type": []}
            for arr, thr, ftype, room in ((fast, 10.0, 7, cap), (edge, float(thr2), 6, None)):
                if room is None:
                    room = cap - len(feats["score"])
                sel = np.flatnonzero(arr["score"] > thr)
                sel = sel[np.argsort(-arr["score"][sel].astype(np.float64), kind="stable")][:max(room, 0)]
                for j in sel:
                    feats["px"].append((float(arr["x"][j]), float(arr["y"][j])))
                    feats["score"].append(float(arr["score"][j]))
                    feats["level"].append(int(arr["level"][j]))
                    a = np.float32(arr["angle"][j])
                    feats["grad"].append((float(np.cos(a)), float(np.sin(a))))  # float overloads, like the reference
                    feats["type"].append(ftype)
            got = {"px": np.array(feats["px"]).reshape(-1, 2), "score": np.array(feats["score"]), "level": np.array(feats["level"], np.int32),
                   "grad": np.array(feats["grad"]).reshape(-1, 2), "type": np.array(feats["type"], np.int32)}
            want = {f: gold[f"det_{ci}_2_{max_n}_{f}"] for f in got}
            assert len(got["score"]) == len(want["score"]), f"case {ci} max_n {max_n}"
            assert np.array_equal(got["score"], want["score"]) and np.array_equal(got["type"], want["type"])
            key = lambda d: sorted(zip(d["score"], d["px"][:, 0], d["px"][:, 1], d["level"]))  # noqa: E731
            if max_n is None:
                assert key(got) == key(want), f"case {ci}"
                # the gradient direction: cosf / sinf of numpy vs glibc may differ in the last float bit
                order_g = np.lexsort((got["px"][:, 1], got["px"][:, 0], got["score"]))
                order_w = np.lexsort((want["px"][:, 1], want["px"][:, 0], want["score"]))
                assert np.allclose(got["grad"][order_g], want["grad"][order_w], atol=1e-6, rtol=0)
                full = set(key(want))
            else:
                # capped list: std::sort is unstable, so WHICH of several equal scores survives the cut is unspecified;
                # every kept feature must be one of the uncapped list's, and everything above the cut score must agree
                assert set(key(got)) <= full, f"case {ci} max_n {max_n}"
                cut = got["score"].min() if len(got["score"]) else 0.0
                assert [k for k in key(got) if k[0] > cut] == [k for k in key(want) if k[0] > cut], f"case {ci} max_n {max_n}"


def test_edgelets_batch_device_arrays_and_occupancy(ctx, orc):
    """A batch of frames with device-resident outputs equals frame-by-frame oracle results; a fully occupied grid yields nothing."""
    import torch
    imgs = np.stack([helpers.detect_image(40 + i, 752, 480, "rect") for i in range(6)])
    p = _gpu_pyr(ctx, imgs, 3)
    n_cells = capi.grid_cells(752, 480, 30)[0]
    occ = (np.random.default_rng(9).random((6, n_cells)) < 0.4).astype(np.uint8)
    occ[5] = 1
    out = torch.zeros((6, n_cells, capi.CORNER_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    capi.edgelet_detect(ctx, p, 100, 8, 30, occupancy=torch.from_numpy(occ).cuda(), corners_out=out)
    ctx.synchronize()
    g = out.cpu().numpy().view(capi.CORNER_DTYPE).reshape(6, n_cells)
    for i in range(6):
        o = orc.edgelet_detector_v2(orc.create_img_pyramid(imgs[i], 3), 100, 8, 30, occ[i])
        _assert_corners_equal(g[i], o, f"frame {i}")
    assert (g[5]["score"] == 100).all() and not g[5]["x"].any()
    part = capi.edgelet_detect(ctx, p, 100, 8, 30, first=2, count=3)
    for i in range(3):
        _assert_corners_equal(part[i], orc.edgelet_detector_v2(orc.create_img_pyramid(imgs[2 + i], 3), 100, 8, 30), f"range frame {i}")


def test_edgelet_argument_errors(ctx):
    p = capi.Pyramid(ctx, 1, 752, 480, 3)
    p1 = capi.Pyramid(ctx, 1, 752, 480, 1)
    n_cells = capi.grid_cells(752, 480, 30)[0]
    out = np.zeros((1, n_cells), capi.CORNER_DTYPE)
    L = capi.lib()
    import ctypes as C
    po = C.c_void_p(out.ctypes.data)
    assert L.svo_cuda_edgelet_detect(ctx._h, p._h, 0, 1, 100, 3, 30, None, po, 0) == -1   # border < 4
    assert L.svo_cuda_edgelet_detect(ctx._h, p1._h, 0, 1, 100, 8, 30, None, po, 0) == -1  # no level 1
    assert L.svo_cuda_edgelet_detect(ctx._h, p._h, 0, 2, 100, 8, 30, None, po, 0) == -1   # frame range
    assert L.svo_cuda_edgelet_detect(ctx._h, p._h, 0, 1, -1, 8, 30, None, po, 0) == -1    # negative threshold
    assert L.svo_cuda_edgelet_detect(ctx._h, p._h, 0, 1, 100, 8, 30, None, None, 0) == -1
    assert L.svo_cuda_edgelet_detect(ctx._h, p._h, 0, 0, 100, 8, 30, None, po, 0) == 0    # empty range is fine
