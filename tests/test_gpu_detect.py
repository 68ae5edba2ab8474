"""GPU parity tests, rows a1-a5: pyramid bytes, FAST corners / scores / non-max and per-cell corners, BIT-EXACT against the
oracle, the committed golden vectors of the reference's own FAST code, and (where it travelled) the compiled reference."""
import os

import numpy as np
import pytest

from svo_pro_universal_b200 import capi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gpu_pyr(ctx, imgs, n_levels, mode=-1):
    imgs = np.ascontiguousarray(imgs)
    if imgs.ndim == 2:
        imgs = imgs[None]
    B, h, w = imgs.shape
    p = capi.Pyramid(ctx, B, w, h, n_levels, mode)
    p.upload(imgs)
    p.build()
    return p


@pytest.mark.parametrize("shape,n_levels", [((480, 752), 5), ((480, 752), 8), ((241, 377), 4), ((31, 47), 2), ((64, 64), 6), ((100, 1030), 3)])
def test_pyramid_bit_exact(ctx, orc, shape, n_levels):
    h, w = shape
    imgs = np.stack([synth.make_image(s, w, h, n_rect=max(8, w * h // 400)) for s in (1, 2, 3)])
    for mode in (-1, 0):
        p = _gpu_pyr(ctx, imgs, n_levels, mode)
        for i in range(3):
            o = orc.create_img_pyramid(imgs[i], n_levels, mode)
            for l in range(n_levels):
                assert np.array_equal(p.download(i, l), o[l]), f"frame {i} level {l} mode {mode}"


def test_pyramid_partial_range_and_idempotence(ctx, orc):
    imgs = np.stack([synth.make_image(s) for s in range(4)])
    p = capi.Pyramid(ctx, 4, 752, 480, 5)
    p.upload(imgs)
    p.build(first=1, count=2)
    assert not p.download(0, 1).any() and not p.download(3, 2).any()  # untouched frames stay zero
    o = orc.create_img_pyramid(imgs[2], 5)
    assert np.array_equal(p.download(2, 4), o[4])
    p.build()
    p.build()  # idempotent
    assert np.array_equal(p.download(2, 4), o[4]) and np.array_equal(p.download(0, 1), orc.create_img_pyramid(imgs[0], 5)[1])


def _check_level_against_lists(score_map, nonmax_map, xy, scores, nm_idx):
    """Dense GPU maps vs the sparse lists fast::* produce."""
    det = np.argwhere(score_map > 0)  # raster order (y, x)
    assert np.array_equal(det[:, ::-1].astype(np.int16), xy), "segment-test corner set differs"
    assert np.array_equal(score_map[xy[:, 1], xy[:, 0]].astype(np.int32), scores), "scores differ"
    nm = np.argwhere(nonmax_map > 0)[:, ::-1].astype(np.int16)
    assert np.array_equal(nm, xy[nm_idx]), "3x3 non-max survivors differ"


def test_fast_stages_match_reference_golden(ctx, orc):
    """a2-a4 against vectors generated from the reference's own sources (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "fast_ref_golden.npz"))
    for seed, w, h, thr in g["cases"]:
        seed, w, h, thr = int(seed), int(w), int(h), int(thr)
        img0 = synth.make_image(seed, w, h, n_rect=max(8, w * h // 400))
        n_levels = 3 if min(w, h) >= 28 else 1
        p = _gpu_pyr(ctx, img0, n_levels)
        for l in range(n_levels):
            sm, nm = capi.fast_level_maps(ctx, p, 0, l, thr, 10)
            _check_level_against_lists(sm, nm, g[f"xy_{seed}_{l}"], g[f"score_{seed}_{l}"].astype(np.int32), g[f"nonmax_{seed}_{l}"])
            sm9, _ = capi.fast_level_maps(ctx, p, 0, l, thr, 9)
            assert np.array_equal(np.argwhere(sm9 > 0)[:, ::-1].astype(np.int16), g[f"xy9_{seed}_{l}"])


def test_fast_corner_lists_match_reference_golden(ctx, orc):
    """The list-shaped leaves of fast.h:32-41 (svo_cuda_fast_corner_list / _corner_score / _nonmax_3x3) against the vectors generated
    from the reference's own fast_corner_detect_10[_sse2] / fast_corner_detect_9 / fast_corner_score_10 / fast_nonmax_3x3."""
    g = np.load(os.path.join(GOLD, "fast_ref_golden.npz"))
    for seed, w, h, thr in g["cases"]:
        seed, w, h, thr = int(seed), int(w), int(h), int(thr)
        img0 = synth.make_image(seed, w, h, n_rect=max(8, w * h // 400))
        n_levels = 3 if min(w, h) >= 28 else 1
        p = _gpu_pyr(ctx, img0, n_levels)
        for l in range(n_levels):
            gxy, gsc, gnm = g[f"xy_{seed}_{l}"], g[f"score_{seed}_{l}"].astype(np.int32), g[f"nonmax_{seed}_{l}"]
            xy, sc, nm = capi.fast_corner_list(ctx, p, 0, l, thr, 10)
            assert np.array_equal(np.stack([xy["x"], xy["y"]], 1), gxy) and np.array_equal(sc, gsc)
            assert np.array_equal(np.flatnonzero(nm), gnm)
            xy9, _, _ = capi.fast_corner_list(ctx, p, 0, l, thr, 9)
            assert np.array_equal(np.stack([xy9["x"], xy9["y"]], 1), g[f"xy9_{seed}_{l}"])
            # the stand-alone leaves on the reference's own lists
            lst = np.zeros(len(gxy), capi.FAST_XY_DTYPE)
            lst["x"], lst["y"] = gxy[:, 0], gxy[:, 1]
            assert np.array_equal(capi.fast_corner_score(ctx, p, 0, l, lst, thr, 10), gsc)
            assert np.array_equal(capi.fast_nonmax_3x3(ctx, lst, gsc), gnm)
            if len(gxy) > 4:  # a truncated list reports the full count
                n_cap = len(gxy) // 2
                xy_t, sc_t, _ = capi.fast_corner_list(ctx, p, 0, l, thr, 10, max_corners=n_cap)
                assert len(xy_t) == n_cap and np.array_equal(sc_t, gsc[:n_cap])


def test_fast_corner_list_leaves_random(ctx, orc):
    """Random images and lists: scores of pixels that are NOT corners (fast_corner_score_10 returns the threshold), non-max on
    arbitrary score lists (ties, neighbours on the row ends), empty lists."""
    rng = np.random.default_rng(23)
    for trial in range(4):
        w, h = int(rng.integers(40, 400)), int(rng.integers(20, 260))
        img = np.kron(rng.integers(0, 256, (h // 2 + 1, w // 2 + 1)), np.ones((2, 2)))[:h, :w].astype(np.uint8)
        p = _gpu_pyr(ctx, img, 1)
        for thr in (5, 40, 200):
            xy, sc, nm = capi.fast_corner_list(ctx, p, 0, 0, thr, 10)
            oxy = orc.fast_detect(img, thr, 10)
            osc = orc.fast_score10(img, oxy, thr)
            assert np.array_equal(np.stack([xy["x"], xy["y"]], 1), oxy) and np.array_equal(sc, osc)
            assert np.array_equal(np.flatnonzero(nm), orc.fast_nonmax3x3(oxy, osc))
            # arbitrary pixels, most of them no corners
            m = 300
            pts = np.unique(np.stack([rng.integers(3, w - 3, m), rng.integers(3, h - 3, m)], 1), axis=0)
            pts = pts[np.lexsort((pts[:, 0], pts[:, 1]))].astype(np.int16)   # raster order
            lst = np.zeros(len(pts), capi.FAST_XY_DTYPE); lst["x"], lst["y"] = pts[:, 0], pts[:, 1]
            assert np.array_equal(capi.fast_corner_score(ctx, p, 0, 0, lst, thr, 10), orc.fast_score10(img, pts, thr))
            # non-max with random scores (many ties) on a dense block of listed pixels
            yy, xx = np.mgrid[3:3 + min(12, h - 6), 3:3 + min(40, w - 6)]
            keep = rng.uniform(size=yy.size) < 0.7
            blk = np.stack([xx.ravel()[keep], yy.ravel()[keep]], 1).astype(np.int16)
            bsc = rng.integers(1, 6, len(blk)).astype(np.int32)
            lst = np.zeros(len(blk), capi.FAST_XY_DTYPE); lst["x"], lst["y"] = blk[:, 0], blk[:, 1]
            assert np.array_equal(capi.fast_nonmax_3x3(ctx, lst, bsc), orc.fast_nonmax3x3(blk, bsc))
    assert len(capi.fast_nonmax_3x3(ctx, np.zeros(0, capi.FAST_XY_DTYPE), np.zeros(0, np.int32))) == 0
    flat = _gpu_pyr(ctx, np.full((64, 64), 90, np.uint8), 1)
    assert len(capi.fast_corner_list(ctx, flat, 0, 0, 10, 10)[0]) == 0


def test_fast_stages_match_oracle_random(ctx, orc):
    rng = np.random.default_rng(11)
    for trial in range(6):
        w, h = int(rng.integers(22, 300)), int(rng.integers(7, 200))
        if trial % 2:
            img = np.kron(rng.integers(0, 256, (h // 3 + 1, w // 3 + 1)), np.ones((3, 3)))[:h, :w].astype(np.uint8)
        else:
            img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        p = _gpu_pyr(ctx, img, 1)
        for thr in (1, 7, 30, 120, 127, 128, 200, 254):  # >= 128 takes the wide-compare quick-reject
            xy = orc.fast_detect(img, thr, 10)
            sc = orc.fast_score10(img, xy, thr)
            nm_idx = orc.fast_nonmax3x3(xy, sc)
            sm, nm = capi.fast_level_maps(ctx, p, 0, 0, thr, 10)
            _check_level_against_lists(sm, nm, xy, sc, nm_idx)
            if orc.ref_lib() is not None:
                assert np.array_equal(xy, orc.fast_detect(img, thr, 10, "ref_sse2"))


@pytest.mark.parametrize("kw", [dict(), dict(threshold=20, border=4, cell_size=25), dict(min_level=1, max_level=3, cell_size=40),
                                dict(threshold=5, max_level=0, cell_size=16)])
def test_fast_detector_cells_bit_exact(ctx, orc, kw):
    """a5: per-cell winners (x, y, level, score) incl. the reference's tie-breaking, on a batch."""
    opt = capi.detector_options(**kw)
    seeds = [0, 3, 8, 13, 21]
    imgs = np.stack([synth.make_image(s) for s in seeds])
    p = _gpu_pyr(ctx, imgs, 5)
    got = capi.fast_detect(ctx, p, opt)
    for i, s in enumerate(seeds):
        exp = orc.fast_detector(imgs[i], 5, -1, opt.threshold, opt.border, opt.min_level, opt.max_level, opt.cell_size)
        for k in ("x", "y", "level", "score", "angle"):
            assert np.array_equal(got[i][k], exp[k]), f"seed {s} field {k}"
        assert (got[i]["score"] > opt.threshold).sum() > 50


def test_fast_detector_subrange_of_odd_sized_frames(ctx, orc):
    """Frames [2, 4) of a batch whose size is no multiple of the tile (377 x 241, levels 0-2): the tile loads take the frame index and
    the out-of-image halo from the level's tensor map, so a wrong frame offset or fill would show here."""
    w, h = 377, 241
    imgs = np.stack([synth.make_image(40 + s, w, h, n_rect=300) for s in range(5)])
    opt = capi.detector_options(threshold=12, border=5, min_level=0, max_level=2, cell_size=24)
    p = _gpu_pyr(ctx, imgs, 4)
    got = capi.fast_detect(ctx, p, opt, first=2, count=2)
    assert got.shape[0] == 2
    for j, i in enumerate((2, 3)):
        exp = orc.fast_detector(imgs[i], 4, -1, opt.threshold, opt.border, opt.min_level, opt.max_level, opt.cell_size)
        for k in ("x", "y", "level", "score"):
            assert np.array_equal(got[j][k], exp[k]), f"frame {i} field {k}"
        assert (got[j]["score"] > opt.threshold).sum() > 20


def test_fast_detector_occupancy_and_flat_image(ctx, orc):
    opt = capi.detector_options()
    img = synth.make_image(2)
    flat = np.full((480, 752), 90, np.uint8)
    p = _gpu_pyr(ctx, np.stack([img, flat]), 5)
    rng = np.random.default_rng(0)
    occ = (rng.uniform(size=(2, 416)) < 0.4).astype(np.uint8)
    got = capi.fast_detect(ctx, p, opt, occupancy=occ)
    exp = orc.fast_detector(img, occupancy=occ[0])
    for k in ("x", "y", "level", "score"):
        assert np.array_equal(got[0][k], exp[k])
    assert (got[0]["score"][occ[0] > 0] == opt.threshold).all()          # occupied cells keep the pre-filled corner
    assert (got[1]["score"] == opt.threshold).all() and not got[1]["x"].any()  # no corners on a flat frame


def test_fused_pyramid_fast_detect_equals_two_step(ctx, orc):
    imgs = np.stack([synth.make_image(s) for s in (4, 5, 6)])
    opt = capi.detector_options()
    p1 = _gpu_pyr(ctx, imgs, 5)
    a = capi.fast_detect(ctx, p1, opt)
    p2 = capi.Pyramid(ctx, 3, 752, 480, 5)
    p2.upload(imgs)
    b = capi.fast_detect(ctx, p2, opt, fused_pyramid=True)
    assert np.array_equal(a, b)
    assert np.array_equal(p2.download(1, 3), p1.download(1, 3))


def test_fast_detect_full_size_properties(ctx):
    """BASELINE config 2 shape at reduced batch: properties that hold at any size — determinism, per-cell containment,
    scores above threshold, winners are local maxima of the score map."""
    B = 64
    imgs = np.stack([synth.make_image(100 + s) for s in range(B)])
    opt = capi.detector_options()
    p = _gpu_pyr(ctx, imgs, 5)
    a = capi.fast_detect(ctx, p, opt)
    b = capi.fast_detect(ctx, p, opt)
    assert np.array_equal(a, b)
    found = a["score"] > opt.threshold
    cell = (a["y"] // opt.cell_size) * 26 + a["x"] // opt.cell_size
    assert (cell[found] == np.broadcast_to(np.arange(416), a.shape)[found]).all()
    for i in (0, B - 1):
        for l in range(3):
            sm, nm = capi.fast_level_maps(ctx, p, i, l, opt.threshold, 10)
            sel = found[i] & (a[i]["level"] == l)
            ys, xs = a[i]["y"][sel] >> l, a[i]["x"][sel] >> l
            assert (nm[ys, xs] == 1).all() and (sm[ys, xs] == a[i]["score"][sel]).all()
