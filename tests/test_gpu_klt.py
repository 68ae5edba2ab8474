"""GPU parity tests, first "next" row: feature_alignment::alignPyr2D — pyramidal KLT with integer
gradients — against the oracle (which is bit-identical to the reference's own compiled alignPyr2D, tests/test_klt_cpu.py)."""
import numpy as np
import pytest

from svo_pro_universal_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _case(seed, spread):
    d = synth.make_align_pair(seed)
    rng = np.random.default_rng(seed)
    px = np.round(d["px"]).astype(np.int32)
    extra = np.stack([rng.integers(0, 752, 60), rng.integers(0, 480, 60)], 1).astype(np.int32)  # incl. border / textureless spots
    px = np.concatenate([px, extra])
    start = px + rng.uniform(-spread, spread, px.shape)
    return d, px, start


@pytest.mark.parametrize("patch_sizes,levels", [([16, 16, 16, 16, 16], (4, 0)), ([8, 8, 8, 8, 8], (3, 1)), ([16, 16, 16, 8, 8], (4, 0)),
                                                ([16, 16, 8, 8, 8], (2, 2))])
def test_align_pyr2d_bit_exact(ctx, orc, patch_sizes, levels):
    d, px, start = _case(3, 6.0)
    rp, cp = orc.create_img_pyramid(d["ref_img"], 5), orc.create_img_pyramid(d["cur_img"], 5)
    ref = capi.Pyramid(ctx, 1, 752, 480, 5); cur = capi.Pyramid(ctx, 1, 752, 480, 5)
    ref.upload(d["ref_img"]); cur.upload(d["cur_img"]); ref.build(); cur.build()
    for n_iter in (30, 3):
        got, st = capi.align_pyr2d(ctx, ref, cur, px, start, levels[0], levels[1], patch_sizes, n_iter=n_iter)
        exp, se = orc.align_pyr2d(rp, cp, px, start, levels[0], levels[1], patch_sizes, n_iter=n_iter, n_threads=8)
        assert np.array_equal(st, se)
        assert np.array_equal(got, exp, equal_nan=True), np.abs(got - exp).max()
    assert 0.2 * len(px) < se.sum() < len(px), "cases must cover converged and failed tracks"


def test_align_pyr2d_batch_with_frame_indices(ctx, orc):
    pairs = [synth.make_align_pair(s) for s in (11, 12, 13)]
    ref = capi.Pyramid(ctx, 3, 752, 480, 4); cur = capi.Pyramid(ctx, 3, 752, 480, 4)
    ref.upload(np.stack([p["ref_img"] for p in pairs])); cur.upload(np.stack([p["cur_img"] for p in pairs])); ref.build(); cur.build()
    rng = np.random.default_rng(1)
    px, start, fi = [], [], []
    for k, p in enumerate(pairs):
        q = np.round(p["px"][:100]).astype(np.int32)
        px.append(q); start.append(q + rng.uniform(-4, 4, q.shape)); fi.append(np.full(len(q), k, np.int32))
    px, start, fi = np.concatenate(px), np.concatenate(start), np.concatenate(fi)
    perm = rng.permutation(len(px))
    px, start, fi = px[perm], start[perm], fi[perm]
    got, st = capi.align_pyr2d(ctx, ref, cur, px, start, 3, 0, [16, 16, 16, 16], ref_frame_idx=fi, cur_frame_idx=fi)
    for k, p in enumerate(pairs):
        sel = fi == k
        exp, se = orc.align_pyr2d(orc.create_img_pyramid(p["ref_img"], 4), orc.create_img_pyramid(p["cur_img"], 4), px[sel], start[sel], 3, 0,
                                  [16, 16, 16, 16])
        assert np.array_equal(st[sel], se) and np.array_equal(got[sel], exp, equal_nan=True)
    # tracked features land on the ground-truth reprojection (the cur image is the ref image warped by a known motion)
    assert st.mean() > 0.8


def test_align_pyr2d_rejects_bad_arguments(ctx):
    ref = capi.Pyramid(ctx, 1, 752, 480, 3)
    with pytest.raises(capi.SvoCudaError):
        capi.align_pyr2d(ctx, ref, ref, np.array([[100, 100]], np.int32), np.array([[100.0, 100.0]]), 2, 0, [12, 12, 12])
    with pytest.raises(capi.SvoCudaError):
        capi.align_pyr2d(ctx, ref, ref, np.array([[100, 100]], np.int32), np.array([[100.0, 100.0]]), 3, 0, [16, 16, 16, 16])
