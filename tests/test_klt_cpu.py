"""CPU tests (-m "not gpu"): the oracle's alignPyr2D against the reference's own compiled alignPyr2D
(oracle/_ref/libdirect_ref.so, src/svo_direct/src/feature_alignment.cpp:761-973) and the committed golden vectors."""
import os

import numpy as np
import pytest

from svo_pro_universal_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = (([16, 16, 16, 16, 16], 4, 0, 30), ([8, 8, 8, 8, 8], 3, 1, 30), ([16, 16, 16, 8, 8], 4, 0, 5), ([16, 16, 8, 8, 8], 2, 2, 30))


def klt_inputs():
    d = synth.make_align_pair(3)
    rng = np.random.default_rng(3)
    px = np.round(d["px"]).astype(np.int32)
    px = np.concatenate([px, np.stack([rng.integers(0, 752, 60), rng.integers(0, 480, 60)], 1).astype(np.int32)])
    return d, px, px + rng.uniform(-6.0, 6.0, px.shape)


def klt_outputs(orc, which):
    d, px, start = klt_inputs()
    rp, cp = orc.create_img_pyramid(d["ref_img"], 5), orc.create_img_pyramid(d["cur_img"], 5)
    out = {}
    for k, (ps, mx, mn, it) in enumerate(CASES):
        p, s = orc.align_pyr2d(rp, cp, px, start, mx, mn, ps, n_iter=it, which=which)
        out[f"px_{k}"], out[f"status_{k}"] = p, s
    return out


def test_oracle_align_pyr2d_equals_reference_golden(orc):
    gold = np.load(os.path.join(GOLD, "klt_ref_golden.npz"))
    mine = klt_outputs(orc, "orc")
    for k in range(len(CASES)):
        assert np.array_equal(mine[f"status_{k}"], gold[f"status_{k}"])
        assert np.array_equal(mine[f"px_{k}"], gold[f"px_{k}"], equal_nan=True), "alignPyr2D must reproduce the reference bit for bit"
        assert 0.5 * len(gold[f"status_{k}"]) < gold[f"status_{k}"].sum() < len(gold[f"status_{k}"])


def test_oracle_align_pyr2d_equals_live_reference(orc):
    if orc.ref_direct_lib() is None:
        pytest.skip("oracle/_ref/libdirect_ref.so not built on this box")
    d = synth.make_align_pair(9)
    rp, cp = orc.create_img_pyramid(d["ref_img"], 4), orc.create_img_pyramid(d["cur_img"], 4)
    rng = np.random.default_rng(9)
    px = np.round(d["px"][:80]).astype(np.int32)
    start = px + rng.uniform(-10, 10, px.shape)
    for ps in ([16, 16, 16, 16], [8, 8, 8, 8]):
        a, sa = orc.align_pyr2d(rp, cp, px, start, 3, 0, ps, which="orc")
        b, sb = orc.align_pyr2d(rp, cp, px, start, 3, 0, ps, which="ref")
        assert np.array_equal(sa, sb) and np.array_equal(a, b, equal_nan=True)
