"""CPU tests of SURVEY §8 row f4 (second half), Point::optimize: the oracle restatement against the committed outputs of the
reference's own compiled point.cpp (and against that library itself where it travelled)."""
import os

import numpy as np

import helpers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_matches_reference_golden(orc):
    g = np.load(os.path.join(GOLD, "point_opt_ref_golden.npz"))
    c = {k: g[k] for k in helpers.POINT_OPT_INPUT_KEYS}  # the stored inputs (see helpers.point_opt_outputs)
    o = helpers.point_opt_outputs(orc, "orc", c=c)
    for sphere in (0, 1):
        assert np.array_equal(o[f"pos_{sphere}"], g[f"pos_{sphere}"])  # bit for bit
        err0 = np.linalg.norm(c["pos0"] - c["pos_true"], axis=1)
        err1 = np.linalg.norm(o[f"pos_{sphere}"] - c["pos_true"], axis=1)
        assert np.median(err1) < 0.25 * np.median(err0)               # the refinement really refines
        single = np.diff(c["obs_begin"]) < 2
        assert single.sum() >= 5 and np.array_equal(o[f"pos_{sphere}"][single], c["pos0"][single])  # < 2 observations: untouched


def test_compiled_reference_agrees_with_golden(orc):
    if orc.ref_point_lib() is None:
        return
    g = np.load(os.path.join(GOLD, "point_opt_ref_golden.npz"))
    r = helpers.point_opt_outputs(orc, "ref", c={k: g[k] for k in helpers.POINT_OPT_INPUT_KEYS})
    for k in r:
        assert np.array_equal(r[k], g[k]), k
