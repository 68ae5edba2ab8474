"""CPU test of SURVEY §8 row f3 (tracker part), FeatureTracker::trackAndDetect: the oracle's restatement of the track bookkeeping
around its alignPyr2D / detectors against the committed outputs of the reference's own compiled tracker."""
import os

import numpy as np

import helpers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_tracker_matches_reference_golden(orc):
    g = np.load(os.path.join(GOLD, "tracker_ref_golden.npz"))
    o = helpers.tracker_outputs(orc, "orc")
    assert set(o) == set(g.files)
    for k in o:
        assert np.array_equal(o[k], g[k]), k
    # the cases really cover termination, reset + re-detection and detection on top of live tracks
    assert sum(int(g[f"n_terminated_0_{k}"]) for k in range(5)) > 3
    assert g["track_id_1_3"].min() > g["track_id_1_2"].max() and len(g["px_2_2"]) > len(g["px_2_1"]) + 50
