"""CPU tests (-m "not gpu"), row f4: the oracle's restatement of PoseOptimizer::run against the outputs of the REFERENCE's own
pose_optimizer.cpp compiled from /root/reference into oracle/_ref/libfrontend_ref.so (tests/golden/pose_opt_ref_golden.npz;
where oracle/_ref travelled the live library is exercised too)."""
import os

import numpy as np
import pytest

import helpers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pose_opt_ref_golden.npz")


def _check(a, g, tol):
    for ci in range(len(helpers.POSE_OPT_CASES)):
        dq, dt = helpers.pose_diff(a[f"p{ci}_T"], g[f"p{ci}_T"])
        assert dq < tol and dt < tol, (ci, dq, dt)
        assert int(a[f"p{ci}_n"]) == int(g[f"p{ci}_n"]) and np.array_equal(a[f"p{ci}_outlier"], g[f"p{ci}_outlier"]), ci
        np.testing.assert_allclose(a[f"p{ci}_stats"], g[f"p{ci}_stats"], rtol=1e-6)   # sigma, medians (float in the reference), iterations


def test_oracle_equals_reference_golden(orc):
    mine = helpers.pose_opt_outputs(orc, "orc")
    _check(mine, np.load(GOLD), 1e-9)
    # the optimisation does its job: close to the true pose, gross outliers flagged
    for ci, spec in enumerate(helpers.POSE_OPT_CASES):
        c, _ = helpers.pose_opt_case(spec)
        dq, dt = helpers.pose_diff(mine[f"p{ci}_T"], c["T_imu_world_true"])
        dq0, dt0 = helpers.pose_diff(c["T_imu_world_init"], c["T_imu_world_true"])
        assert dt < 0.2 * dt0 and dq < 2e-3, (ci, dq, dt)
        assert 0 < mine[f"p{ci}_outlier"].sum() < 0.2 * len(c["px"])


def test_oracle_equals_live_reference(orc):
    if orc.ref_frontend_lib() is None:
        pytest.skip("oracle/_ref/libfrontend_ref.so not built on this box")
    _check(helpers.pose_opt_outputs(orc, "orc"), helpers.pose_opt_outputs(orc, "ref"), 1e-9)
