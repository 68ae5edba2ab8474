"""GPU test of the LINK-TIME SWAP: oracle/_ref/libfrontend_swap.so is the reference's own wrapper and sources compiled against
the reference's own headers, with the hot-path functions — svo::SparseImgAlign::run, svo::Matcher::findMatchDirect /
findEpipolarMatchDirect, depth_filter_utils::updateSeed / updateFilterVogiatzis / updateFilterGaussian / computeTau — coming from
svo_pro_universal_b200/host/ref_swap.cpp, which calls the C ABI of libsvo_cuda.so (oracle/Makefile target `swap`). The same
entry points that produced tests/golden/*.npz from the unmodified reference must reproduce those goldens through the swap:

  * the front-end goldens (SparseImgAlign option sets, prior, radtan, stereo bundle; matcher option sets; updateSeed chains);
  * Reprojector::reprojectFrames and StereoTriangulation::compute — the reference's OWN unmodified reprojector.cpp and
    stereo_triangulation.cpp, whose Matcher / updateSeed calls now land on the GPU (the drop-in the north star asks for:
    src/svo/src/frame_handler_base.cpp:125,135,145 keep `new`-ing the same classes).
"""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROT_TOL, TRANS_TOL, PX_TOL, REL_TOL = 1e-4, 1e-4, 1e-3, 1e-4  # north_star tolerances


@pytest.fixture(scope="module")
def swap(tmp_path_factory):
    """Outputs of every golden-producing wrapper entry point through the swap library, computed in a separate process."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "libfrontend_swap.so")):
        pytest.skip("oracle/_ref/libfrontend_swap.so was not built (needs /root/reference at build time)")
    path = str(tmp_path_factory.mktemp("swap") / "swap_outputs.npz")
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "swap_outputs.py"), path], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, "swap_outputs.py failed:\n" + r.stdout[-2000:] + r.stderr[-4000:]
    return np.load(path)


def _sub(d, prefix):
    return {k[len(prefix):]: d[k] for k in d.files if k.startswith(prefix)}


def test_reference_headers_plus_swap_tu_reproduce_the_frontend_goldens(swap):
    gold = np.load(os.path.join(GOLD, "frontend_ref_golden.npz"))
    mine = _sub(swap, "fe_")
    # ---- svo::SparseImgAlign::run
    a, g = mine["align_rows"], gold["align_rows"]
    assert a.shape == g.shape
    assert np.array_equal(a[:, 10], g[:, 10]), "number of tracked features"
    for i in range(len(a)):
        dq, dt = helpers.pose_diff(a[i, :7], g[i, :7])
        assert dq < ROT_TOL and dt < TRANS_TOL and dq < 1e-8 and dt < 1e-8, (i, dq, dt)
    np.testing.assert_allclose(a[:, 9], g[:, 9], rtol=1e-5)                             # getError()
    np.testing.assert_allclose(mine["align_H"], gold["align_H"], rtol=1e-6, atol=1e-3)  # getHessian()
    dq, dt = helpers.pose_diff(mine["stereo_row"][:7], gold["stereo_row"][:7])
    assert dq < 1e-8 and dt < 1e-8 and mine["stereo_row"][10] == gold["stereo_row"][10]
    for c in range(2):
        dq, dt = helpers.pose_diff(mine["stereo_T_f_w"][c], gold["stereo_T_f_w"][c])
        assert dq < 1e-8 and dt < 1e-8
    # ---- svo::Matcher
    for name in ("fmd_default", "fmd_gain", "epi_sphere", "epi_plane", "epi_a1d", "epi_nosub"):
        res = gold[f"{name}_result"]
        assert np.array_equal(mine[f"{name}_result"], res), name
        ok = res == 0
        assert np.array_equal(mine[f"{name}_search_level"][ok], gold[f"{name}_search_level"][ok])
        assert np.array_equal(mine[f"{name}_patch_with_border"][ok], gold[f"{name}_patch_with_border"][ok]), "warped patch bytes"
        assert np.abs(mine[f"{name}_px_cur"][ok] - gold[f"{name}_px_cur"][ok]).max() < PX_TOL
        np.testing.assert_allclose(mine[f"{name}_f_cur"][ok], gold[f"{name}_f_cur"][ok], atol=1e-6)
        np.testing.assert_allclose(mine[f"{name}_A_cur_ref"][ok], gold[f"{name}_A_cur_ref"][ok], rtol=1e-12, atol=1e-14)
        if name.startswith("epi"):
            np.testing.assert_allclose(mine[f"{name}_depth"][ok], gold[f"{name}_depth"][ok], rtol=1e-4)
            assert np.array_equal(mine[f"{name}_epi_length_pyramid"][ok], gold[f"{name}_epi_length_pyramid"][ok])
            assert np.array_equal(mine[f"{name}_reject"], gold[f"{name}_reject"])
            np.testing.assert_allclose(mine[f"{name}_epi_image"], gold[f"{name}_epi_image"], rtol=1e-12, atol=1e-12)  # Matcher::epi_image_
    # ---- svo::Matcher::scanEpipolarLine called on its own (svo_cuda_scan_epipolar_line behind the reference's signature)
    for name in ("sphere", "plane", "capped", "low_start"):
        assert np.array_equal(mine[f"scan_{name}_zmssd"], gold[f"scan_{name}_zmssd"]), name
        assert np.abs(mine[f"scan_{name}_px"] - gold[f"scan_{name}_px"]).max() < 1e-9, name
    # ---- depth_filter_utils::updateSeed chains
    for name in ("vog", "gauss", "conv"):
        assert int(mine[f"seeds_{name}_n"]) == int(gold[f"seeds_{name}_n"]) > 300
        assert np.array_equal(mine[f"seeds_{name}_types"], gold[f"seeds_{name}_types"])
        assert np.array_equal(mine[f"seeds_{name}_ok"], gold[f"seeds_{name}_ok"])
        np.testing.assert_allclose(mine[f"seeds_{name}_state"], gold[f"seeds_{name}_state"], rtol=REL_TOL)


def test_filter_and_tau_through_the_swap(swap):
    """updateFilterVogiatzis / updateFilterGaussian / computeTau of the swap against the oracle on the same inputs."""
    lf = swap["leaf_filter"]
    assert np.array_equal(lf[:, 0], lf[:, 1]) and np.array_equal(lf[:, 2], lf[:, 3])       # return values
    np.testing.assert_allclose(lf[:, 8:12], lf[:, 4:8], rtol=1e-9)                          # Vogiatzis state
    np.testing.assert_allclose(lf[:, 16:20], lf[:, 12:16], rtol=1e-9)                       # Gaussian state
    np.testing.assert_allclose(swap["leaf_tau"][:, 1], swap["leaf_tau"][:, 0], rtol=1e-9)


def test_the_references_own_reprojector_runs_on_the_swapped_matcher(swap):
    """Reprojector::reprojectFrames, compiled from the reference's unmodified reprojector.cpp: its matchCandidate loop calls
    Matcher::findMatchDirect and depth_filter_utils::updateSeed, which are the swap's. Same outputs as the all-CPU reference."""
    gold = np.load(os.path.join(GOLD, "reproject_ref_golden.npz"))
    mine = _sub(swap, "rp_")
    for ci in range(len(helpers.REPROJECT_FRAMES_CASES)):
        g = {k: gold[f"rf{ci}_{k}"] for k in helpers.REPROJ_FRAMES_KEYS}
        m = {k: mine[f"rf{ci}_{k}"] for k in helpers.REPROJ_FRAMES_KEYS}
        for k in ("type", "level", "point", "seed_feat", "score", "stats", "pt_counters", "feat_type"):
            assert np.array_equal(m[k], g[k]), (ci, k)
        assert np.array_equal(m["occupancy"][:416], g["occupancy"][:416]), ci
        assert np.abs(m["px"] - g["px"]).max() < PX_TOL, ci
        np.testing.assert_allclose(m["state"], g["state"], rtol=REL_TOL, atol=1e-12)
        np.testing.assert_allclose(m["feat_state"], g["feat_state"], rtol=REL_TOL, atol=1e-12)
        assert np.abs(m["f"] - g["f"]).max() < 1e-5 and np.abs(m["grad"] - g["grad"]).max() < 1e-9, ci


def test_the_references_own_stereo_triangulation_runs_on_the_swapped_matcher(swap):
    """StereoTriangulation::compute, compiled from the reference's unmodified stereo_triangulation.cpp (detector, shuffle, sequential
    matching loop with Matcher::findEpipolarMatchDirect from the swap): same triangulated features as the all-CPU reference."""
    mine = _sub(swap, "st_")
    if not mine:
        pytest.skip("wrapper without the stereo entry point")
    gold = np.load(os.path.join(GOLD, "stereo_tri_ref_golden.npz"))
    tol = {"px1": dict(rtol=0, atol=PX_TOL), "f1": dict(rtol=0, atol=1e-5), "grad1": dict(rtol=0, atol=1e-6), "xyz1": dict(rtol=1e-4, atol=1e-6)}
    for k in mine:
        a, g = np.asarray(mine[k]), gold[k]
        base = k.rsplit("_", 1)[0]
        if a.dtype.kind in "iub":
            assert np.array_equal(a, g), k   # the same frame0 features triangulated, in the same order and slots
        else:
            np.testing.assert_allclose(a, g, err_msg=k, **tol.get(base, dict(rtol=1e-9, atol=1e-12)))
