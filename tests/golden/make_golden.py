"""Generates the committed golden fixtures under tests/golden/ (run in the build container, where /root/reference exists).

fast_ref_golden.npz   outputs of the REFERENCE's own FAST code (oracle/_ref/libfast_ref.so, compiled from
                      /root/reference/src/fast_neon/src) on seeded synthetic images: segment test (SSE2 + plain), score and
                      3x3 non-max, per pyramid level. These pin the oracle's rows a2-a4 and the CUDA kernels on boxes where
                      /root/reference does not exist.
oracle_golden.npz     outputs of the CPU oracle (restated reference arithmetic) for pyramid / sparse alignment / matcher /
                      depth filter on seeded inputs: regression pins for the oracle itself ("parity unpinned" rows).
direct_ref_golden.npz outputs of the REFERENCE's own halfSample / align1D / align2D / ZMSSD / Tukey / radtan / seed / grid code
                      (oracle/_ref/libdirect_ref.so, compiled from /root/reference against the container-only shims in
                      oracle/shim) on seeded inputs (tests/helpers.py:direct_cases): pins rows a1, c2-c5, parts of a5, b6,
                      d1 and s1 for the oracle and the CUDA kernels on boxes where /root/reference does not exist.
frontend_ref_golden.npz outputs of the REFERENCE's own SparseImgAlign::run, Matcher::findMatchDirect / findEpipolarMatchDirect and
                      depth_filter_utils::updateSeed (oracle/_ref/libfrontend_ref.so: the reference sources compiled against
                      oracle/shim) on seeded inputs (tests/helpers.py:frontend_outputs): pins rows b, c1, c6-c7, d1-d3.
reproject_ref_golden.npz outputs of the REFERENCE's own reprojector.cpp (getCandidate, sortCandidates*, matchCandidates,
                      matchCandidate; compiled into libfrontend_ref.so) on the cases of tests/helpers.py:REPROJECT_CASES: pins row f1.
pose_opt_ref_golden.npz outputs of the REFERENCE's own PoseOptimizer::run (pose_optimizer.cpp compiled into libfrontend_ref.so) on
                      tests/helpers.py:POSE_OPT_CASES: pins row f4.
detect_ref_golden.npz outputs of the REFERENCE's own detectors (feature_detection.cpp / feature_detection_utils.cpp / fast_neon compiled into
                      oracle/_ref/libdetect_ref.so; its GaussianBlur / Scharr calls resolve to the cv2-pinned restatements of
                      oracle/shim/shim_cv_imgproc.cpp) on tests/helpers.py:DETECT_CASES: pins rows a5, a6 and f2.
cv_imgproc_golden.npz outputs of the REAL OpenCV (python module cv2, version stored inside) for GaussianBlur(3x3, sigma 0) and
                      Scharr(8U -> 16S) on seeded images: pins the two imgproc restatements (oracle + shim) the edgelet detector uses.
stereo_tri_ref_golden.npz what the REFERENCE's own StereoTriangulation::compute (stereo_triangulation.cpp compiled into
                      libfrontend_ref.so, with its detectors, matcher and std::random_shuffle after srand(seed)) leaves in both frames on
                      tests/helpers.py:STEREO_TRI_CASES, plus the visiting orders: pins row f3 (stereo part).
point_opt_ref_golden.npz outputs of the REFERENCE's own Point::optimize (point.h / point.cpp compiled into libpoint_ref.so) on the 400
                      points of tests/helpers.py:point_opt_cases, unit plane and unit sphere: pins row f4 (second half).
tracker_ref_golden.npz what the REFERENCE's own FeatureTracker::trackAndDetect (feature_tracker.cpp + feature_tracking_types / utils compiled
                      into libfrontend_ref.so, with its detectors and alignPyr2D) leaves in every frame of three mono sequences
                      (tests/helpers.py:TRACKER_CASES: plain tracking, reset + re-detection, detection without reset and with the last
                      observation as template): pins row f3 (tracker part).
klt_ref_golden.npz    outputs of the REFERENCE's own alignPyr2D (libdirect_ref.so) on the cases of tests/test_klt_cpu.py.
Usage: python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from svo_pro_universal_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def fast_golden():
    assert orc.ref_lib() is not None, "oracle/_ref/libfast_ref.so missing: run make -C oracle"
    out = {}
    cases = [(0, 752, 480, 10), (1, 752, 480, 20), (2, 160, 120, 5), (3, 47, 30, 7), (4, 21, 13, 10), (5, 22, 9, 10)]
    out["cases"] = np.array(cases, np.int32)
    for seed, w, h, thr in cases:
        img0 = synth.make_image(seed, w, h, n_rect=max(8, w * h // 400))
        pyr = orc.create_img_pyramid(img0, 3 if min(w, h) >= 28 else 1)
        out[f"img_sha_{seed}"] = np.array(sha(img0))
        for l, im in enumerate(pyr):
            xy = orc.fast_detect(im, thr, 10, "ref_sse2")
            xy_plain = orc.fast_detect(im, thr, 10, "ref_plain10")
            assert np.array_equal(xy, xy_plain)
            sc = orc.fast_score10(im, xy, thr, "ref")
            nm = orc.fast_nonmax3x3(xy, sc, "ref")
            xy9 = orc.fast_detect(im, thr, 9, "ref_plain9")
            out[f"xy_{seed}_{l}"] = xy
            out[f"score_{seed}_{l}"] = sc.astype(np.int16)
            out[f"nonmax_{seed}_{l}"] = nm
            out[f"xy9_{seed}_{l}"] = xy9
    np.savez_compressed(os.path.join(HERE, "fast_ref_golden.npz"), **out)
    print("fast_ref_golden.npz", os.path.getsize(os.path.join(HERE, "fast_ref_golden.npz")))


def oracle_golden():
    out = {}
    # pyramid checksums
    img = synth.make_image(11)
    for mode in (-1, 0):
        pyr = orc.create_img_pyramid(img, 5, mode)
        out[f"pyr_sha_mode{mode}"] = np.array([sha(p) for p in pyr])
    # fastDetector per-cell corners
    for seed in (0, 5):
        c = orc.fast_detector(synth.make_image(seed))
        out[f"corners_{seed}"] = c
    # sparse alignment
    for seed in (1, 2, 3):
        d = synth.make_align_pair(seed)
        keep = []
        rp = orc.create_img_pyramid(d["ref_img"], 5)
        cp = orc.create_img_pyramid(d["cur_img"], 5)
        rf = orc.make_frame(rp, d["cam"], d["T_cam_imu"], d["T_imu_world_ref"], d["px"], d["f"], d["depth"], d["eligible"], keep=keep)
        cf = orc.make_frame(cp, d["cam"], d["T_cam_imu"], d["T_imu_world_cur_init"], keep=keep)
        for name, kw in (("default", {}), ("illum_robust", dict(estimate_illumination_gain=1, estimate_illumination_offset=1, robustification=1))):
            r = orc.sparse_align([rf], [cf], orc.default_align_options(**kw))
            out[f"align_{name}_{seed}_T"] = np.array(r.T_icur_iref)
            out[f"align_{name}_{seed}_ab"] = np.array([r.alpha, r.beta, r.chi2])
            out[f"align_{name}_{seed}_iters"] = np.array(list(r.iters), np.int32)
            out[f"align_{name}_{seed}_n"] = np.array(r.n_tracked)
    # Vogiatzis filter known-answer rows
    st = np.array([0.25, 0.0123, 10.0, 10.0])
    rows = []
    for z, tau2 in ((0.26, 1e-4), (0.24, 4e-4), (0.9, 1e-4), (0.255, 2e-5)):
        ok = orc.lib().orc_update_filter_vogiatzis(z, tau2, 1.0 / 1.5, st.ctypes.data_as(orc.f64p))
        rows.append(np.concatenate([[z, tau2, ok], st]))
    out["vogiatzis_rows"] = np.array(rows)
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("oracle_golden.npz", os.path.getsize(os.path.join(HERE, "oracle_golden.npz")))


def direct_golden():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    assert orc.ref_direct_lib() is not None, "oracle/_ref/libdirect_ref.so missing: run make -C oracle"
    out = helpers.direct_outputs(orc, "ref")
    np.savez_compressed(os.path.join(HERE, "direct_ref_golden.npz"), **out)
    print("direct_ref_golden.npz", os.path.getsize(os.path.join(HERE, "direct_ref_golden.npz")))


def frontend_golden():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    assert orc.ref_frontend_lib() is not None, "oracle/_ref/libfrontend_ref.so missing: run make -C oracle"
    out = helpers.frontend_outputs(orc, "ref")
    np.savez_compressed(os.path.join(HERE, "frontend_ref_golden.npz"), **out)
    print("frontend_ref_golden.npz", os.path.getsize(os.path.join(HERE, "frontend_ref_golden.npz")))


def klt_golden():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_klt_cpu
    out = test_klt_cpu.klt_outputs(orc, "ref")
    np.savez_compressed(os.path.join(HERE, "klt_ref_golden.npz"), **out)
    print("klt_ref_golden.npz", os.path.getsize(os.path.join(HERE, "klt_ref_golden.npz")))


def pose_opt_golden():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    assert orc.ref_frontend_lib() is not None, "oracle/_ref/libfrontend_ref.so missing: run make -C oracle"
    np.savez_compressed(os.path.join(HERE, "pose_opt_ref_golden.npz"), **helpers.pose_opt_outputs(orc, "ref"))
    print("pose_opt_ref_golden.npz", os.path.getsize(os.path.join(HERE, "pose_opt_ref_golden.npz")))


def reproject_golden():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    assert orc.ref_frontend_lib() is not None, "oracle/_ref/libfrontend_ref.so missing: run make -C oracle"
    out = helpers.reproject_outputs(orc, "ref")
    out.update(helpers.reproject_frames_reference(orc))   # the whole Reprojector::reprojectFrames, for the C++ facade
    np.savez_compressed(os.path.join(HERE, "reproject_ref_golden.npz"), **out)
    print("reproject_ref_golden.npz", os.path.getsize(os.path.join(HERE, "reproject_ref_golden.npz")))


def stereo_tri_golden():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    assert orc.ref_frontend_lib() is not None
    np.savez_compressed(os.path.join(HERE, "stereo_tri_ref_golden.npz"), **helpers.stereo_tri_reference(orc))
    print("wrote stereo_tri_ref_golden.npz")


def point_opt_golden():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    assert orc.ref_point_lib() is not None
    np.savez_compressed(os.path.join(HERE, "point_opt_ref_golden.npz"), **helpers.point_opt_outputs(orc, "ref"))
    print("wrote point_opt_ref_golden.npz")


def tracker_golden():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    assert orc.ref_frontend_lib() is not None
    np.savez_compressed(os.path.join(HERE, "tracker_ref_golden.npz"), **helpers.tracker_outputs(orc, "ref"))
    print("wrote tracker_ref_golden.npz")


def detect_golden():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    assert orc.ref_detect_lib() is not None, "oracle/_ref/libdetect_ref.so missing: run make -C oracle"
    np.savez_compressed(os.path.join(HERE, "detect_ref_golden.npz"), **helpers.detect_outputs(orc, "ref"))
    print("wrote detect_ref_golden.npz")


def cv_imgproc_golden():
    import cv2
    out = {"cv2_version": np.array(cv2.__version__)}
    rng = np.random.default_rng(77)
    shapes = [(120, 188), (60, 94), (30, 47), (17, 33), (5, 7), (3, 3), (1, 9), (9, 1), (2, 2)]
    out["shapes"] = np.array(shapes, np.int32)
    for i, (h, w) in enumerate(shapes):
        img = rng.integers(0, 256, (h, w)).astype(np.uint8) if i % 2 == 0 else synth.make_image(i, w, h, n_rect=max(4, w * h // 200))
        g = cv2.GaussianBlur(img, (3, 3), 0)
        out[f"img_{i}"] = img
        out[f"blur_{i}"] = g
        out[f"dx_{i}"] = cv2.Scharr(g, cv2.CV_16S, 1, 0, scale=1, delta=0, borderType=cv2.BORDER_DEFAULT)
        out[f"dy_{i}"] = cv2.Scharr(g, cv2.CV_16S, 0, 1, scale=1, delta=0, borderType=cv2.BORDER_DEFAULT)
    np.savez_compressed(os.path.join(HERE, "cv_imgproc_golden.npz"), **out)
    print("wrote cv_imgproc_golden.npz (cv2 " + cv2.__version__ + ")")


if __name__ == "__main__":
    if len(sys.argv) > 1:  # e.g. `make_golden.py reproject`: regenerate one fixture
        for name in sys.argv[1:]:
            globals()[name + "_golden"]()
        sys.exit(0)
    detect_golden()
    cv_imgproc_golden()
    stereo_tri_golden()
    point_opt_golden()
    tracker_golden()
    pose_opt_golden()
    reproject_golden()
    klt_golden()
    frontend_golden()
    fast_golden()
    oracle_golden()
    direct_golden()
