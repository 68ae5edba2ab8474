"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes bindings of the CPU oracle (oracle/liborc.so, restated reference arithmetic) and of the
reference's own FAST sources compiled into oracle/_ref/libfast_ref.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product package (svo_pro_universal_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_LEVELS = 8
MAX_CAMS = 4

u8p = C.POINTER(C.c_uint8)
f64p = C.POINTER(C.c_double)
i32p = C.POINTER(C.c_int)
i16p = C.POINTER(C.c_short)


class Corner(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("level", C.c_int), ("score", C.c_float), ("angle", C.c_float)]


class Frame(C.Structure):
    _fields_ = [
        ("level_data", u8p * MAX_LEVELS),
        ("level_cols", C.c_int * MAX_LEVELS),
        ("level_rows", C.c_int * MAX_LEVELS),
        ("level_step", C.c_int * MAX_LEVELS),
        ("n_levels", C.c_int),
        ("cam", C.c_double * 8),
        ("width", C.c_int),
        ("height", C.c_int),
        ("distortion", C.c_int),
        ("T_cam_imu", C.c_double * 7),
        ("T_imu_world", C.c_double * 7),
        ("n_features", C.c_int),
        ("px", f64p),
        ("f", f64p),
        ("depth", f64p),
        ("eligible", u8p),
    ]


class AlignOptions(C.Structure):
    _fields_ = [
        ("max_level", C.c_int), ("min_level", C.c_int),
        ("estimate_illumination_gain", C.c_int), ("estimate_illumination_offset", C.c_int),
        ("use_distortion_jacobian", C.c_int), ("robustification", C.c_int),
        ("weight_scale", C.c_double),
        ("max_iter", C.c_int),
        ("eps", C.c_double),
        ("alpha_init", C.c_double), ("beta_init", C.c_double),
        ("have_prior", C.c_int),
        ("prior_T", C.c_double * 7),
        ("prior_alpha", C.c_double), ("prior_beta", C.c_double),
        ("lambda_rot", C.c_double), ("lambda_trans", C.c_double),
        ("lambda_alpha", C.c_double), ("lambda_beta", C.c_double),
    ]


class AlignResult(C.Structure):
    _fields_ = [
        ("n_tracked", C.c_int),
        ("T_icur_iref", C.c_double * 7),
        ("alpha", C.c_double), ("beta", C.c_double), ("chi2", C.c_double),
        ("H", C.c_double * 64),
        ("iters", C.c_int * MAX_LEVELS),
        ("T_f_w", (C.c_double * 7) * MAX_CAMS),
        ("stop", C.c_int),
    ]


class Feature(C.Structure):
    _fields_ = [("type", C.c_int), ("px", C.c_double * 2), ("f", C.c_double * 3), ("grad", C.c_double * 2),
                ("level", C.c_int)]


class MatcherOptions(C.Structure):
    _fields_ = [
        ("align_1d", C.c_int), ("align_max_iter", C.c_int),
        ("max_epi_search_steps", C.c_int),
        ("subpix_refinement", C.c_int), ("epi_search_edgelet_filtering", C.c_int), ("scan_on_unit_sphere", C.c_int),
        ("epi_search_edgelet_max_angle", C.c_double),
        ("affine_est_offset", C.c_int), ("affine_est_gain", C.c_int),
        ("max_patch_diff_ratio", C.c_double),
    ]


class MatchOut(C.Structure):
    _fields_ = [
        ("result", C.c_int),
        ("px_cur", C.c_double * 2),
        ("f_cur", C.c_double * 3),
        ("search_level", C.c_int),
        ("A_cur_ref", C.c_double * 4),
        ("h_inv", C.c_double),
        ("epi_length_pyramid", C.c_double),
        ("reject", C.c_int),
        ("depth", C.c_double),
        ("patch_with_border", C.c_uint8 * 100),
        ("epi_image", C.c_double * 2),
    ]


def default_align_options(**kw):
    """SparseImgAlignOptions defaults (sparse_img_align_base.h:37-46) + getDefaultSolverOptions (…base.cpp:35-42)."""
    o = AlignOptions()
    o.max_level, o.min_level = 4, 1
    o.weight_scale = 10.0
    o.max_iter = 10
    o.eps = 0.0005
    for k, v in kw.items():
        if k == "prior_T":
            o.prior_T[:] = list(v)
        else:
            setattr(o, k, v)
    return o


def default_matcher_options(**kw):
    """Matcher::Options defaults (src/svo_direct/include/svo/direct/matcher.h:39-54)."""
    o = MatcherOptions()
    o.align_1d = 0
    o.align_max_iter = 10
    o.max_epi_search_steps = 100
    o.subpix_refinement = 1
    o.epi_search_edgelet_filtering = 1
    o.scan_on_unit_sphere = 1
    o.epi_search_edgelet_max_angle = 0.7
    o.affine_est_offset = 1
    o.affine_est_gain = 0
    o.max_patch_diff_ratio = 2.0
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def build(force=False):
    """Compile liborc.so (and _ref/libfast_ref.so when /root/reference is present)."""
    so = os.path.join(_HERE, "liborc.so")
    if force or not os.path.exists(so) or os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", _HERE, "-s", "all"], check=True)
    return so


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_create_img_pyramid.restype = C.c_size_t
        L.orc_compute_tau.restype = C.c_double
        L.orc_px_error_angle.restype = C.c_double
        L.orc_compute_tau.argtypes = [f64p, f64p, C.c_double, C.c_double]
        L.orc_px_error_angle.argtypes = [C.POINTER(Frame), C.c_double]
        L.orc_update_filter_vogiatzis.argtypes = [C.c_double, C.c_double, C.c_double, f64p]
        L.orc_update_filter_gaussian.argtypes = [C.c_double, C.c_double, f64p]
        L.orc_fast_detect_features.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                               C.c_int, C.c_int, u8p, C.c_int, f64p, f64p, i32p]
        L.orc_find_match_direct.argtypes = [C.POINTER(Frame), C.POINTER(Frame), f64p, C.POINTER(Feature), C.c_double,
                                            f64p, C.POINTER(MatcherOptions), C.POINTER(MatchOut)]
        L.orc_find_epipolar_match_direct.argtypes = [C.POINTER(Frame), C.POINTER(Frame), f64p, C.POINTER(Feature),
                                                     C.c_double, C.c_double, C.c_double, C.POINTER(MatcherOptions),
                                                     C.POINTER(MatchOut)]
        L.orc_get_warp_matrix_affine.argtypes = [C.POINTER(Frame), C.POINTER(Frame), f64p, f64p, C.c_double, f64p,
                                                 C.c_int, f64p]
        L.orc_update_seeds.argtypes = [C.POINTER(Frame), C.c_int, C.POINTER(Frame), f64p, C.c_int, C.POINTER(Feature),
                                       u8p, f64p, C.c_double, C.POINTER(MatcherOptions), C.c_double, C.c_double, C.c_double,
                                       C.c_int, C.c_int, C.c_int, i32p, u8p, C.c_int]
        _lib = L
    return _lib


def ref_lib():
    """The reference's own FAST code (None if it was never built)."""
    global _ref
    if _ref is None:
        so = os.path.join(_HERE, "_ref", "libfast_ref.so")
        if not os.path.exists(so):
            return None
        _ref = C.CDLL(so)
    return _ref


_ref_direct = None


def ref_direct_lib():
    """The reference's own halfSample / align1D / align2D / alignPyr2D / ZMSSD / Tukey / radtan / seed / grid code compiled
    against the container-only shims (oracle/_ref/libdirect_ref.so; None if it was never built)."""
    global _ref_direct
    if _ref_direct is None:
        so = os.path.join(_HERE, "_ref", "libdirect_ref.so")
        if not os.path.exists(so):
            return None
        _ref_direct = C.CDLL(so)
    return _ref_direct


_ref_frontend = None
_frontend_libs = {}
_frontend_variant = "ref"


def use_frontend_lib(variant):
    """Select which library the ref_* front-end entry points go through: "ref" = oracle/_ref/libfrontend_ref.so (the reference's
    own sources), "swap" = oracle/_ref/libfrontend_swap.so (the same wrapper and reference sources with the hot-path functions
    replaced, at link time, by svo_pro_universal_b200/host/ref_swap.cpp calling libsvo_cuda.so — needs a GPU)."""
    global _frontend_variant, _ref_frontend
    assert variant in ("ref", "swap")
    _frontend_variant = variant
    _ref_frontend = _frontend_libs.get(variant)


def ref_frontend_lib():
    """The reference's own SparseImgAlign / Matcher / patch_warp / depth_filter code compiled against the shims
    (oracle/_ref/libfrontend_ref.so; None if it was never built). See use_frontend_lib for the link-time-swap variant."""
    global _ref_frontend
    if _ref_frontend is None:
        so = os.path.join(_HERE, "_ref", "libfrontend_%s.so" % _frontend_variant)
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        _frontend_libs[_frontend_variant] = L
        L.ref_compute_tau.restype = C.c_double
        L.ref_px_error_angle.restype = C.c_double
        L.ref_compute_tau.argtypes = [f64p, f64p, C.c_double, C.c_double]
        L.ref_px_error_angle.argtypes = [C.POINTER(Frame), C.c_double]
        L.ref_update_filter_vogiatzis.argtypes = [C.c_double, C.c_double, C.c_double, f64p]
        L.ref_update_filter_gaussian.argtypes = [C.c_double, C.c_double, f64p]
        L.ref_find_match_direct.argtypes = [C.POINTER(Frame), C.POINTER(Frame), f64p, C.POINTER(Feature), C.c_double, f64p,
                                            C.POINTER(MatcherOptions), C.POINTER(MatchOut)]
        L.ref_find_epipolar_match_direct.argtypes = [C.POINTER(Frame), C.POINTER(Frame), f64p, C.POINTER(Feature), C.c_double,
                                                     C.c_double, C.c_double, C.POINTER(MatcherOptions), C.POINTER(MatchOut)]
        L.ref_get_warp_matrix_affine.argtypes = [C.POINTER(Frame), C.POINTER(Frame), f64p, f64p, C.c_double, f64p, C.c_int, f64p]
        L.ref_update_seeds.argtypes = [C.POINTER(Frame), C.c_int, C.POINTER(Frame), f64p, C.c_int, C.POINTER(Feature), u8p, f64p,
                                       C.c_double, C.POINTER(MatcherOptions), C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                       i32p, u8p]
        _ref_frontend = L
    return _ref_frontend


def _u8(a):
    return a.ctypes.data_as(u8p)


def _f64(a):
    return a.ctypes.data_as(f64p)


def _i32(a):
    return a.ctypes.data_as(i32p)


def pyramid_level_sizes(cols, rows, n_levels):
    out = [(cols, rows)]
    for _ in range(1, n_levels):
        cols, rows = cols // 2, rows // 2
        out.append((cols, rows))
    return out


def create_img_pyramid(img0, n_levels, mode=-1):
    """frame_utils::createImgPyramid -> list of tight uint8 arrays (level 0 is the input)."""
    img0 = np.ascontiguousarray(img0, dtype=np.uint8)
    rows, cols = img0.shape
    sizes = pyramid_level_sizes(cols, rows, n_levels)
    total = sum(c * r for c, r in sizes[1:])
    buf = np.zeros(max(total, 1), np.uint8)
    lib().orc_create_img_pyramid(_u8(img0), cols, rows, n_levels, _u8(buf), mode)
    out, off = [img0], 0
    for c, r in sizes[1:]:
        out.append(buf[off:off + c * r].reshape(r, c).copy())
        off += c * r
    return out


def fast_detect(img, barrier, arc=10, which="orc"):
    """Returns (xy[n,2] int16) of the segment test. which: 'orc' restatement, 'ref_sse2', 'ref_plain10', 'ref_plain9'."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = w * h
    xy = np.zeros((cap, 2), np.int16)
    p = xy.ctypes.data_as(i16p)
    if which == "orc":
        n = lib().orc_fast_detect(_u8(img), w, h, w, barrier, arc, p, cap)
    else:
        sel = {"ref_sse2": 0, "ref_plain10": 1, "ref_plain9": 2}[which]
        n = ref_lib().ref_fast_detect(_u8(img), w, h, w, barrier, sel, p, cap)
    return xy[:n].copy()


def fast_score10(img, xy, threshold, which="orc"):
    img = np.ascontiguousarray(img, np.uint8)
    xy = np.ascontiguousarray(xy, np.int16)
    n = len(xy)
    s = np.zeros(max(n, 1), np.int32)
    fn = lib().orc_fast_score10 if which == "orc" else ref_lib().ref_fast_score10
    fn(_u8(img), img.shape[1], xy.ctypes.data_as(i16p), n, threshold, _i32(s))
    return s[:n]


def fast_nonmax3x3(xy, scores, which="orc"):
    xy = np.ascontiguousarray(xy, np.int16)
    scores = np.ascontiguousarray(scores, np.int32)
    n = len(xy)
    idx = np.zeros(max(n, 1), np.int32)
    fn = lib().orc_fast_nonmax3x3 if which == "orc" else ref_lib().ref_fast_nonmax3x3
    m = fn(xy.ctypes.data_as(i16p), _i32(scores), n, _i32(idx))
    return idx[:m].copy()


def fast_detector(img0, n_levels=5, pyr_mode=-1, threshold=10, border=8, min_level=0, max_level=2, cell_size=30,
                  occupancy=None):
    """feature_detection_utils::fastDetector on a fresh pyramid -> structured array of per-cell Corners."""
    img0 = np.ascontiguousarray(img0, np.uint8)
    rows, cols = img0.shape
    n_cols = -(-cols // cell_size)
    n_rows = -(-rows // cell_size)
    corners = (Corner * (n_cols * n_rows))()
    occ = None if occupancy is None else _u8(np.ascontiguousarray(occupancy, np.uint8))
    lib().orc_fast_detector(_u8(img0), cols, rows, n_levels, pyr_mode, threshold, border, min_level, max_level, cell_size,
                            occ, corners)
    dt = np.dtype([("x", "<i4"), ("y", "<i4"), ("level", "<i4"), ("score", "<f4"), ("angle", "<f4")])
    return np.frombuffer(corners, dtype=dt).copy()


def make_frame(pyr, cam, T_cam_imu=None, T_imu_world=None, px=None, f=None, depth=None, eligible=None, keep=None):
    """Build an orc Frame struct over numpy level arrays. `cam` = dict(fx,fy,cx,cy,k1,k2,p1,p2,width,height,distortion).
    `keep` is a list that receives references keeping the numpy buffers alive."""
    fr = Frame()
    fr.n_levels = len(pyr)
    for i, lv in enumerate(pyr):
        assert lv.dtype == np.uint8 and lv.flags["C_CONTIGUOUS"]
        fr.level_data[i] = _u8(lv)
        fr.level_cols[i] = lv.shape[1]
        fr.level_rows[i] = lv.shape[0]
        fr.level_step[i] = lv.strides[0]
    fr.cam[:] = [cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam.get("k1", 0.0), cam.get("k2", 0.0), cam.get("p1", 0.0),
                 cam.get("p2", 0.0)]
    fr.width, fr.height, fr.distortion = cam["width"], cam["height"], cam.get("distortion", 0)
    ident = [1.0, 0, 0, 0, 0, 0, 0]
    fr.T_cam_imu[:] = list(T_cam_imu) if T_cam_imu is not None else ident
    fr.T_imu_world[:] = list(T_imu_world) if T_imu_world is not None else ident
    bufs = [pyr]
    if px is not None:
        px = np.ascontiguousarray(px, np.float64)
        f = np.ascontiguousarray(f, np.float64)
        depth = np.ascontiguousarray(depth, np.float64)
        n = len(depth)
        eligible = np.ones(n, np.uint8) if eligible is None else np.ascontiguousarray(eligible, np.uint8)
        fr.n_features = n
        fr.px, fr.f, fr.depth, fr.eligible = _f64(px), _f64(f), _f64(depth), _u8(eligible)
        bufs += [px, f, depth, eligible]
    if keep is not None:
        keep.append(bufs)
    else:
        fr._keep = bufs
    return fr


def sparse_align(ref_frames, cur_frames, opt):
    """SparseImgAlign::run on one bundle (lists of Frame structs)."""
    n = len(ref_frames)
    R = (Frame * n)(*ref_frames)
    Cc = (Frame * n)(*cur_frames)
    res = AlignResult()
    lib().orc_sparse_align(n, R, Cc, C.byref(opt), C.byref(res))
    return res


def sparse_align_batch(ref_frames, cur_frames, n_cams, opt, n_threads=1):
    B = len(ref_frames) // n_cams
    R = (Frame * len(ref_frames))(*ref_frames)
    Cc = (Frame * len(cur_frames))(*cur_frames)
    res = (AlignResult * B)()
    lib().orc_sparse_align_batch(B, n_cams, R, Cc, C.byref(opt), res, n_threads)
    return res


def make_features(px, f, grad, ftype, level):
    """Array of orc Feature structs from SoA numpy inputs."""
    n = len(px)
    arr = (Feature * n)()
    for i in range(n):
        arr[i].type = int(ftype[i])
        arr[i].px[:] = [float(px[i][0]), float(px[i][1])]
        arr[i].f[:] = [float(f[i][0]), float(f[i][1]), float(f[i][2])]
        arr[i].grad[:] = [float(grad[i][0]), float(grad[i][1])]
        arr[i].level = int(level[i])
    return arr


MATCH_OUT_NP = np.dtype([("result", "<i4"), ("px_cur", "<f8", 2), ("f_cur", "<f8", 3), ("search_level", "<i4"),
                         ("A_cur_ref", "<f8", 4), ("h_inv", "<f8"), ("epi_length_pyramid", "<f8"), ("reject", "<i4"),
                         ("depth", "<f8"), ("patch_with_border", "u1", 100), ("epi_image", "<f8", 2)], align=True)


def find_match_direct_batch(ref, cur, T_cur_ref, ftrs, ref_depth, px_guess, opt, n_threads=1):
    M = len(ftrs)
    out = (MatchOut * M)()
    T = np.ascontiguousarray(T_cur_ref, np.float64)
    dep = np.ascontiguousarray(ref_depth, np.float64)
    pg = np.ascontiguousarray(px_guess, np.float64)
    lib().orc_find_match_direct_batch(C.byref(ref), C.byref(cur), _f64(T), M, ftrs, _f64(dep), _f64(pg), C.byref(opt), out, n_threads)
    assert C.sizeof(MatchOut) == MATCH_OUT_NP.itemsize
    return np.frombuffer(out, dtype=MATCH_OUT_NP).copy()


def find_epipolar_match_direct_batch(ref, cur, T_cur_ref, ftrs, d_inv3, opt, n_threads=1):
    M = len(ftrs)
    out = (MatchOut * M)()
    T = np.ascontiguousarray(T_cur_ref, np.float64)
    d3 = np.ascontiguousarray(d_inv3, np.float64)
    lib().orc_find_epipolar_match_direct_batch(C.byref(ref), C.byref(cur), _f64(T), M, ftrs, _f64(d3), C.byref(opt), out, n_threads)
    return np.frombuffer(out, dtype=MATCH_OUT_NP).copy()


def scan_epipolar_line(cur, A, B, Cpt, patch64, patch_level, epi_length_pyramid, opt, zmssd_best=2000 * 64, which="orc"):
    """Matcher::scanEpipolarLine on its own: (image_best [2], zmssd_best). which = "orc" (restatement) or "ref" (the compiled
    reference / the swap library, see use_frontend_lib)."""
    if which == "orc":
        fn = lib().orc_scan_epipolar_line
    else:
        fn = ref_frontend_lib().ref_scan_epipolar_line
    fn.restype = None
    fn.argtypes = [C.POINTER(Frame), f64p, f64p, f64p, u8p, C.c_int, C.c_double, C.POINTER(MatcherOptions), f64p, i32p]
    a, b, c = (np.ascontiguousarray(v, np.float64) for v in (A, B, Cpt))
    patch = np.ascontiguousarray(patch64, np.uint8).reshape(64)
    best = np.zeros(2)
    z = np.array([zmssd_best], np.int32)
    fn(C.byref(cur), _f64(a), _f64(b), _f64(c), _u8(patch), int(patch_level), float(epi_length_pyramid), C.byref(opt), _f64(best),
       z.ctypes.data_as(i32p))
    return best, int(z[0])


def update_seeds(ref, cur_frames, T_cur_ref, ftrs, types, states, mu_range, opt, sigma2_thresh=200.0, mappoint_thresh=500.0,
                 px_error_angle=None, check_visibility=1, check_convergence=0, use_vogiatzis=1, n_threads=1):
    """depth_filter_utils::updateSeed for S seeds x n_obs ordered observations. types/states are modified in place."""
    n_obs, S = len(cur_frames), len(ftrs)
    Cf = (Frame * n_obs)(*cur_frames)
    T = np.ascontiguousarray(T_cur_ref, np.float64)
    mr = np.full((n_obs, S), -1, np.int32)
    ok = np.zeros((n_obs, S), np.uint8)
    if px_error_angle is None:
        px_error_angle = lib().orc_px_error_angle(C.byref(cur_frames[0]), 1.0)
    n = lib().orc_update_seeds(C.byref(ref), n_obs, Cf, _f64(T), S, ftrs, _u8(types), _f64(states), mu_range, C.byref(opt),
                               sigma2_thresh, mappoint_thresh, px_error_angle, check_visibility, check_convergence, use_vogiatzis,
                               _i32(mr), _u8(ok), n_threads)
    return n, mr, ok


def pyramid_align_batch(cur_l0_list, ref_frames, cur_frames, opt, n_levels=5, n_threads=1, pyr_mode=-1):
    """Threaded CPU 'frame pair step': pyramid of the new frame + SparseImgAlign::run (bench cpu_baseline / reference arm)."""
    B = len(cur_l0_list)
    rows, cols = cur_l0_list[0].shape
    ptrs = (u8p * B)(*[_u8(a) for a in cur_l0_list])
    R = (Frame * B)(*ref_frames)
    Cc = (Frame * B)(*cur_frames)
    res = (AlignResult * B)()
    lib().orc_pyramid_align_batch(B, n_levels, ptrs, cols, rows, pyr_mode, R, Cc, C.byref(opt), res, n_threads)
    return res


def ref_pyramid_align_batch(cur_l0_list, ref_frames, cur_frames, opt, n_levels=5, n_threads=1):
    """The same frame-pair step on the REFERENCE's own compiled code (vk::halfSample pyramid + SparseImgAlign::run,
    oracle/_ref/libfrontend_ref.so). Returns None when that library was never built."""
    L = ref_frontend_lib()
    if L is None:
        return None
    B = len(cur_l0_list)
    rows, cols = cur_l0_list[0].shape
    ptrs = (u8p * B)(*[_u8(a) for a in cur_l0_list])
    R = (Frame * B)(*ref_frames)
    Cc = (Frame * B)(*cur_frames)
    res = (AlignResult * B)()
    L.ref_pyramid_align_batch(B, n_levels, ptrs, cols, rows, R, Cc, C.byref(opt), res, n_threads)
    return res


# ---- one-function-at-a-time entry points, `which` = "orc" (the restatement) or "ref" (the compiled reference) ------------
def _which(which):
    if which == "orc":
        return lib(), "orc_"
    L = ref_direct_lib()
    if L is None:
        raise RuntimeError("oracle/_ref/libdirect_ref.so not built")
    return L, "ref_"


def ref_create_img_pyramid(img0, n_levels):
    """frame_utils::createImgPyramid through the reference's own vk::halfSample (64-byte aligned continuous buffers)."""
    img0 = np.ascontiguousarray(img0, dtype=np.uint8)
    rows, cols = img0.shape
    sizes = pyramid_level_sizes(cols, rows, n_levels)
    levels = [img0] + [np.zeros((r, c), np.uint8) for c, r in sizes[1:]]
    ptrs = (C.c_void_p * n_levels)(*[l.ctypes.data for l in levels])
    ref_direct_lib().ref_create_img_pyramid(_u8(img0), cols, rows, cols, n_levels, ptrs)
    return levels


def ref_half_sample(img, align_offset=0, extra_stride=0):
    """One vk::halfSample call; align_offset / extra_stride move the input off 16-byte alignment / continuity to reach the
    non-SSE2 branch (vision.cpp:80-87)."""
    img = np.ascontiguousarray(img, np.uint8)
    rows, cols = img.shape
    stride = cols + extra_stride
    raw = np.zeros(stride * rows + 128, np.uint8)
    base = (-raw.ctypes.data) % 64 + align_offset
    view = np.lib.stride_tricks.as_strided(raw[base:], (rows, cols), (stride, 1))
    view[:] = img
    out = np.zeros((rows // 2, cols // 2), np.uint8)
    ref_direct_lib().ref_half_sample(C.c_void_p(raw.ctypes.data + base), cols, rows, stride, _u8(out), cols // 2)
    return out


def align2d(img, patch_with_border, px, n_iter=10, est_offset=True, est_gain=False, which="orc"):
    img = np.ascontiguousarray(img, np.uint8)
    pwb = np.ascontiguousarray(patch_with_border, np.uint8).reshape(100).copy()
    p = np.array(px, np.float64)
    L, pre = _which(which)
    if which == "orc":
        ok = L.orc_align2d(_u8(img), img.shape[1], img.shape[0], img.strides[0], _u8(pwb), n_iter, int(est_offset), int(est_gain), _f64(p))
    else:
        patch = pwb.reshape(10, 10)[1:9, 1:9].copy()
        ok = L.ref_align2d(_u8(img), img.shape[1], img.shape[0], img.strides[0], _u8(pwb), _u8(patch), n_iter, int(est_offset), int(est_gain), _f64(p))
    return bool(ok), p


def align1d(img, direction, patch_with_border, px, n_iter=10, est_offset=True, est_gain=False, which="orc"):
    img = np.ascontiguousarray(img, np.uint8)
    pwb = np.ascontiguousarray(patch_with_border, np.uint8).reshape(100).copy()
    p = np.array(px, np.float64)
    d = np.array(direction, np.float64)
    hinv = C.c_double(0.0)
    L, pre = _which(which)
    if which == "orc":
        ok = L.orc_align1d(_u8(img), img.shape[1], img.shape[0], img.strides[0], _f64(d), _u8(pwb), n_iter, int(est_offset), int(est_gain), _f64(p), C.byref(hinv))
    else:
        patch = pwb.reshape(10, 10)[1:9, 1:9].copy()
        ok = L.ref_align1d(_u8(img), img.shape[1], img.shape[0], img.strides[0], _f64(d), _u8(pwb), _u8(patch), n_iter, int(est_offset), int(est_gain), _f64(p), C.byref(hinv))
    return bool(ok), p, hinv.value


def zmssd(ref_patch64, img, xy, which="orc"):
    """ZMSSD<4> of the 8x8 reference patch against the 8x8 windows of img whose top-left corners are xy [n][2]."""
    img = np.ascontiguousarray(img, np.uint8)
    rp = np.zeros(64 + 16, np.uint8)
    off = (-rp.ctypes.data) % 16  # the SSE paths load the template with aligned loads
    rp[off:off + 64] = np.ascontiguousarray(ref_patch64, np.uint8).reshape(64)
    xy = np.asarray(xy, np.int64).reshape(-1, 2)
    out = np.zeros(len(xy), np.int32)
    if which == "orc":
        for i, (x, y) in enumerate(xy):
            out[i] = lib().orc_zmssd(C.c_void_p(rp.ctypes.data + off), C.c_void_p(img.ctypes.data + int(y) * img.strides[0] + int(x)), img.strides[0])
    else:
        offs = np.ascontiguousarray(xy[:, 1] * img.strides[0] + xy[:, 0], np.int32)
        ref_direct_lib().ref_zmssd(C.c_void_p(rp.ctypes.data + off), _u8(img), img.strides[0], _i32(offs), len(xy), _i32(out), None)
    return out


def tukey_weight(err, b=4.6851, which="orc"):
    e = np.ascontiguousarray(err, np.float32)
    w = np.zeros_like(e)
    L, pre = _which(which)
    getattr(L, pre + "tukey_weight")(C.c_float(b), e.ctypes.data_as(C.c_void_p), len(e), w.ctypes.data_as(C.c_void_p))
    return w


def radtan(k, xy, mode, which="orc"):
    """mode: 'distort' | 'undistort' | 'jacobian' on normalised points xy [n][2]."""
    pts = np.array(xy, np.float64).reshape(-1, 2).copy()
    jac = np.zeros((len(pts), 4), np.float64)
    L, pre = _which(which)
    getattr(L, pre + "radtan")(C.c_double(k[0]), C.c_double(k[1]), C.c_double(k[2]), C.c_double(k[3]),
                               {"distort": 0, "undistort": 1, "jacobian": 2}[mode], _f64(pts), len(pts), _f64(jac))
    return jac if mode == "jacobian" else pts


def seed_helpers(state4, mu_range, thresh, depth, depth_sigma, which="orc"):
    s = np.array(state4, np.float64)
    out = np.zeros(6, np.float64)
    L, pre = _which(which)
    getattr(L, pre + "seed_helpers")(_f64(s), C.c_double(mu_range), C.c_double(thresh), C.c_double(depth), C.c_double(depth_sigma), _f64(out))
    return out


def grid_cell_index(cell_size, n_cols, n_rows, xy, scale, which="orc"):
    xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
    sc = np.ascontiguousarray(scale, np.int32)
    out = np.zeros(len(xy), np.int64)
    if which == "orc":
        lib().orc_grid_cell_index(cell_size, n_cols, _i32(xy), _i32(sc), len(xy), out.ctypes.data_as(C.c_void_p))
    else:
        ref_direct_lib().ref_grid_cell_index(cell_size, n_cols, n_rows, _i32(xy), _i32(sc), len(xy), out.ctypes.data_as(C.c_void_p))
    return out


def patch_from_patch_with_border(pwb, which="orc"):
    pwb = np.ascontiguousarray(pwb, np.uint8).reshape(100)
    out = np.zeros(64, np.uint8)
    L, pre = _which(which)
    getattr(L, pre + "patch_from_patch_with_border")(_u8(pwb), 8, _u8(out))
    return out


# ---- the compiled reference front-end (oracle/_ref/libfrontend_ref.so), same structs as the restatement ---------------------
def ref_sparse_align(ref_frames, cur_frames, opt):
    n = len(ref_frames)
    R = (Frame * n)(*ref_frames)
    Cc = (Frame * n)(*cur_frames)
    res = AlignResult()
    ref_frontend_lib().ref_sparse_align(n, R, Cc, C.byref(opt), C.byref(res))
    return res


def ref_find_match_direct_batch(ref, cur, T_cur_ref, ftrs, ref_depth, px_guess, opt):
    M = len(ftrs)
    out = (MatchOut * M)()
    T = np.ascontiguousarray(T_cur_ref, np.float64)
    L = ref_frontend_lib()
    for i in range(M):
        pg = np.ascontiguousarray(px_guess[i], np.float64)
        L.ref_find_match_direct(C.byref(ref), C.byref(cur), _f64(T), C.byref(ftrs[i]), float(ref_depth[i]), _f64(pg), C.byref(opt), C.byref(out[i]))
    return np.frombuffer(out, dtype=MATCH_OUT_NP).copy()


def ref_find_epipolar_match_direct_batch(ref, cur, T_cur_ref, ftrs, d_inv3, opt):
    M = len(ftrs)
    out = (MatchOut * M)()
    T = np.ascontiguousarray(T_cur_ref, np.float64)
    L = ref_frontend_lib()
    for i in range(M):
        L.ref_find_epipolar_match_direct(C.byref(ref), C.byref(cur), _f64(T), C.byref(ftrs[i]), float(d_inv3[i][0]), float(d_inv3[i][1]),
                                         float(d_inv3[i][2]), C.byref(opt), C.byref(out[i]))
    return np.frombuffer(out, dtype=MATCH_OUT_NP).copy()


def ref_update_seeds(ref, cur_frames, T_cur_ref, ftrs, types, states, mu_range, opt, sigma2_thresh=200.0, mappoint_thresh=500.0,
                     check_visibility=1, check_convergence=0, use_vogiatzis=1):
    n_obs, S = len(cur_frames), len(ftrs)
    Cf = (Frame * n_obs)(*cur_frames)
    T = np.ascontiguousarray(T_cur_ref, np.float64)
    ok = np.zeros((n_obs, S), np.uint8)
    n = ref_frontend_lib().ref_update_seeds(C.byref(ref), n_obs, Cf, _f64(T), S, ftrs, _u8(types), _f64(states), mu_range, C.byref(opt),
                                            sigma2_thresh, mappoint_thresh, check_visibility, check_convergence, use_vogiatzis,
                                            None, _u8(ok))
    return n, ok


def align_pyr2d(ref_pyr, cur_pyr, px_ref_level_0, px_cur, max_level, min_level, patch_sizes, n_iter=30, min_update_squared=0.03 ** 2,
                which="orc", n_threads=1):
    """feature_alignment::alignPyr2D for M features sharing two pyramids (lists of uint8 level arrays).
    Returns (px_cur [M][2] float64, status [M] uint8)."""
    n_levels = len(ref_pyr)
    pr = np.ascontiguousarray(px_ref_level_0, np.int32).reshape(-1, 2)
    pc = np.array(px_cur, np.float64).reshape(-1, 2).copy()
    M = len(pr)
    st = np.zeros(M, np.uint8)
    ps = np.ascontiguousarray(patch_sizes, np.int32)
    assert len(ps) >= n_levels
    if which == "orc":
        cam = dict(fx=1.0, fy=1.0, cx=0.0, cy=0.0, width=ref_pyr[0].shape[1], height=ref_pyr[0].shape[0])
        rf, cf = make_frame(ref_pyr, cam), make_frame(cur_pyr, cam)
        lib().orc_align_pyr2d(C.byref(rf), C.byref(cf), max_level, min_level, _i32(ps), n_iter, C.c_float(min_update_squared), M, _i32(pr),
                              _f64(pc), _u8(st), n_threads)
    else:
        L = ref_direct_lib()
        rl = (C.c_void_p * n_levels)(*[l.ctypes.data for l in ref_pyr])
        cl = (C.c_void_p * n_levels)(*[l.ctypes.data for l in cur_pyr])
        ws = np.array([l.shape[1] for l in ref_pyr], np.int32)
        hs = np.array([l.shape[0] for l in ref_pyr], np.int32)
        ss = np.array([l.strides[0] for l in ref_pyr], np.int32)
        for i in range(M):
            p = pc[i].copy()
            st[i] = L.ref_align_pyr2d(rl, cl, _i32(ws), _i32(hs), _i32(ss), n_levels, max_level, min_level, _i32(ps), len(ps), n_iter,
                                      C.c_float(min_update_squared), _i32(pr[i].copy()), _f64(p))
            pc[i] = p
    return pc, st


# ---- f1: Reprojector candidate matching ----------------------------------------------------------------------------------------
class ReprojMap(C.Structure):
    _fields_ = [("n_kfs", C.c_int), ("kfs", C.POINTER(Frame)), ("kf_seed_mu_range", f64p), ("kf_feat_begin", i32p),
                ("feat", C.POINTER(Feature)), ("feat_score", f64p), ("feat_seed_state", f64p), ("feat_point", i32p), ("feat_kf", i32p),
                ("n_points", C.c_int), ("pt_pos", f64p), ("pt_n_failed", i32p), ("pt_n_succeeded", i32p), ("pt_obs_begin", i32p),
                ("obs_feat", i32p)]


class ReprojOptions(C.Structure):
    _fields_ = [("cell_size", C.c_int), ("max_n_features", C.c_int), ("affine_est_offset", C.c_int), ("affine_est_gain", C.c_int),
                ("sort_by_num_obs", C.c_int), ("seed_sigma2_thresh", C.c_double), ("px_error_angle", C.c_double)]


REPROJ_RESULT_DTYPE = np.dtype([("cur_px", "<f8", 2), ("px", "<f8", 2), ("f", "<f8", 3), ("grad", "<f8", 2), ("seed_state", "<f8", 4),
                                ("status", "<i4"), ("order", "<i4"), ("slot", "<i4"), ("level", "<i4"), ("type_out", "<i4"),
                                ("match_result", "<i4"), ("d_failed", "<i4"), ("d_succeeded", "<i4")])
REPROJ_STATS_DTYPE = np.dtype([("n_candidates", "<i4"), ("n_trials", "<i4"), ("n_matches", "<i4"), ("n_consumed", "<i4")])


def reproject_match(kf_frames, tables, cur_frame, entry_feat, n_features_in, occupancy, opt, which="orc"):
    """One current frame through getCandidate + sort + matchCandidates. kf_frames: Frame structs (pyramid, camera, pose) of the
    keyframes; tables: dict of numpy arrays (see orc_reproj_map); occupancy uint8 [n_cells] updated in place.
    which = "orc" (restatement) or "ref" (the reference's own reprojector.cpp, oracle/_ref/libfrontend_ref.so)."""
    L = lib() if which == "orc" else ref_frontend_lib()
    fn = L.orc_reproject_match if which == "orc" else L.ref_reproject_match
    K = len(kf_frames)
    kfs = (Frame * K)(*kf_frames)
    feats = make_features(tables["feat"]["px"], tables["feat"]["f"], tables["feat"]["grad"], tables["feat"]["type"], tables["feat"]["level"])
    keep = {k: np.ascontiguousarray(tables[k], dt) for k, dt in (
        ("kf_seed_mu_range", np.float64), ("kf_feat_begin", np.int32), ("feat_score", np.float64), ("feat_seed_state", np.float64),
        ("feat_point", np.int32), ("feat_kf", np.int32), ("pt_pos", np.float64), ("pt_n_failed", np.int32),
        ("pt_n_succeeded", np.int32), ("pt_obs_begin", np.int32), ("obs_feat", np.int32))}
    m = ReprojMap(K, kfs, _f64(keep["kf_seed_mu_range"]), _i32(keep["kf_feat_begin"]), feats, _f64(keep["feat_score"]),
                  _f64(keep["feat_seed_state"]), _i32(keep["feat_point"]), _i32(keep["feat_kf"]), int(tables["n_points"]),
                  _f64(keep["pt_pos"]), _i32(keep["pt_n_failed"]), _i32(keep["pt_n_succeeded"]), _i32(keep["pt_obs_begin"]),
                  _i32(keep["obs_feat"]))
    ef = np.ascontiguousarray(entry_feat, np.int32)
    res = np.zeros(len(ef), REPROJ_RESULT_DTYPE)
    st = np.zeros(1, REPROJ_STATS_DTYPE)
    fn(C.byref(m), C.byref(cur_frame), len(ef), _i32(ef), int(n_features_in), _u8(occupancy), C.byref(opt),
       res.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p))
    return res, st[0]


def ref_reproject_frames(kf_frames, tables, n_visible, cur_frame, max_n_features=120, reproject_unconverged_seeds=True,
                         max_unconverged_seeds_ratio=-1.0, min_required_features=0, remove_unconstrained_points=True):
    """The reference's whole Reprojector::reprojectFrames (compiled reprojector.cpp) on the flat tables; the first n_visible
    keyframes are the visible keyframes. Returns a dict of numpy arrays (features appended to the current frame, grid, stats,
    landmark counters, seed states / types of the keyframes afterwards). None when the library was never built."""
    L = ref_frontend_lib()
    if L is None:
        return None
    K = len(kf_frames)
    kfs = (Frame * K)(*kf_frames)
    feats = make_features(tables["feat"]["px"], tables["feat"]["f"], tables["feat"]["grad"], tables["feat"]["type"], tables["feat"]["level"])
    keep = {k: np.ascontiguousarray(tables[k], dt) for k, dt in (
        ("kf_seed_mu_range", np.float64), ("kf_feat_begin", np.int32), ("feat_score", np.float64), ("feat_seed_state", np.float64),
        ("feat_point", np.int32), ("feat_kf", np.int32), ("pt_pos", np.float64), ("pt_n_failed", np.int32),
        ("pt_n_succeeded", np.int32), ("pt_obs_begin", np.int32), ("obs_feat", np.int32))}
    m = ReprojMap(K, kfs, _f64(keep["kf_seed_mu_range"]), _i32(keep["kf_feat_begin"]), feats, _f64(keep["feat_score"]),
                  _f64(keep["feat_seed_state"]), _i32(keep["feat_point"]), _i32(keep["feat_kf"]), int(tables["n_points"]),
                  _f64(keep["pt_pos"]), _i32(keep["pt_n_failed"]), _i32(keep["pt_n_succeeded"]), _i32(keep["pt_obs_begin"]),
                  _i32(keep["obs_feat"]))
    cap = int(tables["n_feat"]) + 8
    o = dict(type=np.zeros(cap, np.int32), px=np.zeros((cap, 2)), level=np.zeros(cap, np.int32), point=np.zeros(cap, np.int32),
             seed_feat=np.zeros(cap, np.int32), state=np.zeros((cap, 4)), f=np.zeros((cap, 3)), grad=np.zeros((cap, 2)),
             score=np.zeros(cap), occupancy=np.zeros(4096, np.uint8), stats=np.zeros(3, np.int32),
             pt_counters=np.zeros((max(int(tables["n_points"]), 1), 2), np.int32), feat_state=np.zeros((int(tables["n_feat"]), 4)),
             feat_type=np.zeros(int(tables["n_feat"]), np.int32))
    L.ref_reproject_frames.argtypes = [C.POINTER(ReprojMap), C.c_int, C.POINTER(Frame), C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                       i32p, f64p, i32p, i32p, i32p, f64p, f64p, f64p, f64p, u8p, i32p, i32p, f64p, i32p]
    n = L.ref_reproject_frames(C.byref(m), int(n_visible), C.byref(cur_frame), int(max_n_features), int(reproject_unconverged_seeds),
                               float(max_unconverged_seeds_ratio), int(min_required_features), int(remove_unconstrained_points),
                               _i32(o["type"]), _f64(o["px"]), _i32(o["level"]), _i32(o["point"]), _i32(o["seed_feat"]), _f64(o["state"]),
                               _f64(o["f"]), _f64(o["grad"]), _f64(o["score"]), _u8(o["occupancy"]), _i32(o["stats"]),
                               _i32(o["pt_counters"]), _f64(o["feat_state"]), _i32(o["feat_type"]))
    for k in ("type", "px", "level", "point", "seed_feat", "state", "f", "grad", "score"):
        o[k] = o[k][:n]
    o["n"] = n
    return o


# ---- f4: PoseOptimizer ---------------------------------------------------------------------------------------------------------
class PoseOptOptions(C.Structure):
    _fields_ = [("err_type", C.c_int), ("max_iter", C.c_int), ("eps", C.c_double), ("reproj_thresh_px", C.c_double),
                ("have_prior", C.c_int), ("prior_q", C.c_double * 4), ("prior_lambda", C.c_double)]


def pose_opt_options(err_type=0, max_iter=10, eps=0.000001, reproj_thresh_px=2.0, prior_q=None, prior_lambda=0.0):
    """PoseOptimizer::getDefaultSolverOptions (pose_optimizer.cpp:22-29), poseoptim_thresh default 2.0 px."""
    o = PoseOptOptions(err_type, max_iter, eps, reproj_thresh_px, 0, (C.c_double * 4)(1, 0, 0, 0), prior_lambda)
    if prior_q is not None:
        o.have_prior = 1
        o.prior_q[:] = list(prior_q)
    return o


def pose_optimize(case, opt, which="orc"):
    """PoseOptimizer::run on a synth.make_pose_opt_case dict. Returns (n_final, T_imu_world[7], outlier[N], stats[6])."""
    L = lib() if which == "orc" else ref_frontend_lib()
    fn = L.orc_pose_optimize if which == "orc" else L.ref_pose_optimize
    keep = []
    dummy = [np.zeros((8, 8), np.uint8)]
    frames = [make_frame(dummy, case["cam"], T, case["T_imu_world_init"], keep=keep) for T in case["T_cam_imu"]]
    n_cams, N = len(frames), len(case["px"])
    fr = (Frame * n_cams)(*frames)
    ft = make_features(case["px"], case["f"], case["grad"], case["type"], case["level"])
    T = np.zeros(7); outl = np.zeros(N, np.uint8); stats = np.zeros(6)
    xyz = np.ascontiguousarray(case["xyz_world"], np.float64)
    n = fn(n_cams, fr, N, ft, _i32(np.ascontiguousarray(case["feat_cam"], np.int32)), _f64(xyz), _u8(np.ascontiguousarray(case["has_xyz"], np.uint8)),
           C.byref(opt), _f64(T), _u8(outl), _f64(stats))
    return n, T, outl, stats


# ---- f2: edgelet detector / detector classes (restatement = liborc.so, compiled reference = oracle/_ref/libdetect_ref.so) ---------
_ref_detect = None
CORNER_DT = np.dtype([("x", "<i4"), ("y", "<i4"), ("level", "<i4"), ("score", "<f4"), ("angle", "<f4")])
DETECTOR_FAST, DETECTOR_FAST_GRAD, DETECTOR_GRID_GRAD = 0, 2, 5  # svo::DetectorType


def ref_detect_lib():
    """The reference's own feature_detection.cpp / feature_detection_utils.cpp / fast_neon sources compiled against the shims
    (oracle/_ref/libdetect_ref.so; None if it was never built)."""
    global _ref_detect
    if _ref_detect is None:
        so = os.path.join(_HERE, "_ref", "libdetect_ref.so")
        if not os.path.exists(so):
            return None
        _ref_detect = C.CDLL(so)
    return _ref_detect


def _detect_fn(name, which):
    L = lib() if which == "orc" else ref_detect_lib()
    if L is None:
        raise RuntimeError("oracle/_ref/libdetect_ref.so not built")
    fn = getattr(L, ("orc_" if which == "orc" else "ref_") + name)
    return fn


def _pyr_args(pyr, keep):
    lv = [np.ascontiguousarray(p, np.uint8) for p in pyr]
    keep.extend(lv)
    n = len(lv)
    data = (C.c_void_p * n)(*[p.ctypes.data for p in lv])
    cols = (C.c_int * n)(*[p.shape[1] for p in lv])
    rows = (C.c_int * n)(*[p.shape[0] for p in lv])
    step = (C.c_int * n)(*[p.strides[0] for p in lv])
    return n, data, cols, rows, step


def gaussian_blur3x3(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().orc_gaussian_blur3x3(_u8(img), img.shape[1], img.shape[0], img.strides[0], _u8(out))
    return out


def scharr3x3(img):
    img = np.ascontiguousarray(img, np.uint8)
    dx = np.empty(img.shape, np.int16)
    dy = np.empty(img.shape, np.int16)
    lib().orc_scharr3x3(_u8(img), img.shape[1], img.shape[0], img.strides[0], dx.ctypes.data_as(i16p), dy.ctypes.data_as(i16p))
    return dx, dy


def edgelet_detector_v2(pyr, threshold=100, border=8, cell_size=30, occupancy=None, which="orc"):
    """feature_detection_utils::edgeletDetector_V2 on a pyramid (list of level images) -> per-cell Corners."""
    keep = []
    n, data, cols, rows, step = _pyr_args(pyr, keep)
    n_cells = (-(-pyr[0].shape[1] // cell_size)) * (-(-pyr[0].shape[0] // cell_size))
    corners = (Corner * n_cells)()
    occ = None if occupancy is None else _u8(np.ascontiguousarray(occupancy, np.uint8))
    _detect_fn("edgelet_detector_v2", which)(n, data, cols, rows, step, int(threshold), int(border), int(cell_size), occ, corners)
    return np.frombuffer(corners, dtype=CORNER_DT).copy()


def fast_detector_pyr(pyr, threshold=10, border=8, min_level=0, max_level=2, cell_size=30, occupancy=None, which="orc"):
    """feature_detection_utils::fastDetector on a given pyramid -> per-cell Corners."""
    keep = []
    n, data, cols, rows, step = _pyr_args(pyr, keep)
    n_cells = (-(-pyr[0].shape[1] // cell_size)) * (-(-pyr[0].shape[0] // cell_size))
    corners = (Corner * n_cells)()
    occ = None if occupancy is None else _u8(np.ascontiguousarray(occupancy, np.uint8))
    name = "fast_detector_pyr" if which == "orc" else "fast_detector"
    _detect_fn(name, which)(n, data, cols, rows, step, int(threshold), int(border), int(min_level), int(max_level), int(cell_size),
                            occ, corners)
    return np.frombuffer(corners, dtype=CORNER_DT).copy()


def angle_at_pixel_histogram(img, x, y, halfpatch=4, which="orc"):
    img = np.ascontiguousarray(img, np.uint8)
    fn = _detect_fn("angle_at_pixel_histogram", which)
    fn.restype = C.c_double
    return fn(_u8(img), img.shape[1], img.shape[0], img.strides[0], int(x), int(y), int(halfpatch))


def angle_histogram_bin(gx, gy):
    return lib().orc_angle_histogram_bin(int(gx), int(gy))


def detect_features(detector_type, pyr, threshold_primary=10.0, threshold_secondary=100.0, border=8, min_level=0, max_level=2,
                    cell_size=30, occupancy=None, max_n=None, which="orc"):
    """AbstractDetector::detect of the detector makeDetector builds (DETECTOR_FAST / DETECTOR_FAST_GRAD / DETECTOR_GRID_GRAD) with an
    empty mask -> dict(px [n,2], score [n], level [n], grad [n,2], type [n])."""
    keep = []
    n, data, cols, rows, step = _pyr_args(pyr, keep)
    n_cells = (-(-pyr[0].shape[1] // cell_size)) * (-(-pyr[0].shape[0] // cell_size))
    max_n = n_cells if max_n is None else int(max_n)
    cap = max(2 * n_cells, max_n, 1)
    px = np.zeros((cap, 2)); sc = np.zeros(cap); lv = np.zeros(cap, np.int32); gr = np.zeros((cap, 2)); ty = np.zeros(cap, np.int32)
    occ = None if occupancy is None else _u8(np.ascontiguousarray(occupancy, np.uint8))
    fn = _detect_fn("detect_features", which)
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                   C.c_int, u8p, C.c_int, f64p, f64p, i32p, f64p, i32p]
    m = fn(int(detector_type), n, data, cols, rows, step, float(threshold_primary), float(threshold_secondary), int(border),
           int(min_level), int(max_level), int(cell_size), occ, max_n, _f64(px), _f64(sc), _i32(lv), _f64(gr), _i32(ty))
    return {"px": px[:m].copy(), "score": sc[:m].copy(), "level": lv[:m].copy(), "grad": gr[:m].copy(), "type": ty[:m].copy()}


# ---- f3: StereoTriangulation::compute ---------------------------------------------------------------------------------------------
STEREO_RESULT_DT = np.dtype([("px_cur", "<f8", 2), ("f_cur", "<f8", 3), ("grad_cur", "<f8", 2), ("xyz_world", "<f8", 3), ("depth", "<f8"),
                             ("status", "<i4"), ("slot", "<i4"), ("match_result", "<i4"), ("level", "<i4"), ("type", "<i4"), ("_pad", "<i4")])


def stereo_triangulate(frame0, frame1, ftrs, n_desired, n_features_in_frame1=0, mean_depth_inv=1.0 / 3.0, min_depth_inv=1.0,
                       max_depth_inv=1.0 / 50.0):
    """The matching loop of StereoTriangulation::compute on features in visiting order (make_features array)
    -> (results [n] STEREO_RESULT_DT, n_succeeded, n_failed)."""
    n = len(ftrs)
    res = np.zeros(n, STEREO_RESULT_DT)
    nf = C.c_int(0)
    fn = lib().orc_stereo_triangulate
    fn.argtypes = [C.POINTER(Frame), C.POINTER(Frame), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                   C.c_void_p, C.POINTER(C.c_int)]
    ns = fn(C.byref(frame0), C.byref(frame1), n, C.cast(ftrs, C.c_void_p), int(n_desired), int(n_features_in_frame1), mean_depth_inv,
            min_depth_inv, max_depth_inv, res.ctypes.data, C.byref(nf))
    return res, ns, nf.value


def ref_stereo_shuffle_order(seed, n_old, n_corners, n_new):
    """The visiting order the reference's two std::random_shuffle calls produce after srand(seed) (frame0 feature indices)."""
    order = np.zeros(n_new, np.int32)
    ref_frontend_lib().ref_stereo_shuffle_order(C.c_uint(seed), int(n_old), int(n_corners), int(n_new), _i32(order))
    return order


def ref_stereo_triangulation_compute(frame0, frame1, detector_type=DETECTOR_FAST_GRAD, threshold_primary=10.0, threshold_secondary=100.0,
                                     triangulate_n_features=120, mean_depth_inv=1.0 / 3.0, min_depth_inv=1.0, max_depth_inv=1.0 / 50.0,
                                     seed=1, cap=1024):
    """The reference's own StereoTriangulation::compute on two feature-less frames -> dict of frame0's and frame1's new columns."""
    L = ref_frontend_lib()
    fn = L.ref_stereo_triangulation_compute
    fn.argtypes = [C.POINTER(Frame), C.POINTER(Frame), C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                   C.c_uint, C.c_int, C.POINTER(C.c_int)] + [C.c_void_p] * 13
    n0 = C.c_int(0)
    px0 = np.zeros((cap, 2)); level0 = np.zeros(cap, np.int32); type0 = np.zeros(cap, np.int32); score0 = np.zeros(cap); grad0 = np.zeros((cap, 2))
    px1 = np.zeros((cap, 2)); f1 = np.zeros((cap, 3)); grad1 = np.zeros((cap, 2)); level1 = np.zeros(cap, np.int32)
    type1 = np.zeros(cap, np.int32); score1 = np.zeros(cap); xyz1 = np.zeros((cap, 3)); ref1 = np.zeros(cap, np.int32)
    n1 = fn(C.byref(frame0), C.byref(frame1), int(detector_type), threshold_primary, threshold_secondary, int(triangulate_n_features),
            mean_depth_inv, min_depth_inv, max_depth_inv, C.c_uint(seed), cap, C.byref(n0),
            *[a.ctypes.data for a in (px0, level0, type0, score0, grad0, px1, f1, grad1, level1, type1, score1, xyz1, ref1)])
    a, b = n0.value, n1
    return dict(n0=a, px0=px0[:a], level0=level0[:a], type0=type0[:a], score0=score0[:a], grad0=grad0[:a], n1=b, px1=px1[:b], f1=f1[:b],
                grad1=grad1[:b], level1=level1[:b], type1=type1[:b], score1=score1[:b], xyz1=xyz1[:b], ref_index1=ref1[:b])


# ---- f4 (second half): Point::optimize ------------------------------------------------------------------------------------------------
_ref_point = None


def ref_point_lib():
    """The reference's own point.h / point.cpp compiled against the shims (oracle/_ref/libpoint_ref.so; None if never built)."""
    global _ref_point
    if _ref_point is None:
        so = os.path.join(_HERE, "_ref", "libpoint_ref.so")
        if not os.path.exists(so):
            return None
        _ref_point = C.CDLL(so)
    return _ref_point


def point_optimize(T_f_w, f, pos, n_iter=5, using_bearing_vector=False, which="orc"):
    """Point::optimize on one point: T_f_w [n_obs, 7], f [n_obs, 3], pos [3] -> (optimised pos, iterations started or None)."""
    T = np.ascontiguousarray(T_f_w, np.float64).reshape(-1, 7)
    fv = np.ascontiguousarray(f, np.float64).reshape(-1, 3)
    p = np.array(pos, np.float64).reshape(3).copy()
    fn = lib().orc_point_optimize if which == "orc" else ref_point_lib().ref_point_optimize
    it = fn(len(T), _f64(T), _f64(fv), _f64(p), int(n_iter), int(bool(using_bearing_vector)))
    return p, (it if which == "orc" else None)


# ---- f3 (tracker part): FeatureTracker::trackAndDetect ----------------------------------------------------------------------------
KLT_PATCH_SIZES = (16, 16, 16, 8, 8)  # FeatureTrackerOptions defaults (feature_tracking_types.h:11-42)


def feature_tracker_sequence(pyrs, detector_type=DETECTOR_FAST, threshold_primary=10.0, threshold_secondary=100.0, min_tracks_to_detect=50,
                             reset_before_detection=True, template_is_first=True, klt_max_level=4, klt_min_level=0, klt_max_iter=30,
                             klt_min_update_squared=0.001):
    """FeatureTracker::trackAndDetect over a mono sequence of pyramids (src/svo_tracker/src/feature_tracker.cpp:32-187): the host
    bookkeeping restated around the oracle's alignPyr2D and detectors. Returns per frame dict(px, track_id, score, n_active,
    n_terminated, disparity)."""
    tracks = []          # each: dict(id, obs=[(frame index, feature index)])
    frames = []          # per frame: dict(px, score, track_id)
    out, next_id = [], 0
    for k, pyr in enumerate(pyrs):
        cur = dict(px=np.zeros((0, 2)), score=np.zeros(0), track_id=np.zeros(0, np.int64))
        frames.append(cur)
        # trackFrameBundle (:52-127)
        terminated = 0
        if tracks:
            ref_obs = [t["obs"][0] if template_is_first else t["obs"][-1] for t in tracks]
            last_obs = [t["obs"][-1] for t in tracks]
            ref_px = np.array([frames[a]["px"][b] for a, b in ref_obs]).astype(np.int32)      # getPx().cast<int>()
            cur_px = np.array([frames[a]["px"][b] for a, b in last_obs])
            new_px, ok = np.zeros_like(cur_px), np.zeros(len(tracks), bool)
            for fi in sorted(set(a for a, _ in ref_obs)):                                        # one batch per template frame
                sel = np.flatnonzero([a == fi for a, _ in ref_obs])
                p, s = align_pyr2d(pyrs[fi], pyr, ref_px[sel], cur_px[sel], klt_max_level, klt_min_level, KLT_PATCH_SIZES, klt_max_iter,
                                   klt_min_update_squared)
                new_px[sel], ok[sel] = p, s.astype(bool)
            kept, px, sc, ids = [], [], [], []
            for t, (ra, rb), p, good in zip(tracks, ref_obs, new_px, ok):
                if good:
                    px.append(p); sc.append(frames[ra]["score"][rb]); ids.append(t["id"])
                    t["obs"].append((k, len(px) - 1))
                    kept.append(t)
                else:
                    terminated += 1
            tracks = kept
            cur["px"], cur["score"], cur["track_id"] = np.array(px).reshape(-1, 2), np.array(sc), np.array(ids, np.int64)
        # trackAndDetect (:32-50)
        if len(tracks) < min_tracks_to_detect:
            if reset_before_detection:
                tracks = []
                cur["px"], cur["score"], cur["track_id"] = np.zeros((0, 2)), np.zeros(0), np.zeros(0, np.int64)
            # initializeNewTracks (:129-187): grid filled with the frame's keypoints, detect, append, one track per new feature
            n_cols, n_rows = -(-pyr[0].shape[1] // 30), -(-pyr[0].shape[0] // 30)
            occ = np.zeros(n_cols * n_rows, np.uint8)
            for x, y in cur["px"]:
                occ[(int(y) // 30) * n_cols + int(x) // 30] = 1
            det = detect_features(detector_type, pyr, threshold_primary, threshold_secondary, occupancy=occ)
            n_old = len(cur["px"])
            cur["px"] = np.concatenate([cur["px"], det["px"]])
            cur["score"] = np.concatenate([cur["score"], det["score"]])
            new_ids = np.arange(next_id, next_id + len(det["px"]))
            next_id += len(det["px"])
            cur["track_id"] = np.concatenate([cur["track_id"], new_ids])
            for j, tid in enumerate(new_ids):
                tracks.append(dict(id=int(tid), obs=[(k, n_old + j)]))
        disp = [np.linalg.norm(frames[t["obs"][0][0]]["px"][t["obs"][0][1]] - frames[t["obs"][-1][0]]["px"][t["obs"][-1][1]]) for t in tracks]
        pivot = int(np.floor(0.5 * len(disp))) if disp else 0
        out.append(dict(px=cur["px"].copy(), track_id=cur["track_id"].copy(), score=cur["score"].copy(), n_active=len(tracks),
                        n_terminated=terminated, disparity=float(np.sort(disp)[::-1][pivot]) if disp else 0.0))
    return out


def ref_feature_tracker_sequence(frames, detector_type=DETECTOR_FAST, threshold_primary=10.0, threshold_secondary=100.0,
                                 min_tracks_to_detect=50, reset_before_detection=True, template_is_first=True, cap=1024):
    """The reference's own FeatureTracker::trackAndDetect over a mono sequence of oracle frames (oracle/_ref/libfrontend_ref.so)."""
    n = len(frames)
    arr = (Frame * n)(*frames)
    nf = np.zeros(n, np.int32); px = np.zeros((n, cap, 2)); tid = np.zeros((n, cap), np.int32); sc = np.zeros((n, cap))
    na = np.zeros(n, np.int32); nt = np.zeros(n, np.int32); disp = np.zeros(n)
    fn = ref_frontend_lib().ref_feature_tracker_sequence
    fn.argtypes = [C.c_int, C.POINTER(Frame), C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7
    fn(n, arr, int(detector_type), threshold_primary, threshold_secondary, int(min_tracks_to_detect), int(reset_before_detection),
       int(template_is_first), cap, *[a.ctypes.data for a in (nf, px, tid, sc, na, nt, disp)])
    return [dict(px=px[k, :nf[k]].copy(), track_id=tid[k, :nf[k]].astype(np.int64), score=sc[k, :nf[k]].copy(), n_active=int(na[k]),
                 n_terminated=int(nt[k]), disparity=float(disp[k])) for k in range(n)]
