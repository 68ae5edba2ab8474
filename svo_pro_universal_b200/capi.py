"""ctypes binding of libsvo_cuda.so (include/svo_cuda.h) — the C ABI of the B200-native front-end hot path.

This is plumbing for tests, bench.py and Python users; the product is the CUDA library behind the C ABI and the C++
facades in svo_pro_universal_b200/host/. There is NO CPU fallback: if the shared library is missing or no CUDA device is
visible, loading / context creation raises.

Array arguments may be numpy arrays (host memory, SVO_MEM_HOST) or torch CUDA tensors (SVO_MEM_DEVICE); one call must not
mix the two.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvo_cuda.so")

MAX_LEVELS = 8
MAX_CAMS = 4
MEM_HOST, MEM_DEVICE = 0, 1


class SvoCudaError(RuntimeError):
    pass


class Camera(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("k1", C.c_double), ("k2", C.c_double), ("p1", C.c_double), ("p2", C.c_double),
                ("width", C.c_int), ("height", C.c_int), ("distortion", C.c_int), ("_pad", C.c_int)]

    @classmethod
    def from_dict(cls, d):
        return cls(d["fx"], d["fy"], d["cx"], d["cy"], d.get("k1", 0.0), d.get("k2", 0.0), d.get("p1", 0.0),
                   d.get("p2", 0.0), d["width"], d["height"], d.get("distortion", 0), 0)


class DetectorOptions(C.Structure):
    _fields_ = [("threshold", C.c_int), ("border", C.c_int), ("min_level", C.c_int), ("max_level", C.c_int),
                ("cell_size", C.c_int), ("arc_length", C.c_int)]


def detector_options(threshold=10, border=8, min_level=0, max_level=2, cell_size=30, arc_length=10):
    """svo::DetectorOptions defaults (feature_detection_types.h:49-84)."""
    return DetectorOptions(threshold, border, min_level, max_level, cell_size, arc_length)


class SparseAlignOptions(C.Structure):
    _fields_ = [("max_level", C.c_int), ("min_level", C.c_int),
                ("estimate_illumination_gain", C.c_int), ("estimate_illumination_offset", C.c_int),
                ("use_distortion_jacobian", C.c_int), ("robustification", C.c_int),
                ("weight_scale", C.c_double),
                ("max_iter", C.c_int), ("_pad", C.c_int),
                ("eps", C.c_double),
                ("alpha_init", C.c_double), ("beta_init", C.c_double),
                ("lambda_rot", C.c_double), ("lambda_trans", C.c_double),
                ("lambda_alpha", C.c_double), ("lambda_beta", C.c_double)]


def sparse_align_options(**kw):
    """SparseImgAlignOptions defaults (sparse_img_align_base.h:37-46) + getDefaultSolverOptions (…base.cpp:35-42)."""
    o = SparseAlignOptions()
    o.max_level, o.min_level = 4, 1
    o.weight_scale = 10.0
    o.max_iter = 10
    o.eps = 0.0005
    for k, v in kw.items():
        setattr(o, k, v)
    return o


ALIGN_PRIOR_DTYPE = np.dtype([("T", "<f8", 7), ("alpha", "<f8"), ("beta", "<f8")])
ALIGN_RESULT_DTYPE = np.dtype([("T_icur_iref", "<f8", 7), ("T_f_w", "<f8", (MAX_CAMS, 7)), ("alpha", "<f8"), ("beta", "<f8"),
                               ("chi2", "<f8"), ("H", "<f8", 64), ("n_tracked", "<i4"), ("iters", "<i4", MAX_LEVELS),
                               ("stop", "<i4")], align=True)
CORNER_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("level", "<i4"), ("score", "<f4"), ("angle", "<f4")])
FEATURE_DTYPE = np.dtype([("px", "<f8", 2), ("f", "<f8", 3), ("grad", "<f8", 2), ("type", "<i4"), ("level", "<i4")])
MATCH_OUT_DTYPE = np.dtype([("px_cur", "<f8", 2), ("f_cur", "<f8", 3), ("A_cur_ref", "<f8", 4), ("h_inv", "<f8"),
                            ("epi_length_pyramid", "<f8"), ("depth", "<f8"), ("result", "<i4"), ("search_level", "<i4"),
                            ("reject", "<i4"), ("_pad", "<i4"), ("epi_image", "<f8", 2)])
FAST_XY_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2")])  # svo_fast_xy = fast::fast_xy


class MatcherOptions(C.Structure):
    _fields_ = [("align_1d", C.c_int), ("align_max_iter", C.c_int),
                ("max_epi_search_steps", C.c_int),
                ("subpix_refinement", C.c_int), ("epi_search_edgelet_filtering", C.c_int), ("scan_on_unit_sphere", C.c_int),
                ("epi_search_edgelet_max_angle", C.c_double),
                ("affine_est_offset", C.c_int), ("affine_est_gain", C.c_int),
                ("max_patch_diff_ratio", C.c_double)]


def matcher_options(**kw):
    """svo::Matcher::Options defaults (matcher.h:39-54)."""
    o = MatcherOptions(0, 10, 100, 1, 1, 1, 0.7, 1, 0, 2.0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class DepthFilterOptions(C.Structure):
    _fields_ = [("seed_convergence_sigma2_thresh", C.c_double), ("mappoint_convergence_sigma2_thresh", C.c_double),
                ("px_error_angle", C.c_double),
                ("check_visibility", C.c_int), ("check_convergence", C.c_int), ("use_vogiatzis_update", C.c_int),
                ("_pad", C.c_int)]


def depth_filter_options(**kw):
    """DepthFilterOptions defaults (depth_filter.h:27-60) and updateSeed's call-site flags (depth_filter.cpp:225-226)."""
    o = DepthFilterOptions(200.0, 500.0, 0.0, 1, 0, 1, 0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class ReprojMap(C.Structure):
    """svo_reproj_map: flat map tables (keyframe feature columns + landmark bookkeeping)."""
    _fields_ = [("n_kfs", C.c_int), ("n_feat", C.c_int), ("n_points", C.c_int), ("n_obs", C.c_int),
                ("kf_T_f_w", C.c_void_p), ("kf_seed_mu_range", C.c_void_p), ("kf_frame_idx", C.c_void_p), ("feat", C.c_void_p),
                ("feat_score", C.c_void_p), ("feat_seed_state", C.c_void_p), ("feat_point", C.c_void_p), ("feat_kf", C.c_void_p),
                ("pt_pos", C.c_void_p), ("pt_n_failed", C.c_void_p), ("pt_n_succeeded", C.c_void_p), ("pt_obs_begin", C.c_void_p),
                ("obs_feat", C.c_void_p)]


class ReprojectorOptions(C.Structure):
    _fields_ = [("cell_size", C.c_int), ("max_n_features", C.c_int), ("affine_est_offset", C.c_int), ("affine_est_gain", C.c_int),
                ("sort_by_num_obs", C.c_int), ("_pad", C.c_int), ("seed_sigma2_thresh", C.c_double), ("px_error_angle", C.c_double)]


def reprojector_options(**kw):
    """ReprojectorOptions defaults (reprojector.h:27-70): cell 30, 120 features, affine offset on / gain off, sigma2 thresh 200."""
    o = ReprojectorOptions(30, 120, 1, 0, 0, 0, 200.0, 0.0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class PoseOptimizerOptions(C.Structure):
    _fields_ = [("err_type", C.c_int), ("max_iter", C.c_int), ("eps", C.c_double), ("reproj_thresh_px", C.c_double),
                ("prior_lambda", C.c_double)]


def pose_optimizer_options(**kw):
    """PoseOptimizer::getDefaultSolverOptions (pose_optimizer.cpp:22-29), kUnitPlane, poseoptim_thresh 2.0 px."""
    o = PoseOptimizerOptions(0, 10, 0.000001, 2.0, 0.0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


POSE_OPT_RESULT_DTYPE = np.dtype([("T_imu_world", "<f8", 7), ("T_f_w", "<f8", (MAX_CAMS, 7)), ("measurement_sigma", "<f8"),
                                  ("reproj_error_before", "<f8"), ("reproj_error_after", "<f8"), ("chi2", "<f8"),
                                  ("n_meas_final", "<i4"), ("n_meas", "<i4"), ("iters", "<i4"), ("stop", "<i4")])
REPROJ_RESULT_DTYPE = np.dtype([("cur_px", "<f8", 2), ("px", "<f8", 2), ("f", "<f8", 3), ("grad", "<f8", 2), ("seed_state", "<f8", 4),
                                ("status", "<i4"), ("order", "<i4"), ("slot", "<i4"), ("level", "<i4"), ("type_out", "<i4"),
                                ("match_result", "<i4"), ("d_failed", "<i4"), ("d_succeeded", "<i4")])
REPROJ_STATS_DTYPE = np.dtype([("n_candidates", "<i4"), ("n_trials", "<i4"), ("n_matches", "<i4"), ("n_consumed", "<i4")])
STEREO_RESULT_DTYPE = np.dtype([("px_cur", "<f8", 2), ("f_cur", "<f8", 3), ("grad_cur", "<f8", 2), ("xyz_world", "<f8", 3), ("depth", "<f8"),
                                ("status", "<i4"), ("slot", "<i4"), ("match_result", "<i4"), ("level", "<i4"), ("type", "<i4"), ("_pad", "<i4")])
STEREO_STATS_DTYPE = np.dtype([("n_succeeded", "<i4"), ("n_failed", "<i4")])
STEREO_NOT_REACHED, STEREO_FAILED, STEREO_SUCCESS = range(3)
REPROJ_NOT_CANDIDATE, REPROJ_NOT_REACHED, REPROJ_SKIPPED, REPROJ_FAILED, REPROJ_MATCHED = range(5)

_lib = None


def lib():
    """Load libsvo_cuda.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        path = os.environ.get("SVO_CUDA_LIB", LIB_PATH)  # the override is for A/B experiments with alternative builds
        if not os.path.exists(path):
            raise SvoCudaError(f"{path} is missing: build it with `make -C svo_pro_universal_b200/csrc` (there is no CPU fallback)")
        L = C.CDLL(path)
        L.svo_cuda_last_error.restype = C.c_char_p
        L.svo_cuda_last_error.argtypes = [C.c_void_p]
        L.svo_cuda_launch_count.restype = C.c_longlong
        L.svo_cuda_launch_count.argtypes = [C.c_void_p]
        vp, ci, cd, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
        L.svo_cuda_sizeof.argtypes = [C.c_char_p]
        L.svo_cuda_ctx_create.argtypes = [ci, C.POINTER(vp)]
        L.svo_cuda_ctx_destroy.argtypes = [vp]
        L.svo_cuda_ctx_set_stream.argtypes = [vp, vp]
        L.svo_cuda_ctx_synchronize.argtypes = [vp]
        if hasattr(L, "svo_cuda_host_alloc"):
            L.svo_cuda_host_alloc.argtypes = [vp, sz, ci, C.POINTER(vp)]
            L.svo_cuda_host_free.argtypes = [vp, vp]
        L.svo_cuda_grid_cells.argtypes = [ci, ci, ci, C.POINTER(ci), C.POINTER(ci)]
        L.svo_cuda_pyr_create.argtypes = [vp, ci, ci, ci, ci, ci, C.POINTER(vp)]
        L.svo_cuda_pyr_destroy.argtypes = [vp, vp]
        L.svo_cuda_pyr_upload.argtypes = [vp, vp, ci, ci, vp, sz, sz, ci]
        L.svo_cuda_pyr_build.argtypes = [vp, vp, ci, ci]
        L.svo_cuda_pyr_download.argtypes = [vp, vp, ci, ci, vp, sz, ci]
        L.svo_cuda_pyr_level_info.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci), C.POINTER(sz), C.POINTER(sz), C.POINTER(vp)]
        L.svo_cuda_fast_detect.argtypes = [vp, vp, ci, ci, C.POINTER(DetectorOptions), vp, vp, ci]
        L.svo_cuda_pyramid_fast_detect.argtypes = [vp, vp, ci, ci, C.POINTER(DetectorOptions), vp, vp, ci]
        L.svo_cuda_fast_level_maps.argtypes = [vp, vp, ci, ci, ci, ci, vp, vp, ci]
        L.svo_cuda_sparse_align.argtypes = [vp, ci, C.POINTER(vp), C.POINTER(vp), vp, vp, C.POINTER(Camera), vp, ci, vp, vp,
                                            vp, ci, vp, vp, vp, vp, C.POINTER(SparseAlignOptions), vp, vp, ci]
        L.svo_cuda_align2d.argtypes = [vp, vp, vp, vp, ci, vp, ci, ci, ci, vp, vp, ci]
        L.svo_cuda_align1d.argtypes = [vp, vp, vp, vp, ci, vp, vp, ci, ci, ci, vp, vp, vp, ci]
        L.svo_cuda_align_pyr2d.argtypes = [vp, vp, vp, vp, vp, ci, vp, vp, ci, ci, vp, ci, C.c_float, vp, ci]
        L.svo_cuda_warp_affine.argtypes = [vp, vp, vp, C.POINTER(Camera), C.POINTER(Camera), vp, vp, ci, vp, vp, vp, vp, vp, vp, ci]
        L.svo_cuda_find_match_direct.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Camera), C.POINTER(Camera), vp, vp, ci, vp, vp,
                                                 vp, C.POINTER(MatcherOptions), vp, ci]
        L.svo_cuda_find_epipolar_match_direct.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Camera), C.POINTER(Camera), vp, vp, ci,
                                                          vp, vp, C.POINTER(MatcherOptions), vp, ci]
        L.svo_cuda_update_filter_vogiatzis.argtypes = [vp, ci, vp, vp, vp, vp, vp, ci]
        if hasattr(L, "svo_cuda_update_filter_seq"):  # absent from older A/B builds selected through SVO_CUDA_LIB
            L.svo_cuda_update_filter_seq.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, ci, ci]
        L.svo_cuda_edgelet_detect.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp, ci]
        L.svo_cuda_fastgrad_detect.argtypes = [vp, vp, ci, ci, C.POINTER(DetectorOptions), ci, ci, vp, vp, vp, ci]
        L.svo_cuda_angle_histogram_bins.argtypes = [vp, vp, ci]
        L.svo_cuda_stereo_triangulate.argtypes = [vp] * 5 + [C.POINTER(Camera), C.POINTER(Camera), vp, vp, ci, vp, ci, vp, vp, vp, cd, cd, cd,
                                                             C.POINTER(MatcherOptions), vp, vp, ci]
        L.svo_cuda_optimize_points.argtypes = [vp, ci, vp, vp, ci, vp, vp, ci, vp, ci, ci, vp, ci]
        L.svo_cuda_compute_tau.argtypes = [vp, ci, vp, vp, vp, cd, vp, ci]
        L.svo_cuda_update_seeds.argtypes = [vp, vp, vp, C.POINTER(Camera), C.POINTER(Camera), ci, vp, vp, vp, vp, vp, ci, vp, vp,
                                            vp, C.POINTER(MatcherOptions), C.POINTER(DepthFilterOptions), vp, vp, ci]
        L.svo_cuda_reproject_match.argtypes = [vp, vp, vp, C.POINTER(Camera), C.POINTER(Camera), C.POINTER(ReprojMap), ci, vp, vp, vp,
                                               vp, ci, vp, vp, C.POINTER(ReprojectorOptions), vp, vp, ci]
        L.svo_cuda_pose_optimize.argtypes = [vp, ci, C.POINTER(Camera), vp, ci, vp, vp, ci, vp, vp, vp, vp, vp,
                                             C.POINTER(PoseOptimizerOptions), vp, vp, ci]
        if hasattr(L, "svo_cuda_fast_corner_list"):
            L.svo_cuda_fast_corner_list.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp, vp, C.POINTER(ci), ci]
            L.svo_cuda_fast_corner_score.argtypes = [vp, vp, ci, ci, ci, vp, ci, ci, vp, ci]
            L.svo_cuda_fast_nonmax_3x3.argtypes = [vp, ci, vp, vp, vp, C.POINTER(ci), ci]
            L.svo_cuda_scan_epipolar_line.argtypes = [vp, vp, vp, C.POINTER(Camera), ci, vp, vp, vp, vp, vp, vp, C.POINTER(MatcherOptions),
                                                      vp, vp, ci]
        _lib = L
    return _lib


EXPORTED_SYMBOLS = [
    "svo_cuda_ctx_create", "svo_cuda_ctx_destroy", "svo_cuda_ctx_set_stream", "svo_cuda_ctx_synchronize", "svo_cuda_last_error",
    "svo_cuda_launch_count", "svo_cuda_device_count", "svo_cuda_sizeof", "svo_cuda_pyr_create", "svo_cuda_pyr_destroy", "svo_cuda_pyr_upload",
    "svo_cuda_pyr_build", "svo_cuda_pyr_download", "svo_cuda_pyr_level_info", "svo_cuda_grid_cells", "svo_cuda_fast_detect",
    "svo_cuda_pyramid_fast_detect", "svo_cuda_fast_level_maps", "svo_cuda_sparse_align", "svo_cuda_align2d", "svo_cuda_align1d",
    "svo_cuda_warp_affine", "svo_cuda_find_match_direct", "svo_cuda_find_epipolar_match_direct",
    "svo_cuda_update_filter_vogiatzis", "svo_cuda_compute_tau", "svo_cuda_update_seeds", "svo_cuda_align_pyr2d",
    "svo_cuda_reproject_match", "svo_cuda_pose_optimize", "svo_cuda_edgelet_detect", "svo_cuda_fastgrad_detect",
    "svo_cuda_angle_histogram_bins", "svo_cuda_stereo_triangulate", "svo_cuda_optimize_points", "svo_cuda_update_filter_seq",
    "svo_cuda_host_alloc", "svo_cuda_host_free", "svo_cuda_fast_corner_list", "svo_cuda_fast_corner_score", "svo_cuda_fast_nonmax_3x3",
    "svo_cuda_scan_epipolar_line",
]


def _is_torch(a):
    return hasattr(a, "data_ptr")


_TORCH_NAMES = {"float64": "f8", "float32": "f4", "int32": "i4", "int64": "i8", "uint8": "u1", "int8": "i1", "int16": "i2"}


def _check(a, code, n_items=None, what="array"):
    """The C ABI takes raw pointers: verify the element type (numpy dtype / torch dtype; "u1" also accepts a byte view of a
    struct array) and, when given, the minimum number of items, so that a float32 / int64 / short array is an error, not garbage."""
    if a is None:
        return a
    if _is_torch(a):
        got = _TORCH_NAMES.get(str(a.dtype).replace("torch.", ""), str(a.dtype))
        n = a.numel()
    else:
        assert isinstance(a, np.ndarray), f"{what}: expected a numpy array or torch tensor"
        got = a.dtype.str.lstrip("<|=") if a.dtype.fields is None else "struct"
        n = a.size
    if code == "struct":  # POD records: a numpy structured array or a uint8 tensor holding the same bytes
        assert got in ("struct", "u1"), f"{what}: expected a record array or its byte view, got {got}"
    else:
        assert got == code, f"{what}: expected element type {code}, got {got}"
        if n_items is not None:
            assert n >= n_items, f"{what}: expected at least {n_items} items, got {n}"
    return a


def _ptr(a):
    """(void*, mem kind) of a numpy array, torch tensor or None."""
    if a is None:
        return None, None
    if _is_torch(a):
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr()), (MEM_DEVICE if a.is_cuda else MEM_HOST)
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "arrays must be C-contiguous numpy arrays or torch tensors"
    return C.c_void_p(a.ctypes.data), MEM_HOST


def _ptrs(*arrs):
    ps, kinds = [], set()
    for a in arrs:
        p, k = _ptr(a)
        ps.append(p)
        if k is not None:
            kinds.add(k)
    assert len(kinds) <= 1, "one call must not mix host and device arrays"
    return ps, (kinds.pop() if kinds else MEM_HOST)


class Context:
    """svo_cuda_ctx: one GPU + one stream."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().svo_cuda_ctx_create(device, C.byref(self._h))
        if rc != 0:
            raise SvoCudaError(f"svo_cuda_ctx_create(device={device}) failed with status {rc} "
                               "(no CUDA device? this library has no CPU fallback)")
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise SvoCudaError(f"status {rc}: {lib().svo_cuda_last_error(self._h).decode()}")

    def set_stream(self, cuda_stream_ptr):
        self.check(lib().svo_cuda_ctx_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self.check(lib().svo_cuda_ctx_synchronize(self._h))

    @property
    def launches(self):
        return int(lib().svo_cuda_launch_count(self._h))

    def close(self):
        if self._h:
            lib().svo_cuda_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostBuffer:
    """Page-locked host memory from svo_cuda_host_alloc, exposed as a numpy array (`.array`): the staging buffer of SVO_MEM_HOST calls.
    write_combined=True: write-only for the CPU (uncached reads), read by the copy engine without cache snooping."""

    def __init__(self, ctx, shape, dtype=np.uint8, write_combined=False):
        self.ctx = ctx
        dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        self._p = C.c_void_p()
        ctx.check(lib().svo_cuda_host_alloc(ctx._h, n, int(bool(write_combined)), C.byref(self._p)))
        buf = (C.c_uint8 * max(n, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.uint8, count=n).view(dtype).reshape(shape)
        self.nbytes = n

    def close(self):
        if self._p:
            self.array = None
            lib().svo_cuda_host_free(self.ctx._h, self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pyramid:
    """svo_cuda_pyr: a device-resident batch of image pyramids (svo::Frame::img_pyr_ for n_frames frames)."""

    def __init__(self, ctx, n_frames, width, height, n_levels, halfsample_mode=-1):
        self.ctx = ctx
        self._h = C.c_void_p()
        ctx.check(lib().svo_cuda_pyr_create(ctx._h, n_frames, width, height, n_levels, halfsample_mode, C.byref(self._h)))
        self.n_frames, self.width, self.height, self.n_levels = n_frames, width, height, n_levels

    def upload(self, images, first=0, sync=None):
        """images: [count, H, W] uint8 numpy (host) or torch tensor (pinned host or cuda). sync: wait for the copy (default: only for
        numpy input, whose pageable buffer must not be freed under the copy; pass False for page-locked HostBuffer arrays)."""
        if not _is_torch(images):
            images = np.ascontiguousarray(images, dtype=np.uint8)
        if images.ndim == 2:
            images = images.reshape((1,) + tuple(images.shape))
        count, h, w = images.shape
        assert (h, w) == (self.height, self.width)
        p, kind = _ptr(images)
        self.ctx.check(lib().svo_cuda_pyr_upload(self.ctx._h, self._h, first, count, p, w, w * h, kind))
        if (kind == MEM_HOST and not _is_torch(images)) if sync is None else sync:
            self.ctx.synchronize()  # pageable numpy buffer: do not let it be freed under the copy

    def build(self, first=0, count=None):
        self.ctx.check(lib().svo_cuda_pyr_build(self.ctx._h, self._h, first, self.n_frames - first if count is None else count))

    def level_info(self, level):
        c, r, p, s, d = C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t(), C.c_void_p()
        self.ctx.check(lib().svo_cuda_pyr_level_info(self._h, level, C.byref(c), C.byref(r), C.byref(p), C.byref(s), C.byref(d)))
        return dict(cols=c.value, rows=r.value, pitch=p.value, frame_stride=s.value, device_ptr=d.value)

    def download(self, frame, level):
        info = self.level_info(level)
        out = np.empty((info["rows"], info["cols"]), np.uint8)
        self.ctx.check(lib().svo_cuda_pyr_download(self.ctx._h, self._h, frame, level, C.c_void_p(out.ctypes.data), info["cols"], MEM_HOST))
        return out

    def close(self):
        if self._h:
            lib().svo_cuda_pyr_destroy(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def grid_cells(width, height, cell_size):
    nc, nr = C.c_int(), C.c_int()
    n = lib().svo_cuda_grid_cells(width, height, cell_size, C.byref(nc), C.byref(nr))
    return n, nc.value, nr.value


def fast_detect(ctx, pyr, opt, first=0, count=None, occupancy=None, corners_out=None, fused_pyramid=False):
    """fastDetector for frames [first, first+count): returns corners [count, n_cells] (CORNER_DTYPE)."""
    count = pyr.n_frames - first if count is None else count
    n_cells, _, _ = grid_cells(pyr.width, pyr.height, opt.cell_size)
    if corners_out is None:
        corners_out = np.zeros((count, n_cells), CORNER_DTYPE)
    (po, pc), kind = _ptrs(occupancy, corners_out)
    fn = lib().svo_cuda_pyramid_fast_detect if fused_pyramid else lib().svo_cuda_fast_detect
    ctx.check(fn(ctx._h, pyr._h, first, count, C.byref(opt), po, pc, kind))
    return corners_out


def edgelet_detect(ctx, pyr, threshold=100, border=8, cell_size=30, first=0, count=None, occupancy=None, corners_out=None):
    """edgeletDetector_V2 for frames [first, first+count): per-cell edgelets [count, n_cells] (CORNER_DTYPE)."""
    count = pyr.n_frames - first if count is None else count
    n_cells, _, _ = grid_cells(pyr.width, pyr.height, cell_size)
    if corners_out is None:
        corners_out = np.zeros((count, n_cells), CORNER_DTYPE)
    (po, pc), kind = _ptrs(occupancy, corners_out)
    ctx.check(lib().svo_cuda_edgelet_detect(ctx._h, pyr._h, first, count, int(threshold), int(border), int(cell_size), po, pc, kind))
    return corners_out


def fastgrad_detect(ctx, pyr, opt, threshold_secondary=100, max_n_features=None, first=0, count=None, occupancy=None,
                    corners_out=None, edgelets_out=None):
    """FastGradDetector::detect up to fillFeatures' sort: (FAST corners, edgelets), both [count, n_cells] (CORNER_DTYPE)."""
    count = pyr.n_frames - first if count is None else count
    n_cells, _, _ = grid_cells(pyr.width, pyr.height, opt.cell_size)
    max_n_features = n_cells if max_n_features is None else int(max_n_features)
    if corners_out is None:
        corners_out = np.zeros((count, n_cells), CORNER_DTYPE)
    if edgelets_out is None:
        edgelets_out = np.zeros((count, n_cells), CORNER_DTYPE) if not _is_torch(corners_out) else corners_out.new_zeros(corners_out.shape)
    (po, pc, pe), kind = _ptrs(occupancy, corners_out, edgelets_out)
    ctx.check(lib().svo_cuda_fastgrad_detect(ctx._h, pyr._h, first, count, C.byref(opt), int(threshold_secondary), max_n_features,
                                             po, pc, pe, kind))
    return corners_out, edgelets_out


def angle_histogram_bins(ctx):
    """Bins of angle_hist::angleHistogram for every gradient (gx, gy) in [-255, 255]^2: int8 [511, 511], row gy + 255."""
    out = np.zeros((511, 511), np.int8)
    ctx.check(lib().svo_cuda_angle_histogram_bins(ctx._h, C.c_void_p(out.ctypes.data), MEM_HOST))
    return out


def fast_level_maps(ctx, pyr, frame, level, threshold=10, arc_length=10):
    info = pyr.level_info(level)
    score = np.zeros((info["rows"], info["cols"]), np.int16)
    nonmax = np.zeros((info["rows"], info["cols"]), np.uint8)
    ctx.check(lib().svo_cuda_fast_level_maps(ctx._h, pyr._h, frame, level, threshold, arc_length,
                                             C.c_void_p(score.ctypes.data), C.c_void_p(nonmax.ctypes.data), MEM_HOST))
    return score, nonmax


def fast_corner_list(ctx, pyr, frame, level, threshold=10, arc_length=10, max_corners=None):
    """fast_corner_detect_10/9 + fast_corner_score_10 + fast_nonmax_3x3 of one level: (xy FAST_XY_DTYPE [n], scores int32 [n],
    nonmax uint8 [n]) in raster order (host arrays)."""
    info = pyr.level_info(level)
    cap = int(max_corners) if max_corners is not None else max(1024, info["rows"] * info["cols"] // 16)
    while True:
        xy = np.zeros(cap, FAST_XY_DTYPE)
        sc = np.zeros(cap, np.int32)
        nm = np.zeros(cap, np.uint8)
        n = C.c_int(0)
        ctx.check(lib().svo_cuda_fast_corner_list(ctx._h, pyr._h, int(frame), int(level), int(threshold), int(arc_length), cap,
                                                  C.c_void_p(xy.ctypes.data), C.c_void_p(sc.ctypes.data), C.c_void_p(nm.ctypes.data),
                                                  C.byref(n), MEM_HOST))
        if n.value <= cap or max_corners is not None:
            m = min(n.value, cap)
            return xy[:m], sc[:m], nm[:m]
        cap = n.value


def fast_corner_score(ctx, pyr, frame, level, xy, threshold=10, arc_length=10):
    """fast_corner_score_10 of a caller-supplied corner list (FAST_XY_DTYPE): int32 [n]."""
    xy = np.ascontiguousarray(xy, FAST_XY_DTYPE)
    out = np.zeros(len(xy), np.int32)
    ctx.check(lib().svo_cuda_fast_corner_score(ctx._h, pyr._h, int(frame), int(level), len(xy), C.c_void_p(xy.ctypes.data), int(threshold),
                                               int(arc_length), C.c_void_p(out.ctypes.data), MEM_HOST))
    return out


def fast_nonmax_3x3(ctx, xy, scores):
    """fast_nonmax_3x3 on a raster-ordered list: indices (int32, ascending) of the surviving corners."""
    xy = np.ascontiguousarray(xy, FAST_XY_DTYPE)
    scores = np.ascontiguousarray(scores, np.int32)
    assert len(xy) == len(scores)
    idx = np.zeros(max(len(xy), 1), np.int32)
    n = C.c_int(0)
    ctx.check(lib().svo_cuda_fast_nonmax_3x3(ctx._h, len(xy), C.c_void_p(xy.ctypes.data), C.c_void_p(scores.ctypes.data),
                                             C.c_void_p(idx.ctypes.data), C.byref(n), MEM_HOST))
    return idx[:n.value]


def scan_epipolar_line(ctx, cur_pyr, cam_cur, A, B, Cpt, patch, patch_level, epi_length_pyramid, opt, zmssd_best=None, cur_frame_idx=None):
    """Matcher::scanEpipolarLine for M scans (host arrays): returns (image_best [M, 2], zmssd_best [M])."""
    A = np.ascontiguousarray(A, np.float64).reshape(-1, 3)
    M = len(A)
    B = np.ascontiguousarray(B, np.float64).reshape(M, 3)
    Cpt = np.ascontiguousarray(Cpt, np.float64).reshape(M, 3)
    patch = np.ascontiguousarray(patch, np.uint8).reshape(M, 64)
    patch_level = np.ascontiguousarray(patch_level, np.int32).reshape(M)
    epi = np.ascontiguousarray(epi_length_pyramid, np.float64).reshape(M)
    z = np.full(M, 2000 * 64, np.int32) if zmssd_best is None else np.ascontiguousarray(zmssd_best, np.int32).copy()
    best = np.zeros((M, 2))
    cf = None if cur_frame_idx is None else np.ascontiguousarray(cur_frame_idx, np.int32)
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    ctx.check(lib().svo_cuda_scan_epipolar_line(ctx._h, cur_pyr._h, vp(cf), C.byref(cam_cur), M, vp(A), vp(B), vp(Cpt), vp(patch),
                                                vp(patch_level), vp(epi), C.byref(opt), vp(best), vp(z), MEM_HOST))
    return best, z


def sparse_align(ctx, ref_pyrs, cur_pyrs, cams, T_cam_imu, T_imu_world_ref, T_imu_world_cur, n_features, px, f, depth, eligible,
                 opt, priors=None, ref_frame_idx=None, cur_frame_idx=None, results=None):
    """svo_cuda_sparse_align. ref_pyrs/cur_pyrs: lists (one Pyramid per camera). Feature arrays are
    [B, n_cams, max_features, ...]. Returns results (ALIGN_RESULT_DTYPE numpy, or the torch uint8 tensor passed in)."""
    n_cams = len(ref_pyrs)
    B = T_imu_world_ref.shape[0]
    max_features = px.shape[-2]
    cam_arr = (Camera * n_cams)(*cams)
    rp = (C.c_void_p * n_cams)(*[p._h for p in ref_pyrs])
    cp = (C.c_void_p * n_cams)(*[p._h for p in cur_pyrs])
    T_cam_imu = np.ascontiguousarray(T_cam_imu, np.float64).reshape(n_cams, 7)
    if results is None:
        results = np.zeros(B, ALIGN_RESULT_DTYPE)
    nf = B * n_cams * max_features
    _check(ref_frame_idx, "i4", B * n_cams, "ref_frame_idx"); _check(cur_frame_idx, "i4", B * n_cams, "cur_frame_idx")
    _check(T_imu_world_ref, "f8", 7 * B, "T_imu_world_ref"); _check(T_imu_world_cur, "f8", 7 * B, "T_imu_world_cur")
    _check(n_features, "i4", B * n_cams, "n_features"); _check(px, "f8", 2 * nf, "px"); _check(f, "f8", 3 * nf, "f")
    _check(depth, "f8", nf, "depth"); _check(eligible, "u1", nf, "eligible"); _check(priors, "struct", None, "priors")
    _check(results, "struct", None, "results")
    ps, kind = _ptrs(ref_frame_idx, cur_frame_idx, T_imu_world_ref, T_imu_world_cur, n_features, px, f, depth, eligible, priors, results)
    ctx.check(lib().svo_cuda_sparse_align(ctx._h, n_cams, rp, cp, ps[0], ps[1], cam_arr, C.c_void_p(T_cam_imu.ctypes.data), B,
                                          ps[2], ps[3], ps[4], max_features, ps[5], ps[6], ps[7], ps[8], C.byref(opt), ps[9],
                                          ps[10], kind))
    return results


def align2d(ctx, pyr, frame_idx, level, patch_with_border, px, n_iter=10, est_offset=True, est_gain=False):
    M = len(px)
    px = np.ascontiguousarray(px, np.float64).copy()
    conv = np.zeros(M, np.uint8)
    ps, kind = _ptrs(np.ascontiguousarray(frame_idx, np.int32), np.ascontiguousarray(level, np.int32),
                     np.ascontiguousarray(patch_with_border, np.uint8), px, conv)
    ctx.check(lib().svo_cuda_align2d(ctx._h, pyr._h, ps[0], ps[1], M, ps[2], n_iter, int(est_offset), int(est_gain), ps[3], ps[4], kind))
    return px, conv


def align1d(ctx, pyr, frame_idx, level, direction, patch_with_border, px, n_iter=10, est_offset=True, est_gain=False):
    M = len(px)
    px = np.ascontiguousarray(px, np.float64).copy()
    conv = np.zeros(M, np.uint8)
    hinv = np.zeros(M, np.float64)
    ps, kind = _ptrs(np.ascontiguousarray(frame_idx, np.int32), np.ascontiguousarray(level, np.int32),
                     np.ascontiguousarray(direction, np.float64), np.ascontiguousarray(patch_with_border, np.uint8), px, hinv, conv)
    ctx.check(lib().svo_cuda_align1d(ctx._h, pyr._h, ps[0], ps[1], M, ps[2], ps[3], n_iter, int(est_offset), int(est_gain),
                                     ps[4], ps[5], ps[6], kind))
    return px, conv, hinv


def align_pyr2d(ctx, ref_pyr, cur_pyr, px_ref_level_0, px_cur, max_level, min_level, patch_sizes, n_iter=30, min_update_squared=0.03 ** 2,
                ref_frame_idx=None, cur_frame_idx=None):
    """svo_cuda_align_pyr2d (feature_alignment::alignPyr2DVec). Returns (px_cur [M][2] float64, status [M] uint8)."""
    dev = _is_torch(px_cur)
    M = (px_cur.numel() // 2) if dev else len(px_cur)
    if dev:
        import torch
        pc = px_cur.clone()
        status = torch.zeros(M, dtype=torch.uint8, device=px_cur.device)
        pr = px_ref_level_0
    else:
        pc = np.ascontiguousarray(px_cur, np.float64).reshape(-1, 2).copy()
        status = np.zeros(M, np.uint8)
        pr = np.ascontiguousarray(px_ref_level_0, np.int32).reshape(-1, 2)
        ref_frame_idx = None if ref_frame_idx is None else np.ascontiguousarray(ref_frame_idx, np.int32)
        cur_frame_idx = None if cur_frame_idx is None else np.ascontiguousarray(cur_frame_idx, np.int32)
    ps_arr = np.zeros(8, np.int32)
    ps_arr[:len(patch_sizes)] = patch_sizes
    ptrs, kind = _ptrs(pr, pc, status)
    rfi = _ptr(ref_frame_idx)[0] if ref_frame_idx is not None else None
    cfi = _ptr(cur_frame_idx)[0] if cur_frame_idx is not None else None
    ctx.check(lib().svo_cuda_align_pyr2d(ctx._h, ref_pyr._h, cur_pyr._h, rfi, cfi, M, ptrs[0], ptrs[1], max_level, min_level,
                                         ps_arr.ctypes.data_as(C.c_void_p), n_iter, C.c_float(min_update_squared), ptrs[2], kind))
    return pc, status


def _n_features(ftrs):
    """Number of svo_feature records in a FEATURE_DTYPE numpy array or a torch byte tensor holding the same bytes."""
    if _is_torch(ftrs):
        return ftrs.numel() * ftrs.element_size() // FEATURE_DTYPE.itemsize
    return len(ftrs)


def make_features(px, f, grad, ftype, level):
    n = len(px)
    a = np.zeros(n, FEATURE_DTYPE)
    a["px"], a["f"], a["grad"], a["type"], a["level"] = px, f, grad, ftype, level
    return a


def warp_affine(ctx, ref_pyr, cam_ref, cam_cur, T_cur_ref, ftrs, depth, ref_frame_idx=None, T_idx=None):
    M = _n_features(ftrs)
    A = np.zeros((M, 4))
    sl = np.zeros(M, np.int32)
    pwb = np.zeros((M, 100), np.uint8)
    ok = np.zeros(M, np.uint8)
    T = np.ascontiguousarray(T_cur_ref, np.float64).reshape(-1, 7)
    ps, kind = _ptrs(ref_frame_idx, T, T_idx, ftrs, np.ascontiguousarray(depth, np.float64), A, sl, pwb, ok)
    ctx.check(lib().svo_cuda_warp_affine(ctx._h, ref_pyr._h, ps[0], C.byref(cam_ref), C.byref(cam_cur), ps[1], ps[2], M, ps[3], ps[4],
                                         ps[5], ps[6], ps[7], ps[8], kind))
    return A, sl, pwb, ok


def find_match_direct(ctx, ref_pyr, cur_pyr, cam_ref, cam_cur, T_cur_ref, ftrs, ref_depth, px_guess, opt, ref_frame_idx=None,
                      cur_frame_idx=None, T_idx=None, out=None):
    M = _n_features(ftrs)
    if out is None:
        out = np.zeros(M, MATCH_OUT_DTYPE)
    if not _is_torch(T_cur_ref):
        T_cur_ref = np.ascontiguousarray(T_cur_ref, np.float64).reshape(-1, 7)
    ps, kind = _ptrs(ref_frame_idx, cur_frame_idx, T_cur_ref, T_idx, ftrs, ref_depth, px_guess, out)
    ctx.check(lib().svo_cuda_find_match_direct(ctx._h, ref_pyr._h, cur_pyr._h, ps[0], ps[1], C.byref(cam_ref), C.byref(cam_cur), ps[2],
                                               ps[3], M, ps[4], ps[5], ps[6], C.byref(opt), ps[7], kind))
    return out


def find_epipolar_match_direct(ctx, ref_pyr, cur_pyr, cam_ref, cam_cur, T_cur_ref, ftrs, d_inv, opt, ref_frame_idx=None,
                               cur_frame_idx=None, T_idx=None, out=None):
    M = _n_features(ftrs)
    if out is None:
        out = np.zeros(M, MATCH_OUT_DTYPE)
    if not _is_torch(T_cur_ref):
        T_cur_ref = np.ascontiguousarray(T_cur_ref, np.float64).reshape(-1, 7)
    ps, kind = _ptrs(ref_frame_idx, cur_frame_idx, T_cur_ref, T_idx, ftrs, d_inv, out)
    ctx.check(lib().svo_cuda_find_epipolar_match_direct(ctx._h, ref_pyr._h, cur_pyr._h, ps[0], ps[1], C.byref(cam_ref), C.byref(cam_cur),
                                                        ps[2], ps[3], M, ps[4], ps[5], C.byref(opt), ps[6], kind))
    return out


def update_filter_vogiatzis(ctx, z, tau2, mu_range, state, ok=None):
    """In-place on `state` [n,4]. Returns ok [n] uint8."""
    n = state.shape[0]
    if ok is None and not _is_torch(state):
        ok = np.zeros(n, np.uint8)
    ps, kind = _ptrs(z, tau2, mu_range, state, ok)
    ctx.check(lib().svo_cuda_update_filter_vogiatzis(ctx._h, n, ps[0], ps[1], ps[2], ps[3], ps[4], kind))
    return ok


def update_filter_seq(ctx, z, tau2, mu_range, state, ok=None, gaussian=False):
    """svo_cuda_update_filter_seq: n_obs ordered updates per seed in one launch; z, tau2 [n_obs, n]; `state` [n, 4] in place.
    Returns ok [n_obs, n] uint8 (numpy) or the tensor passed in (None on the device path when not requested)."""
    n = state.shape[0]
    n_obs = z.shape[0] if z.ndim == 2 else 1
    if ok is None and not _is_torch(state):
        ok = np.zeros((n_obs, n), np.uint8)
    _check(z, "f8", n * n_obs, "z"); _check(tau2, "f8", n * n_obs, "tau2"); _check(mu_range, "f8", n, "mu_range")
    _check(state, "f8", 4 * n, "state"); _check(ok, "u1", n * n_obs, "ok")
    ps, kind = _ptrs(z, tau2, mu_range, state, ok)
    ctx.check(lib().svo_cuda_update_filter_seq(ctx._h, n, n_obs, ps[0], ps[1], ps[2], ps[3], ps[4], int(bool(gaussian)), kind))
    return ok


def compute_tau(ctx, T_ref_cur, f, z, px_error_angle):
    n = len(z)
    tau = np.zeros(n)
    ps, kind = _ptrs(np.ascontiguousarray(T_ref_cur, np.float64), np.ascontiguousarray(f, np.float64),
                     np.ascontiguousarray(z, np.float64), tau)
    ctx.check(lib().svo_cuda_compute_tau(ctx._h, n, ps[0], ps[1], ps[2], px_error_angle, ps[3], kind))
    return tau


def update_seeds(ctx, ref_pyr, cur_pyr, cam_ref, cam_cur, ftrs, types, state, seed_mu_range, obs_frame_idx, obs_T_idx, T_cur_ref,
                 mopt, dopt, ref_frame_idx=None, want_match_results=True):
    """svo_cuda_update_seeds; types/state updated in place. Returns (n_success, match_results [n_obs,S] or None)."""
    S = _n_features(ftrs)
    n_obs = obs_frame_idx.shape[0]
    dev = _is_torch(state)
    if dev:
        import torch
        n_success = torch.zeros(1, dtype=torch.int32, device=state.device)
        mr = torch.full((n_obs, S), -1, dtype=torch.int32, device=state.device) if want_match_results else None
    else:
        n_success = np.zeros(1, np.int32)
        mr = np.full((n_obs, S), -1, np.int32) if want_match_results else None
    _check(ref_frame_idx, "i4", S, "ref_frame_idx"); _check(ftrs, "struct", None, "ftrs"); _check(types, "u1", S, "types")
    _check(state, "f8", 4 * S, "state"); _check(seed_mu_range, "f8", S, "seed_mu_range"); _check(obs_frame_idx, "i4", n_obs * S, "obs_frame_idx")
    _check(obs_T_idx, "i4", n_obs * S, "obs_T_idx"); _check(T_cur_ref, "f8", 7, "T_cur_ref")
    ps, kind = _ptrs(ref_frame_idx, ftrs, types, state, seed_mu_range, obs_frame_idx, obs_T_idx, T_cur_ref, n_success, mr)
    ctx.check(lib().svo_cuda_update_seeds(ctx._h, ref_pyr._h, cur_pyr._h, C.byref(cam_ref), C.byref(cam_cur), S, ps[0], ps[1], ps[2],
                                          ps[3], ps[4], n_obs, ps[5], ps[6], ps[7], C.byref(mopt), C.byref(dopt), ps[8], ps[9], kind))
    return n_success, mr


_REPROJ_MAP_ARRAYS = ("kf_T_f_w", "kf_seed_mu_range", "kf_frame_idx", "feat", "feat_score", "feat_seed_state", "feat_point", "feat_kf",
                      "pt_pos", "pt_n_failed", "pt_n_succeeded", "pt_obs_begin", "obs_feat")


def reproject_match(ctx, ref_pyr, cur_pyr, cam_ref, cam_cur, tables, cur_T_f_w, n_features_in, entry_begin, entry_feat, occupancy, opt,
                    cur_frame_idx=None, results=None, stats=None):
    """svo_cuda_reproject_match. `tables`: dict with the svo_reproj_map arrays (numpy or torch cuda; kf_frame_idx optional) plus
    n_kfs / n_feat / n_points / n_obs. occupancy [F, n_cells] uint8 is updated in place. Returns (results, stats)."""
    arrs = [tables.get(k) for k in _REPROJ_MAP_ARRAYS]
    dev = _is_torch(occupancy) and occupancy.is_cuda
    F = int(cur_T_f_w.shape[0])
    n_entries = int(entry_feat.shape[0])
    if results is None:
        if dev:
            import torch
            results = torch.zeros(max(n_entries, 1) * REPROJ_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=occupancy.device)
            stats = torch.zeros(F * REPROJ_STATS_DTYPE.itemsize, dtype=torch.uint8, device=occupancy.device)
        else:
            results = np.zeros(n_entries, REPROJ_RESULT_DTYPE)
            stats = np.zeros(F, REPROJ_STATS_DTYPE)
    ps, kind = _ptrs(*arrs, cur_frame_idx, cur_T_f_w, n_features_in, entry_begin, entry_feat, occupancy, results, stats)
    m = ReprojMap(int(tables["n_kfs"]), int(tables["n_feat"]), int(tables["n_points"]), int(tables["n_obs"]),
                  *[(p.value if p is not None else None) for p in ps[:len(arrs)]])
    q = ps[len(arrs):]
    ctx.check(lib().svo_cuda_reproject_match(ctx._h, ref_pyr._h, cur_pyr._h, C.byref(cam_ref), C.byref(cam_cur), C.byref(m), F, q[0], q[1],
                                             q[2], q[3], n_entries, q[4], q[5], C.byref(opt), q[6], q[7], kind))
    return results, stats


def stereo_triangulate(ctx, pyr0, pyr1, cam0, cam1, T_f1f0, T_world_cam0, feat_begin, ftrs, n_desired, n_features_in_frame1, mopt,
                       mean_depth_inv=1.0 / 3.0, min_depth_inv=1.0, max_depth_inv=1.0 / 50.0, frame0_idx=None, frame1_idx=None,
                       results=None, stats=None):
    """svo_cuda_stereo_triangulate: B stereo pairs, features in visiting order. Arrays numpy (host) or torch cuda (T_f1f0 always a
    host 7-vector). Returns (results [n] STEREO_RESULT_DTYPE, stats [B] STEREO_STATS_DTYPE) or the tensors passed in."""
    T_f1f0 = np.ascontiguousarray(T_f1f0, np.float64).reshape(7)
    B = int(T_world_cam0.shape[0])
    N = _n_features(ftrs)
    if results is None:
        if _is_torch(T_world_cam0) and T_world_cam0.is_cuda:
            import torch
            results = torch.zeros(max(N, 1) * STEREO_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=T_world_cam0.device)
            stats = torch.zeros(B * STEREO_STATS_DTYPE.itemsize, dtype=torch.uint8, device=T_world_cam0.device)
        else:
            results = np.zeros(N, STEREO_RESULT_DTYPE)
            stats = np.zeros(B, STEREO_STATS_DTYPE)
    ps, kind = _ptrs(frame0_idx, frame1_idx, T_world_cam0, feat_begin, ftrs, n_desired, n_features_in_frame1, results, stats)
    pf0, pf1, pT, pb, pf, pn, ps1, pr, pst = ps
    fn = lib().svo_cuda_stereo_triangulate
    ctx.check(fn(ctx._h, pyr0._h, pyr1._h, pf0, pf1, C.byref(cam0), C.byref(cam1), C.c_void_p(T_f1f0.ctypes.data), pT, B, pb, N, pf, pn, ps1,
                 mean_depth_inv, min_depth_inv, max_depth_inv, C.byref(mopt), pr, pst, kind))
    return results, stats


def optimize_points(ctx, pos, obs_begin, obs_frame, obs_f, T_f_w, n_iter=5, using_bearing_vector=False, iters_out=None):
    """svo_cuda_optimize_points: Point::optimize for P points; pos [P, 3] is updated in place (numpy or torch cuda). Returns iters_out."""
    P = int(pos.shape[0])
    n_obs, n_frames = int(obs_frame.shape[0]), int(T_f_w.shape[0])
    if iters_out is None:
        if _is_torch(pos) and pos.is_cuda:
            import torch
            iters_out = torch.zeros(P, dtype=torch.int32, device=pos.device)
        else:
            iters_out = np.zeros(P, np.int32)
    (pp, pb, pfr, pf, pT, pi), kind = _ptrs(pos, obs_begin, obs_frame, obs_f, T_f_w, iters_out)
    fn = lib().svo_cuda_optimize_points
    ctx.check(fn(ctx._h, P, pp, pb, n_obs, pfr, pf, n_frames, pT, int(n_iter), int(bool(using_bearing_vector)), pi, kind))
    return iters_out


def pose_optimize(ctx, cams, T_cam_imu, T_imu_world, feat_begin, ftrs, feat_cam, xyz_world, has_xyz, opt, prior_q=None, results=None,
                  outlier=None):
    """svo_cuda_pose_optimize: B bundles. Arrays numpy (host) or torch cuda. Returns (results, outlier)."""
    n_cams = len(cams)
    cam_arr = (Camera * n_cams)(*cams)
    T_cam_imu = np.ascontiguousarray(T_cam_imu, np.float64).reshape(n_cams, 7)
    B = int(T_imu_world.shape[0])
    N = _n_features(ftrs)
    dev = _is_torch(T_imu_world) and T_imu_world.is_cuda
    if results is None:
        if dev:
            import torch
            results = torch.zeros(B * POSE_OPT_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=T_imu_world.device)
            outlier = torch.zeros(max(N, 1), dtype=torch.uint8, device=T_imu_world.device)
        else:
            results = np.zeros(B, POSE_OPT_RESULT_DTYPE)
            outlier = np.zeros(N, np.uint8)
    _check(T_imu_world, "f8", 7 * B, "T_imu_world"); _check(feat_begin, "i4", B + 1, "feat_begin"); _check(ftrs, "struct", None, "ftrs")
    _check(feat_cam, "i4", N, "feat_cam"); _check(xyz_world, "f8", 3 * N, "xyz_world"); _check(has_xyz, "u1", N, "has_xyz")
    _check(prior_q, "f8", 4 * B, "prior_q")
    ps, kind = _ptrs(T_imu_world, feat_begin, ftrs, feat_cam, xyz_world, has_xyz, prior_q, results, outlier)
    ctx.check(lib().svo_cuda_pose_optimize(ctx._h, n_cams, cam_arr, C.c_void_p(T_cam_imu.ctypes.data), B, ps[0], ps[1], N, ps[2], ps[3],
                                           ps[4], ps[5], ps[6], C.byref(opt), ps[7], ps[8], kind))
    return results, outlier
