"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol include/svo_cuda.h declares, its POD structs match
the numpy/ctypes mirrors, and it fails loudly (no CPU fallback) when no CUDA device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from svo_pro_universal_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "svo_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svo_cuda_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/svo_cuda.h but not exported by libsvo_cuda.so"
    assert sorted(capi.EXPORTED_SYMBOLS) == names, "capi.EXPORTED_SYMBOLS is out of sync with the header"


def test_struct_layouts_match_header():
    """Python mirrors of the POD structs have the size the C compiler gives include/svo_cuda.h."""
    sz = lambda n: capi.lib().svo_cuda_sizeof(n.encode())
    assert capi.ALIGN_RESULT_DTYPE.itemsize == sz("svo_align_result") == 7 * 8 + 4 * 7 * 8 + 3 * 8 + 64 * 8 + 4 + 8 * 4 + 4
    assert capi.ALIGN_PRIOR_DTYPE.itemsize == sz("svo_align_prior")
    assert capi.CORNER_DTYPE.itemsize == sz("svo_corner")
    assert capi.FEATURE_DTYPE.itemsize == sz("svo_feature")
    assert capi.MATCH_OUT_DTYPE.itemsize == sz("svo_match_out")
    assert C.sizeof(capi.Camera) == sz("svo_camera")
    assert C.sizeof(capi.SparseAlignOptions) == sz("svo_sparse_align_options")
    assert C.sizeof(capi.MatcherOptions) == sz("svo_matcher_options")
    assert C.sizeof(capi.DepthFilterOptions) == sz("svo_depth_filter_options")
    assert C.sizeof(capi.DetectorOptions) == sz("svo_detector_options")
    assert capi.STEREO_RESULT_DTYPE.itemsize == sz("svo_stereo_result") and capi.STEREO_STATS_DTYPE.itemsize == sz("svo_stereo_stats")
    assert sz("nope") == -1


def test_grid_cells_matches_reference_grid():
    # AbstractDetector grid: ceil(752/30) x ceil(480/30) = 26 x 16 (feature_detection.cpp:27-37)
    assert capi.grid_cells(752, 480, 30) == (416, 26, 16)
    assert capi.grid_cells(47, 30, 30) == (2, 2, 1)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert capi.lib().svo_cuda_device_count() == 0
    with pytest.raises(capi.SvoCudaError):
        capi.Context(0)
    h = C.c_void_p()
    assert capi.lib().svo_cuda_ctx_create(0, C.byref(h)) == -3  # SVO_ERR_NO_DEVICE
    assert not h.value


def test_option_defaults_mirror_reference():
    o = capi.sparse_align_options()
    assert (o.max_level, o.min_level, o.max_iter) == (4, 1, 10) and o.eps == 0.0005 and o.weight_scale == 10.0
    assert not (o.estimate_illumination_gain or o.estimate_illumination_offset or o.robustification or o.use_distortion_jacobian)
    m = capi.matcher_options()
    assert (m.align_max_iter, m.max_epi_search_steps, m.scan_on_unit_sphere, m.affine_est_offset, m.affine_est_gain) == (10, 100, 1, 1, 0)
    d = capi.detector_options()
    assert (d.threshold, d.border, d.min_level, d.max_level, d.cell_size) == (10, 8, 0, 2, 30)
    f = capi.depth_filter_options()
    assert (f.seed_convergence_sigma2_thresh, f.mappoint_convergence_sigma2_thresh) == (200.0, 500.0)
    assert np.dtype(capi.CORNER_DTYPE).names == ("x", "y", "level", "score", "angle")
