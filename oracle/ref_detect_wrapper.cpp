// ORACLE — TEST INFRASTRUCTURE ONLY.
// C wrapper around the REFERENCE's own feature detectors, compiled from where they lie under /root/reference into
// oracle/_ref/libdetect_ref.so by oracle/Makefile:
//   src/svo_direct/src/feature_detection.cpp        (FastDetector / GradientDetectorGrid / FastGradDetector ::detect)
//   src/svo_direct/src/feature_detection_utils.cpp  (fastDetector, edgeletDetector_V2, fillFeatures, angle histogram)
//   src/fast_neon/src/*.cpp                         (fast_corner_detect_10[_sse2], fast_corner_score_10, fast_nonmax_3x3)
// Eigen / glog / the OpenCV containers resolve to the stand-ins in oracle/shim. The only OpenCV *arithmetic* on this path
// (GaussianBlur 3x3 and Scharr inside edgeletDetector_V2) is restated in shim/shim_cv_imgproc.cpp and pinned against the real
// OpenCV (cv2 4.13) by tests/golden/cv_imgproc_golden.npz. No reference source is copied here.
#include <svo/direct/feature_detection.h>
#include <svo/direct/feature_detection_types.h>
#include <svo/direct/feature_detection_utils.h>
#include <svo/common/camera.h>
#include <svo/common/occupancy_grid_2d.h>
#include <vikit/cameras/camera_geometry.h>
#include <vikit/cameras/pinhole_projection.h>
#include <vikit/cameras/no_distortion.h>
#include <cstring>
#include "orc_capi.h"

namespace {
svo::ImgPyr makePyr(int n_levels, const uint8_t* const* data, const int* cols, const int* rows, const int* step) {
  svo::ImgPyr pyr;
  for (int l = 0; l < n_levels; ++l) pyr.emplace_back(rows[l], cols[l], CV_8UC1, const_cast<uint8_t*>(data[l]), (size_t)step[l]);
  return pyr;
}
}  // namespace

extern "C" {

// feature_detection_utils::fastDetector (a5): corners_out [n_cells] pre-filled with score = threshold like FastDetector::detect.
void ref_fast_detector(int n_levels, const uint8_t* const* data, const int* cols, const int* rows, const int* step, int threshold,
                       int border, int min_level, int max_level, int cell_size, const uint8_t* occupancy, orc_corner* corners_out) {
  svo::ImgPyr pyr = makePyr(n_levels, data, cols, rows, step);
  svo::OccupandyGrid2D grid(cell_size, svo::OccupandyGrid2D::getNCell(cols[0], cell_size), svo::OccupandyGrid2D::getNCell(rows[0], cell_size));
  for (size_t k = 0; k < grid.occupancy_.size(); ++k) grid.occupancy_[k] = occupancy && occupancy[k];
  svo::Corners corners(grid.n_cols * grid.n_rows, svo::Corner(0, 0, threshold, 0, 0.0f));
  svo::feature_detection_utils::fastDetector(pyr, threshold, border, min_level, max_level, corners, grid);
  for (size_t k = 0; k < corners.size(); ++k)
    corners_out[k] = orc_corner{corners[k].x, corners[k].y, corners[k].level, corners[k].score, corners[k].angle};
}

// feature_detection_utils::edgeletDetector_V2: corners_out [n_cells] pre-filled with score = threshold like GradientDetectorGrid::detect.
void ref_edgelet_detector_v2(int n_levels, const uint8_t* const* data, const int* cols, const int* rows, const int* step,
                             int threshold, int border, int cell_size, const uint8_t* occupancy, orc_corner* corners_out) {
  svo::ImgPyr pyr = makePyr(n_levels, data, cols, rows, step);
  svo::OccupandyGrid2D grid(cell_size, svo::OccupandyGrid2D::getNCell(cols[0], cell_size), svo::OccupandyGrid2D::getNCell(rows[0], cell_size));
  for (size_t k = 0; k < grid.occupancy_.size(); ++k) grid.occupancy_[k] = occupancy && occupancy[k];
  svo::Corners corners(grid.n_cols * grid.n_rows, svo::Corner(0, 0, threshold, 0, 0.0f));
  svo::feature_detection_utils::edgeletDetector_V2(pyr, threshold, border, 0, 0, corners, grid);
  for (size_t k = 0; k < corners.size(); ++k)
    corners_out[k] = orc_corner{corners[k].x, corners[k].y, corners[k].level, corners[k].score, corners[k].angle};
}

double ref_angle_at_pixel_histogram(const uint8_t* img, int cols, int rows, int step, int x, int y, int halfpatch_size) {
  cv::Mat m(rows, cols, CV_8UC1, const_cast<uint8_t*>(img), (size_t)step);
  return svo::feature_detection_utils::getAngleAtPixelUsingHistogram(m, Eigen::Vector2i(x, y), (size_t)halfpatch_size);
}

// AbstractDetector::detect(img_pyr, mask = empty, max_n_features, ...) of the detector makeDetector builds for detector_type
// (svo::DetectorType: 0 kFast, 2 kFastGrad, 5 kGridGrad, feature_detection_types.h:32-44). occupancy [n_cells] is copied into
// grid_ before the call (what the callers do through fillWithKeypoints). Returns the number of features; outputs have room for
// max_n entries.
int ref_detect_features(int detector_type, int n_levels, const uint8_t* const* data, const int* cols, const int* rows,
                        const int* step, double threshold_primary, double threshold_secondary, int border, int min_level,
                        int max_level, int cell_size, const uint8_t* occupancy, int max_n, double* px_out, double* score_out,
                        int* level_out, double* grad_out, int* type_out) {
  using namespace vk::cameras;
  svo::ImgPyr pyr = makePyr(n_levels, data, cols, rows, step);
  typedef PinholeProjection<NoDistortion> P;
  svo::CameraPtr cam = std::make_shared<CameraGeometry<P>>(cols[0], rows[0], P(300.0, 300.0, cols[0] / 2.0, rows[0] / 2.0, NoDistortion()));
  svo::DetectorOptions o;
  o.cell_size = cell_size; o.max_level = max_level; o.min_level = min_level; o.border = border;
  o.detector_type = static_cast<svo::DetectorType>(detector_type);
  o.threshold_primary = threshold_primary; o.threshold_secondary = threshold_secondary;
  svo::AbstractDetector::Ptr det = svo::feature_detection_utils::makeDetector(o, cam);
  for (size_t k = 0; k < det->grid_.occupancy_.size(); ++k) det->grid_.occupancy_[k] = occupancy && occupancy[k];
  svo::Keypoints px; svo::Scores sc; svo::Levels lv; svo::Gradients gr; svo::FeatureTypes ty;
  det->detect(pyr, cv::Mat(), (size_t)max_n, px, sc, lv, gr, ty);
  const int n = (int)px.cols();
  for (int i = 0; i < n && i < max_n; ++i) {
    px_out[2 * i] = px(0, i); px_out[2 * i + 1] = px(1, i);
    score_out[i] = sc(i); level_out[i] = lv(i);
    grad_out[2 * i] = gr(0, i); grad_out[2 * i + 1] = gr(1, i);
    type_out[i] = (int)ty[i];
  }
  return n;
}

}  // extern "C"
