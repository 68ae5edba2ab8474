// svo_b200.h — C++ host facades of the B200 front-end hot path, mirroring the reference's class surface.
//
// The reference wires these classes behind src/svo/include/svo/svo_factory.h by new-ing them in FrameHandlerBase
// (src/svo/src/frame_handler_base.cpp:125,135,145). The facades keep the reference's names, method signatures, public data
// members and error behaviour for the hot path and forward the work to the C ABI of include/svo_cuda.h — there is no CPU
// implementation behind them. In the reference tree the argument types are Eigen / OpenCV / minkindr types; neither library
// is installed in this build environment, so this header carries minimal stand-ins with the SAME member names the facades
// touch (Frame::img_pyr_, px_vec_, f_vec_, T_f_w_, invmu_sigma2_a_b_vec_, ...). INTEGRATION.md lists the one-line adapters from
// the real types (cv::Mat -> Image, Eigen::Matrix<double,2,Dynamic> -> Keypoints, kindr::minimal::QuatTransformation ->
// Transformation).
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <condition_variable>
#include <mutex>
#include <queue>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/svo_cuda.h"

namespace svo {

using FloatType = double;  // src/svo_common/include/svo/common/types.h:16

// ---- stand-ins for the container types (see header comment) ------------------------------------------------------------
struct Image {  // cv::Mat as used on this path: 8-bit, data / cols / rows / step
  std::vector<uint8_t> storage;
  uint8_t* data = nullptr;
  int cols = 0, rows = 0;
  size_t step = 0;
  Image() = default;
  Image(int rows_, int cols_) : storage(size_t(rows_) * cols_), data(storage.data()), cols(cols_), rows(rows_), step(cols_) {}
  bool empty() const { return data == nullptr; }
};
using ImgPyr = std::vector<Image>;

struct Transformation {  // kindr::minimal::QuatTransformation: rotation quaternion (w,x,y,z) + position
  std::array<double, 4> q{{1, 0, 0, 0}};
  std::array<double, 3> t{{0, 0, 0}};
  void toArray(double* a) const { for (int i = 0; i < 4; ++i) a[i] = q[i]; for (int i = 0; i < 3; ++i) a[4 + i] = t[i]; }
  static Transformation fromArray(const double* a) { Transformation T; for (int i = 0; i < 4; ++i) T.q[i] = a[i]; for (int i = 0; i < 3; ++i) T.t[i] = a[4 + i]; return T; }
  Transformation inverse() const;
  Transformation operator*(const Transformation& rhs) const;
};

using Keypoint = std::array<double, 2>;
using BearingVector = std::array<double, 3>;
using GradientVector = std::array<double, 2>;
using SeedState = std::array<double, 4>;  // inv-mu, sigma2, a, b (src/svo_common/include/svo/common/seed.h)

enum class FeatureType : uint8_t {  // src/svo_common/include/svo/common/types.h:60-73
  kEdgeletSeed = 0, kCornerSeed = 1, kMapPointSeed = 2, kEdgeletSeedConverged = 3, kCornerSeedConverged = 4,
  kMapPointSeedConverged = 5, kEdgelet = 6, kCorner = 7, kMapPoint = 8, kFixedLandmark = 9, kOutlier = 10
};
inline bool isSeed(FeatureType t) { return static_cast<uint8_t>(t) < 6; }
inline bool isCornerEdgeletSeed(FeatureType t) {
  return t == FeatureType::kEdgeletSeedConverged || t == FeatureType::kCornerSeedConverged || t == FeatureType::kEdgeletSeed || t == FeatureType::kCornerSeed;
}
inline bool isConvergedCornerEdgeletSeed(FeatureType t) { return t == FeatureType::kEdgeletSeedConverged || t == FeatureType::kCornerSeedConverged; }
inline bool isUnconvergedCornerEdgeletSeed(FeatureType t) { return t == FeatureType::kEdgeletSeed || t == FeatureType::kCornerSeed; }
inline bool isEdgelet(FeatureType t) { return t == FeatureType::kEdgelet || t == FeatureType::kEdgeletSeed || t == FeatureType::kEdgeletSeedConverged; }
inline bool isFixedLandmark(FeatureType t) { return t == FeatureType::kFixedLandmark; }
inline bool isMapPoint(FeatureType t) { return t == FeatureType::kMapPoint || t == FeatureType::kMapPointSeed || t == FeatureType::kMapPointSeedConverged; }

struct Camera {  // vk::cameras::CameraGeometry<PinholeProjection<...>>
  svo_camera model{};
  int imageWidth() const { return model.width; }
  int imageHeight() const { return model.height; }
};
using CameraPtr = std::shared_ptr<Camera>;

namespace b200 { class GpuPyramid; }

struct Frame;
using FramePtr = std::shared_ptr<Frame>;
struct KeypointIdentifier {  // src/svo_common/include/svo/common/point.h:36-60
  std::weak_ptr<Frame> frame;
  int frame_id;
  size_t keypoint_index_;
  KeypointIdentifier(const FramePtr& _frame, const size_t _feature_index);
};
// svo::Point reduced to the members the Reprojector reads / writes (point.h:82-91)
struct Point {
  int id_ = -1;
  std::array<double, 3> pos_{{0, 0, 0}};
  std::vector<KeypointIdentifier> obs_;
  std::array<int, 8> last_projected_kf_id_;
  int n_failed_reproj_ = 0;
  int n_succeeded_reproj_ = 0;
  int last_structure_optim_ = 0;  // point.h:89: frame id of the last Point::optimize
  Point() { last_projected_kf_id_.fill(-1); }
  // Point::optimize (point.h:155; point.cpp:248-325) for this one point: ONE svo_cuda_optimize_points call with P = 1
  // (optimizeStructure below batches all landmarks of a frame into one call).
  void optimize(const size_t n_iter, bool using_bearing_vector = false);
  int id() const { return id_; }
};
using PointPtr = std::shared_ptr<Point>;
struct SeedRef {  // src/svo_common/include/svo/common/feature_wrapper.h:15-26
  FramePtr keyframe;
  int seed_id = -1;
};

// svo::Frame reduced to the members the hot path reads / writes (src/svo_common/include/svo/common/frame.h:30-424)
struct Frame {
  using Ptr = std::shared_ptr<Frame>;
  int id_ = 0;
  CameraPtr cam_;
  ImgPyr img_pyr_;
  Transformation T_f_w_;     // frame (camera) from world
  Transformation T_cam_imu_; // T_cam_imu(); T_imu_cam() is its inverse
  size_t num_features_ = 0;
  std::vector<Keypoint> px_vec_;
  std::vector<BearingVector> f_vec_;
  std::vector<GradientVector> grad_vec_;
  std::vector<double> score_vec_;
  std::vector<int> level_vec_;
  std::vector<FeatureType> type_vec_;
  // Distance of the feature's landmark / seed from this camera's centre, < 0 when the feature has neither
  // (replaces the landmark_vec_ / seed_ref_vec_ pointer walk of sparse_img_align.cpp:282-293, done by the adapter).
  std::vector<double> depth_vec_;
  std::vector<SeedState> invmu_sigma2_a_b_vec_;
  std::vector<int> track_id_vec_;        // frame.h:73 (KLT track ids, written by the FeatureTracker)
  std::vector<PointPtr> landmark_vec_;   // used by the Reprojector; may stay empty for the other facades
  std::vector<SeedRef> seed_ref_vec_;
  double seed_mu_range_ = 0.0;
  mutable std::shared_ptr<b200::GpuPyramid> gpu_;  // device-resident copy of img_pyr_ (the analogue of the reference's FrameGpu)

  const CameraPtr& cam() const { return cam_; }
  Transformation T_imu_world() const { return T_cam_imu_.inverse() * T_f_w_; }
  std::array<double, 3> pos() const { return T_f_w_.inverse().t; }  // T_world_cam().getPosition()
  bool isValidLandmark(size_t i) const { return i < landmark_vec_.size() && landmark_vec_[i] != nullptr; }
  size_t numLandmarks() const { size_t n = 0; for (const PointPtr& p : landmark_vec_) n += p != nullptr; return n; }  // frame.h:176-181
  size_t numFeatures() const { return num_features_; }
  size_t numTrackedFeatures() const;  // frame.h:153-163
  const Transformation& T_cam_imu() const { return T_cam_imu_; }
  void clearFeatureStorage();
};
struct FrameBundle {  // frame.h:426-543
  using Ptr = std::shared_ptr<FrameBundle>;
  std::vector<FramePtr> frames_;
  size_t size() const { return frames_.size(); }
  bool empty() const { return frames_.empty(); }
  const FramePtr& at(size_t i) const { return frames_.at(i); }
};

namespace frame_utils {
// src/svo_common/src/frame.cpp:372-386 — builds the pyramid on the GPU (svo_cuda_pyr_build), keeps it resident in *gpu and
// mirrors the levels into `pyr` for host code that still reads them.
void createImgPyramid(const Image& img_level_0, int n_levels, ImgPyr& pyr, std::shared_ptr<b200::GpuPyramid>* gpu = nullptr);
}  // namespace frame_utils

// ---- (b) SparseImgAlign ------------------------------------------------------------------------------------------------
struct SparseImgAlignOptions {  // src/svo_img_align/include/svo/img_align/sparse_img_align_base.h:37-46
  int max_level = 4;
  int min_level = 1;
  bool estimate_illumination_gain = false;
  bool estimate_illumination_offset = false;
  bool use_distortion_jacobian = false;
  bool robustification = false;
  double weight_scale = 10;
};
namespace solver {
struct MiniLeastSquaresSolverOptions {  // src/vikit/vikit_solver/include/vikit/solver/mini_least_squares_solver.h:20-47
  size_t max_iter = 15;
  double eps = 0.0000000001;
  bool verbose = false;
};
}  // namespace solver

class SparseImgAlignBase {
 public:
  using Ptr = std::shared_ptr<SparseImgAlignBase>;
  using SolverOptions = solver::MiniLeastSquaresSolverOptions;
  SolverOptions solver_options_;
  SparseImgAlignBase(SolverOptions optimization_options, SparseImgAlignOptions options)
      : solver_options_(optimization_options), options_(options) {}
  virtual ~SparseImgAlignBase() = default;
  virtual size_t run(const FrameBundle::Ptr& ref_frames, const FrameBundle::Ptr& cur_frames) = 0;
  static SolverOptions getDefaultSolverOptions();  // sparse_img_align_base.cpp:35-42
  void setWeightedPrior(const Transformation& T_cur_ref_prior, double alpha_prior, double beta_prior, double lambda_rot,
                        double lambda_trans, double lambda_alpha, double lambda_beta);
  void reset();  // MiniLeastSquaresSolver::reset (mini_least_squares_solver.hpp:240-250)
  // sparse_img_align_base.h:85-93. The device kernel is built for the reference's own 4x4 patches (sparse_img_align.cpp:31: the only
  // size SparseImgAlign ever sets); any other size is rejected instead of silently aligning with a different patch.
  template <class derived>
  void setPatchSize(size_t patch_size) {
    if (patch_size != 4) throw std::invalid_argument("SparseImgAlign (B200): only patch_size 4 is implemented (sparse_img_align.cpp:31)");
    patch_size_ = int(patch_size); border_size_ = 1;
    patch_size_with_border_ = patch_size_ + 2 * border_size_;
    patch_area_ = patch_size_ * patch_size_;
  }
  inline void setMaxNumFeaturesToAlign(int num) { max_num_features_ = num; }
  inline void setAlphaInitialValue(double alpha_init) { alpha_init_ = alpha_init; }
  inline void setBetaInitialValue(double beta_init) { beta_init_ = beta_init; }
  inline void setCompensation(const bool do_compensation) {
    options_.estimate_illumination_gain = do_compensation;
    options_.estimate_illumination_offset = do_compensation;
  }
  double getError() const { return chi2_; }
  const std::array<double, 64>& getHessian() const { return H_; }  // row-major 8x8

 protected:
  SparseImgAlignOptions options_;
  bool have_prior_ = false;
  svo_align_prior prior_{};
  double prior_lambda_rot_ = 0, prior_lambda_trans_ = 0, prior_lambda_alpha_ = 0, prior_lambda_beta_ = 0;
  int max_num_features_ = -1;  // stored but unused by the CPU reference as well (sparse_img_align_base.h:102-105)
  double alpha_init_ = 0.0, beta_init_ = 0.0;
  double chi2_ = 0.0;
  std::array<double, 64> H_{};
  int patch_size_ = 4, border_size_ = 1, patch_size_with_border_ = 6, patch_area_ = 16;  // sparse_img_align_base.h:132-135
};

class SparseImgAlign : public SparseImgAlignBase {  // src/svo_img_align/include/svo/img_align/sparse_img_align.h:30-77
 public:
  using Ptr = std::shared_ptr<SparseImgAlign>;
  SparseImgAlign(SolverOptions optimization_options, SparseImgAlignOptions options)
      : SparseImgAlignBase(optimization_options, options) {}
  // Returns the number of tracked features; writes T_f_w_ of every cur frame (sparse_img_align.cpp:34-113).
  size_t run(const FrameBundle::Ptr& ref_frames, const FrameBundle::Ptr& cur_frames) override;
};

// ---- (c) feature alignment + Matcher ------------------------------------------------------------------------------------
namespace feature_alignment {  // src/svo_direct/include/svo/direct/feature_alignment.h:23-43
bool align1D(const Image& cur_img, const GradientVector& dir, uint8_t* ref_patch_with_border, uint8_t* ref_patch, const int n_iter,
             const bool affine_est_offset, const bool affine_est_gain, Keypoint* cur_px_estimate, double* h_inv = nullptr);
bool align2D(const Image& cur_img, uint8_t* ref_patch_with_border, uint8_t* ref_patch, const int n_iter,
             const bool affine_est_offset, const bool affine_est_gain, Keypoint& cur_px_estimate, bool no_simd = false);
}  // namespace feature_alignment

struct FeatureWrapper {  // src/svo_common/include/svo/common/feature_wrapper.h:34-45 (the fields the matcher reads)
  FeatureType type;
  Keypoint px;
  BearingVector f;
  GradientVector grad;
  int level;
};

namespace patch_score {
// src/svo_direct/include/svo/direct/patch_score.h:43-109: the reference patch and its sums; the scoring itself (computeScore) runs on
// the device inside the epipolar scan.
template <int HALF_PATCH_SIZE>
class ZMSSD {
 public:
  static const int patch_size_ = 2 * HALF_PATCH_SIZE;
  static const int patch_area_ = patch_size_ * patch_size_;
  static const int threshold_ = 2000 * patch_area_;
  uint8_t* ref_patch_;
  int sumA_, sumAA_;
  explicit ZMSSD(uint8_t* ref_patch) : ref_patch_(ref_patch) {
    uint32_t a = 0, aa = 0;
    for (int r = 0; r < patch_area_; ++r) { const uint32_t n = ref_patch_[r]; a += n; aa += n * n; }
    sumA_ = int(a); sumAA_ = int(aa);
  }
  static int threshold() { return threshold_; }
};
}  // namespace patch_score

class Matcher {  // src/svo_direct/include/svo/direct/matcher.h:28-140
 public:
  static const int kHalfPatchSize = 4;
  static const int kPatchSize = 8;
  typedef patch_score::ZMSSD<kHalfPatchSize> PatchScore;  // matcher.h:36
  struct Options {
    bool align_1d = false;
    int align_max_iter = 10;
    double max_epi_length_optim = 2.0;
    size_t max_epi_search_steps = 100;
    bool subpix_refinement = true;
    bool epi_search_edgelet_filtering = true;
    bool scan_on_unit_sphere = true;
    double epi_search_edgelet_max_angle = 0.7;
    bool verbose = false;
    bool use_affine_warp_ = true;
    bool affine_est_offset_ = true;
    bool affine_est_gain_ = false;
    double max_patch_diff_ratio = 2.0;
  } options_;
  enum class MatchResult { kSuccess, kFailScore, kFailTriangulation, kFailVisibility, kFailWarp, kFailAlignment, kFailRange,
                           kFailAngle, kFailCloseView, kFailLock, kFailTooFar };
  alignas(16) uint8_t patch_[kPatchSize * kPatchSize] = {};                           // matcher.h:70: the warped reference patch
  alignas(16) uint8_t patch_with_border_[(kPatchSize + 2) * (kPatchSize + 2)] = {};   // matcher.h:71
  std::array<double, 4> A_cur_ref_{};  // row-major 2x2
  std::array<double, 2> epi_image_{};  // matcher.h:73: vector from epipolar start to end on the image plane
  double epi_length_pyramid_ = 0;
  double h_inv_ = 0;
  int search_level_ = 0;
  bool reject_ = false;
  Keypoint px_cur_{};
  BearingVector f_cur_{};

  MatchResult findMatchDirect(const Frame& ref_frame, const Frame& cur_frame, const FeatureWrapper& ref_ftr, const FloatType& ref_depth,
                              Keypoint& px_cur);
  MatchResult findEpipolarMatchDirect(const Frame& ref_frame, const Frame& cur_frame, const FeatureWrapper& ref_ftr,
                                      const double d_estimate_inv, const double d_min_inv, const double d_max_inv, double& depth);
  MatchResult findEpipolarMatchDirect(const Frame& ref_frame, const Frame& cur_frame, const Transformation& T_cur_ref,
                                      const FeatureWrapper& ref_ftr, const double d_estimate_inv, const double d_min_inv,
                                      const double d_max_inv, double& depth);
  // matcher.h:111-122: the ZMSSD scan of the segment A~C~B on its own (reads options_ and epi_length_pyramid_, as the reference does)
  void scanEpipolarLine(const Frame& frame, const BearingVector& A, const BearingVector& B, const BearingVector& C,
                        const PatchScore& patch_score, const int patch_level, Keypoint* image_best, int* zmssd_best);
  static std::string getResultString(const MatchResult& result);
  svo_matcher_options cOptions() const;
};

// ---- (d) DepthFilter -----------------------------------------------------------------------------------------------------
struct DepthFilterOptions {  // src/svo_direct/include/svo/direct/depth_filter.h:27-60
  double seed_convergence_sigma2_thresh = 200.0;
  double mappoint_convergence_sigma2_thresh = 500.0;
  bool scan_epi_unit_sphere = false;
  bool affine_est_offset = true;
  bool affine_est_gain = false;
  size_t max_n_seeds_per_frame = 200;  // depth_filter.h:61
};

class AbstractDetector;
struct DetectorOptions;

namespace depth_filter_utils {  // depth_filter.h:179-236
// depth_filter.h:181-188; depth_filter.cpp:254-365: detect new features in the cells the detector's grid leaves free (device call
// through the detector facade), append them as corner / edgelet seeds with mu = 1 / depth_mean, sigma2 = (1 / depth_min)^2 / 36,
// a = b = 10, and set the frame's seed_mu_range_.
void initializeSeeds(const FramePtr& frame, const std::shared_ptr<AbstractDetector>& feature_detector, const size_t max_n_seeds,
                     const float depth_min, const float depth_max, const float depth_mean);
bool updateSeed(const Frame& cur_frame, Frame& ref_frame, const size_t& seed_index, Matcher& matcher,
                const FloatType sigma2_convergence_threshold, const bool check_visibility = true, const bool check_convergence = false,
                const bool use_vogiatzis_update = true);
bool updateFilterVogiatzis(const FloatType z, const FloatType tau2, const FloatType z_range, SeedState& seed);
bool updateFilterGaussian(const FloatType z, const FloatType tau2, SeedState& seed);  // depth_filter.h:207-211; depth_filter.cpp:554-579
double computeTau(const Transformation& T_ref_cur, const BearingVector& f, const FloatType z, const FloatType px_error_angle);
}  // namespace depth_filter_utils

class DepthFilter {
 public:
  DepthFilterOptions options_;
  // depth_filter.h:162-165 ("need public access to set grid occupancy"): callers lock feature_detector_mut_ around grid updates
  std::mutex feature_detector_mut_;
  std::shared_ptr<AbstractDetector> feature_detector_;      // set by the detector-building constructor
  std::shared_ptr<AbstractDetector> sec_feature_detector_;  // extra Shi-Tomasi points for loop closing (extra_map_points): never built here
  explicit DepthFilter(const DepthFilterOptions& options);
  // depth_filter.h:80-84: the constructor that builds its own detector through makeDetector
  DepthFilter(const DepthFilterOptions& options, const DetectorOptions& detector_options, const CameraPtr& cam);
  ~DepthFilter();  // stops the thread if necessary (depth_filter.cpp:59-63)
  DepthFilter(const DepthFilter&) = delete;
  DepthFilter& operator=(const DepthFilter&) = delete;
  // depth_filter.cpp:65-88: seed initialisation and seed updates in a parallel thread. The worker owns its own svo_cuda context
  // (= its own stream, b200::context() is per host thread), so its launches overlap with the pipeline thread's — the GPU analogue of
  // the reference's hand-off: addKeyframe / updateSeeds then only enqueue and return.
  void startThread();
  void stopThread();
  // DepthFilter::addKeyframe (depth_filter.cpp:89-133): initializeSeeds on the new keyframe, at once or — threaded — as a job that
  // first drops every queued job ("this one has priority", :129-131)
  void addKeyframe(const FramePtr& frame, const double depth_mean, const double depth_min, const double depth_max);
  void reset();  // depth_filter.cpp:135-144: drops the queued jobs
  // DepthFilter::updateSeeds (depth_filter.cpp:200-249): every seed of every ref frame against cur_frame. Not threaded: one batched
  // launch per ref frame, returns the number of successful updates. Threaded: one job per ref frame is queued (the reference queues one
  // per seed and processes them one at a time; a batch of one per launch would idle the GPU) and 0 is returned, as in the reference.
  size_t updateSeeds(const std::vector<FramePtr>& ref_frames_with_seeds, const FramePtr& cur_frame);
  Matcher& getMatcher() { return matcher_; }
  // (not in the reference) block until the worker has drained its queue: lets a caller read the seed states at a defined point
  void waitForJobs();

 private:
  struct Job {  // depth_filter.h:68-101
    enum Type { UPDATE, SEED_INIT } type = UPDATE;
    FramePtr cur_frame, ref_frame;
    double min_depth = 0, max_depth = 0, mean_depth = 0;
  };
  void updateSeedsLoop();
  size_t updateSeedsOfRefFrame(const FramePtr& ref_frame, const FramePtr& cur_frame);
  Matcher matcher_;
  std::mutex jobs_mut_;
  std::condition_variable jobs_condvar_, idle_condvar_;
  std::queue<Job> jobs_;
  bool quit_thread_ = false, busy_ = false;
  std::unique_ptr<std::thread> thread_;
};

// ---- (a) FAST detector -----------------------------------------------------------------------------------------------------
struct Corner {  // src/svo_direct/include/svo/direct/feature_detection_types.h:17-29
  int x, y, level;
  float score, angle;
  Corner(int _x, int _y, float _score, int _level, float _angle) : x(_x), y(_y), level(_level), score(_score), angle(_angle) {}
};
using Corners = std::vector<Corner>;

enum class DetectorType {  // feature_detection_types.h:33-45 (the grid detectors of this path are implemented)
  kFast, kGrad, kFastGrad, kShiTomasi, kShiTomasiGrad, kGridGrad, kAll, kGradHuangMumford, kCanny, kSobel
};

struct DetectorOptions {  // feature_detection_types.h:49-84
  size_t cell_size = 30;
  int max_level = 2;
  int min_level = 0;
  int border = 8;
  DetectorType detector_type = DetectorType::kFast;
  double threshold_primary = 10.0;
  double threshold_secondary = 100.0;
  size_t sec_grid_fineness = 1;  // feature_detection_types.h:79-80
};

class OccupandyGrid2D {  // src/svo_common/include/svo/common/occupancy_grid_2d.h:10-110
 public:
  const int cell_size, n_cols, n_rows;
  std::vector<uint8_t> occupancy_;
  OccupandyGrid2D(int cell_size_, int n_cols_, int n_rows_)
      : cell_size(cell_size_), n_cols(n_cols_), n_rows(n_rows_), occupancy_(size_t(n_cols_) * n_rows_, 0) {}
  void reset() { std::fill(occupancy_.begin(), occupancy_.end(), 0); }
  size_t size() const { return occupancy_.size(); }
  size_t getCellIndex(int x, int y, int scale = 1) const { return size_t((scale * y) / cell_size) * n_cols + size_t((scale * x) / cell_size); }
  void fillWithKeypoints(const Keypoint& px) { occupancy_.at(getCellIndex(int(px[0]), int(px[1]), 1)) = 1; }  // occupancy_grid_2d.h:40-48 (one column)
};

using Keypoints = std::vector<Keypoint>;
using Gradients = std::vector<GradientVector>;
using Scores = std::vector<double>;
using Levels = std::vector<int>;
using FeatureTypes = std::vector<FeatureType>;

namespace feature_detection_utils {
// src/svo_direct/include/svo/direct/feature_detection_utils.h:43-50; `gpu` is the frame's device pyramid.
void fastDetector(const b200::GpuPyramid& gpu, const int threshold, const int border, const size_t min_level, const size_t max_level,
                  Corners& corners, OccupandyGrid2D& grid);
// feature_detection_utils.h:75-82 (runs on pyramid level 1, reports level 0; min_level / max_level are unused there too)
void edgeletDetector_V2(const b200::GpuPyramid& gpu, const int threshold, const int border, const int min_level, const int max_level,
                        Corners& corners, OccupandyGrid2D& grid);
// feature_detection_utils.h:27-38 without the mask (the facades run with an empty mask): score > threshold, grid marking, sort by
// score, at most max_n_features appended.
void fillFeatures(const Corners& corners, const FeatureType& type, const double& threshold, const size_t max_n_features,
                  Keypoints& keypoints, Scores& scores, Levels& levels, Gradients& gradients, FeatureTypes& types, OccupandyGrid2D& grid);
}  // namespace feature_detection_utils

class AbstractDetector {  // src/svo_direct/include/svo/direct/feature_detection.h:20-64
 public:
  using Ptr = std::shared_ptr<AbstractDetector>;
  DetectorOptions options_;
  OccupandyGrid2D grid_;
  // feature_detection.h:59-61: the finer grid a secondary detector checks new features against (filled by the caller through
  // fillWithKeypoints; only the Shi-Tomasi detector, which is not on this path, reads it)
  OccupandyGrid2D closeness_check_grid_;
  AbstractDetector(const DetectorOptions& options, const CameraPtr& cam);
  virtual ~AbstractDetector() = default;
  // AbstractDetector::detect(const FramePtr&) (feature_detection.cpp:40-50): appends the features to the frame's SoA arrays and
  // computes their unit bearing vectors. max_n_features = grid size.
  void detect(const FramePtr& frame);
  // The virtual detect of the reference (feature_detection.h:41-49) with the frame's device pyramid in place of (img_pyr, mask).
  virtual void detect(const b200::GpuPyramid& gpu, const size_t max_n_features, Keypoints& px_vec, Scores& score_vec, Levels& level_vec,
                      Gradients& grad_vec, FeatureTypes& types_vec) = 0;
  void resetGrid() { grid_.reset(); closeness_check_grid_.reset(); }  // feature_detection.h:51-55
};

class FastDetector : public AbstractDetector {  // feature_detection.h:66-81; feature_detection.cpp:53-74
 public:
  using AbstractDetector::AbstractDetector;
  using AbstractDetector::detect;
  void detect(const b200::GpuPyramid& gpu, const size_t max_n_features, Keypoints& px_vec, Scores& score_vec, Levels& level_vec,
              Gradients& grad_vec, FeatureTypes& types_vec) override;
};

class GradientDetectorGrid : public AbstractDetector {  // feature_detection.h:105-122; feature_detection.cpp:130-151
 public:
  using AbstractDetector::AbstractDetector;
  using AbstractDetector::detect;
  void detect(const b200::GpuPyramid& gpu, const size_t max_n_features, Keypoints& px_vec, Scores& score_vec, Levels& level_vec,
              Gradients& grad_vec, FeatureTypes& types_vec) override;
};

class FastGradDetector : public AbstractDetector {  // feature_detection.h:133-150; feature_detection.cpp:154-194 (the default detector)
 public:
  using AbstractDetector::AbstractDetector;
  using AbstractDetector::detect;
  void detect(const b200::GpuPyramid& gpu, const size_t max_n_features, Keypoints& px_vec, Scores& score_vec, Levels& level_vec,
              Gradients& grad_vec, FeatureTypes& types_vec) override;
};

namespace feature_detection_utils {
// feature_detection_utils.h:21-25: kFast, kFastGrad and kGridGrad are available; the other types throw b200::Error.
AbstractDetector::Ptr makeDetector(const DetectorOptions& options, const CameraPtr& cam);
}  // namespace feature_detection_utils

// ---- (f1) Reprojector ---------------------------------------------------------------------------------------------------------
struct ReprojectorOptions {  // src/svo/include/svo/reprojector.h:27-70 (without the global-map options)
  size_t max_n_features_per_frame = 120;
  size_t cell_size = 30;
  bool reproject_unconverged_seeds = true;
  double max_unconverged_seeds_ratio = -1.0;
  size_t min_required_features = 0;
  double seed_sigma2_thresh = 200;
  bool remove_unconstrained_points = true;
  bool affine_est_offset = true;
  bool affine_est_gain = false;
};

class Reprojector {  // reprojector.h:77-166
 public:
  using Ptr = std::shared_ptr<Reprojector>;
  ReprojectorOptions options_;
  struct Statistics {
    size_t n_matches = 0, n_trials = 0;
    void reset() { n_matches = 0; n_trials = 0; }
    void add(const Statistics s) { n_matches += s.n_matches; n_trials += s.n_trials; }
    double successRate() const { return n_trials == 0 ? 0.0 : n_matches / (1.0 * n_trials); }
  } stats_;
  std::unique_ptr<OccupandyGrid2D> grid_;
  size_t camera_index_;
  Reprojector(const ReprojectorOptions& options, size_t camera_index) : options_(options), camera_index_(camera_index) {}
  // Project the landmarks and seeds of visible_kfs into cur_frame and match at most one per grid cell
  // (reprojector.cpp:28-310): up to three svo_cuda_reproject_match calls (landmarks, converged seeds, unconverged seeds).
  void reprojectFrames(const FramePtr& cur_frame, const std::vector<FramePtr>& visible_kfs, std::vector<PointPtr>& trash_points);

 private:
  bool doesFrameHaveEnoughFeatures(const FramePtr& frame) const {
    return options_.max_n_features_per_frame > 0 && frame->numTrackedFeatures() >= options_.max_n_features_per_frame;
  }
};

// ---- (f3) FeatureTracker (pyramidal KLT tracks) --------------------------------------------------------------------------------
namespace feature_alignment {
// feature_alignment::alignPyr2DVec (src/svo_direct/include/svo/direct/feature_alignment.h:68-80; .cpp:761-798) for features whose
// templates all live in ref_frame: ONE svo_cuda_align_pyr2d call. px_ref are truncated to integers as the tracker does.
void alignPyr2DVec(const Frame& ref_frame, const Frame& cur_frame, int max_level, int min_level, const std::vector<int>& patch_sizes,
                   int n_iter, float min_update_squared, const std::vector<std::array<int, 2>>& px_ref, std::vector<Keypoint>& px_cur,
                   std::vector<uint8_t>& status);
}  // namespace feature_alignment

struct PointIdProvider {  // src/svo_common/include/svo/common/point.h:22-34
  static int getNewPointId();
};

struct FeatureTrackerOptions {  // src/svo_tracker/include/svo/tracker/feature_tracking_types.h:9-43
  int klt_max_level = 4;
  int klt_min_level = 0;
  std::vector<int> klt_patch_sizes = {16, 16, 16, 8, 8};
  int klt_max_iter = 30;
  double klt_min_update_squared = 0.001;
  bool klt_template_is_first_observation = true;
  size_t min_tracks_to_detect_new_features = 50;
  bool reset_before_detection = true;
};

class FeatureRef {  // feature_tracking_types.h:46-77
 public:
  FeatureRef(const FrameBundle::Ptr& frame_bundle, size_t frame_index, size_t feature_index)
      : frame_bundle_(frame_bundle), frame_index_(frame_index), feature_index_(feature_index) {}
  const FrameBundle::Ptr getFrameBundle() const { return frame_bundle_; }
  size_t getFrameIndex() const { return frame_index_; }
  size_t getFeatureIndex() const { return feature_index_; }
  const Keypoint& getPx() const { return frame_bundle_->frames_.at(frame_index_)->px_vec_.at(feature_index_); }
  const BearingVector& getBearing() const { return frame_bundle_->frames_.at(frame_index_)->f_vec_.at(feature_index_); }
  const FramePtr getFrame() const { return frame_bundle_->frames_.at(frame_index_); }
 private:
  FrameBundle::Ptr frame_bundle_;
  size_t frame_index_, feature_index_;
};

class FeatureTrack {  // feature_tracking_types.h:83-146
 public:
  explicit FeatureTrack(int track_id) : track_id_(track_id) { feature_track_.reserve(10); }
  int getTrackId() const { return track_id_; }
  const std::vector<FeatureRef>& getFeatureTrack() const { return feature_track_; }
  size_t size() const { return feature_track_.size(); }
  bool empty() const { return feature_track_.empty(); }
  const FeatureRef& front() const { return feature_track_.front(); }
  const FeatureRef& back() const { return feature_track_.back(); }
  const FeatureRef& at(size_t i) const { return feature_track_.at(i); }
  void pushBack(const FrameBundle::Ptr& frame_bundle, size_t frame_index, size_t feature_index) {
    feature_track_.emplace_back(frame_bundle, frame_index, feature_index);
  }
  double getDisparity() const;  // feature_tracking_types.cpp:38-41
 private:
  int track_id_;
  std::vector<FeatureRef> feature_track_;
};
using FeatureTracks = std::vector<FeatureTrack>;

namespace feature_tracking_utils {
double getTracksDisparityPercentile(const FeatureTracks& tracks, double pivot_ratio);  // feature_tracking_utils.cpp:11-33
}

class FeatureTracker {  // src/svo_tracker/include/svo/tracker/feature_tracker.h:14-84
 public:
  // `cams` stands for the CameraBundle: one camera per frame of the bundles that will be tracked.
  FeatureTracker(const FeatureTrackerOptions& options, const DetectorOptions& detector_options, const std::vector<CameraPtr>& cams);
  void trackAndDetect(const FrameBundle::Ptr& nframe_kp1);         // feature_tracker.cpp:32-50
  // :52-127 — every active track of a camera goes through ONE svo_cuda_align_pyr2d call per template frame (one call in all
  // when the templates are first observations of the same detection, the default)
  size_t trackFrameBundle(const FrameBundle::Ptr& nframe_kp1);
  size_t initializeNewTracks(const FrameBundle::Ptr& nframe_k);    // :129-187
  const FeatureTracks& getActiveTracks(size_t frame_index) const { return active_tracks_.at(frame_index); }
  size_t getTotalActiveTracks() const;
  void getNumTrackedAndDisparityPerFrame(double pivot_ratio, std::vector<size_t>* num_tracked, std::vector<double>* disparity) const;
  FrameBundle::Ptr getOldestFrameInTrack(size_t frame_index) const { return active_tracks_.at(frame_index).front().at(0).getFrameBundle(); }
  void resetActiveTracks() { for (auto& t : active_tracks_) t.clear(); }
  void resetTerminatedTracks() { for (auto& t : terminated_tracks_) t.clear(); }
  void reset();
  FeatureTrackerOptions options_;
  const size_t bundle_size_;
  std::vector<AbstractDetector::Ptr> detectors_;
  std::vector<FeatureTracks> active_tracks_, terminated_tracks_;
};

// ---- (f3) StereoTriangulation ---------------------------------------------------------------------------------------------------
struct StereoTriangulationOptions {  // src/svo/include/svo/stereo_triangulation.h:12-18
  size_t triangulate_n_features = 120;
  double mean_depth_inv = 1.0 / 3.0;
  double min_depth_inv = 1.0 / 1.0;
  double max_depth_inv = 1.0 / 50.0;
};

class StereoTriangulation {  // stereo_triangulation.h:20-37
 public:
  typedef std::shared_ptr<StereoTriangulation> Ptr;
  StereoTriangulationOptions options_;
  AbstractDetector::Ptr feature_detector_;
  StereoTriangulation(const StereoTriangulationOptions& options, const AbstractDetector::Ptr& feature_detector)
      : options_(options), feature_detector_(feature_detector) {}
  // src/svo/src/stereo_triangulation.cpp:23-139: detects new features in frame0, visits them in the order of the reference's two
  // std::random_shuffle calls (libstdc++'s std::rand() based algorithm, corners first), matches them along the epipolar line in
  // frame1 (ONE svo_cuda_stereo_triangulate call for all of them) and creates a Point + the frame1 feature for the first
  // triangulate_n_features - numLandmarks() successes.
  void compute(const FramePtr& frame0, const FramePtr& frame1);
};

// FrameHandlerBase::optimizeStructure (src/svo/src/frame_handler_base.cpp:785-825): Point::optimize for the non-edgelet landmarks of
// every frame of the bundle — one svo_cuda_optimize_points call per frame — and last_structure_optim_ = frame id. As in the
// reference, max_n_pts only reorders the points (its loop runs over all of them, :812-816); 0 returns at once, -1 means all.
void optimizeStructure(const FrameBundle::Ptr& frames, int max_n_pts, int max_iter, bool optimize_on_sphere = false);

// ---- (f4) PoseOptimizer ---------------------------------------------------------------------------------------------------------
class PoseOptimizer {  // src/svo/include/svo/pose_optimizer.h:20-103
 public:
  using Ptr = std::shared_ptr<PoseOptimizer>;
  using SolverOptions = solver::MiniLeastSquaresSolverOptions;
  enum class ErrorType { kUnitPlane, kBearingVectorDiff, kImagePlane };
  struct Statistics {
    double reproj_error_after = 0.0;
    double reproj_error_before = 0.0;
  } stats_;
  explicit PoseOptimizer(SolverOptions solver_options) : solver_options_(solver_options) {}
  static SolverOptions getDefaultSolverOptions();  // pose_optimizer.cpp:22-29: GaussNewton, max_iter 10, eps 1e-6
  // Optimises frame->T_f_w_ of every frame of the bundle over its landmark / seed observations, marks outliers
  // (type kOutlier, landmark and seed reference dropped) and returns the number of remaining measurements (:39-94).
  size_t run(const FrameBundle::Ptr& frame_bundle, double reproj_thresh_px);
  void setRotationPrior(const std::array<double, 4>& R_frame_world, double lambda);  // quaternion (w,x,y,z), :30-37
  void reset() { have_prior_ = false; iter_ = 0; }  // MiniLeastSquaresSolver::reset
  size_t iterCount() const { return iter_; }
  void setErrorType(ErrorType type) { err_type_ = type; }
  double measurement_sigma_ = 1.0;
  ErrorType err_type_ = ErrorType::kUnitPlane;
  SolverOptions solver_options_;

 private:
  bool have_prior_ = false;
  std::array<double, 4> prior_q_{{1, 0, 0, 0}};
  double prior_lambda_ = 0.0;
  size_t iter_ = 0;
};

// ---- device plumbing --------------------------------------------------------------------------------------------------------
namespace b200 {
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
// The calling thread's context on the current device (created on first use; throws b200::Error when no GPU is present).
svo_cuda_ctx* context();
class GpuPyramid {  // RAII owner of a one-frame svo_cuda_pyr
 public:
  GpuPyramid(int width, int height, int n_levels);
  ~GpuPyramid();
  GpuPyramid(const GpuPyramid&) = delete;
  GpuPyramid& operator=(const GpuPyramid&) = delete;
  svo_cuda_pyr* handle() const { return pyr_; }
  int n_levels() const { return n_levels_; }
  int width() const { return width_; }
  int height() const { return height_; }
 private:
  svo_cuda_pyr* pyr_ = nullptr;
  int width_, height_, n_levels_;
};
// Device pyramid of a frame: reuses frame.gpu_ or uploads frame.img_pyr_[0] and builds the levels.
const GpuPyramid& ensureGpu(const Frame& frame);
}  // namespace b200

}  // namespace svo

// ---- (a2-a4) the fast:: leaves with their list-shaped results (src/fast_neon/include/fast/fast.h:11-41) ---------------------------------
// Same signatures as the reference; each call uploads the image it is given, runs the level kernel and compacts the corners in raster
// order on the device (svo_cuda_fast_corner_list / _corner_score / _nonmax_3x3). The detectors above do not go through these: they keep
// everything on the device and only read back one corner per grid cell.
namespace fast {
struct fast_xy {
  short x, y;
  fast_xy(short x_, short y_) : x(x_), y(y_) {}
};
typedef unsigned char fast_byte;
void fast_corner_detect_9(const fast_byte* img, int imgWidth, int imgHeight, int widthStep, short barrier, std::vector<fast_xy>& corners);
void fast_corner_detect_9_sse2(const fast_byte* img, int imgWidth, int imgHeight, int widthStep, short barrier, std::vector<fast_xy>& corners);
void fast_corner_detect_10(const fast_byte* img, int imgWidth, int imgHeight, int widthStep, short barrier, std::vector<fast_xy>& corners);
void fast_corner_detect_10_sse2(const fast_byte* img, int imgWidth, int imgHeight, int widthStep, short barrier, std::vector<fast_xy>& corners);
// NOTE: the reference takes no image size here; the pixels it reads lie within 3 of the listed corners, and so does the upload.
void fast_corner_score_10(const fast_byte* img, const int img_stride, const std::vector<fast_xy>& corners, const int threshold, std::vector<int>& scores);
void fast_nonmax_3x3(const std::vector<fast_xy>& corners, const std::vector<int>& scores, std::vector<int>& nonmax_corners);
}  // namespace fast
