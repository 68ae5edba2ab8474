// TEST INFRASTRUCTURE ONLY (oracle/): a reduced svo::Frame / svo::FrameBundle with the members and accessors the
// reference's direct front-end sources read (same names, types and meaning as src/svo_common/include/svo/common/frame.h),
// so that sparse_img_align.cpp, matcher.cpp, patch_warp.cpp and depth_filter.cpp compile from /root/reference without the
// map / bundle-adjustment / serialisation parts of the real class. The three Jacobian helpers are restated from
// frame.h:342-357 (jacobian_xyz2uv_imu) and frame.cpp:264-290 (getErrorMultiplier, getAngleError, jacobian_xyz2image_imu).
#pragma once
#include <algorithm>
#include <memory>
#include <vector>
#include <opencv2/core/core.hpp>
#include <glog/logging.h>
#include <svo/common/types.h>
#include <svo/common/transformation.h>
#include <svo/common/camera.h>
#include <svo/common/feature_wrapper.h>
#include <svo/common/point.h>
#include <svo/common/seed.h>

namespace vk {
inline Eigen::Matrix3d skew(const Eigen::Vector3d& v) {  // vikit/math_utils.h:85-92
  Eigen::Matrix3d m;
  m << 0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0;
  return m;
}
}  // namespace vk

namespace svo {

class Frame {
 public:
  using Landmarks = std::vector<PointPtr>;
  using SeedRefs = std::vector<SeedRef>;

  int id_ = 0;
  CameraPtr cam_;
  Transformation T_f_w_;
  ImgPyr img_pyr_;
  Transformation T_body_cam_;
  Transformation T_cam_body_;

  size_t num_features_ = 0u;
  Keypoints px_vec_;
  Bearings f_vec_;
  Scores score_vec_;
  Levels level_vec_;
  Gradients grad_vec_;
  FeatureTypes type_vec_;
  Landmarks landmark_vec_;
  TrackIds track_id_vec_;
  SeedRefs seed_ref_vec_;
  SeedStates invmu_sigma2_a_b_vec_;
  std::vector<bool> in_ba_graph_vec_;
  FloatType seed_mu_range_ = 0;

  Frame() {}
  Frame(const Frame&) = delete;
  Frame& operator=(const Frame&) = delete;

  void resizeFeatureStorage(size_t num) {  // frame.cpp:94-123 (contents are kept, new slots get the reference's initial values)
    if (static_cast<size_t>(px_vec_.cols()) < num) {
      const size_t n_old = px_vec_.cols();
      px_vec_.conservativeResize(Eigen::NoChange, num); f_vec_.conservativeResize(Eigen::NoChange, num);
      score_vec_.conservativeResize(num); level_vec_.conservativeResize(num); grad_vec_.conservativeResize(Eigen::NoChange, num);
      invmu_sigma2_a_b_vec_.conservativeResize(Eigen::NoChange, num); track_id_vec_.conservativeResize(num);
      type_vec_.resize(num, FeatureType::kCorner); landmark_vec_.resize(num, nullptr); seed_ref_vec_.resize(num);
      in_ba_graph_vec_.resize(num, false);
      for (size_t i = n_old; i < num; ++i) { level_vec_(i) = 0; track_id_vec_(i) = -1; score_vec_(i) = -1; }
    }
  }
  void clearFeatureStorage() {  // frame.cpp:125-139
    px_vec_.resize(Eigen::NoChange, 0); f_vec_.resize(Eigen::NoChange, 0); score_vec_.resize(0); level_vec_.resize(0);
    grad_vec_.resize(Eigen::NoChange, 0); invmu_sigma2_a_b_vec_.resize(Eigen::NoChange, 0); track_id_vec_.resize(0);
    type_vec_.clear(); landmark_vec_.clear(); seed_ref_vec_.clear();
    num_features_ = 0;
  }
  const cv::Mat& getMask() const { return cam_->getMask(); }
  FeatureWrapper getFeatureWrapper(size_t i) {
    return FeatureWrapper(type_vec_[i], px_vec_.col(i), f_vec_.col(i), grad_vec_.col(i), score_vec_(i), level_vec_(i),
                          landmark_vec_[i], seed_ref_vec_[i], track_id_vec_(i));
  }
  FeatureWrapper getEmptyFeatureWrapper() { return getFeatureWrapper(num_features_); }  // frame.cpp:166-169
  inline bool isValidLandmark(size_t i) const { return (landmark_vec_.at(i) != nullptr); }  // frame.h:148-150
  inline size_t numTrackedFeatures() const {  // frame.h:153-163
    size_t count = 0;
    for (size_t i = 0; i < num_features_; ++i)
      if ((isValidLandmark(i) && !isFixedLandmark(type_vec_[i]) && !isMapPoint(type_vec_[i])) || isCornerEdgeletSeed(type_vec_[i])) ++count;
    return count;
  }
  inline size_t numLandmarks() const {  // frame.h:176-181
    return static_cast<size_t>(std::count_if(landmark_vec_.begin(), landmark_vec_.end(), [](const PointPtr& p) { return p != nullptr; }));
  }
  inline size_t numFixedLandmarks() const {  // frame.h:183-191
    size_t count = 0;
    for (size_t i = 0; i < num_features_; ++i)
      if (isValidLandmark(i) && isFixedLandmark(type_vec_[i])) ++count;
    return count;
  }
  inline FloatType getSeedDepth(size_t idx) const { return seed::getDepth(invmu_sigma2_a_b_vec_.col(idx)); }
  inline Position getSeedPosInFrame(size_t idx) const { return f_vec_.col(idx) * getSeedDepth(idx); }
  inline size_t numFeatures() const { return num_features_; }
  inline const cv::Mat& img() const { return img_pyr_[0]; }
  inline int id() const { return id_; }
  inline const Transformation& T_imu_cam() const { return T_body_cam_; }
  inline const Transformation& T_cam_imu() const { return T_cam_body_; }
  inline const Transformation& T_cam_world() const { return T_f_w_; }
  inline Transformation T_world_cam() const { return T_f_w_.inverse(); }
  inline Transformation T_world_imu() const { return (T_imu_cam() * T_f_w_).inverse(); }
  inline Transformation T_imu_world() const { return T_imu_cam() * T_f_w_; }
  inline void set_T_cam_imu(const Transformation& T_cam_imu) {
    T_cam_body_ = T_cam_imu;
    T_body_cam_ = T_cam_imu.inverse();
  }
  inline void set_T_w_imu(const Transformation& T_w_imu) { T_f_w_ = (T_w_imu * T_body_cam_).inverse(); }
  inline const CameraPtr& cam() const { return cam_; }
  inline Eigen::Vector3d pos() const { return T_world_cam().getPosition(); }
  inline Eigen::Vector3d imuPos() const { return T_world_imu().getPosition(); }
  double getErrorMultiplier() const { return cam_->errorMultiplier(); }
  double getAngleError(double img_err) const { return cam_->getAngleError(img_err); }

  // frame.cpp:229-257
  bool isVisible(const Eigen::Vector3d& xyz_w, Eigen::Vector2d* px = nullptr) const {
    Eigen::Vector3d xyz_f = T_f_w_ * xyz_w;
    if (cam()->getType() == Camera::Type::kPinhole) {
      if (xyz_f.z() < 0.0) return false;  // point is behind the camera
      Eigen::Vector2d px_top_left(0.0, 0.0);
      Eigen::Vector3d f_top_left;
      cam()->backProject3(px_top_left, &f_top_left);
      f_top_left.normalize();
      const Eigen::Vector3d z(0.0, 0.0, 1.0);
      const double min_cos_in_cam = f_top_left.dot(z);
      const double cur_cos_angle = xyz_f.normalized().dot(z);
      if (cur_cos_angle < min_cos_in_cam) return false;
    }
    if (px != nullptr) return cam()->project3(xyz_f, px).isKeypointVisible();
    Eigen::Vector2d px_temp;
    return cam()->project3(xyz_f, &px_temp).isKeypointVisible();
  }

  // frame.h:342-357
  inline static void jacobian_xyz2uv_imu(const Transformation& T_cam_imu, const Eigen::Vector3d& p_in_imu,
                                         Eigen::Matrix<double, 2, 6>& J) {
    Eigen::Matrix<double, 3, 6> G_x;
    G_x.block<3, 3>(0, 0) = Eigen::Matrix3d::Identity();
    G_x.block<3, 3>(0, 3) = -vk::skew(p_in_imu);
    const Eigen::Vector3d p_in_cam = T_cam_imu * p_in_imu;
    Eigen::Matrix<double, 2, 3> J_proj;
    J_proj << 1, 0, -p_in_cam[0] / p_in_cam[2], 0, 1, -p_in_cam[1] / p_in_cam[2];
    J = -1.0 / p_in_cam[2] * J_proj * T_cam_imu.getRotation().getRotationMatrix() * G_x;
  }
  // frame.h:360-371
  inline static void jacobian_xyz2img_imu(const Transformation& T_cam_imu, const Eigen::Vector3d& p_in_imu,
                                          const Eigen::Matrix<double, 2, 3>& J_cam, Eigen::Matrix<double, 2, 6>& J) {
    Eigen::Matrix<double, 3, 6> G_x;
    G_x.block<3, 3>(0, 0) = Eigen::Matrix3d::Identity();
    G_x.block<3, 3>(0, 3) = -vk::skew(p_in_imu);
    J = J_cam * T_cam_imu.getRotation().getRotationMatrix() * G_x;
  }
  // frame.h:374-397
  inline static void jacobian_xyz2f_imu(const Transformation& T_cam_imu, const Eigen::Vector3d& p_in_imu, Eigen::Matrix<double, 3, 6>& J) {
    Eigen::Matrix<double, 3, 6> G_x;
    G_x.block<3, 3>(0, 0) = Eigen::Matrix3d::Identity();
    G_x.block<3, 3>(0, 3) = -vk::skew(p_in_imu);
    const Eigen::Vector3d p_in_cam = T_cam_imu * p_in_imu;
    Eigen::Matrix<double, 3, 3> J_normalize;
    double x2 = p_in_cam[0] * p_in_cam[0];
    double y2 = p_in_cam[1] * p_in_cam[1];
    double z2 = p_in_cam[2] * p_in_cam[2];
    double xy = p_in_cam[0] * p_in_cam[1];
    double yz = p_in_cam[1] * p_in_cam[2];
    double zx = p_in_cam[2] * p_in_cam[0];
    J_normalize << y2 + z2, -xy, -zx, -xy, x2 + z2, -yz, -zx, -yz, x2 + y2;
    J_normalize *= 1 / std::pow(x2 + y2 + z2, 1.5);
    J = J_normalize * T_cam_imu.getRotationMatrix() * G_x;
  }
  // frame.cpp:274-290
  static void jacobian_xyz2image_imu(const Camera& cam, const Transformation& T_cam_imu, const Eigen::Vector3d& p_in_imu,
                                     Eigen::Matrix<double, 2, 6>& J) {
    Eigen::Matrix<double, 3, 6> G_x;
    G_x.block<3, 3>(0, 0) = Eigen::Matrix3d::Identity();
    G_x.block<3, 3>(0, 3) = -vk::skew(p_in_imu);
    const Eigen::Vector3d p_in_cam = T_cam_imu * p_in_imu;
    Eigen::Matrix<double, 2, 3> J_proj;
    Eigen::Vector2d out_point;
    cam.project3(p_in_cam, &out_point, &J_proj);
    J = J_proj * T_cam_imu.getRotation().getRotationMatrix() * G_x;
  }
};

#ifndef SVO_SHIM_REAL_POINT  // libpoint_ref.so compiles the reference's own point.h / point.cpp, which define these themselves
inline KeypointIdentifier::KeypointIdentifier(const FramePtr& _frame, const size_t _feature_index)  // point.cpp:19-23
    : frame(_frame), frame_id(_frame->id_), keypoint_index_(_feature_index) {}

// point.cpp:83-129
inline void Point::addObservation(const FramePtr& frame, const size_t feature_index) {  // point.cpp:41-58
  CHECK_NOTNULL(frame.get());
  const auto id = frame->id();
  auto it = std::find_if(obs_.begin(), obs_.end(), [&](const KeypointIdentifier& i) { return i.frame_id == id; });
  if (it == obs_.end()) obs_.emplace_back(KeypointIdentifier(frame, feature_index));
  else CHECK_EQ(it->keypoint_index_, feature_index);
}
inline bool Point::getCloseViewObs(const Eigen::Vector3d& framepos, FramePtr& ref_frame, size_t& ref_feature_index) const {
  double min_cos_angle = 0.0;
  Eigen::Vector3d obs_dir(framepos - pos_);
  obs_dir.normalize();
  for (const KeypointIdentifier& obs : obs_) {
    if (FramePtr frame = obs.frame.lock()) {
      Eigen::Vector3d dir(frame->pos() - pos_);
      dir.normalize();
      const double cos_angle = obs_dir.dot(dir);
      if (cos_angle > min_cos_angle) {
        min_cos_angle = cos_angle;
        ref_frame = frame;
        ref_feature_index = obs.keypoint_index_;
      }
    } else {
      return false;
    }
  }
  if (min_cos_angle < 0.4) return false;  // observations more than 60 degrees away are useless
  return true;
}
#endif  // SVO_SHIM_REAL_POINT

class FrameBundle {
 public:
  typedef std::shared_ptr<FrameBundle> Ptr;
  typedef std::vector<FramePtr> FrameList;
  explicit FrameBundle(const std::vector<FramePtr>& frames) : frames_(frames) {}
  inline const FramePtr& at(size_t i) const { return frames_.at(i); }
  inline size_t size() const { return frames_.size(); }
  inline bool empty() const { return frames_.empty(); }
  size_t numFeatures() const { size_t n = 0; for (const FramePtr& f : frames_) n += f->numFeatures(); return n; }  // frame.cpp:313-319
  Transformation get_T_W_B() const { return frames_[0]->T_world_imu(); }
  void set_T_W_B(const Transformation& T_W_B) {
    for (const FramePtr& frame : frames_) frame->T_f_w_ = (T_W_B * frame->T_body_cam_).inverse();
  }
  FrameList frames_;
};
using FrameBundlePtr = std::shared_ptr<FrameBundle>;

namespace frame_utils {
// frame.cpp:427-439: back-project every keypoint and normalise the bearing vectors
inline void computeNormalizedBearingVectors(const Keypoints& px_vec, const Camera& cam, Bearings* f_vec) {
  std::vector<bool> success;
  cam.backProject3(px_vec, f_vec, &success);
  for (const bool s : success) CHECK(s);
  for (int i = 0; i < f_vec->cols(); ++i) {
    const Eigen::Vector3d f = f_vec->col(i);
    f_vec->col(i) = f / f.norm();
  }
}
}  // namespace frame_utils

}  // namespace svo
