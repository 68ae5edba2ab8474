// TEST INFRASTRUCTURE ONLY (oracle/): the out-of-line members of vk::cameras::CameraGeometryBase that the direct front-end
// needs, restated from src/vikit/vikit_cameras/src/camera_geometry_base.cpp:15-19, 46-59, 61-66, 77-87. The reference file
// itself cannot be compiled here because it also holds the YAML loader (yaml-cpp is not installed); loading from YAML or
// from a mask file aborts.
#include <cstdlib>
#include <vikit/cameras/camera_geometry_base.h>

namespace vk {
namespace cameras {

CameraGeometryBase::CameraGeometryBase(const int width, const int height) : width_(width), height_(height) {}

CameraGeometryBase::Ptr CameraGeometryBase::loadFromYaml(const std::string&) { std::abort(); }
void CameraGeometryBase::loadMask(const std::string&) { std::abort(); }

void CameraGeometryBase::backProject3(const Eigen::Ref<const Eigen::Matrix2Xd>& keypoints, Eigen::Matrix3Xd* out_bearing_vectors,
                                      std::vector<bool>* success) const {
  const int num_keypoints = keypoints.cols();
  out_bearing_vectors->resize(Eigen::NoChange, num_keypoints);
  success->resize(num_keypoints);
  for (int i = 0; i < num_keypoints; ++i) {
    Eigen::Vector3d bearing_vector;
    (*success)[i] = backProject3(keypoints.col(i), &bearing_vector);
    out_bearing_vectors->col(i) = bearing_vector;
  }
}

void CameraGeometryBase::setMask(const cv::Mat& mask) {
  CHECK_EQ(height_, mask.rows);
  CHECK_EQ(width_, mask.cols);
  mask_ = mask;
}

bool CameraGeometryBase::isMasked(const Eigen::Ref<const Eigen::Vector2d>& keypoint) const {
  return keypoint[0] < 0.0 || keypoint[0] >= static_cast<double>(width_) || keypoint[1] < 0.0 ||
         keypoint[1] >= static_cast<double>(height_) ||
         (!mask_.empty() && mask_.at<uint8_t>(static_cast<int>(keypoint[1]), static_cast<int>(keypoint[0])) == 0);
}

Eigen::Vector2d CameraGeometryBase::createRandomKeypoint() const { std::abort(); }

}  // namespace cameras
}  // namespace vk
