// TEST INFRASTRUCTURE ONLY (oracle/): replaces the reference's svo/common/camera.h, which also pulls in NCamera and its
// YAML loader (yaml-cpp is not installed). The camera classes themselves are the reference's own headers.
#pragma once
#include <memory>
#include <vector>
#include <svo/common/camera_fwd.h>
#include <vikit/cameras/camera_geometry_base.h>
#include <vikit/cameras/camera_geometry.h>
#include <vikit/cameras/no_distortion.h>
#include <vikit/cameras/radial_tangential_distortion.h>
#include <vikit/cameras/pinhole_projection.h>

namespace vk {
namespace cameras {
// the slice of NCamera (a rig of cameras) the depth filter's constructor touches
class NCamera {
 public:
  typedef std::shared_ptr<NCamera> Ptr;
  explicit NCamera(const std::vector<std::shared_ptr<CameraGeometryBase>>& cams) : cams_(cams) {}
  std::shared_ptr<CameraGeometryBase> getCameraShared(size_t i) const { return cams_.at(i); }
  size_t getNumCameras() const { return cams_.size(); }
 private:
  std::vector<std::shared_ptr<CameraGeometryBase>> cams_;
};
}  // namespace cameras
}  // namespace vk
