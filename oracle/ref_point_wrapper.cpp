// ORACLE — TEST INFRASTRUCTURE ONLY.
// C wrapper around the REFERENCE's own svo::Point::optimize, compiled from where it lies under /root/reference
// (src/svo_common/include/svo/common/point.h, src/svo_common/src/point.cpp) into oracle/_ref/libpoint_ref.so by oracle/Makefile.
// svo::Frame is the reduced class of oracle/shim/svo_fake (pose + bearing vectors are all Point::optimize reads from it).
#include <svo/common/point.h>
#include <svo/common/frame.h>
#include <vector>

extern "C" int ref_point_optimize(int n_obs, const double* T_f_w, const double* f, double pos[3], int n_iter, int using_bearing_vector) {
  using svo::Transformation;
  std::vector<svo::FramePtr> frames;
  svo::Point pt(Eigen::Vector3d(pos[0], pos[1], pos[2]));
  for (int i = 0; i < n_obs; ++i) {
    auto fr = std::make_shared<svo::Frame>();
    fr->id_ = i + 1;
    const double* a = T_f_w + 7 * i;
    fr->T_f_w_ = Transformation(svo::Quaternion(a[0], a[1], a[2], a[3]), Eigen::Vector3d(a[4], a[5], a[6]));
    fr->resizeFeatureStorage(1);
    fr->num_features_ = 1;
    fr->f_vec_.col(0) = Eigen::Vector3d(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
    frames.push_back(fr);
    pt.addObservation(fr, 0);
  }
  pt.optimize((size_t)n_iter, using_bearing_vector != 0);
  for (int k = 0; k < 3; ++k) pos[k] = pt.pos_[k];
  return 0;
}
