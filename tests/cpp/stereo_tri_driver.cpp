// Test driver for the svo::StereoTriangulation facade (svo_pro_universal_b200/host/svo_b200.h): reads one synthetic stereo pair
// written by tests/test_gpu_host_facade.py (two level-0 images, camera, extrinsics, pose, options, srand seed), runs
// makeDetector + StereoTriangulation::compute and writes what it left in both frames back as raw doubles. The Python test compares
// them with what the REFERENCE's own compiled StereoTriangulation::compute left there (tests/golden/stereo_tri_ref_golden.npz).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "../../svo_pro_universal_b200/host/svo_b200.h"

using namespace svo;

template <class T>
static std::vector<T> rd(std::ifstream& f, size_t n) {
  std::vector<T> v(n);
  f.read(reinterpret_cast<char*>(v.data()), sizeof(T) * n);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: stereo_tri_driver in.bin out.bin\n"); return 2; }
  try {
    std::ifstream in(argv[1], std::ios::binary);
    const auto hdr = rd<int32_t>(in, 6);
    const int n_levels = hdr[0], detector_type = hdr[1], triangulate_n = hdr[2], w = hdr[4], h = hdr[5];
    const unsigned seed = unsigned(hdr[3]);
    const auto camv = rd<double>(in, 8);
    const auto cami = rd<int32_t>(in, 3);
    const auto T_cam_imu0 = rd<double>(in, 7), T_cam_imu1 = rd<double>(in, 7), T_imu_world = rd<double>(in, 7), dinv = rd<double>(in, 3);
    auto img0 = rd<uint8_t>(in, size_t(w) * h), img1 = rd<uint8_t>(in, size_t(w) * h);
    auto cam = std::make_shared<Camera>();
    cam->model = svo_camera{camv[0], camv[1], camv[2], camv[3], camv[4], camv[5], camv[6], camv[7], cami[0], cami[1], cami[2], 0};
    auto mk = [&](std::vector<uint8_t>& img, const std::vector<double>& T_cam_imu, int id) {
      auto f = std::make_shared<Frame>();
      f->id_ = id;
      f->cam_ = cam;
      Image im; im.data = img.data(); im.cols = w; im.rows = h; im.step = w;
      frame_utils::createImgPyramid(im, n_levels, f->img_pyr_, &f->gpu_);
      f->T_cam_imu_ = Transformation::fromArray(T_cam_imu.data());
      f->T_f_w_ = f->T_cam_imu_ * Transformation::fromArray(T_imu_world.data());
      return f;
    };
    FramePtr frame0 = mk(img0, T_cam_imu0, 1), frame1 = mk(img1, T_cam_imu1, 2);
    DetectorOptions o;
    o.detector_type = DetectorType(detector_type);
    StereoTriangulationOptions so;
    so.triangulate_n_features = size_t(triangulate_n);
    so.mean_depth_inv = dinv[0]; so.min_depth_inv = dinv[1]; so.max_depth_inv = dinv[2];
    StereoTriangulation st(so, feature_detection_utils::makeDetector(o, cam));
    std::srand(seed);
    st.compute(frame0, frame1);
    std::vector<double> out;
    out.push_back(double(frame0->num_features_)); out.push_back(double(frame1->num_features_));
    for (size_t i = 0; i < frame0->num_features_; ++i) {
      out.push_back(frame0->px_vec_[i][0]); out.push_back(frame0->px_vec_[i][1]); out.push_back(double(int(frame0->type_vec_[i])));
      out.push_back(frame0->landmark_vec_[i] ? 1.0 : 0.0);
    }
    for (size_t i = 0; i < frame1->num_features_; ++i) {
      out.push_back(frame1->px_vec_[i][0]); out.push_back(frame1->px_vec_[i][1]);
      for (int k = 0; k < 3; ++k) out.push_back(frame1->f_vec_[i][k]);
      out.push_back(frame1->grad_vec_[i][0]); out.push_back(frame1->grad_vec_[i][1]);
      out.push_back(frame1->level_vec_[i]); out.push_back(double(int(frame1->type_vec_[i]))); out.push_back(frame1->score_vec_[i]);
      const PointPtr& p = frame1->landmark_vec_[i];
      for (int k = 0; k < 3; ++k) out.push_back(p->pos_[k]);
      out.push_back(double(p->obs_.at(0).keypoint_index_));
      out.push_back(double(p->obs_.size()));
    }
    // FrameHandlerBase::optimizeStructure on the bundle: every new landmark (2 observations) through Point::optimize, 5 iterations
    auto bundle = std::make_shared<FrameBundle>();
    bundle->frames_ = {frame0, frame1};
    optimizeStructure(bundle, -1, 5);
    for (size_t i = 0; i < frame1->num_features_; ++i) {
      const PointPtr& p = frame1->landmark_vec_[i];
      for (int k = 0; k < 3; ++k) out.push_back(p->pos_[k]);
      out.push_back(double(p->last_structure_optim_));
    }
    std::ofstream of(argv[2], std::ios::binary);
    of.write(reinterpret_cast<const char*>(out.data()), sizeof(double) * out.size());
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "stereo_tri_driver: %s\n", e.what());
    return 1;
  }
}
