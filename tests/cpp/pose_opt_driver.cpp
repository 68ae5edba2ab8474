// Test driver for the svo::PoseOptimizer facade (svo_pro_universal_b200/host/svo_b200.h): reads one synthetic frame bundle
// written by tests/test_gpu_host_facade.py (cameras, start pose, features with landmarks), runs PoseOptimizer::run and writes
// the optimised pose, the outlier marks and the statistics back as raw doubles. The Python test compares them with the outputs
// of the REFERENCE's own compiled PoseOptimizer::run (tests/golden/pose_opt_ref_golden.npz).
#include <cstdio>
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
  if (argc < 3) { std::fprintf(stderr, "usage: pose_opt_driver in.bin out.bin\n"); return 2; }
  try {
    std::ifstream in(argv[1], std::ios::binary);
    const auto hdr = rd<int32_t>(in, 4);
    const int n_cams = hdr[0], N = hdr[1], err_type = hdr[2], have_prior = hdr[3];
    const auto camv = rd<double>(in, 8);
    const auto cami = rd<int32_t>(in, 3);
    const auto T_cam_imu = rd<double>(in, size_t(n_cams) * 7), T0 = rd<double>(in, 7), prior_q = rd<double>(in, 4);
    const auto px = rd<double>(in, size_t(N) * 2), fv = rd<double>(in, size_t(N) * 3), grad = rd<double>(in, size_t(N) * 2), xyz = rd<double>(in, size_t(N) * 3);
    const auto level = rd<int32_t>(in, N), type = rd<int32_t>(in, N), feat_cam = rd<int32_t>(in, N);
    const auto has = rd<uint8_t>(in, N);
    auto cam = std::make_shared<Camera>();
    cam->model = svo_camera{camv[0], camv[1], camv[2], camv[3], camv[4], camv[5], camv[6], camv[7], cami[0], cami[1], cami[2], 0};
    auto bundle = std::make_shared<FrameBundle>();
    const Transformation T_imu_world = Transformation::fromArray(T0.data());
    std::vector<std::vector<int>> idx(n_cams);
    for (int i = 0; i < N; ++i) idx[feat_cam[i]].push_back(i);
    for (int c = 0; c < n_cams; ++c) {
      auto f = std::make_shared<Frame>();
      f->cam_ = cam;
      f->T_cam_imu_ = Transformation::fromArray(T_cam_imu.data() + 7 * c);
      f->T_f_w_ = f->T_cam_imu_ * T_imu_world;
      for (int i : idx[c]) {
        f->px_vec_.push_back({px[2 * i], px[2 * i + 1]});
        f->f_vec_.push_back({fv[3 * i], fv[3 * i + 1], fv[3 * i + 2]});
        f->grad_vec_.push_back({grad[2 * i], grad[2 * i + 1]});
        f->level_vec_.push_back(level[i]);
        f->type_vec_.push_back(FeatureType(type[i]));
        PointPtr p;
        if (has[i]) { p = std::make_shared<Point>(); p->pos_ = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}; }
        f->landmark_vec_.push_back(p);
        f->seed_ref_vec_.push_back(SeedRef());
      }
      f->num_features_ = idx[c].size();
      bundle->frames_.push_back(f);
    }
    PoseOptimizer po(PoseOptimizer::getDefaultSolverOptions());
    po.reset();
    po.setErrorType(PoseOptimizer::ErrorType(err_type));
    if (have_prior) po.setRotationPrior({prior_q[0], prior_q[1], prior_q[2], prior_q[3]}, 0.5);
    const size_t n = po.run(bundle, 2.0);
    std::vector<double> out;
    out.push_back(double(n));
    double T[7];
    bundle->at(0)->T_imu_world().toArray(T);
    out.insert(out.end(), T, T + 7);
    out.push_back(po.measurement_sigma_); out.push_back(po.stats_.reproj_error_before); out.push_back(po.stats_.reproj_error_after);
    out.push_back(double(po.iterCount()));
    std::vector<double> outl(N, 0.0);
    for (int c = 0; c < n_cams; ++c)
      for (size_t j = 0; j < idx[c].size(); ++j)
        outl[idx[c][j]] = (bundle->at(c)->type_vec_[j] == FeatureType::kOutlier && type[idx[c][j]] != int(FeatureType::kOutlier) &&
                           bundle->at(c)->landmark_vec_[j] == nullptr) ? 1.0 : 0.0;
    out.insert(out.end(), outl.begin(), outl.end());
    std::ofstream of(argv[2], std::ios::binary);
    of.write(reinterpret_cast<const char*>(out.data()), sizeof(double) * out.size());
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "pose_opt_driver: %s\n", e.what());
    return 1;
  }
}
