// Test driver for the svo::Reprojector facade (svo_pro_universal_b200/host/svo_b200.h): reads the flat map tables written by
// tests/test_gpu_host_facade.py, rebuilds svo::Frame / svo::Point objects (landmarks with their observation lists, seeds),
// runs Reprojector::reprojectFrames and writes what the call appended to the current frame, the grid, the statistics and the
// mutated landmark counters / seed states back as raw doubles. The Python test compares them with the outputs of the
// REFERENCE's own Reprojector::reprojectFrames (tests/golden/reproject_ref_golden.npz).
#include <cstdio>
#include <fstream>
#include <iostream>
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
  if (argc < 3) { std::fprintf(stderr, "usage: reproject_driver in.bin out.bin\n"); return 2; }
  try {
    std::ifstream in(argv[1], std::ios::binary);
    const auto hdr = rd<int32_t>(in, 12);
    const int W = hdr[0], H = hdr[1], n_levels = hdr[2], K = hdr[3], n_visible = hdr[4], NF = hdr[5], NP = hdr[6], NO = hdr[7];
    ReprojectorOptions opt;
    opt.max_n_features_per_frame = size_t(hdr[8]);
    opt.reproject_unconverged_seeds = hdr[9] != 0;
    opt.min_required_features = size_t(hdr[10]);
    opt.remove_unconstrained_points = hdr[11] != 0;
    opt.max_unconverged_seeds_ratio = rd<double>(in, 1)[0];
    const auto camv = rd<double>(in, 8);
    const auto cami = rd<int32_t>(in, 3);
    auto cam = std::make_shared<Camera>();
    cam->model = svo_camera{camv[0], camv[1], camv[2], camv[3], camv[4], camv[5], camv[6], camv[7], cami[0], cami[1], cami[2], 0};
    auto mk = [&](int id) {
      Image img(H, W);
      in.read(reinterpret_cast<char*>(img.data), size_t(W) * H);
      auto f = std::make_shared<Frame>();
      f->id_ = id;
      f->cam_ = cam;
      frame_utils::createImgPyramid(img, n_levels, f->img_pyr_, &f->gpu_);
      return f;
    };
    std::vector<FramePtr> kfs;
    for (int k = 0; k < K; ++k) kfs.push_back(mk(k + 1));
    FramePtr cur = mk(1000);
    const auto kf_T = rd<double>(in, size_t(K) * 7), cur_T = rd<double>(in, 7), mu_range = rd<double>(in, K);
    const auto begin = rd<int32_t>(in, K + 1);
    const auto px = rd<double>(in, size_t(NF) * 2), fv = rd<double>(in, size_t(NF) * 3), grad = rd<double>(in, size_t(NF) * 2);
    const auto type = rd<int32_t>(in, NF), level = rd<int32_t>(in, NF);
    const auto score = rd<double>(in, NF), state = rd<double>(in, size_t(NF) * 4);
    const auto point = rd<int32_t>(in, NF);
    const auto pt_pos = rd<double>(in, size_t(NP) * 3);
    const auto pt_failed = rd<int32_t>(in, NP), pt_succ = rd<int32_t>(in, NP), obs_begin = rd<int32_t>(in, NP + 1), obs_feat = rd<int32_t>(in, NO);
    cur->T_f_w_ = Transformation::fromArray(cur_T.data());
    std::vector<PointPtr> pts(NP);
    for (int p = 0; p < NP; ++p) {
      pts[p] = std::make_shared<Point>();
      pts[p]->id_ = p;
      pts[p]->pos_ = {pt_pos[3 * p], pt_pos[3 * p + 1], pt_pos[3 * p + 2]};
      pts[p]->n_failed_reproj_ = pt_failed[p];
      pts[p]->n_succeeded_reproj_ = pt_succ[p];
    }
    std::vector<int> feat_kf(NF);
    for (int k = 0; k < K; ++k) {
      Frame& f = *kfs[k];
      f.T_f_w_ = Transformation::fromArray(kf_T.data() + 7 * k);
      f.seed_mu_range_ = mu_range[k];
      for (int i = begin[k]; i < begin[k + 1]; ++i) {
        feat_kf[i] = k;
        f.px_vec_.push_back({px[2 * i], px[2 * i + 1]});
        f.f_vec_.push_back({fv[3 * i], fv[3 * i + 1], fv[3 * i + 2]});
        f.grad_vec_.push_back({grad[2 * i], grad[2 * i + 1]});
        f.type_vec_.push_back(FeatureType(type[i]));
        f.level_vec_.push_back(level[i]);
        f.score_vec_.push_back(score[i]);
        f.depth_vec_.push_back(-1.0);
        f.invmu_sigma2_a_b_vec_.push_back({state[4 * i], state[4 * i + 1], state[4 * i + 2], state[4 * i + 3]});
        f.landmark_vec_.push_back(point[i] >= 0 ? pts[point[i]] : nullptr);
        f.seed_ref_vec_.push_back(SeedRef());
      }
      f.num_features_ = size_t(begin[k + 1] - begin[k]);
    }
    for (int p = 0; p < NP; ++p)
      for (int o = obs_begin[p]; o < obs_begin[p + 1]; ++o) {
        const int fi = obs_feat[o], k = feat_kf[fi];
        pts[p]->obs_.emplace_back(kfs[k], size_t(fi - begin[k]));
      }

    Reprojector rp(opt, 0);
    std::vector<FramePtr> visible(kfs.begin(), kfs.begin() + n_visible);
    std::vector<PointPtr> trash;
    rp.reprojectFrames(cur, visible, trash);

    std::vector<double> out;
    out.push_back(double(cur->num_features_));
    for (size_t s = 0; s < cur->num_features_; ++s) {
      out.push_back(double(int(cur->type_vec_[s])));
      out.push_back(cur->px_vec_[s][0]); out.push_back(cur->px_vec_[s][1]);
      out.push_back(cur->level_vec_[s]);
      out.push_back(cur->landmark_vec_[s] ? cur->landmark_vec_[s]->id() : -1);
      int seed_feat = -1;
      if (cur->seed_ref_vec_[s].keyframe) seed_feat = begin[cur->seed_ref_vec_[s].keyframe->id_ - 1] + cur->seed_ref_vec_[s].seed_id;
      out.push_back(seed_feat);
      for (int c = 0; c < 4; ++c) out.push_back(cur->invmu_sigma2_a_b_vec_[s][c]);
      for (int c = 0; c < 3; ++c) out.push_back(cur->f_vec_[s][c]);
      out.push_back(cur->grad_vec_[s][0]); out.push_back(cur->grad_vec_[s][1]);
      out.push_back(cur->score_vec_[s]);
    }
    out.push_back(double(rp.grid_->occupancy_.size()));
    for (uint8_t o : rp.grid_->occupancy_) out.push_back(o);
    out.push_back(double(rp.stats_.n_trials)); out.push_back(double(rp.stats_.n_matches)); out.push_back(double(trash.size()));
    for (int p = 0; p < NP; ++p) { out.push_back(pts[p]->n_failed_reproj_); out.push_back(pts[p]->n_succeeded_reproj_); }
    for (int i = 0; i < NF; ++i) {
      const Frame& f = *kfs[feat_kf[i]];
      const int j = i - begin[feat_kf[i]];
      for (int c = 0; c < 4; ++c) out.push_back(f.invmu_sigma2_a_b_vec_[j][c]);
      out.push_back(double(int(f.type_vec_[j])));
    }
    std::ofstream of(argv[2], std::ios::binary);
    of.write(reinterpret_cast<const char*>(out.data()), sizeof(double) * out.size());
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "reproject_driver: %s\n", e.what());
    return 1;
  }
}
