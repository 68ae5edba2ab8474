// Test driver for the C++ facades (svo_pro_universal_b200/host/svo_b200.h): reads one synthetic frame pair written by
// tests/test_gpu_host_facade.py, runs the reference-shaped calls (createImgPyramid, FastDetector::detect,
// SparseImgAlign::run, Matcher::findMatchDirect / findEpipolarMatchDirect, DepthFilter::updateSeeds) and writes the results
// back as raw doubles. The Python test compares them with the oracle.
#include <algorithm>
#include <cstdio>
#include <mutex>
#include <stdexcept>
#include <cstdlib>
#include <fstream>
#include <cstring>
#include <iostream>
#include <thread>
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
  if (argc < 3) { std::fprintf(stderr, "usage: facade_driver in.bin out.bin\n"); return 2; }
  try {
    std::ifstream in(argv[1], std::ios::binary);
    const auto hdr = rd<int32_t>(in, 4);
    const int W = hdr[0], H = hdr[1], n_levels = hdr[2], N = hdr[3];
    Image ref0(H, W), cur0(H, W);
    in.read(reinterpret_cast<char*>(ref0.data), size_t(W) * H);
    in.read(reinterpret_cast<char*>(cur0.data), size_t(W) * H);
    const auto camv = rd<double>(in, 8);
    const auto cami = rd<int32_t>(in, 3);
    const auto T_cam_imu = rd<double>(in, 7), T_ref = rd<double>(in, 7), T_cur_init = rd<double>(in, 7), T_cur_true = rd<double>(in, 7);
    const auto px = rd<double>(in, 2 * N), fv = rd<double>(in, 3 * N), depth = rd<double>(in, N), grad = rd<double>(in, 2 * N),
               guess = rd<double>(in, 2 * N);
    const auto type = rd<int32_t>(in, N), level = rd<int32_t>(in, N);

    auto cam = std::make_shared<Camera>();
    cam->model = svo_camera{camv[0], camv[1], camv[2], camv[3], camv[4], camv[5], camv[6], camv[7], cami[0], cami[1], cami[2], 0};
    auto mk = [&](const Image& img, const std::vector<double>& T_f_w, int id) {
      auto f = std::make_shared<Frame>();
      f->id_ = id;
      f->cam_ = cam;
      frame_utils::createImgPyramid(img, n_levels, f->img_pyr_, &f->gpu_);
      f->T_f_w_ = Transformation::fromArray(T_f_w.data());
      f->T_cam_imu_ = Transformation::fromArray(T_cam_imu.data());
      return f;
    };
    FramePtr ref = mk(ref0, T_ref, 1), cur = mk(cur0, T_cur_init, 2);
    std::vector<double> out;

    // (a) FastDetector::detect on the ref frame
    {
      FastDetector det(DetectorOptions(), cam);
      det.detect(ref);
      out.push_back(double(ref->num_features_));
      for (size_t i = 0; i < ref->num_features_; ++i) {
        out.push_back(ref->px_vec_[i][0]); out.push_back(ref->px_vec_[i][1]); out.push_back(ref->score_vec_[i]); out.push_back(ref->level_vec_[i]);
      }
      double cs = 0;  // checksum of the pyramid mirrored back to the host
      for (const Image& im : ref->img_pyr_) for (int y = 0; y < im.rows; ++y) for (int x = 0; x < im.cols; ++x) cs += im.data[y * im.step + x] * double((x + 3 * y) % 7 + 1);
      out.push_back(cs);
      ref->clearFeatureStorage();
    }
    // (a') makeDetector(kFastGrad) / makeDetector(kGridGrad) ->detect on the ref frame (the reference's default detector)
    for (DetectorType t : {DetectorType::kFastGrad, DetectorType::kGridGrad}) {
      DetectorOptions o;
      o.detector_type = t;
      AbstractDetector::Ptr det = feature_detection_utils::makeDetector(o, cam);
      det->detect(ref);
      out.push_back(double(ref->num_features_));
      for (size_t i = 0; i < ref->num_features_; ++i) {
        out.push_back(ref->px_vec_[i][0]); out.push_back(ref->px_vec_[i][1]); out.push_back(ref->score_vec_[i]); out.push_back(ref->level_vec_[i]);
        out.push_back(double(int(ref->type_vec_[i]))); out.push_back(ref->grad_vec_[i][0]); out.push_back(ref->grad_vec_[i][1]);
        out.push_back(ref->f_vec_[i][2]);
      }
      ref->clearFeatureStorage();
    }
    // (d4) DepthFilter::addKeyframe -> depth_filter_utils::initializeSeeds on the ref frame (FastGrad detector, at most 200 seeds)
    {
      DetectorOptions o;
      o.detector_type = DetectorType::kFastGrad;
      DepthFilter df(DepthFilterOptions(), o, cam);
      df.addKeyframe(ref, 3.0, 1.0, 10.0);
      out.push_back(double(ref->num_features_));
      out.push_back(ref->seed_mu_range_);
      for (size_t i = 0; i < ref->num_features_; ++i) {
        out.push_back(ref->px_vec_[i][0]); out.push_back(ref->px_vec_[i][1]); out.push_back(ref->score_vec_[i]); out.push_back(ref->level_vec_[i]);
        out.push_back(double(int(ref->type_vec_[i])));
        for (int k = 0; k < 4; ++k) out.push_back(ref->invmu_sigma2_a_b_vec_[i][k]);
      }
      df.addKeyframe(ref, 3.0, 1.0, 10.0);  // already 200 features: nothing is added
      out.push_back(double(ref->num_features_));
      ref->clearFeatureStorage();
    }
    // features of the test set
    for (int i = 0; i < N; ++i) {
      ref->px_vec_.push_back({px[2 * i], px[2 * i + 1]});
      ref->f_vec_.push_back({fv[3 * i], fv[3 * i + 1], fv[3 * i + 2]});
      ref->grad_vec_.push_back({grad[2 * i], grad[2 * i + 1]});
      ref->score_vec_.push_back(0);
      ref->level_vec_.push_back(level[i]);
      ref->type_vec_.push_back(FeatureType(type[i]));
      ref->depth_vec_.push_back(depth[i]);
      ref->invmu_sigma2_a_b_vec_.push_back({0.25, (1 / 1.5) * (1 / 1.5) / 36.0, 10.0, 10.0});
    }
    ref->num_features_ = N;
    ref->seed_mu_range_ = 1 / 1.5;

    // (b) SparseImgAlign::run
    {
      SparseImgAlign align(SparseImgAlign::getDefaultSolverOptions(), SparseImgAlignOptions());
      auto rb = std::make_shared<FrameBundle>(), cb = std::make_shared<FrameBundle>();
      rb->frames_ = {ref}; cb->frames_ = {cur};
      align.reset();
      const size_t n = align.run(rb, cb);
      out.push_back(double(n));
      double T[7]; cur->T_f_w_.toArray(T);
      for (double v : T) out.push_back(v);
      out.push_back(align.getError());
    }
    // (c) Matcher with the true relative pose
    cur->T_f_w_ = Transformation::fromArray(T_cur_true.data());
    const size_t seq_matcher_at = out.size();
    {
      Matcher m;
      for (int i = 0; i < N; ++i) {
        FeatureWrapper fw{FeatureType(type[i]), ref->px_vec_[i], ref->f_vec_[i], ref->grad_vec_[i], level[i]};
        Keypoint pc{guess[2 * i], guess[2 * i + 1]};
        const auto r = m.findMatchDirect(*ref, *cur, fw, depth[i], pc);
        out.push_back(double(int(r))); out.push_back(pc[0]); out.push_back(pc[1]);
        double d = 0;
        const auto r2 = m.findEpipolarMatchDirect(*ref, *cur, fw, 1.0 / depth[i], 1.3 / depth[i], 0.7 / depth[i], d);
        out.push_back(double(int(r2))); out.push_back(d); out.push_back(m.px_cur_[0]); out.push_back(m.px_cur_[1]);
      }
    }
    // (d) DepthFilter::updateSeeds: every feature becomes a seed of its kind
    {
      for (int i = 0; i < N; ++i) ref->type_vec_[i] = (FeatureType(type[i]) == FeatureType::kEdgelet) ? FeatureType::kEdgeletSeed : FeatureType::kCornerSeed;
      DepthFilterOptions o;
      o.scan_epi_unit_sphere = true;
      DepthFilter df(o);
      const size_t n = df.updateSeeds({ref}, cur);
      out.push_back(double(n));
      for (int i = 0; i < N; ++i) {
        for (double v : ref->invmu_sigma2_a_b_vec_[i]) out.push_back(v);
        out.push_back(double(int(ref->type_vec_[i])));
      }
    }
    // (e) two host threads on frames that have NO device copy yet: both race into the lazy upload (b200::ensureGpu), each thread
    // runs on its own context / stream; every result must equal the sequential results of (c)
    {
      auto bare = [&](const FramePtr& src) {
        auto f = std::make_shared<Frame>();
        f->id_ = src->id_ + 10;
        f->cam_ = cam;
        for (const Image& im : src->img_pyr_) {
          Image c(im.rows, im.cols);
          for (int y = 0; y < im.rows; ++y) std::memcpy(c.data + size_t(y) * c.step, im.data + size_t(y) * im.step, size_t(im.cols));
          f->img_pyr_.push_back(std::move(c));
        }
        for (Image& im : f->img_pyr_) im.data = im.storage.data();
        f->T_f_w_ = src->T_f_w_;
        f->T_cam_imu_ = src->T_cam_imu_;
        return f;
      };
      FramePtr ref2 = bare(ref), cur2 = bare(cur);
      std::vector<double> ra(3 * size_t(N)), rb(4 * size_t(N));
      std::string err;
      std::thread ta([&] {
        try {
          Matcher m;
          for (int i = 0; i < N; ++i) {
            FeatureWrapper fw{FeatureType(type[i]), ref->px_vec_[i], ref->f_vec_[i], ref->grad_vec_[i], level[i]};
            Keypoint pc{guess[2 * i], guess[2 * i + 1]};
            const auto r = m.findMatchDirect(*ref2, *cur2, fw, depth[i], pc);
            ra[3 * i] = double(int(r)); ra[3 * i + 1] = pc[0]; ra[3 * i + 2] = pc[1];
          }
        } catch (const std::exception& e) { err = e.what(); }
      });
      std::thread tb([&] {
        try {
          Matcher m;
          for (int i = 0; i < N; ++i) {
            FeatureWrapper fw{FeatureType(type[i]), ref->px_vec_[i], ref->f_vec_[i], ref->grad_vec_[i], level[i]};
            double d = 0;
            const auto r2 = m.findEpipolarMatchDirect(*ref2, *cur2, fw, 1.0 / depth[i], 1.3 / depth[i], 0.7 / depth[i], d);
            rb[4 * i] = double(int(r2)); rb[4 * i + 1] = d; rb[4 * i + 2] = m.px_cur_[0]; rb[4 * i + 3] = m.px_cur_[1];
          }
        } catch (const std::exception& e) { err = e.what(); }
      });
      ta.join(); tb.join();
      if (!err.empty()) throw std::runtime_error("two-thread section: " + err);
      size_t mismatches = 0;
      for (int i = 0; i < N; ++i) {
        const double* s7 = &out[seq_matcher_at + 7 * size_t(i)];
        for (int k = 0; k < 3; ++k) mismatches += ra[3 * i + k] != s7[k];
        for (int k = 0; k < 4; ++k) mismatches += rb[4 * i + k] != s7[3 + k];
      }
      out.push_back(double(mismatches));
    }
    // (f) the rest of the reference's public surface on this path
    {
      // fast:: leaves with list-shaped results (fast.h:20-41) on level 0 of the reference frame
      const Image& l0 = ref->img_pyr_[0];
      std::vector<fast::fast_xy> corners, corners9;
      fast::fast_corner_detect_10_sse2(l0.data, l0.cols, l0.rows, int(l0.step), 10, corners);
      fast::fast_corner_detect_9(l0.data, l0.cols, l0.rows, int(l0.step), 10, corners9);
      std::vector<int> scores, nm;
      fast::fast_corner_score_10(l0.data, int(l0.step), corners, 10, scores);
      fast::fast_nonmax_3x3(corners, scores, nm);
      out.push_back(double(corners.size()));
      for (size_t i = 0; i < corners.size(); ++i) { out.push_back(corners[i].x); out.push_back(corners[i].y); out.push_back(scores[i]); }
      out.push_back(double(nm.size()));
      for (int i : nm) out.push_back(i);
      out.push_back(double(corners9.size()));
      const size_t before = corners.size();
      fast::fast_corner_detect_10(l0.data, l0.cols, l0.rows, int(l0.step), 10, corners);  // appends, as the reference's push_back
      out.push_back(double(corners.size() == 2 * before));
      // Matcher::epi_image_ / patch_ / patch_with_border_ and scanEpipolarLine on its own
      Matcher m;
      const int n_chk = std::min(N, 40);
      out.push_back(double(n_chk));
      for (int i = 0; i < n_chk; ++i) {
        FeatureWrapper fw{FeatureType(type[i]), ref->px_vec_[i], ref->f_vec_[i], ref->grad_vec_[i], level[i]};
        double d = 0;
        const auto r2 = m.findEpipolarMatchDirect(*ref, *cur, fw, 1.0 / depth[i], 1.3 / depth[i], 0.7 / depth[i], d);
        out.push_back(double(int(r2)));
        out.push_back(m.epi_image_[0]); out.push_back(m.epi_image_[1]);
        out.push_back(double(m.search_level_)); out.push_back(m.epi_length_pyramid_);
        for (int k = 0; k < 100; ++k) out.push_back(m.patch_with_border_[k]);
        for (int k = 0; k < 64; ++k) out.push_back(m.patch_[k]);
        // the scan alone, on the segment the driver's caller also forms (A, B, C follow the patch in the output)
        const Transformation T_cr = cur->T_f_w_ * ref->T_f_w_.inverse();
        BearingVector Rf;
        {  // q * f * q^-1
          const double w = T_cr.q[0], x = T_cr.q[1], y = T_cr.q[2], z = T_cr.q[3];
          const double ux = 2 * (y * fw.f[2] - z * fw.f[1]), uy = 2 * (z * fw.f[0] - x * fw.f[2]), uz = 2 * (x * fw.f[1] - y * fw.f[0]);
          Rf = {fw.f[0] + w * ux + (y * uz - z * uy), fw.f[1] + w * uy + (z * ux - x * uz), fw.f[2] + w * uz + (x * uy - y * ux)};
        }
        BearingVector A, B, C;
        for (int k = 0; k < 3; ++k) {
          A[k] = Rf[k] + T_cr.t[k] * (1.3 / depth[i]); B[k] = Rf[k] + T_cr.t[k] * (0.7 / depth[i]); C[k] = Rf[k] + T_cr.t[k] * (1.0 / depth[i]);
        }
        Matcher::PatchScore ps(m.patch_);
        Keypoint best{0, 0};
        int z = Matcher::PatchScore::threshold();
        m.scanEpipolarLine(*cur, A, B, C, ps, m.search_level_, &best, &z);
        for (int k = 0; k < 3; ++k) { out.push_back(A[k]); out.push_back(B[k]); out.push_back(C[k]); }
        out.push_back(best[0]); out.push_back(best[1]); out.push_back(double(z));
      }
      // depth_filter_utils::updateFilterGaussian
      SeedState st{0.31, 0.004, 10.0, 10.0};
      const bool okg = depth_filter_utils::updateFilterGaussian(0.33, 0.0007, st);
      out.push_back(double(okg));
      for (double v : st) out.push_back(v);
      // SparseImgAlignBase::setPatchSize: 4 is the reference's (and the kernel's) size, anything else is refused
      SparseImgAlign align(SparseImgAlign::getDefaultSolverOptions(), SparseImgAlignOptions());
      align.setPatchSize<SparseImgAlign>(4);
      bool refused = false;
      try { align.setPatchSize<SparseImgAlign>(8); } catch (const std::invalid_argument&) { refused = true; }
      out.push_back(double(refused));
      // AbstractDetector::closeness_check_grid_ and the DepthFilter's public detector members
      DetectorOptions dopt;
      dopt.sec_grid_fineness = 2;
      FastDetector fd(dopt, cam);
      out.push_back(double(fd.grid_.size())); out.push_back(double(fd.closeness_check_grid_.size()));
      fd.closeness_check_grid_.fillWithKeypoints(Keypoint{100.0, 50.0});
      out.push_back(double(std::count(fd.closeness_check_grid_.occupancy_.begin(), fd.closeness_check_grid_.occupancy_.end(), uint8_t(1))));
      fd.resetGrid();
      out.push_back(double(std::count(fd.closeness_check_grid_.occupancy_.begin(), fd.closeness_check_grid_.occupancy_.end(), uint8_t(1))));
      DepthFilter df(DepthFilterOptions(), DetectorOptions(), cam);
      {
        std::lock_guard<std::mutex> lock(df.feature_detector_mut_);
        out.push_back(double(df.feature_detector_ != nullptr)); out.push_back(double(df.sec_feature_detector_ == nullptr));
      }
    }
    // (g) the depth filter's parallel thread (depth_filter.cpp:65-88, 145-198): addKeyframe / updateSeeds only enqueue, the worker runs the
    // same batched calls on its own context; results must equal the synchronous run of (d) bit for bit
    {
      auto clone = [&](const FramePtr& src) {
        auto f = std::make_shared<Frame>();
        f->id_ = src->id_ + 20;
        f->cam_ = cam;
        for (const Image& im : src->img_pyr_) {
          Image c(im.rows, im.cols);
          for (int y = 0; y < im.rows; ++y) std::memcpy(c.data + size_t(y) * c.step, im.data + size_t(y) * im.step, size_t(im.cols));
          f->img_pyr_.push_back(std::move(c));
        }
        for (Image& im : f->img_pyr_) im.data = im.storage.data();
        f->T_f_w_ = src->T_f_w_;
        f->T_cam_imu_ = src->T_cam_imu_;
        return f;
      };
      FramePtr ref3 = clone(ref), cur3 = clone(cur);
      for (int i = 0; i < N; ++i) {
        ref3->px_vec_.push_back(ref->px_vec_[i]); ref3->f_vec_.push_back(ref->f_vec_[i]); ref3->grad_vec_.push_back(ref->grad_vec_[i]);
        ref3->score_vec_.push_back(0); ref3->level_vec_.push_back(level[i]); ref3->depth_vec_.push_back(depth[i]);
        ref3->type_vec_.push_back((FeatureType(type[i]) == FeatureType::kEdgelet) ? FeatureType::kEdgeletSeed : FeatureType::kCornerSeed);
        ref3->invmu_sigma2_a_b_vec_.push_back({0.25, (1 / 1.5) * (1 / 1.5) / 36.0, 10.0, 10.0});
      }
      ref3->num_features_ = N;
      ref3->seed_mu_range_ = 1 / 1.5;
      DepthFilterOptions o;
      o.scan_epi_unit_sphere = true;
      DepthFilter df(o, DetectorOptions(), cam);
      df.startThread();
      df.startThread();  // "Thread already started!": no second thread
      const size_t queued = df.updateSeeds({ref3}, cur3);  // returns at once with 0, as the reference's threaded branch
      FramePtr kf = clone(cur);
      df.waitForJobs();
      size_t mismatches = 0;
      for (int i = 0; i < N; ++i) {
        for (int k = 0; k < 4; ++k) mismatches += ref3->invmu_sigma2_a_b_vec_[i][k] != ref->invmu_sigma2_a_b_vec_[i][k];
        mismatches += ref3->type_vec_[i] != ref->type_vec_[i];
      }
      out.push_back(double(queued)); out.push_back(double(mismatches));
      // a queued update is dropped by the keyframe job that follows it ("clear all other jobs, this one has priority")
      df.addKeyframe(kf, 3.0, 1.0, 10.0);
      df.waitForJobs();
      out.push_back(double(kf->num_features_));
      df.reset();
      df.stopThread();
      df.stopThread();
    }
    std::ofstream o(argv[2], std::ios::binary);
    o.write(reinterpret_cast<const char*>(out.data()), sizeof(double) * out.size());
    std::printf("facade_driver ok: %zu doubles\n", out.size());
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "facade_driver: %s\n", e.what());
    return 1;
  }
}
