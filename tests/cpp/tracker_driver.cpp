// Test driver for the svo::FeatureTracker facade (svo_pro_universal_b200/host/svo_b200.h): reads a mono image sequence written by
// tests/test_gpu_host_facade.py, runs FeatureTracker::trackAndDetect on every frame and writes each frame's columns (px, track id,
// score), the active / terminated track counts and the median disparity back as raw doubles. The Python test compares them with what
// the REFERENCE's own compiled FeatureTracker left there (tests/golden/tracker_ref_golden.npz).
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
  if (argc < 3) { std::fprintf(stderr, "usage: tracker_driver in.bin out.bin\n"); return 2; }
  try {
    std::ifstream in(argv[1], std::ios::binary);
    const auto hdr = rd<int32_t>(in, 8);
    const int n_frames = hdr[0], detector_type = hdr[1], min_tracks = hdr[2], reset = hdr[3], first = hdr[4], w = hdr[5], h = hdr[6], n_levels = hdr[7];
    const auto camv = rd<double>(in, 8);
    const auto cami = rd<int32_t>(in, 3);
    auto cam = std::make_shared<Camera>();
    cam->model = svo_camera{camv[0], camv[1], camv[2], camv[3], camv[4], camv[5], camv[6], camv[7], cami[0], cami[1], cami[2], 0};
    FeatureTrackerOptions to;
    to.min_tracks_to_detect_new_features = size_t(min_tracks);
    to.reset_before_detection = reset != 0;
    to.klt_template_is_first_observation = first != 0;
    DetectorOptions o;
    o.detector_type = DetectorType(detector_type);
    FeatureTracker tracker(to, o, {cam});
    std::vector<double> out;
    std::vector<FrameBundle::Ptr> keep;
    std::vector<std::vector<uint8_t>> imgs;
    int first_id = -1;
    for (int k = 0; k < n_frames; ++k) {
      imgs.push_back(rd<uint8_t>(in, size_t(w) * h));
      auto f = std::make_shared<Frame>();
      f->id_ = k + 1;
      f->cam_ = cam;
      Image im; im.data = imgs.back().data(); im.cols = w; im.rows = h; im.step = w;
      frame_utils::createImgPyramid(im, n_levels, f->img_pyr_, &f->gpu_);
      auto bundle = std::make_shared<FrameBundle>();
      bundle->frames_ = {f};
      keep.push_back(bundle);
      tracker.trackAndDetect(bundle);
      out.push_back(double(f->num_features_));
      for (size_t i = 0; i < f->num_features_; ++i) {
        if (first_id < 0) first_id = f->track_id_vec_[i];
        out.push_back(f->px_vec_[i][0]); out.push_back(f->px_vec_[i][1]);
        out.push_back(double(f->track_id_vec_[i] - first_id)); out.push_back(f->score_vec_[i]);
      }
      std::vector<size_t> nt; std::vector<double> disp;
      tracker.getNumTrackedAndDisparityPerFrame(0.5, &nt, &disp);
      out.push_back(double(tracker.getTotalActiveTracks())); out.push_back(double(tracker.terminated_tracks_.at(0).size())); out.push_back(disp.at(0));
    }
    std::ofstream of(argv[2], std::ios::binary);
    of.write(reinterpret_cast<const char*>(out.data()), sizeof(double) * out.size());
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "tracker_driver: %s\n", e.what());
    return 1;
  }
}
