// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header for the rules).
//
// Row f1 of SURVEY.md §8: the Reprojector's candidate flow — reprojector_utils::getCandidate, projectPointAndCheckVisibility,
// sortCandidatesByReprojStats / sortCandidatesByNumObs, matchCandidates, matchCandidate (src/svo/src/reprojector.cpp:310-543),
// Frame::isVisible (src/svo_common/src/frame.cpp:229-257), Point::getCloseViewObs (src/svo_common/src/point.cpp:83-129).
// Parity status: pinned — the reference's own reprojector.cpp compiles into oracle/_ref/libfrontend_ref.so
// (ref_reproject_match in ref_frontend_wrapper.cpp) and this restatement is checked against it.
// The reference sorts with std::sort (order of equal candidates unspecified); this restatement uses a stable sort, i.e. equal
// candidates keep their visiting order — the pin uses candidates without ties.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#include "orc_capi.h"
#include "orc_math.hpp"
#include "orc_matcher.hpp"
#include "orc_depth_filter.hpp"

namespace orc {

// ref: src/svo_common/include/svo/common/types.h:84-130
inline bool isConvergedCornerEdgeletSeed(FeatureType t) { return t == FeatureType::kEdgeletSeedConverged || t == FeatureType::kCornerSeedConverged; }
inline bool isConvergedMapPointSeed(FeatureType t) { return t == FeatureType::kMapPointSeedConverged; }
inline bool isUnconvergedCornerEdgeletSeed(FeatureType t) { return t == FeatureType::kEdgeletSeed || t == FeatureType::kCornerSeed; }
inline bool isUnconvergedMapPointSeed(FeatureType t) { return t == FeatureType::kMapPointSeed; }

inline FeatureRef reprojFeatureOf(const orc_feature* f) {
  FeatureRef r;
  r.type = static_cast<FeatureType>(f->type);
  r.px = {f->px[0], f->px[1]};
  r.f = {f->f[0], f->f[1], f->f[2]};
  r.grad = {f->grad[0], f->grad[1]};
  r.level = f->level;
  return r;
}

struct ReprojFrame {  // the parts of svo::Frame the reprojector reads
  MatchFrame mf;
  SE3 T_f_w;
  V3 pos() const { return inverse(T_f_w).t; }  // Frame::pos(): T_world_cam().getPosition()
};

// ref: src/svo_common/src/frame.cpp:229-257
inline bool frameIsVisible(const ReprojFrame& fr, const V3& xyz_w, V2* px) {
  const V3 xyz_f = fr.T_f_w * xyz_w;
  // every camera on this path is a pinhole (cam()->getType() == kPinhole)
  if (xyz_f.z < 0.0) return false;
  V3 f_top_left = fr.mf.cam.backProject3({0.0, 0.0});
  f_top_left = normalized(f_top_left);
  const V3 z{0.0, 0.0, 1.0};
  const double min_cos_in_cam = dot(f_top_left, z);
  const double cur_cos_angle = dot(normalized(xyz_f), z);
  if (cur_cos_angle < min_cos_in_cam) return false;
  *px = fr.mf.cam.project3(xyz_f);
  return fr.mf.cam.isKeypointVisible(px->x, px->y);
}

// ref: reprojector.cpp:520-543
inline bool projectPointAndCheckVisibility(const ReprojFrame& frame, const V3& xyz, V2* px) {
  if (!frameIsVisible(frame, xyz, px)) return false;
  const int pxi0 = int(px->x), pxi1 = int(px->y);  // px->cast<int>()
  constexpr int kPatchSize = 8;
  return frame.mf.cam.isKeypointVisibleWithMarginInt(pxi0, pxi1, kPatchSize);
}

struct Candidate {  // reprojector.h:118-145
  int entry;     // which entry produced it (stands for ref_frame + ref_index)
  int feat;      // global feature index
  V2 cur_px;
  int n_reproj;
  double score;
  FeatureType type;
  size_t n_obs;
};

// OccupandyGrid2D::getCellIndex(int x, int y, 1): occupancy_grid_2d.h:82-95 (the Candidate's double pixel is truncated to int by the call)
inline size_t reprojCellIndex(const V2& cur_px, int cell_size, int n_cols) {
  const int x = int(cur_px.x), y = int(cur_px.y);
  return size_t(std::floor(double(y) / cell_size) * n_cols + std::floor(double(x) / cell_size));
}

inline void reprojectMatch(const orc_reproj_map& map, const std::vector<ReprojFrame>& kfs, const ReprojFrame& cur, int E,
                           const int* entry_feat, int n_features_in, uint8_t* occupancy, const orc_reproj_options& opt,
                           orc_reproj_result* results, orc_reproj_stats* stats) {
  const int n_cols = int(std::ceil(double(cur.mf.cam.width) / opt.cell_size));  // OccupandyGrid2D::getNCell
  std::vector<Candidate> candidates;
  // ---- getCandidate (reprojector.cpp:487-518) for every entry in visiting order
  for (int e = 0; e < E; ++e) {
    const int fi = entry_feat[e];
    orc_reproj_result& r = results[e];
    r = orc_reproj_result();
    r.status = ORC_REPROJ_NOT_CANDIDATE; r.order = -1; r.slot = -1; r.match_result = -1;
    r.type_out = map.feat[fi].type;
    for (int k = 0; k < 4; ++k) r.seed_state[k] = map.feat_seed_state[4 * size_t(fi) + k];
    const int kf = map.feat_kf[fi];
    const int pt = map.feat_point[fi];
    V3 xyz_world{0, 0, 0};
    int n_reproj = 0;
    if (pt >= 0) {
      xyz_world = {map.pt_pos[3 * pt], map.pt_pos[3 * pt + 1], map.pt_pos[3 * pt + 2]};
      n_reproj = map.pt_n_succeeded[pt] - map.pt_n_failed[pt];
    } else {
      const V3 f{map.feat[fi].f[0], map.feat[fi].f[1], map.feat[fi].f[2]};
      xyz_world = inverse(kfs[kf].T_f_w) * (f * seed::getDepth(map.feat_seed_state + 4 * size_t(fi)));
    }
    V2 px;
    if (!projectPointAndCheckVisibility(cur, xyz_world, &px)) continue;
    r.cur_px[0] = px.x; r.cur_px[1] = px.y;
    Candidate c;
    c.entry = e; c.feat = fi; c.cur_px = px; c.n_reproj = n_reproj; c.score = map.feat_score[fi];
    c.type = static_cast<FeatureType>(map.feat[fi].type);
    c.n_obs = pt >= 0 ? size_t(map.pt_obs_begin[pt + 1] - map.pt_obs_begin[pt]) : 0u;
    candidates.push_back(c);
  }
  // ---- sortCandidatesByReprojStats / sortCandidatesByNumObs (reprojector.cpp:312-341)
  if (opt.sort_by_num_obs)
    std::stable_sort(candidates.begin(), candidates.end(), [](const Candidate& lhs, const Candidate& rhs) {
      return lhs.n_obs > rhs.n_obs || (lhs.n_obs == rhs.n_obs && lhs.n_reproj > rhs.n_reproj)
             || (lhs.n_obs == rhs.n_obs && lhs.n_reproj == rhs.n_reproj && lhs.score > rhs.score);
    });
  else
    std::stable_sort(candidates.begin(), candidates.end(), [](const Candidate& lhs, const Candidate& rhs) {
      return lhs.type > rhs.type || (lhs.type == rhs.type && lhs.n_reproj > rhs.n_reproj)
             || (lhs.type == rhs.type && lhs.n_reproj == rhs.n_reproj && lhs.score > rhs.score);
    });
  for (size_t p = 0; p < candidates.size(); ++p) {
    results[candidates[p].entry].order = int(p);
    results[candidates[p].entry].status = ORC_REPROJ_NOT_REACHED;
  }
  // ---- matchCandidates (reprojector.cpp:342-381)
  Matcher matcher;
  matcher.options_.affine_est_offset_ = opt.affine_est_offset != 0;
  matcher.options_.affine_est_gain_ = opt.affine_est_gain != 0;
  const size_t max_n = size_t(opt.max_n_features);
  size_t num_features = size_t(n_features_in);
  stats->n_candidates = int(candidates.size());
  stats->n_trials = 0; stats->n_matches = 0;
  int i = 0;
  for (Candidate& c : candidates) {
    ++i;
    orc_reproj_result& r = results[c.entry];
    const size_t grid_index = reprojCellIndex(c.cur_px, opt.cell_size, n_cols);
    if (max_n > 0 && occupancy[grid_index]) { r.status = ORC_REPROJ_SKIPPED; continue; }
    ++stats->n_trials;
    r.status = ORC_REPROJ_FAILED;
    // ---- matchCandidate (reprojector.cpp:384-485)
    bool ok = false;
    V2 grad_ref{0, 0};
    const int kf = map.feat_kf[c.feat];
    const int pt = map.feat_point[c.feat];
    if (pt < 0) {
      FeatureRef ref_ftr = reprojFeatureOf(&map.feat[c.feat]);
      const SE3 T_cur_ref = cur.T_f_w * inverse(kfs[kf].T_f_w);
      if (isConvergedCornerEdgeletSeed(c.type) || isConvergedMapPointSeed(c.type)) {
        const double ref_depth = seed::getDepth(r.seed_state);
        V2 px = c.cur_px;
        const Matcher::MatchResult res = matcher.findMatchDirect(kfs[kf].mf, cur.mf, T_cur_ref, ref_ftr, ref_depth, px);
        r.match_result = int(res);
        ok = res == Matcher::MatchResult::kSuccess;
      } else if (isUnconvergedCornerEdgeletSeed(c.type) || isUnconvergedMapPointSeed(c.type)) {
        FeatureType type = c.type;
        ok = updateSeed(cur.mf, kfs[kf].mf, T_cur_ref, ref_ftr, type, r.seed_state, map.kf_seed_mu_range[kf], matcher,
                        opt.seed_sigma2_thresh, opt.px_error_angle, false, false, true, &r.match_result);
        r.type_out = int(type);
      }
      grad_ref = ref_ftr.grad;
    } else {
      // Point::getCloseViewObs (point.cpp:83-129)
      const V3 pos{map.pt_pos[3 * pt], map.pt_pos[3 * pt + 1], map.pt_pos[3 * pt + 2]};
      double min_cos_angle = 0.0;
      const V3 obs_dir = normalized(cur.pos() - pos);
      int best = -1;
      for (int o = map.pt_obs_begin[pt]; o < map.pt_obs_begin[pt + 1]; ++o) {
        const V3 dir = normalized(kfs[map.feat_kf[map.obs_feat[o]]].pos() - pos);
        const double cos_angle = dot(obs_dir, dir);
        if (cos_angle > min_cos_angle) { min_cos_angle = cos_angle; best = map.obs_feat[o]; }
      }
      if (!(min_cos_angle < 0.4)) {
        const ReprojFrame& ref_frame = kfs[map.feat_kf[best]];
        FeatureRef ref_ftr = reprojFeatureOf(&map.feat[best]);
        const double ref_depth = norm(ref_frame.pos() - pos);
        const SE3 T_cur_ref = cur.T_f_w * inverse(ref_frame.T_f_w);
        V2 px = c.cur_px;
        const Matcher::MatchResult res = matcher.findMatchDirect(ref_frame.mf, cur.mf, T_cur_ref, ref_ftr, ref_depth, px);
        r.match_result = int(res);
        if (res != Matcher::MatchResult::kSuccess) {
          r.d_failed = 1;
        } else {
          r.d_succeeded = 1;
          grad_ref = ref_ftr.grad;
          ok = true;
        }
      }
    }
    if (!ok) continue;
    if (isEdgelet(c.type)) {
      V2 g{matcher.A_cur_ref_[0][0] * grad_ref.x + matcher.A_cur_ref_[0][1] * grad_ref.y,
           matcher.A_cur_ref_[1][0] * grad_ref.x + matcher.A_cur_ref_[1][1] * grad_ref.y};
      g = normalized(g);
      r.grad[0] = g.x; r.grad[1] = g.y;
    }
    r.px[0] = matcher.px_cur_.x; r.px[1] = matcher.px_cur_.y;
    r.f[0] = matcher.f_cur_.x; r.f[1] = matcher.f_cur_.y; r.f[2] = matcher.f_cur_.z;
    r.level = matcher.search_level_;
    r.status = ORC_REPROJ_MATCHED;
    r.slot = int(num_features);
    ++stats->n_matches;
    ++num_features;
    occupancy[grid_index] = 1;
    if (max_n > 0 && num_features >= max_n) break;
  }
  stats->n_consumed = i;
}

}  // namespace orc
