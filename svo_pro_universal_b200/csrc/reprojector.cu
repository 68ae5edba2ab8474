// (f1) Reprojector candidate matching: project the map into the current frame, bucket the candidates in grid cells, and
// find at most one match per cell in the reference's candidate order.
//
// ref: src/svo/src/reprojector.cpp:310-341 (sortCandidatesByReprojStats / ByNumObs), :342-381 (matchCandidates),
//      :384-485 (matchCandidate), :487-518 (getCandidate), :520-543 (projectPointAndCheckVisibility)
//      src/svo_common/src/frame.cpp:229-257 (Frame::isVisible), src/svo_common/src/point.cpp:83-129 (getCloseViewObs)
//      src/svo_common/include/svo/common/occupancy_grid_2d.h:82-95 (getCellIndex)
//
// The reference walks ONE sorted candidate list sequentially: skip a candidate whose cell is occupied, else try to match it,
// mark the cell on success and stop after max_n_features. A match attempt depends only on the candidate itself, so the walk
// factorises: per cell, the first candidate (in list order) that matches is the cell's winner; the global stop is the
// position of the Q-th winner in list order. The launch sequence per batch of F current frames:
//   reproj_candidates_kernel  one thread per entry: world point -> isVisible -> 8 px margin -> cur_px, grid cell, sort key
//   reproj_sort_kernel        one CTA per frame: bitonic sort of the candidates by the reference's comparator (ties keep the
//                             visiting order), then a second sort by (cell, position) that yields per-cell candidate queues
//   reproj_match_kernel       one 8-lane group per (frame, cell): walks the cell's queue until the first match, in two
//                             passes: <false> findMatchDirect only (landmarks, converged seeds); <true> continues the queues
//                             that reached an unconverged seed without a winner (updateSeed with the epipolar search)
//   reproj_commit_kernel      one CTA per frame: prefix count of the winners in list order, stop position, statuses, slots,
//                             occupancy, statistics; attempts behind the stop position are rolled back to "not reached"
// No host round trip between the stages.
#include "depth_filter_dev.cuh"
#include <algorithm>

using namespace svo_dev;

namespace {

constexpr int kMaxPerFrame = 4096;  // entries per current frame and grid cells (shared-memory sort capacity)
// CTA shape / occupancy target of the match stage (macros for A/B builds; measured on the B200, see profiles/).
#ifndef SVO_REPROJ_THREADS
#define SVO_REPROJ_THREADS 128
#endif
#ifndef SVO_REPROJ_MINB
#define SVO_REPROJ_MINB 3
#endif
constexpr int kThreads = SVO_REPROJ_THREADS;
constexpr int kGroupsPerCta = kThreads / kGroup;
constexpr int kSortThreads = 512;

struct ReprojParams {
  PyrView ref_pyr, cur_pyr;
  svo_camera cam_ref, cam_cur;
  svo_reproj_map map;
  svo_reprojector_options opt;
  svo_matcher_options mopt;
  double px_error_angle;
  int F, n_cells, n_cols;
  int frame0;          // first frame of this launch (grids with the frame in blockIdx.y cover at most 65535 frames per launch)
  const int* cur_frame_idx;
  const double* cur_T_f_w;
  const int* n_features_in;
  const int* entry_begin;
  const int* entry_feat;
  uint8_t* occupancy;
  svo_reproj_result* results;
  svo_reproj_stats* stats;
  // scratch
  int* entry_cell;     // [E] grid cell of a candidate, -1 = not a candidate
  int* sorted_entry;   // [E] frame-local position -> global entry index
  unsigned* cell_list; // [E] (cell << 13 | position), grouped by cell, positions ascending
  int* cell_begin;     // [F][n_cells] first index into cell_list of the frame, -1 = empty cell
  int* cell_success;   // [F][n_cells] position of the cell's winner, -1 = none
  int* cell_items;     // [F][n_cells] the non-empty cells of the frame, compacted, ordered by the list position of their first candidate
  int* n_nonempty;     // [F]
  int* work_item;      // [F * n_cells] second matching pass: compact list of (frame * n_cells + cell) that may still hold a winner before the stop
  int* work_count;     // [1]
  int* resume_cell;    // [F * n_cells] compact list of (frame * n_cells + cell) handed over by the direct-match pass to the full pass
  int* resume_q;       // [F * n_cells] queue index where each of them continues
  int* resume_count;   // [1]
  int* n_cand;         // [F]
};

SVO_D int frameEntries(const ReprojParams& P, int j, int* base) {
  *base = P.entry_begin[j];
  return min(P.entry_begin[j + 1] - *base, kMaxPerFrame);
}

// total order on doubles as unsigned integers (-0.0 is folded onto +0.0 first: the reference compares with operator>)
SVO_D unsigned long long orderedDouble(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// ---- stage 1: getCandidate ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reproj_candidates_kernel(const ReprojParams P) {
  const int j = P.frame0 + (int)blockIdx.y;
  int base;
  const int n = frameEntries(P, j, &base);
  const int n_all = P.entry_begin[j + 1] - base;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_all) return;
  const int e = base + i;
  const int fi = P.entry_feat[e];
  const svo_feature ft = P.map.feat[fi];
  svo_reproj_result r;
  memset(&r, 0, sizeof(r));
  r.status = SVO_REPROJ_NOT_CANDIDATE; r.order = -1; r.slot = -1; r.match_result = -1;
  r.type_out = ft.type;
  for (int k = 0; k < 4; ++k) r.seed_state[k] = P.map.feat_seed_state[4 * (size_t)fi + k];
  int cell = -1;
  if (i < n) {
    const int pt = P.map.feat_point[fi];
    V3d xyz_world;
    if (pt >= 0) {
      xyz_world = V3d{P.map.pt_pos[3 * (size_t)pt], P.map.pt_pos[3 * (size_t)pt + 1], P.map.pt_pos[3 * (size_t)pt + 2]};
    } else {  // ref_frame->T_world_cam() * ref_frame->getSeedPosInFrame(ref_index)
      const SE3d T_w_ref = se3Inv(se3Load(P.map.kf_T_f_w + 7 * (size_t)P.map.feat_kf[fi]));
      xyz_world = se3Apply(T_w_ref, V3d{ft.f[0], ft.f[1], ft.f[2]} * (1.0 / r.seed_state[0]));
    }
    // Frame::isVisible (frame.cpp:229-257), pinhole branch
    const V3d xyz_f = se3Apply(se3Load(P.cur_T_f_w + 7 * (size_t)j), xyz_world);
    bool vis = !(xyz_f.z < 0.0);
    if (vis) {
      const V3d f_top_left = normalized3(camBackProject3(P.cam_cur, 0.0, 0.0));
      const V3d z{0.0, 0.0, 1.0};
      vis = !(dot3(normalized3(xyz_f), z) < dot3(f_top_left, z));
    }
    if (vis) {
      const V2d px = camProject3(P.cam_cur, xyz_f);
      vis = px.x >= 0.0 && px.y >= 0.0 && px.x < (double)P.cam_cur.width && px.y < (double)P.cam_cur.height;
      if (vis) {  // projectPointAndCheckVisibility: isKeypointVisibleWithMargin(px.cast<int>(), 8)
        const int x = (int)px.x, y = (int)px.y;
        constexpr int kPatchSize = 8;
        vis = x >= kPatchSize && y >= kPatchSize && x < P.cam_cur.width - kPatchSize && y < P.cam_cur.height - kPatchSize;
        if (vis) {
          r.cur_px[0] = px.x; r.cur_px[1] = px.y;
          // OccupandyGrid2D::getCellIndex(int x, int y, 1)
          cell = (int)(floor((double)y / P.opt.cell_size) * P.n_cols + floor((double)x / P.opt.cell_size));
        }
      }
    }
  }
  P.entry_cell[e] = cell;
  P.results[e] = r;
}

// ---- stage 2: sort + per-cell queues -----------------------------------------------------------------------------------
struct SortSmem {
  unsigned long long* hi;  // primary << 32 | biased n_reproj (descending); 0 for non-candidates
  unsigned long long* lo;  // ordered score (descending)
  int* idx;                // frame-local entry index (ascending tie-break); bit 30 set = not a candidate
};
SVO_D bool before(const SortSmem& s, int a, int b) {
  const bool va = !(s.idx[a] & 0x40000000), vb = !(s.idx[b] & 0x40000000);
  if (va != vb) return va;
  if (s.hi[a] != s.hi[b]) return s.hi[a] > s.hi[b];
  if (s.lo[a] != s.lo[b]) return s.lo[a] > s.lo[b];
  return s.idx[a] < s.idx[b];
}

__global__ void __launch_bounds__(kSortThreads) reproj_sort_kernel(const ReprojParams P) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ int s_count, s_nz;
  const int j = blockIdx.x, tid = threadIdx.x;
  int base;
  const int n = frameEntries(P, j, &base);
  int npad = 32;
  while (npad < n) npad <<= 1;
  SortSmem s;
  s.hi = reinterpret_cast<unsigned long long*>(s_raw);
  s.lo = s.hi + npad;
  s.idx = reinterpret_cast<int*>(s.lo + npad);
  if (tid == 0) { s_count = 0; s_nz = 0; }
  for (int c = tid; c < P.n_cells; c += kSortThreads) {
    P.cell_begin[(size_t)j * P.n_cells + c] = -1;
    P.cell_success[(size_t)j * P.n_cells + c] = -1;
  }
  __syncthreads();
  int mine = 0;
  for (int i = tid; i < npad; i += kSortThreads) {
    unsigned long long hi = 0, lo = 0;
    int idx = i | 0x40000000;
    if (i < n && P.entry_cell[base + i] >= 0) {
      const int fi = P.entry_feat[base + i];
      const int pt = P.map.feat_point[fi];
      const int n_reproj = pt >= 0 ? P.map.pt_n_succeeded[pt] - P.map.pt_n_failed[pt] : 0;
      const unsigned primary = P.opt.sort_by_num_obs ? (pt >= 0 ? (unsigned)(P.map.pt_obs_begin[pt + 1] - P.map.pt_obs_begin[pt]) : 0u)
                                                     : (unsigned)P.map.feat[fi].type;
      hi = ((unsigned long long)primary << 32) | (unsigned)(n_reproj + 0x80000000u);
      lo = orderedDouble(P.map.feat_score[fi]);
      idx = i;
      ++mine;
    }
    s.hi[i] = hi; s.lo[i] = lo; s.idx[i] = idx;
  }
  if (mine) atomicAdd(&s_count, mine);
  __syncthreads();
  const int n_cand = s_count;
  for (int k = 2; k <= npad; k <<= 1)
    for (int d = k >> 1; d > 0; d >>= 1) {
      for (int i = tid; i < npad; i += kSortThreads) {
        const int x = i ^ d;
        if (x > i) {
          const bool up = (i & k) == 0;
          if (up ? before(s, x, i) : before(s, i, x)) {
            const unsigned long long h = s.hi[i], l = s.lo[i];
            const int t = s.idx[i];
            s.hi[i] = s.hi[x]; s.lo[i] = s.lo[x]; s.idx[i] = s.idx[x];
            s.hi[x] = h; s.lo[x] = l; s.idx[x] = t;
          }
        }
      }
      __syncthreads();
    }
  // positions -> entries; (cell, position) keys for the queues
  unsigned* ck = reinterpret_cast<unsigned*>(s.hi);  // reuse: the sort keys are dead once idx[] is final
  int* order = s.idx;
  __syncthreads();
  for (int p = tid; p < npad; p += kSortThreads) {
    unsigned key = 0xFFFFFFFFu;
    if (p < n_cand) {
      const int e = base + order[p];
      P.sorted_entry[base + p] = e;
      P.results[e].order = p;
      P.results[e].status = SVO_REPROJ_NOT_REACHED;
      key = ((unsigned)P.entry_cell[e] << 13) | (unsigned)p;
    }
    ck[p] = key;
  }
  __syncthreads();
  for (int k = 2; k <= npad; k <<= 1)
    for (int d = k >> 1; d > 0; d >>= 1) {
      for (int i = tid; i < npad; i += kSortThreads) {
        const int x = i ^ d;
        if (x > i) {
          const bool up = (i & k) == 0;
          const unsigned a = ck[i], b = ck[x];
          if (up ? b < a : a < b) { ck[i] = b; ck[x] = a; }
        }
      }
      __syncthreads();
    }
  // the non-empty cells in the order of their first candidate's list position (the matching is progressive in that order):
  // head[position] = cell for queue heads, then an ordered compaction over the positions
  int* head = reinterpret_cast<int*>(s.lo);  // the score keys are dead
  for (int p2 = tid; p2 < npad; p2 += kSortThreads) head[p2] = -1;
  __syncthreads();
  for (int q = tid; q < n_cand; q += kSortThreads) {
    const unsigned key = ck[q];
    P.cell_list[base + q] = key;
    if (q == 0 || (ck[q - 1] >> 13) != (key >> 13)) {
      P.cell_begin[(size_t)j * P.n_cells + (key >> 13)] = q;
      head[key & 8191u] = (int)(key >> 13);
    }
  }
  __syncthreads();
  {
    __shared__ int s_warp[kSortThreads / 32];
    const int per = npad / kSortThreads > 0 ? npad / kSortThreads : 1;  // npad and kSortThreads are powers of two
    const int lo_ = tid * per;
    int cnt = 0;
    if (lo_ < npad)
      for (int k = 0; k < per; ++k) cnt += head[lo_ + k] >= 0;
    int inc = cnt;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int off = inc - cnt;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (lo_ < npad)
      for (int k = 0; k < per; ++k)
        if (head[lo_ + k] >= 0) P.cell_items[(size_t)j * P.n_cells + off++] = head[lo_ + k];
    if (tid == kSortThreads - 1) s_nz = off;
  }
  __syncthreads();
  if (tid == 0) { P.n_cand[j] = n_cand; P.n_nonempty[j] = s_nz; }
}

// ---- stage 3: matchCandidate per cell queue ----------------------------------------------------------------------------
// Point::getCloseViewObs (point.cpp:83-129): the observation whose viewing direction is closest to the current one.
SVO_D int closeViewObs(const svo_reproj_map& map, int pt, const V3d& pos, const V3d& framepos) {
  double min_cos_angle = 0.0;
  const V3d obs_dir = normalized3(framepos - pos);
  int best = -1;
  for (int o = map.pt_obs_begin[pt]; o < map.pt_obs_begin[pt + 1]; ++o) {
    const int fo = map.obs_feat[o];
    const V3d kf_pos = se3Inv(se3Load(map.kf_T_f_w + 7 * (size_t)map.feat_kf[fo])).t;  // Frame::pos()
    const double cos_angle = dot3(obs_dir, normalized3(kf_pos - pos));
    if (cos_angle > min_cos_angle) { min_cos_angle = cos_angle; best = fo; }
  }
  return min_cos_angle < 0.4 ? -1 : best;  // observations more than 60 degrees away are useless
}

// FULL = false: the direct-match pass — findMatchDirect only (landmarks and converged seeds, which the reference's order puts
// first). A queue that reaches an unconverged seed before it has a winner is handed over (cell_resume) to the FULL pass,
// which also carries updateSeed with the epipolar search. Splitting keeps the common pass small: with everything inlined in
// one kernel half of the warp stalls were instruction-fetch misses (profiles/).
// Cells of the first matching pass of a frame: its quota of new features plus a margin for failed attempts (96 % of the attempts of the
// bench scenes succeed); the cells are visited in the order of their first candidate's list position.
SVO_D int firstPassCells(const ReprojParams& P, int j) {
  const int quota = max(1, P.opt.max_n_features - P.n_features_in[j]);
  return min(P.n_nonempty[j], quota + quota / 4 + 8);
}

// mode 0: grid (items, frame) — every candidate (max_n_features == 0) or the first-pass cells of the frame; mode 1: the hand-over list
// of the direct-match pass (resume_cell / resume_q); mode 2: the work list of the second pass (work_item). The list modes run as a
// grid-stride loop over lists whose length is known only on the device.
template <bool FULL>
#ifndef SVO_REPROJ_DIRECT_MINB
#define SVO_REPROJ_DIRECT_MINB 5   // chain stage per 16384 frames: 3 CTAs / SM 9.62 ms, 4 (128 registers) 8.32 ms, 5 (96) 8.04 ms
#endif
__global__ void __launch_bounds__(kThreads, FULL ? SVO_REPROJ_MINB : SVO_REPROJ_DIRECT_MINB) reproj_match_kernel(const ReprojParams P, int items_per_frame, int mode) {
  __shared__ __align__(16) uint8_t s_pwb[kGroupsPerCta * kPwbPitch];
  const Group g = makeGroup();
  const int gi = threadIdx.x / kGroup;
  uint8_t* pwb = s_pwb + gi * kPwbPitch;
  const bool unlimited = P.opt.max_n_features <= 0;  // matchCandidates ignores the grid when max_n_features_per_frame == 0
  const int total = mode == 1 ? *P.resume_count : (mode == 2 ? *P.work_count : 0);
  for (int k0 = blockIdx.x * kGroupsPerCta; mode == 0 || k0 < total; k0 += gridDim.x * kGroupsPerCta) {
  int item = k0 + gi;
  int j = mode == 0 ? P.frame0 + (int)blockIdx.y : 0;
  int q = 0, q_end = 0;
  unsigned cell = 0;
  bool occupied = false;
  int base = 0, n_cand = 0;
  if (mode != 0) {
    const int k = item;
    item = -1;  // no work unless the list holds an entry for this group
    if (k < total) {
      const int jc = mode == 1 ? P.resume_cell[k] : P.work_item[k];
      j = jc / P.n_cells;
      item = jc - j * P.n_cells;
      frameEntries(P, j, &base);
      n_cand = P.n_cand[j];
      q = mode == 1 ? P.resume_q[k] : P.cell_begin[(size_t)j * P.n_cells + item];
      q_end = q < 0 ? 0 : n_cand;
      q = max(q, 0);
      cell = (unsigned)item;
      occupied = P.occupancy[(size_t)j * P.n_cells + item] != 0;
    }
  } else {
    frameEntries(P, j, &base);
    n_cand = P.n_cand[j];
    if (item < items_per_frame) {
      if (unlimited) {
        q = item; q_end = min(item + 1, n_cand);
      } else if (item < firstPassCells(P, j)) {
        item = P.cell_items[(size_t)j * P.n_cells + item];  // item-th non-empty cell of the frame, in list order of the queue heads
        q = P.cell_begin[(size_t)j * P.n_cells + item];
        q_end = q < 0 ? 0 : n_cand;
        q = max(q, 0);
        cell = (unsigned)item;
        occupied = P.occupancy[(size_t)j * P.n_cells + item] != 0;
      }
    }
  }
  const SE3d T_cur_w = se3Load(P.cur_T_f_w + 7 * (size_t)j);
  const int cf = P.cur_frame_idx ? P.cur_frame_idx[j] : j;
  bool found = false;
  // Rounds: every queue of the warp first skips ahead to its next candidate that needs an attempt (cheap, divergent), then
  // the whole warp reconverges on the vote and the attempts of the round run in lock-step.
  for (;;) {
    int p = -1;
    while (q < q_end) {
      int pp = q;
      if (!unlimited) {
        const unsigned key = P.cell_list[base + q];
        if ((key >> 13) != cell) { q = q_end; break; }
        pp = (int)(key & 8191u);
      }
      if (occupied || found) {
        if (g.r == 0) P.results[P.sorted_entry[base + pp]].status = SVO_REPROJ_SKIPPED;
        ++q;
        continue;
      }
      if (!FULL) {
        const int t = P.map.feat[P.entry_feat[P.sorted_entry[base + pp]]].type;
        if (P.map.feat_point[P.entry_feat[P.sorted_entry[base + pp]]] < 0 && (t == kEdgeletSeed || t == kCornerSeed || t == kMapPointSeed)) {
          if (g.r == 0) {  // the full pass continues here
            const int k = atomicAdd(P.resume_count, 1);
            P.resume_cell[k] = j * P.n_cells + item;
            P.resume_q[k] = q;
          }
          q = q_end;
          break;
        }
      }
      p = pp;
      ++q;
      break;
    }
    if (!__any_sync(0xffffffffu, p >= 0)) break;
    if (p < 0) continue;
    const int e = P.sorted_entry[base + p];
    svo_reproj_result* r = P.results + e;
    const int fi = P.entry_feat[e];
    const int ctype = P.map.feat[fi].type;
    const int pt = P.map.feat_point[fi];
    const double guess_x = r->cur_px[0], guess_y = r->cur_px[1];
    MatchState m;
    initMatchState(m);
    int mr = -1, type_out = ctype, d_failed = 0, d_succeeded = 0;
    double st[4] = {P.map.feat_seed_state[4 * (size_t)fi], P.map.feat_seed_state[4 * (size_t)fi + 1],
                    P.map.feat_seed_state[4 * (size_t)fi + 2], P.map.feat_seed_state[4 * (size_t)fi + 3]};
    double grad_x = 0.0, grad_y = 0.0;
    bool ok = false;
    // which reference feature is matched, from which keyframe, at which depth — then ONE call site per matcher entry point
    // (the four cell queues of a warp run in lock-step as long as they are in the same code)
    int mode = 0;  // 1 = findMatchDirect, 2 = updateSeed
    int fr = fi;   // the feature whose patch is warped
    double ref_depth = 0.0;
    V3d pos{0, 0, 0};
    if (pt < 0) {
      if (ctype == kEdgeletSeedConverged || ctype == kCornerSeedConverged || ctype == kMapPointSeedConverged) {
        mode = 1;
        ref_depth = 1.0 / st[0];  // getSeedDepth
      } else if (ctype == kEdgeletSeed || ctype == kCornerSeed || ctype == kMapPointSeed) {
        mode = 2;
      }
    } else {
      pos = V3d{P.map.pt_pos[3 * (size_t)pt], P.map.pt_pos[3 * (size_t)pt + 1], P.map.pt_pos[3 * (size_t)pt + 2]};
      fr = closeViewObs(P.map, pt, pos, se3Inv(T_cur_w).t);
      if (fr >= 0) mode = 1;
    }
    if (mode) {
      svo_feature ft = P.map.feat[fr];
      const int kf = P.map.feat_kf[fr];
      const int rf = P.map.kf_frame_idx ? P.map.kf_frame_idx[kf] : kf;
      const SE3d T_ref_w = se3Load(P.map.kf_T_f_w + 7 * (size_t)kf);
      if (pt >= 0) ref_depth = norm3(se3Inv(T_ref_w).t - pos);  // (ref_frame->pos() - landmark->pos()).norm()
      const SE3d T = se3Mul(T_cur_w, se3Inv(T_ref_w));
      if (mode == 1) {
        mr = findMatchDirect(g, P.ref_pyr, rf, P.cur_pyr, cf, P.cam_ref, P.cam_cur, T, ft, ref_depth, guess_x, guess_y, P.mopt, pwb, m);
        ok = mr == kSuccess;
        if (pt >= 0) { d_failed = !ok; d_succeeded = ok; }
      } else if (FULL) {
        ok = updateSeedOnce(g, P.ref_pyr, rf, P.cur_pyr, cf, P.cam_ref, P.cam_cur, T, ft, type_out, st, P.map.kf_seed_mu_range[kf],
                            P.opt.seed_sigma2_thresh, P.px_error_angle, false, false, true, P.mopt, pwb, m, &mr);
      }
      grad_x = ft.grad[0]; grad_y = ft.grad[1];
    }
    __syncwarp(g.mask);
    if (g.r == 0) {
      r->match_result = mr;
      r->type_out = type_out;
      r->d_failed = d_failed; r->d_succeeded = d_succeeded;
      for (int k = 0; k < 4; ++k) r->seed_state[k] = st[k];
      r->status = ok ? SVO_REPROJ_MATCHED : SVO_REPROJ_FAILED;
      if (ok) {
        if (isEdgeletType(ctype)) {  // feature.grad = (matcher.A_cur_ref_ * grad_ref).normalized()
          const V2d gp = normalized2(V2d{m.A[0][0] * grad_x + m.A[0][1] * grad_y, m.A[1][0] * grad_x + m.A[1][1] * grad_y});
          r->grad[0] = gp.x; r->grad[1] = gp.y;
        }
        r->px[0] = m.px_x; r->px[1] = m.px_y;
        r->f[0] = m.f_cur.x; r->f[1] = m.f_cur.y; r->f[2] = m.f_cur.z;
        r->level = m.search_level;
        if (!unlimited) P.cell_success[(size_t)j * P.n_cells + item] = p;
      }
    }
    if (ok && !unlimited) found = true;
  }
  if (mode == 0) break;
  }
}

// ---- stage 3b: after the first matching pass — which of the remaining cells can still matter? --------------------------------
// The reference stops at the quota-th success in list order. With the winners found so far that stop is at position p1 (or nowhere
// yet); more winners can only move it forward, so a cell whose first candidate lies behind p1 can never be reached and is left out
// (its entries stay "not reached"). The other unvisited cells go on the work list of the second pass.
__global__ void __launch_bounds__(256) reproj_progress_kernel(const ReprojParams P) {
  __shared__ int s_acc[kMaxPerFrame];
  __shared__ int s_part[256];
  __shared__ int s_stop;
  const int j = blockIdx.x, tid = threadIdx.x;
  int base;
  frameEntries(P, j, &base);
  const int n_nz = P.n_nonempty[j], k1 = firstPassCells(P, j);
  if (k1 >= n_nz) return;  // the first pass visited every cell
  const int quota = max(1, P.opt.max_n_features - P.n_features_in[j]);
  for (int p = tid; p < kMaxPerFrame; p += 256) s_acc[p] = 0;
  if (tid == 0) s_stop = 0x7fffffff;
  __syncthreads();
  for (int c = tid; c < P.n_cells; c += 256) {
    const int p = P.cell_success[(size_t)j * P.n_cells + c];
    if (p >= 0) s_acc[p] = 1;
  }
  __syncthreads();
  constexpr int kChunk = kMaxPerFrame / 256;
  int sum = 0;
  for (int k = 0; k < kChunk; ++k) sum += s_acc[tid * kChunk + k];
  s_part[tid] = sum;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    const int v = tid >= d ? s_part[tid - d] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int excl = s_part[tid] - sum;
  for (int k = 0; k < kChunk; ++k) {
    const int p = tid * kChunk + k;
    if (s_acc[p]) { if (excl + 1 == quota) s_stop = p; ++excl; }
  }
  __syncthreads();
  const int stop = s_stop;
  for (int k = k1 + tid; k < n_nz; k += 256) {
    const int cell = P.cell_items[(size_t)j * P.n_cells + k];
    const int head = (int)(P.cell_list[base + P.cell_begin[(size_t)j * P.n_cells + cell]] & 8191u);
    if (head < stop) P.work_item[atomicAdd(P.work_count, 1)] = j * P.n_cells + cell;
  }
}

// ---- stage 4: stop position, slots, occupancy, statistics --------------------------------------------------------------
__global__ void __launch_bounds__(256) reproj_commit_kernel(const ReprojParams P) {
  __shared__ int s_acc[kMaxPerFrame];   // 1 where the candidate at that position is a winner; then the exclusive prefix count
  __shared__ int s_part[256];
  __shared__ int s_break, s_trials, s_matches;
  const int j = blockIdx.x, tid = threadIdx.x;
  int base;
  frameEntries(P, j, &base);
  const int n_cand = P.n_cand[j];
  const bool unlimited = P.opt.max_n_features <= 0;
  const int n_in = P.n_features_in[j];
  // matchCandidates stops when frame->num_features_ >= max after a success: at least one success is always taken
  const int quota = unlimited ? 0x7fffffff : max(1, P.opt.max_n_features - n_in);
  for (int p = tid; p < kMaxPerFrame; p += 256) s_acc[p] = 0;
  if (tid == 0) { s_break = n_cand - 1; s_trials = 0; s_matches = 0; }
  __syncthreads();
  if (unlimited) {
    for (int p = tid; p < n_cand; p += 256) s_acc[p] = P.results[P.sorted_entry[base + p]].status == SVO_REPROJ_MATCHED;
  } else {
    for (int c = tid; c < P.n_cells; c += 256) {
      const int p = P.cell_success[(size_t)j * P.n_cells + c];
      if (p >= 0) s_acc[p] = 1;
    }
  }
  __syncthreads();
  // exclusive prefix sum over positions: 16 consecutive positions per thread
  constexpr int kChunk = kMaxPerFrame / 256;
  int local[kChunk], sum = 0;
  for (int k = 0; k < kChunk; ++k) { local[k] = sum; sum += s_acc[tid * kChunk + k]; }
  s_part[tid] = sum;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    const int v = tid >= d ? s_part[tid - d] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  const int offset = s_part[tid] - sum;
  for (int k = 0; k < kChunk; ++k) {
    const int p = tid * kChunk + k;
    const int excl = offset + local[k];
    if (s_acc[p] && excl + 1 == quota) s_break = p;  // the quota-th winner: matchCandidates breaks right after it
    local[k] = excl;
  }
  __syncthreads();
  const int i_break = s_break;
  int trials = 0, matches = 0;
  for (int k = 0; k < kChunk; ++k) {
    const int p = tid * kChunk + k;
    if (p >= n_cand) break;
    const int e = P.sorted_entry[base + p];
    svo_reproj_result* r = P.results + e;
    if (p > i_break) {  // never visited by the reference's loop: undo the speculative attempt
      const int fi = P.entry_feat[e];
      r->status = SVO_REPROJ_NOT_REACHED;
      r->match_result = -1; r->d_failed = 0; r->d_succeeded = 0; r->level = 0; r->slot = -1;
      r->type_out = P.map.feat[fi].type;
      for (int q = 0; q < 4; ++q) r->seed_state[q] = P.map.feat_seed_state[4 * (size_t)fi + q];
      r->px[0] = r->px[1] = r->f[0] = r->f[1] = r->f[2] = r->grad[0] = r->grad[1] = 0.0;
      continue;
    }
    const int status = r->status;
    if (status == SVO_REPROJ_MATCHED) {
      r->slot = n_in + local[k];
      ++matches; ++trials;
      P.occupancy[(size_t)j * P.n_cells + P.entry_cell[e]] = 1;
    } else if (status == SVO_REPROJ_FAILED) {
      ++trials;
    }
  }
  if (trials) atomicAdd(&s_trials, trials);
  if (matches) atomicAdd(&s_matches, matches);
  __syncthreads();
  if (tid == 0) {
    svo_reproj_stats st;
    st.n_candidates = n_cand; st.n_trials = s_trials; st.n_matches = s_matches;
    st.n_consumed = n_cand > 0 ? i_break + 1 : 0;
    P.stats[j] = st;
  }
}

}  // namespace

extern "C" int svo_cuda_reproject_match(svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const svo_cuda_pyr* cur_pyr,
                                        const svo_camera* cam_ref, const svo_camera* cam_cur, const svo_reproj_map* map, int F,
                                        const int* cur_frame_idx, const double* cur_T_f_w, const int* n_features_in,
                                        const int* entry_begin, int n_entries, const int* entry_feat, uint8_t* occupancy,
                                        const svo_reprojector_options* opt, svo_reproj_result* results, svo_reproj_stats* stats,
                                        svo_mem mem) {
  if (!ctx || !ref_pyr || !cur_pyr || !cam_ref || !cam_cur || !map || F < 0 || !cur_T_f_w || !n_features_in || !entry_begin ||
      n_entries < 0 || (n_entries > 0 && (!entry_feat || !results)) || !occupancy || !opt || !stats)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_reproject_match: bad arguments");
  if (map->n_kfs < 0 || map->n_feat < 0 || map->n_points < 0 || map->n_obs < 0 || !map->kf_T_f_w || !map->kf_seed_mu_range ||
      (map->n_feat > 0 && (!map->feat || !map->feat_score || !map->feat_seed_state || !map->feat_point || !map->feat_kf)) ||
      (map->n_points > 0 && (!map->pt_pos || !map->pt_n_failed || !map->pt_n_succeeded || !map->pt_obs_begin)) ||
      (map->n_obs > 0 && !map->obs_feat) || opt->cell_size <= 0)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_reproject_match: bad map tables / options");
  if (F == 0) return SVO_OK;
  cudaSetDevice(ctx->device);
  ReprojParams P;
  memset(&P, 0, sizeof(P));
  P.n_cells = svo_cuda_grid_cells(cam_cur->width, cam_cur->height, opt->cell_size, &P.n_cols, nullptr);
  if (P.n_cells <= 0 || P.n_cells > kMaxPerFrame)
    return SVO_FAIL(ctx, SVO_ERR_TOO_MANY_FEATURES, "svo_cuda_reproject_match: more than 4096 grid cells");
  if (mem == SVO_MEM_HOST) {
    if (entry_begin[F] != n_entries) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_reproject_match: n_entries != entry_begin[F]");
    for (int j = 0; j < F; ++j)
      if (entry_begin[j + 1] - entry_begin[j] > kMaxPerFrame || entry_begin[j + 1] < entry_begin[j])
        return SVO_FAIL(ctx, SVO_ERR_TOO_MANY_FEATURES, "svo_cuda_reproject_match: more than 4096 entries for one frame");
  }
  P.ref_pyr = makeView(ref_pyr);
  P.cur_pyr = makeView(cur_pyr);
  P.cam_ref = *cam_ref;
  P.cam_cur = *cam_cur;
  P.opt = *opt;
  // `Matcher matcher;` of matchCandidates: the defaults of Matcher::Options (matcher.h:39-54) + the two affine flags
  P.mopt.align_1d = 0; P.mopt.align_max_iter = 10; P.mopt.max_epi_search_steps = 100; P.mopt.subpix_refinement = 1;
  P.mopt.epi_search_edgelet_filtering = 1; P.mopt.scan_on_unit_sphere = 1; P.mopt.epi_search_edgelet_max_angle = 0.7;
  P.mopt.affine_est_offset = opt->affine_est_offset; P.mopt.affine_est_gain = opt->affine_est_gain;
  P.mopt.max_patch_diff_ratio = 2.0;
  P.px_error_angle = opt->px_error_angle > 0.0 ? opt->px_error_angle : camAngleError(*cam_cur, 1.0);
  P.F = F;
  Stager st(ctx, mem);
  P.map = *map;
  P.map.kf_T_f_w = st.in(map->kf_T_f_w, (size_t)map->n_kfs * 7);
  P.map.kf_seed_mu_range = st.in(map->kf_seed_mu_range, (size_t)map->n_kfs);
  P.map.kf_frame_idx = st.in(map->kf_frame_idx, (size_t)map->n_kfs);
  P.map.feat = st.in(map->feat, (size_t)map->n_feat);
  P.map.feat_score = st.in(map->feat_score, (size_t)map->n_feat);
  P.map.feat_seed_state = st.in(map->feat_seed_state, (size_t)map->n_feat * 4);
  P.map.feat_point = st.in(map->feat_point, (size_t)map->n_feat);
  P.map.feat_kf = st.in(map->feat_kf, (size_t)map->n_feat);
  P.map.pt_pos = st.in(map->pt_pos, (size_t)map->n_points * 3);
  P.map.pt_n_failed = st.in(map->pt_n_failed, (size_t)map->n_points);
  P.map.pt_n_succeeded = st.in(map->pt_n_succeeded, (size_t)map->n_points);
  P.map.pt_obs_begin = st.in(map->pt_obs_begin, (size_t)map->n_points + 1);
  P.map.obs_feat = st.in(map->obs_feat, (size_t)map->n_obs);
  P.cur_frame_idx = st.in(cur_frame_idx, (size_t)F);
  P.cur_T_f_w = st.in(cur_T_f_w, (size_t)F * 7);
  P.n_features_in = st.in(n_features_in, (size_t)F);
  P.entry_begin = st.in(entry_begin, (size_t)F + 1);
  P.entry_feat = st.in(entry_feat, (size_t)n_entries);
  P.occupancy = st.inout(occupancy, (size_t)F * P.n_cells);
  P.results = st.out(results, (size_t)n_entries);
  P.stats = st.out(stats, (size_t)F);
  const size_t ne = (size_t)(n_entries > 0 ? n_entries : 1);
  P.entry_cell = (int*)st.scratch(ne * sizeof(int));
  P.sorted_entry = (int*)st.scratch(ne * sizeof(int));
  P.cell_list = (unsigned*)st.scratch(ne * sizeof(unsigned));
  P.cell_begin = (int*)st.scratch((size_t)F * P.n_cells * sizeof(int));
  P.cell_success = (int*)st.scratch((size_t)F * P.n_cells * sizeof(int));
  P.cell_items = (int*)st.scratch((size_t)F * P.n_cells * sizeof(int));
  P.n_nonempty = (int*)st.scratch((size_t)F * sizeof(int));
  P.resume_cell = (int*)st.scratch((size_t)F * P.n_cells * sizeof(int));
  P.resume_q = (int*)st.scratch((size_t)F * P.n_cells * sizeof(int));
  P.resume_count = (int*)st.scratch(sizeof(int));
  P.work_item = (int*)st.scratch((size_t)F * P.n_cells * sizeof(int));
  P.work_count = (int*)st.scratch(sizeof(int));
  P.n_cand = (int*)st.scratch((size_t)F * sizeof(int));
  if (!st.send()) return st.finish();

  const int per_frame_cap = mem == SVO_MEM_HOST ? [&] { int m = 1; for (int j = 0; j < F; ++j) m = std::max(m, entry_begin[j + 1] - entry_begin[j]); return m; }()
                                                : kMaxPerFrame;
  if (n_entries > 0) {
    for (int f0 = 0; f0 < F; f0 += 65535) {  // grid.y limit
      P.frame0 = f0;
      reproj_candidates_kernel<<<dim3((per_frame_cap + 255) / 256, std::min(65535, F - f0)), 256, 0, ctx->stream>>>(P);
      SVO_LAUNCH_CHECK(ctx);
    }
    P.frame0 = 0;
  }
  int npad = 32;
  while (npad < per_frame_cap) npad <<= 1;
  const size_t sort_smem = (size_t)npad * (8 + 8 + 4);
  if (!ctx->attr_reproj) {
    SVO_CUDA_TRY(ctx, cudaFuncSetAttribute(reproj_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxPerFrame * 20));
    ctx->attr_reproj = true;
  }
  reproj_sort_kernel<<<F, kSortThreads, sort_smem, ctx->stream>>>(P);
  SVO_LAUNCH_CHECK(ctx);
  const int items = opt->max_n_features <= 0 ? per_frame_cap : P.n_cells;
  const dim3 mgrid((items + kGroupsPerCta - 1) / kGroupsPerCta, 1);
  if (opt->max_n_features <= 0) {  // unlimited: one candidate per item, every kind of candidate
    for (int f0 = 0; f0 < F; f0 += 65535) {
      P.frame0 = f0;
      reproj_match_kernel<true><<<dim3(mgrid.x, std::min(65535, F - f0)), kThreads, 0, ctx->stream>>>(P, items, 0);
      SVO_LAUNCH_CHECK(ctx);
    }
    P.frame0 = 0;
  } else {
    // Progressive matching. Pass 1 visits, per frame, the cells whose first candidates come first in list order — as many as the
    // frame's quota of new features plus a margin — with the direct-match kernel, then the full kernel on the queues that reached an
    // unconverged seed; the progress kernel lists the unvisited cells that may still hold a winner before the stop; pass 2 repeats both
    // kernels on that list. The reference makes ~125 attempts per frame on the bench scenes (120 successes); matching every non-empty
    // cell at once made ~330.
    const int first_items = std::min(P.n_cells, opt->max_n_features + opt->max_n_features / 4 + 8);
    const dim3 grid1((first_items + kGroupsPerCta - 1) / kGroupsPerCta, 1);
    const int list_grid = ctx->sm_count * 8;  // grid-stride over lists whose length is known only on the device
    SVO_CUDA_TRY(ctx, cudaMemsetAsync(P.resume_count, 0, sizeof(int), ctx->stream));
    SVO_CUDA_TRY(ctx, cudaMemsetAsync(P.work_count, 0, sizeof(int), ctx->stream));
    for (int f0 = 0; f0 < F; f0 += 65535) {
      P.frame0 = f0;
      reproj_match_kernel<false><<<dim3(grid1.x, std::min(65535, F - f0)), kThreads, 0, ctx->stream>>>(P, first_items, 0);
      SVO_LAUNCH_CHECK(ctx);
    }
    P.frame0 = 0;
    reproj_match_kernel<true><<<list_grid, kThreads, 0, ctx->stream>>>(P, 0, 1);
    SVO_LAUNCH_CHECK(ctx);
    reproj_progress_kernel<<<F, 256, 0, ctx->stream>>>(P);
    SVO_LAUNCH_CHECK(ctx);
    SVO_CUDA_TRY(ctx, cudaMemsetAsync(P.resume_count, 0, sizeof(int), ctx->stream));
    reproj_match_kernel<false><<<list_grid, kThreads, 0, ctx->stream>>>(P, 0, 2);
    SVO_LAUNCH_CHECK(ctx);
    reproj_match_kernel<true><<<list_grid, kThreads, 0, ctx->stream>>>(P, 0, 1);
    SVO_LAUNCH_CHECK(ctx);
  }
  reproj_commit_kernel<<<F, 256, 0, ctx->stream>>>(P);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}
