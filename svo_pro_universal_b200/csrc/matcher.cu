// (c) Batched feature alignment and patch matching: align2D / align1D, affine warp, Matcher::findMatchDirect and
// Matcher::findEpipolarMatchDirect. One 8-lane group per feature (see matcher_dev.cuh); 128-thread CTAs hold 16 features.
#include "matcher_dev.cuh"

using namespace svo_dev;

namespace {

constexpr int kThreads = 128;
constexpr int kGroupsPerCta = kThreads / kGroup;

struct MatchParams {
  PyrView ref_pyr, cur_pyr;
  svo_camera cam_ref, cam_cur;
  const int* ref_frame_idx;
  const int* cur_frame_idx;
  const double* T_cur_ref;
  const int* T_idx;
  int M;
  const svo_feature* ftrs;
  const double* depth;     // findMatchDirect: ref depth [M]; epipolar: d_inv [M][3]
  const double* px_guess;  // [M][2]
  svo_matcher_options opt;
  svo_match_out* out;
  // warp-only outputs
  double* A_out;
  int* search_level_out;
  uint8_t* pwb_out;
  uint8_t* ok_out;
  int align1d_from_type;  // findEpipolarMatchDirect: align_1d = isEdgelet(type) per feature (StereoTriangulation::compute)
  int depth_shared;        // one (estimate, min, max) inverse-depth triple for all features
  // progressive matching (svo_cuda_stereo_triangulate): entry i belongs to list entry_pair[i]; only the entries at list positions
  // [chunk_lo, chunk_hi) of lists that are not done yet are matched by this launch
  const int* perm;         // group k works on feature perm[k] (features grouped by type and level), or null: k itself
  const int* entry_pair;
  const int* pair_begin;
  const uint8_t* pair_done;
  int chunk_lo, chunk_hi;
};

SVO_D void writeOut(const Group& g, svo_match_out* out, int i, const MatchState& m, int result, double depth) {
  if (g.r != 0) return;
  svo_match_out o;
  o.px_cur[0] = m.px_x; o.px_cur[1] = m.px_y;
  o.f_cur[0] = m.f_cur.x; o.f_cur[1] = m.f_cur.y; o.f_cur[2] = m.f_cur.z;
  o.A_cur_ref[0] = m.A[0][0]; o.A_cur_ref[1] = m.A[0][1]; o.A_cur_ref[2] = m.A[1][0]; o.A_cur_ref[3] = m.A[1][1];
  o.h_inv = m.h_inv;
  o.epi_length_pyramid = m.epi_length_pyramid;
  o.depth = depth;
  o.result = result;
  o.search_level = m.search_level;
  o.reject = m.reject;
  o._pad = 0;
  o.epi_image[0] = m.epi_x; o.epi_image[1] = m.epi_y;
  out[i] = o;
}

// mode 0: findMatchDirect, 1: findEpipolarMatchDirect, 2: warp only; SCAN: the epipolar scan compiled in (1 sphere, 0 plane)
template <int MODE, int SCAN = 2>
// Resident CTAs per SM the register allocation is held to, per mode (512 k features on the B200, work grouped by type and level):
// findMatchDirect 1.084 / 0.973 / 0.948 ms at 4 / 5 / 6, epipolar search 1.342 / 1.289 / 1.335 ms (round 1, ungrouped work: 4 was best).
#ifndef SVO_MATCH_MINB
#define SVO_MATCH_MINB (MODE == 0 ? 6 : 5)
#endif
__global__ void __launch_bounds__(kThreads, SVO_MATCH_MINB) match_kernel(const MatchParams P) {
  __shared__ __align__(16) uint8_t s_pwb[kGroupsPerCta * kPwbPitch];
  const Group g = makeGroup();
  const int gi = threadIdx.x / kGroup;
  const int k_item = blockIdx.x * kGroupsPerCta + gi;
  if (k_item >= P.M) return;
  const int i = P.perm ? P.perm[k_item] : k_item;
  if (MODE == 1 && P.entry_pair) {
    const int b = P.entry_pair[i], pos = i - P.pair_begin[b];
    if (pos < P.chunk_lo || pos >= P.chunk_hi || P.pair_done[b]) return;
  }
  uint8_t* pwb = s_pwb + gi * kPwbPitch;
  const svo_feature ft = P.ftrs[i];
  const int rf = P.ref_frame_idx ? P.ref_frame_idx[i] : 0;
  const int cf = P.cur_frame_idx ? P.cur_frame_idx[i] : 0;
  const SE3d T = se3Load(P.T_cur_ref + 7 * (size_t)(P.T_idx ? P.T_idx[i] : 0));
  MatchState m;
  initMatchState(m);
  if (MODE == 0) {
    m.px_x = P.px_guess[2 * i]; m.px_y = P.px_guess[2 * i + 1];
    const int res = findMatchDirect(g, P.ref_pyr, rf, P.cur_pyr, cf, P.cam_ref, P.cam_cur, T, ft, P.depth[i], P.px_guess[2 * i],
                                    P.px_guess[2 * i + 1], P.opt, pwb, m);
    writeOut(g, P.out, i, m, res, 0.0);
  } else if (MODE == 1) {
    if (P.align1d_from_type && ft.type < 0) {  // a hole in a fixed-shape entry list (svo_cuda_stereo_triangulate): nothing to match
      if (g.r == 0) { svo_match_out o; memset(&o, 0, sizeof(o)); o.result = -2; P.out[i] = o; }
      return;
    }
    double depth = 0.0;
    const double* dd = P.depth + (P.depth_shared ? 0 : 3 * (size_t)i);
    const bool a1d = P.align1d_from_type ? isEdgeletType(ft.type) : P.opt.align_1d != 0;
    const int res = findEpipolarMatchDirect<SCAN>(g, P.ref_pyr, rf, P.cur_pyr, cf, P.cam_ref, P.cam_cur, T, ft, dd[0], dd[1], dd[2], P.opt, a1d,
                                            pwb, m, &depth);
    writeOut(g, P.out, i, m, res, depth);
  } else {
    const V3d f_ref{ft.f[0], ft.f[1], ft.f[2]};
    getWarpMatrixAffine(P.cam_ref, P.cam_cur, ft.px[0], ft.px[1], f_ref, P.depth[i], T, ft.level, m.A);
    m.search_level = getBestSearchLevel(m.A, P.ref_pyr.n_levels - 1);
    for (int k = g.r; k < 100; k += kGroup) pwb[k] = 0;
    const bool ok = warpAffine10(g, m.A, levelView(P.ref_pyr, rf, ft.level), ft.px[0], ft.px[1], ft.level, m.search_level, pwb);
    if (g.r == 0) {
      P.A_out[4 * i] = m.A[0][0]; P.A_out[4 * i + 1] = m.A[0][1]; P.A_out[4 * i + 2] = m.A[1][0]; P.A_out[4 * i + 3] = m.A[1][1];
      P.search_level_out[i] = m.search_level;
      P.ok_out[i] = ok ? 1 : 0;
    }
    for (int k = g.r; k < 100; k += kGroup) P.pwb_out[100 * (size_t)i + k] = pwb[k];
  }
}

// ---- work order of a large matcher call --------------------------------------------------------------------------------------
// The four features of a warp run in lock-step only while they are in the same code: an edgelet (1-D alignment, edgelet filter) next to
// a corner (2-D alignment) makes the warp execute both, and features of different pyramid levels diverge in the warp / scan loops.
// Measured on the B200 (512 k features, 25 % edgelets, levels 0-2): findMatchDirect 1.195 ms in the caller's order, 0.954 ms grouped by
// (type, level); the epipolar search 1.314 -> 1.184 ms. The grouping is a stable counting sort of the feature INDICES over 16 buckets
// (order inside a bucket = the caller's order, which keeps the features of a frame together: a random order costs the epipolar search
// 16 %); results are written by feature index, so callers see no difference.
constexpr int kOrderChunk = 1024, kOrderBuckets = 16, kOrderMinFeatures = 16384;
SVO_D int orderBucket(const svo_feature& f) { return (isEdgeletType(f.type) ? 0 : 8) + min(max(f.level, 0), 7); }

__global__ void __launch_bounds__(256) order_count_kernel(const svo_feature* __restrict__ ftrs, int M, int n_chunks, int* __restrict__ counts) {
  __shared__ int s_c[kOrderBuckets];
  if (threadIdx.x < kOrderBuckets) s_c[threadIdx.x] = 0;
  __syncthreads();
  for (int j = 0; j < kOrderChunk / 256; ++j) {
    const int i = blockIdx.x * kOrderChunk + j * 256 + threadIdx.x;
    if (i < M) atomicAdd(&s_c[orderBucket(ftrs[i])], 1);
  }
  __syncthreads();
  if (threadIdx.x < kOrderBuckets) counts[threadIdx.x * n_chunks + blockIdx.x] = s_c[threadIdx.x];  // bucket-major: scanned in (bucket, chunk) order
}

// exclusive scan of a[0..n) in place (one CTA)
__global__ void __launch_bounds__(1024) order_scan_kernel(int* a, int n) {
  __shared__ int s_w[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const int v = i < n ? a[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      s_w[lane] = w;
    }
    __syncthreads();
    const int carry = s_carry;
    if (i < n) a[i] = carry + (warp ? s_w[warp - 1] : 0) + inc - v;
    __syncthreads();
    if (tid == 1023) s_carry = carry + s_w[31];
    __syncthreads();
  }
}

// stable scatter: thread t of a chunk owns 4 consecutive features; per bucket an exclusive prefix over the threads (four buckets at a
// time in the 16-bit lanes of one 64-bit word: a chunk holds at most 1024 features of a bucket)
__global__ void __launch_bounds__(256) order_scatter_kernel(const svo_feature* __restrict__ ftrs, int M, int n_chunks, const int* __restrict__ offsets,
                                                            int* __restrict__ perm) {
  __shared__ unsigned long long s_w[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * kOrderChunk + 4 * tid;
  int key[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) key[j] = i0 + j < M ? orderBucket(ftrs[i0 + j]) : -1;
  for (int q = 0; q < kOrderBuckets / 4; ++q) {
    unsigned long long c = 0ull;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (key[j] >= 0 && (key[j] >> 2) == q) c += 1ull << (16 * (key[j] & 3));
    unsigned long long inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();  // s_w of the previous round has been read
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    unsigned long long base = inc - c;
    for (int w = 0; w < warp; ++w) base += s_w[w];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (key[j] >= 0 && (key[j] >> 2) == q) {
        const int b = key[j], sh = 16 * (b & 3);
        const int rank = (int)((base >> sh) & 0xFFFFull);
        perm[offsets[b * n_chunks + blockIdx.x] + rank] = i0 + j;
        base += 1ull << sh;
      }
  }
}

struct AlignOnlyParams {
  PyrView pyr;
  const int* frame_idx;
  const int* level;
  int M;
  const double* dir;  // null -> align2D
  const uint8_t* pwb;
  int n_iter, est_offset, est_gain;
  double* px;
  double* h_inv;
  uint8_t* converged;
};

__global__ void __launch_bounds__(kThreads) align_only_kernel(const AlignOnlyParams P) {
  __shared__ __align__(16) uint8_t s_pwb[kGroupsPerCta * kPwbPitch];
  const Group g = makeGroup();
  const int gi = threadIdx.x / kGroup;
  const int i = blockIdx.x * kGroupsPerCta + gi;
  if (i >= P.M) return;
  uint8_t* pwb = s_pwb + gi * kPwbPitch;
  for (int k = g.r; k < 100; k += kGroup) pwb[k] = P.pwb[100 * (size_t)i + k];
  __syncwarp(g.mask);
  const int frame = P.frame_idx ? P.frame_idx[i] : 0;
  const int level = P.level ? P.level[i] : 0;
  const ImgView img = levelView(P.pyr, frame, level);
  double x = P.px[2 * i], y = P.px[2 * i + 1];
  bool conv;
  double hinv = 0.0;
  if (P.dir) conv = align1D(g, img, P.dir[2 * i], P.dir[2 * i + 1], pwb, P.n_iter, P.est_offset != 0, P.est_gain != 0, x, y, &hinv);
  else conv = align2D(g, img, pwb, P.n_iter, P.est_offset != 0, P.est_gain != 0, x, y);
  if (g.r == 0) {
    P.px[2 * i] = x; P.px[2 * i + 1] = y;
    P.converged[i] = conv ? 1 : 0;
    if (P.h_inv) P.h_inv[i] = hinv;
  }
}

// Matcher::scanEpipolarLine on its own (matcher.cpp:324-488): one group per scan, the caller supplies the segment, the 8x8 reference
// patch and the members the scan reads (epi_length_pyramid_, options_).
struct ScanParams {
  PyrView cur_pyr;
  svo_camera cam_cur;
  const int* cur_frame_idx;
  int M;
  const double* A;
  const double* B;
  const double* C;
  const uint8_t* patch;       // [M][64]
  const int* patch_level;
  const double* epi_length_pyramid;
  svo_matcher_options opt;
  double* image_best;         // [M][2]
  int* zmssd_best;            // [M] in / out
};

__global__ void __launch_bounds__(kThreads) scan_epipolar_kernel(const ScanParams P) {
  __shared__ __align__(16) uint8_t s_pwb[kGroupsPerCta * kPwbPitch];
  const Group g = makeGroup();
  const int gi = threadIdx.x / kGroup;
  const int i = blockIdx.x * kGroupsPerCta + gi;
  if (i >= P.M) return;
  uint8_t* pwb = s_pwb + gi * kPwbPitch;
  // row r of the 8x8 patch into the interior of the 10x10 bordered layout makeZmssdRef reads
  for (int x = 0; x < 8; ++x) pwb[(g.r + 1) * 10 + 1 + x] = P.patch[64 * (size_t)i + 8 * g.r + x];
  __syncwarp(g.mask);
  const ZmssdRef zref = makeZmssdRef(g, pwb);
  const int pl = P.patch_level[i];
  EpiSetup e;
  memset(&e, 0, sizeof(e));
  e.epi_length_pyramid = P.epi_length_pyramid[i];
  const V3d A{P.A[3 * i], P.A[3 * i + 1], P.A[3 * i + 2]}, B{P.B[3 * i], P.B[3 * i + 1], P.B[3 * i + 2]};
  const V3d C{P.C[3 * i], P.C[3 * i + 1], P.C[3 * i + 2]};
  epiScanSetup(A, B, C, P.opt, e);
  const ImgView cur = levelView(P.cur_pyr, P.cur_frame_idx ? P.cur_frame_idx[i] : 0, pl);
  double px_x = 0.0, px_y = 0.0;
  const int z0 = P.zmssd_best[i];
  const int z = P.opt.scan_on_unit_sphere ? scanEpipolarUnitSphere(g, e, P.cam_cur, cur, pl, zref, px_x, px_y, z0)
                                          : scanEpipolarUnitPlane(g, e, P.cam_cur, cur, pl, zref, px_x, px_y, z0);
  if (g.r == 0) {
    P.image_best[2 * i] = px_x; P.image_best[2 * i + 1] = px_y;
    P.zmssd_best[i] = z;
  }
}

__global__ void stereo_entry_frames_kernel(const int* feat_begin, const int* frame0_idx, const int* frame1_idx, int* e0, int* e1, int* ep) {
  const int b = blockIdx.x;
  const int f0 = frame0_idx ? frame0_idx[b] : b, f1 = frame1_idx ? frame1_idx[b] : b;
  for (int i = feat_begin[b] + threadIdx.x; i < feat_begin[b + 1]; i += blockDim.x) { e0[i] = f0; e1[i] = f1; ep[i] = b; }
}

// After the entries at list positions [chunk_lo, chunk_hi) were matched: lists that have their n_desired successes are done, the
// reference's loop would have stopped inside this chunk (everything behind it stays unmatched = "not reached").
__global__ void stereo_progress_kernel(const svo_match_out* match, const int* feat_begin, const int* n_desired, int chunk_lo, int chunk_hi,
                                       int* succ, uint8_t* done) {
  __shared__ int s_cnt;
  const int b = blockIdx.x;
  if (done[b]) return;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const long long lo = (long long)feat_begin[b] + chunk_lo;
  const long long hi = min((long long)feat_begin[b + 1], (long long)feat_begin[b] + chunk_hi);
  int cnt = 0;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) cnt += match[i].result == 0;
  if (cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0) {
    succ[b] += s_cnt;
    if (succ[b] >= n_desired[b]) done[b] = 1;
  }
}

// StereoTriangulation::compute's sequential bookkeeping (stereo_triangulation.cpp:93-133) on top of speculative matches of ALL
// entries: one CTA per stereo pair scans its entries in visiting order; entry i is accepted iff it matched and fewer than n_desired
// entries before it did; everything behind the n_desired-th success is "not reached" (the reference breaks out of its loop).
__global__ void stereo_commit_kernel(const svo_match_out* match, const svo_feature* ftrs, const int* feat_begin, const int* n_desired,
                                     const int* n_features_in_frame1, const double* T_world_cam0, svo_stereo_result* out,
                                     svo_stereo_stats* stats) {
  __shared__ int s_warp[32];
  __shared__ int s_base, s_failed;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int lo = feat_begin[b], hi = feat_begin[b + 1], want = n_desired[b], slot0 = n_features_in_frame1[b];
  const SE3d T_w_c = se3Load(T_world_cam0 + 7 * (size_t)b);
  if (tid == 0) { s_base = 0; s_failed = 0; }
  __syncthreads();
  for (int c0 = lo; c0 < hi; c0 += blockDim.x) {
    const int i = c0 + tid;
    const bool ok = i < hi && match[i].result == 0;  // Matcher::MatchResult::kSuccess
    const bool hole = i < hi && match[i].result == -2;  // entry with a negative type: not part of the list
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_base;  // successes before this entry
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    before += __popc(bal & ((1u << lane) - 1u));
    int chunk = 0;
    for (int w = 0; w < nwarp; ++w) chunk += s_warp[w];
    if (i < hi) {
      svo_stereo_result r;
      memset(&r, 0, sizeof(r));
      r.slot = -1;
      r.match_result = -1;
      const bool reached = before < want;  // the loop is still running when it gets to entry i
      if (reached && !hole) {
        const svo_match_out m = match[i];
        r.match_result = m.result;
        if (ok) {
          const svo_feature ft = ftrs[i];
          const V3d xyz = se3Apply(T_w_c, V3d{ft.f[0] * m.depth, ft.f[1] * m.depth, ft.f[2] * m.depth});
          const V2d gc = normalized2(V2d{m.A_cur_ref[0] * ft.grad[0] + m.A_cur_ref[1] * ft.grad[1],
                                         m.A_cur_ref[2] * ft.grad[0] + m.A_cur_ref[3] * ft.grad[1]});
          r.status = SVO_STEREO_SUCCESS;
          r.slot = slot0 + before;
          r.depth = m.depth;
          r.xyz_world[0] = xyz.x; r.xyz_world[1] = xyz.y; r.xyz_world[2] = xyz.z;
          r.px_cur[0] = m.px_cur[0]; r.px_cur[1] = m.px_cur[1];
          r.f_cur[0] = m.f_cur[0]; r.f_cur[1] = m.f_cur[1]; r.f_cur[2] = m.f_cur[2];
          r.grad_cur[0] = gc.x; r.grad_cur[1] = gc.y;
          r.level = ft.level; r.type = ft.type;
        } else {
          r.status = SVO_STEREO_FAILED;
          atomicAdd(&s_failed, 1);
        }
      }
      out[i] = r;
    }
    __syncthreads();
    if (tid == 0) s_base += chunk;
    __syncthreads();
  }
  if (tid == 0) {
    stats[b].n_succeeded = min(s_base, want);
    stats[b].n_failed = s_failed;
  }
}

[[maybe_unused]] int checkLevels(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, const char* who) {
  if (!pyr) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, who);
  return SVO_OK;
}

}  // namespace

extern "C" {

int svo_cuda_align2d(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, const int* frame_idx, const int* level, int M,
                     const uint8_t* patch_with_border, int n_iter, int affine_est_offset, int affine_est_gain, double* px,
                     uint8_t* converged, svo_mem mem) {
  if (!ctx || !pyr || M < 0 || !patch_with_border || !px || !converged)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_align2d: bad arguments");
  if (M == 0) return SVO_OK;
  cudaSetDevice(ctx->device);
  Stager st(ctx, mem);
  AlignOnlyParams P;
  P.pyr = makeView(pyr);
  P.frame_idx = st.in(frame_idx, (size_t)M);
  P.level = st.in(level, (size_t)M);
  P.M = M;
  P.dir = nullptr;
  P.pwb = st.in(patch_with_border, (size_t)M * 100);
  P.n_iter = n_iter; P.est_offset = affine_est_offset; P.est_gain = affine_est_gain;
  P.px = st.inout(px, (size_t)M * 2);
  P.h_inv = nullptr;
  P.converged = st.out(converged, (size_t)M);
  if (!st.send()) return st.finish();
  align_only_kernel<<<(M + kGroupsPerCta - 1) / kGroupsPerCta, kThreads, 0, ctx->stream>>>(P);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

int svo_cuda_align1d(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, const int* frame_idx, const int* level, int M, const double* dir,
                     const uint8_t* patch_with_border, int n_iter, int affine_est_offset, int affine_est_gain, double* px,
                     double* h_inv, uint8_t* converged, svo_mem mem) {
  if (!ctx || !pyr || M < 0 || !dir || !patch_with_border || !px || !converged)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_align1d: bad arguments");
  if (M == 0) return SVO_OK;
  cudaSetDevice(ctx->device);
  Stager st(ctx, mem);
  AlignOnlyParams P;
  P.pyr = makeView(pyr);
  P.frame_idx = st.in(frame_idx, (size_t)M);
  P.level = st.in(level, (size_t)M);
  P.M = M;
  P.dir = st.in(dir, (size_t)M * 2);
  P.pwb = st.in(patch_with_border, (size_t)M * 100);
  P.n_iter = n_iter; P.est_offset = affine_est_offset; P.est_gain = affine_est_gain;
  P.px = st.inout(px, (size_t)M * 2);
  P.h_inv = st.out(h_inv, (size_t)M);
  P.converged = st.out(converged, (size_t)M);
  if (!st.send()) return st.finish();
  align_only_kernel<<<(M + kGroupsPerCta - 1) / kGroupsPerCta, kThreads, 0, ctx->stream>>>(P);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

static int matchCommon(int mode, svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const svo_cuda_pyr* cur_pyr, const int* ref_frame_idx,
                       const int* cur_frame_idx, const svo_camera* cam_ref, const svo_camera* cam_cur, const double* T_cur_ref,
                       const int* T_idx, int n_T, int M, const svo_feature* ftrs, const double* depth, size_t depth_per,
                       const double* px_guess, const svo_matcher_options* opt, svo_match_out* out, double* A_out, int* sl_out,
                       uint8_t* pwb_out, uint8_t* ok_out, svo_mem mem) {
  if (M == 0) return SVO_OK;
  cudaSetDevice(ctx->device);
  Stager st(ctx, mem);
  MatchParams P;
  memset(&P, 0, sizeof(P));
  P.ref_pyr = makeView(ref_pyr);
  P.cur_pyr = makeView(cur_pyr ? cur_pyr : ref_pyr);
  P.cam_ref = *cam_ref;
  P.cam_cur = *cam_cur;
  P.ref_frame_idx = st.in(ref_frame_idx, (size_t)M);
  P.cur_frame_idx = st.in(cur_frame_idx, (size_t)M);
  P.T_cur_ref = st.in(T_cur_ref, (size_t)n_T * 7);
  P.T_idx = st.in(T_idx, (size_t)M);
  P.M = M;
  P.ftrs = st.in(ftrs, (size_t)M);
  P.depth = st.in(depth, (size_t)M * depth_per);
  P.px_guess = st.in(px_guess, (size_t)M * 2);
  if (opt) P.opt = *opt;
  P.out = st.outWriteOnly(out, (size_t)M);
  P.A_out = st.out(A_out, (size_t)M * 4);
  P.search_level_out = st.out(sl_out, (size_t)M);
  P.pwb_out = st.out(pwb_out, (size_t)M * 100);
  P.ok_out = st.out(ok_out, (size_t)M);
  int* d_perm = nullptr;
  int* d_counts = nullptr;
  const int n_chunks = (M + kOrderChunk - 1) / kOrderChunk;
  const bool ordered = (mode == 0 || mode == 1) && M >= kOrderMinFeatures;
  if (ordered) {
    d_perm = (int*)st.scratch(sizeof(int) * (size_t)M);
    d_counts = (int*)st.scratch(sizeof(int) * (size_t)kOrderBuckets * n_chunks);
  }
  if (!st.send() || (ordered && (!d_perm || !d_counts))) return st.finish();
  if (ordered) {
    order_count_kernel<<<n_chunks, 256, 0, ctx->stream>>>(P.ftrs, M, n_chunks, d_counts);
    SVO_LAUNCH_CHECK(ctx);
    order_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, kOrderBuckets * n_chunks);
    SVO_LAUNCH_CHECK(ctx);
    order_scatter_kernel<<<n_chunks, 256, 0, ctx->stream>>>(P.ftrs, M, n_chunks, d_counts, d_perm);
    SVO_LAUNCH_CHECK(ctx);
    P.perm = d_perm;
  }
  const int grid = (M + kGroupsPerCta - 1) / kGroupsPerCta;
  if (mode == 0) match_kernel<0><<<grid, kThreads, 0, ctx->stream>>>(P);
  else if (mode == 1 && P.opt.scan_on_unit_sphere) match_kernel<1, 1><<<grid, kThreads, 0, ctx->stream>>>(P);
  else if (mode == 1) match_kernel<1, 0><<<grid, kThreads, 0, ctx->stream>>>(P);
  else match_kernel<2><<<grid, kThreads, 0, ctx->stream>>>(P);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

// Number of transformations referenced by T_idx (host copy needed only to size the staging of T_cur_ref).
static int countT(const int* T_idx, int M, svo_mem mem, int* n_T) {
  if (!T_idx) { *n_T = 1; return SVO_OK; }
  if (mem == SVO_MEM_DEVICE) { *n_T = 0; return SVO_OK; }  // device arrays are used in place, no size needed
  int mx = 0;
  for (int i = 0; i < M; ++i) mx = T_idx[i] > mx ? T_idx[i] : mx;
  *n_T = mx + 1;
  return SVO_OK;
}

int svo_cuda_scan_epipolar_line(svo_cuda_ctx* ctx, const svo_cuda_pyr* cur_pyr, const int* cur_frame_idx, const svo_camera* cam_cur, int M,
                                const double* A, const double* B, const double* C, const uint8_t* patch, const int* patch_level,
                                const double* epi_length_pyramid, const svo_matcher_options* opt, double* image_best, int* zmssd_best,
                                svo_mem mem) {
  if (!ctx || !cur_pyr || !cam_cur || M < 0 || !A || !B || !C || !patch || !patch_level || !epi_length_pyramid || !opt || !image_best ||
      !zmssd_best)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_scan_epipolar_line: bad arguments");
  if (M == 0) return SVO_OK;
  SVO_BIND(ctx);
  Stager st(ctx, mem);
  ScanParams P;
  memset(&P, 0, sizeof(P));
  P.cur_pyr = makeView(cur_pyr);
  P.cam_cur = *cam_cur;
  P.cur_frame_idx = st.in(cur_frame_idx, (size_t)M);
  P.M = M;
  P.A = st.in(A, (size_t)M * 3); P.B = st.in(B, (size_t)M * 3); P.C = st.in(C, (size_t)M * 3);
  P.patch = st.in(patch, (size_t)M * 64);
  P.patch_level = st.in(patch_level, (size_t)M);
  P.epi_length_pyramid = st.in(epi_length_pyramid, (size_t)M);
  P.opt = *opt;
  P.image_best = st.outWriteOnly(image_best, (size_t)M * 2);
  P.zmssd_best = st.inout(zmssd_best, (size_t)M);
  if (!st.send()) return st.finish();
  scan_epipolar_kernel<<<(M + kGroupsPerCta - 1) / kGroupsPerCta, kThreads, 0, ctx->stream>>>(P);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

int svo_cuda_warp_affine(svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const int* ref_frame_idx, const svo_camera* cam_ref,
                         const svo_camera* cam_cur, const double* T_cur_ref, const int* T_idx, int M, const svo_feature* ftrs,
                         const double* depth, double* A_out, int* search_level_out, uint8_t* patch_with_border_out, uint8_t* ok_out,
                         svo_mem mem) {
  if (!ctx || !ref_pyr || !cam_ref || !cam_cur || !T_cur_ref || M < 0 || !ftrs || !depth || !A_out || !search_level_out ||
      !patch_with_border_out || !ok_out)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_warp_affine: bad arguments");
  int n_T;
  countT(T_idx, M, mem, &n_T);
  return matchCommon(2, ctx, ref_pyr, nullptr, ref_frame_idx, nullptr, cam_ref, cam_cur, T_cur_ref, T_idx, n_T, M, ftrs, depth, 1,
                     nullptr, nullptr, nullptr, A_out, search_level_out, patch_with_border_out, ok_out, mem);
}

int svo_cuda_find_match_direct(svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const svo_cuda_pyr* cur_pyr, const int* ref_frame_idx,
                               const int* cur_frame_idx, const svo_camera* cam_ref, const svo_camera* cam_cur, const double* T_cur_ref,
                               const int* T_idx, int M, const svo_feature* ftrs, const double* ref_depth, const double* px_cur_guess,
                               const svo_matcher_options* opt, svo_match_out* out, svo_mem mem) {
  if (!ctx || !ref_pyr || !cur_pyr || !cam_ref || !cam_cur || !T_cur_ref || M < 0 || !ftrs || !ref_depth || !px_cur_guess || !opt || !out)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_find_match_direct: bad arguments");
  int n_T;
  countT(T_idx, M, mem, &n_T);
  return matchCommon(0, ctx, ref_pyr, cur_pyr, ref_frame_idx, cur_frame_idx, cam_ref, cam_cur, T_cur_ref, T_idx, n_T, M, ftrs, ref_depth,
                     1, px_cur_guess, opt, out, nullptr, nullptr, nullptr, nullptr, mem);
}

int svo_cuda_find_epipolar_match_direct(svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const svo_cuda_pyr* cur_pyr,
                                        const int* ref_frame_idx, const int* cur_frame_idx, const svo_camera* cam_ref,
                                        const svo_camera* cam_cur, const double* T_cur_ref, const int* T_idx, int M,
                                        const svo_feature* ftrs, const double* d_inv, const svo_matcher_options* opt, svo_match_out* out,
                                        svo_mem mem) {
  if (!ctx || !ref_pyr || !cur_pyr || !cam_ref || !cam_cur || !T_cur_ref || M < 0 || !ftrs || !d_inv || !opt || !out)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_find_epipolar_match_direct: bad arguments");
  int n_T;
  countT(T_idx, M, mem, &n_T);
  return matchCommon(1, ctx, ref_pyr, cur_pyr, ref_frame_idx, cur_frame_idx, cam_ref, cam_cur, T_cur_ref, T_idx, n_T, M, ftrs, d_inv, 3,
                     nullptr, opt, out, nullptr, nullptr, nullptr, nullptr, mem);
}

int svo_cuda_stereo_triangulate(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr0, const svo_cuda_pyr* pyr1, const int* frame0_idx,
                                const int* frame1_idx, const svo_camera* cam0, const svo_camera* cam1, const double* T_f1f0,
                                const double* T_world_cam0, int B, const int* feat_begin, int n_features, const svo_feature* ftrs,
                                const int* n_desired, const int* n_features_in_frame1, double mean_depth_inv, double min_depth_inv,
                                double max_depth_inv, const svo_matcher_options* mopt, svo_stereo_result* results,
                                svo_stereo_stats* stats, svo_mem mem) {
  if (!ctx || !pyr0 || !pyr1 || !cam0 || !cam1 || !T_f1f0 || !T_world_cam0 || B < 0 || !feat_begin || n_features < 0 || !n_desired ||
      !n_features_in_frame1 || !mopt || !stats || (n_features > 0 && (!ftrs || !results)))
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_stereo_triangulate: bad arguments");
  if (B == 0) return SVO_OK;
  cudaSetDevice(ctx->device);
  Stager st(ctx, mem);
  const int* d_begin = st.in(feat_begin, (size_t)B + 1);
  const svo_feature* d_ftrs = st.in(ftrs, (size_t)n_features);
  const int* d_want = st.in(n_desired, (size_t)B);
  const int* d_slot0 = st.in(n_features_in_frame1, (size_t)B);
  const double* d_Twc = st.in(T_world_cam0, (size_t)B * 7);
  const int* d_f0 = st.in(frame0_idx, (size_t)B);
  const int* d_f1 = st.in(frame1_idx, (size_t)B);
  svo_stereo_result* d_res = st.out(results, (size_t)n_features);
  svo_stereo_stats* d_stats = st.out(stats, (size_t)B);
  // per-entry frame indices (entry -> pair) and the shared transformation / depth triple, built on the device side of the call
  const double h_shared[10] = {T_f1f0[0], T_f1f0[1], T_f1f0[2], T_f1f0[3], T_f1f0[4], T_f1f0[5], T_f1f0[6], mean_depth_inv, min_depth_inv,
                               max_depth_inv};
  double* d_shared = (double*)st.scratch(sizeof(h_shared));
  int* d_e0 = (int*)st.scratch(sizeof(int) * (size_t)(n_features > 0 ? n_features : 1));
  int* d_e1 = (int*)st.scratch(sizeof(int) * (size_t)(n_features > 0 ? n_features : 1));
  svo_match_out* d_match = (svo_match_out*)st.scratch(sizeof(svo_match_out) * (size_t)(n_features > 0 ? n_features : 1));
  int* d_ep = (int*)st.scratch(sizeof(int) * (size_t)(n_features > 0 ? n_features : 1));
  int* d_succ = (int*)st.scratch(sizeof(int) * (size_t)B);
  uint8_t* d_done = (uint8_t*)st.scratch((size_t)B);
  if (!st.send() || !d_shared || !d_e0 || !d_e1 || !d_match || !d_ep || !d_succ || !d_done) return st.finish();
  SVO_CUDA_TRY(ctx, cudaMemcpyAsync(d_shared, h_shared, sizeof(h_shared), cudaMemcpyHostToDevice, ctx->stream));
  if (n_features > 0) {
    stereo_entry_frames_kernel<<<B, 128, 0, ctx->stream>>>(d_begin, d_f0, d_f1, d_e0, d_e1, d_ep);
    SVO_LAUNCH_CHECK(ctx);
    // unmatched entries read as result = -1 (neither a success nor a hole)
    SVO_CUDA_TRY(ctx, cudaMemsetAsync(d_match, 0xFF, sizeof(svo_match_out) * (size_t)n_features, ctx->stream));
    SVO_CUDA_TRY(ctx, cudaMemsetAsync(d_succ, 0, sizeof(int) * (size_t)B, ctx->stream));
    SVO_CUDA_TRY(ctx, cudaMemsetAsync(d_done, 0, (size_t)B, ctx->stream));
    MatchParams P;
    memset(&P, 0, sizeof(P));
    P.ref_pyr = makeView(pyr0);
    P.cur_pyr = makeView(pyr1);
    P.cam_ref = *cam0;
    P.cam_cur = *cam1;
    P.ref_frame_idx = d_e0;
    P.cur_frame_idx = d_e1;
    P.T_cur_ref = d_shared;
    P.T_idx = nullptr;
    P.M = n_features;
    P.ftrs = d_ftrs;
    P.depth = d_shared + 7;
    P.depth_shared = 1;
    P.align1d_from_type = 1;
    P.opt = *mopt;
    P.out = d_match;
    P.entry_pair = d_ep;
    P.pair_begin = d_begin;
    P.pair_done = d_done;
    // Progressive matching: the reference stops a list at its n_desired-th success, so the lists are matched in chunks of list
    // positions and a list drops out once it has its successes (4 chunks, then the rest in one launch). Matching everything at
    // once visits ~3x the entries of the reference's loop at the default 120 of ~390 features.
#ifndef SVO_STEREO_CHUNK
#define SVO_STEREO_CHUNK 160  // A/B on the B200 (tools/time_stereo.py, chain keyframe stage per 2048 pairs): 96 -> 4.09 ms, 128 -> 4.81, 160 -> 3.66, 224 -> 4.42
#endif
    constexpr int kChunk = SVO_STEREO_CHUNK, kChunks = 4;
    for (int c = 0; c <= kChunks; ++c) {
      P.chunk_lo = c * kChunk;
      P.chunk_hi = c < kChunks ? (c + 1) * kChunk : 0x7FFFFFFF;
      if (P.opt.scan_on_unit_sphere) match_kernel<1, 1><<<(n_features + kGroupsPerCta - 1) / kGroupsPerCta, kThreads, 0, ctx->stream>>>(P);
      else match_kernel<1, 0><<<(n_features + kGroupsPerCta - 1) / kGroupsPerCta, kThreads, 0, ctx->stream>>>(P);
      SVO_LAUNCH_CHECK(ctx);
      if (c < kChunks) {
        stereo_progress_kernel<<<B, 128, 0, ctx->stream>>>(d_match, d_begin, d_want, P.chunk_lo, P.chunk_hi, d_succ, d_done);
        SVO_LAUNCH_CHECK(ctx);
      }
    }
  }
  stereo_commit_kernel<<<B, 256, 0, ctx->stream>>>(d_match, d_ftrs, d_begin, d_want, d_slot0, d_Twc, d_res, d_stats);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

}  // extern "C"
