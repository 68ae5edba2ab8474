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
  out[i] = o;
}

// mode 0: findMatchDirect, 1: findEpipolarMatchDirect, 2: warp only
template <int MODE>
// 126 registers, 4 CTAs/SM: fastest of 4/5/6 on the B200 (1.58 / 2.94 ms for 512 k features)
__global__ void __launch_bounds__(kThreads, 4) match_kernel(const MatchParams P) {
  __shared__ __align__(16) uint8_t s_pwb[kGroupsPerCta * kPwbPitch];
  const Group g = makeGroup();
  const int gi = threadIdx.x / kGroup;
  const int i = blockIdx.x * kGroupsPerCta + gi;
  if (i >= P.M) return;
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
    double depth = 0.0;
    const int res = findEpipolarMatchDirect(g, P.ref_pyr, rf, P.cur_pyr, cf, P.cam_ref, P.cam_cur, T, ft, P.depth[3 * i],
                                            P.depth[3 * i + 1], P.depth[3 * i + 2], P.opt, P.opt.align_1d != 0, pwb, m, &depth);
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

int checkLevels(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, const char* who) {
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
  if (st.failed()) return st.finish();
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
  if (st.failed()) return st.finish();
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
  P.out = st.out(out, (size_t)M);
  P.A_out = st.out(A_out, (size_t)M * 4);
  P.search_level_out = st.out(sl_out, (size_t)M);
  P.pwb_out = st.out(pwb_out, (size_t)M * 100);
  P.ok_out = st.out(ok_out, (size_t)M);
  if (st.failed()) return st.finish();
  const int grid = (M + kGroupsPerCta - 1) / kGroupsPerCta;
  if (mode == 0) match_kernel<0><<<grid, kThreads, 0, ctx->stream>>>(P);
  else if (mode == 1) match_kernel<1><<<grid, kThreads, 0, ctx->stream>>>(P);
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

}  // extern "C"
