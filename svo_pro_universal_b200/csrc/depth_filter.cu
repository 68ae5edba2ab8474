// (d) Depth-filter seed update: Vogiatzis Gaussian x Beta filter on inverse depth, driven by the epipolar matcher.
//
// ref: src/svo_direct/src/depth_filter.cpp:200-249 (updateSeeds), :367-499 (updateSeed), :501-552 (updateFilterVogiatzis),
//      :554-578 (updateFilterGaussian), :580-596 (computeTau)
//      src/svo_common/include/svo/common/seed.h:110-169 (inverse-depth parametrisation)
//      src/vikit/vikit_common/include/vikit/math_utils.h:186-194 (normPdf)
//
// Kernels:
//   vogiatzis_kernel      one thread per independent update; state is streamed as 2 x double2 (80 B per update in+out).
//   compute_tau_kernel    one thread per (T_ref_cur, f, z).
//   update_seeds_kernel   one 8-lane group per seed walks that seed's observations IN ORDER (the filter is sequential per
//                         seed, seeds are independent): visibility gate, epipolar match (matcher_dev.cuh), tau, filter
//                         update, convergence flag — the whole depth_filter_utils::updateSeed without leaving the device.
#include "depth_filter_dev.cuh"

using namespace svo_dev;

namespace {

__global__ void __launch_bounds__(256) vogiatzis_kernel(int n, const double* __restrict__ z, const double* __restrict__ tau2,
                                                        const double* __restrict__ mu_range, double* __restrict__ state,
                                                        uint8_t* __restrict__ ok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2* sp = reinterpret_cast<double2*>(state) + 2 * (size_t)i;
  const double2 s01 = sp[0], s23 = sp[1];
  double s[4] = {s01.x, s01.y, s23.x, s23.y};
  const bool r = updateFilterVogiatzis(z[i], tau2[i], mu_range[i], s);
  sp[0] = make_double2(s[0], s[1]);
  sp[1] = make_double2(s[2], s[3]);
  if (ok) ok[i] = r ? 1 : 0;
}

__global__ void __launch_bounds__(256) compute_tau_kernel(int n, const double* __restrict__ T_ref_cur, const double* __restrict__ f,
                                                          const double* __restrict__ z, double px_error_angle, double* __restrict__ tau) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3d t{T_ref_cur[7 * (size_t)i + 4], T_ref_cur[7 * (size_t)i + 5], T_ref_cur[7 * (size_t)i + 6]};
  const V3d fv{f[3 * (size_t)i], f[3 * (size_t)i + 1], f[3 * (size_t)i + 2]};
  tau[i] = computeTau(t, fv, z[i], px_error_angle);
}

constexpr int kThreads = 128;
constexpr int kGroupsPerCta = kThreads / kGroup;

struct SeedParams {
  PyrView ref_pyr, cur_pyr;
  svo_camera cam_ref, cam_cur;
  int S, n_obs;
  const int* ref_frame_idx;
  const svo_feature* ftrs;
  uint8_t* types;
  double* state;
  const double* seed_mu_range;
  const int* obs_frame_idx;
  const int* obs_T_idx;
  const double* T_cur_ref;
  svo_matcher_options mopt;
  svo_depth_filter_options dopt;
  double px_error_angle;
  int* n_success;
  int* match_results;
};

// 80 registers: measured 27.4 ms vs 28.9 ms at the compiler default (150 registers) for 3.2 M seed-observations
__global__ void __launch_bounds__(kThreads, 6) update_seeds_kernel(const SeedParams P) {
  __shared__ __align__(16) uint8_t s_pwb[kGroupsPerCta * kPwbPitch];
  const Group g = makeGroup();
  const int gi = threadIdx.x / kGroup;
  const int s = blockIdx.x * kGroupsPerCta + gi;
  if (s >= P.S) return;
  uint8_t* pwb = s_pwb + gi * kPwbPitch;
  svo_feature ft = P.ftrs[s];
  int type = P.types[s];
  double st[4] = {P.state[4 * (size_t)s], P.state[4 * (size_t)s + 1], P.state[4 * (size_t)s + 2], P.state[4 * (size_t)s + 3]};
  const double mu_range = P.seed_mu_range[s];
  const int rf = P.ref_frame_idx ? P.ref_frame_idx[s] : 0;
  int n_ok = 0;
  for (int o = 0; o < P.n_obs; ++o) {
    const size_t oi = (size_t)o * P.S + s;
    int mr = -1;
    const int cf = P.obs_frame_idx[oi];
    if (cf >= 0) {  // depth_filter.cpp:377-381 (cur frame == ref frame): the caller marks such observations with a negative index
      const SE3d T = se3Load(P.T_cur_ref + 7 * (size_t)P.obs_T_idx[oi]);
      // DepthFilter::updateSeeds picks the threshold by type (:214-221)
      const double cur_thresh = (type == kMapPointSeed || type == kMapPointSeedConverged) ? P.dopt.mappoint_convergence_sigma2_thresh
                                                                                             : P.dopt.seed_convergence_sigma2_thresh;
      MatchState m;
      if (updateSeedOnce(g, P.ref_pyr, rf, P.cur_pyr, cf, P.cam_ref, P.cam_cur, T, ft, type, st, mu_range, cur_thresh, P.px_error_angle,
                         P.dopt.check_visibility != 0, P.dopt.check_convergence != 0, P.dopt.use_vogiatzis_update != 0, P.mopt, pwb, m, &mr))
        ++n_ok;
    }
    if (P.match_results && g.r == 0) P.match_results[oi] = mr;
  }
  if (g.r == 0) {
    P.types[s] = (uint8_t)type;
    P.state[4 * (size_t)s] = st[0]; P.state[4 * (size_t)s + 1] = st[1];
    P.state[4 * (size_t)s + 2] = st[2]; P.state[4 * (size_t)s + 3] = st[3];
    if (n_ok) atomicAdd(P.n_success, n_ok);
  }
}

}  // namespace

extern "C" {

int svo_cuda_update_filter_vogiatzis(svo_cuda_ctx* ctx, int n, const double* z, const double* tau2, const double* mu_range,
                                     double* state, uint8_t* ok, svo_mem mem) {
  if (!ctx || n < 0 || !z || !tau2 || !mu_range || !state)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_update_filter_vogiatzis: bad arguments");
  if (n == 0) return SVO_OK;
  cudaSetDevice(ctx->device);
  Stager st(ctx, mem);
  const double* dz = st.in(z, (size_t)n);
  const double* dt = st.in(tau2, (size_t)n);
  const double* dm = st.in(mu_range, (size_t)n);
  double* ds = st.inout(state, (size_t)n * 4);
  uint8_t* dok = st.out(ok, (size_t)n);
  if (st.failed()) return st.finish();
  vogiatzis_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, dz, dt, dm, ds, dok);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

int svo_cuda_compute_tau(svo_cuda_ctx* ctx, int n, const double* T_ref_cur, const double* f, const double* z, double px_error_angle,
                         double* tau, svo_mem mem) {
  if (!ctx || n < 0 || !T_ref_cur || !f || !z || !tau) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_compute_tau: bad arguments");
  if (n == 0) return SVO_OK;
  cudaSetDevice(ctx->device);
  Stager st(ctx, mem);
  const double* dT = st.in(T_ref_cur, (size_t)n * 7);
  const double* df = st.in(f, (size_t)n * 3);
  const double* dz = st.in(z, (size_t)n);
  double* dtau = st.out(tau, (size_t)n);
  if (st.failed()) return st.finish();
  compute_tau_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, dT, df, dz, px_error_angle, dtau);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

int svo_cuda_update_seeds(svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const svo_cuda_pyr* cur_pyr, const svo_camera* cam_ref,
                          const svo_camera* cam_cur, int S, const int* ref_frame_idx, const svo_feature* ftrs, uint8_t* types,
                          double* state, const double* seed_mu_range, int n_obs, const int* obs_frame_idx, const int* obs_T_idx,
                          const double* T_cur_ref, const svo_matcher_options* mopt, const svo_depth_filter_options* dopt, int* n_success,
                          int* match_results, svo_mem mem) {
  if (!ctx || !ref_pyr || !cur_pyr || !cam_ref || !cam_cur || S < 0 || !ftrs || !types || !state || !seed_mu_range || n_obs < 0 ||
      !obs_frame_idx || !obs_T_idx || !T_cur_ref || !mopt || !dopt || !n_success)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_update_seeds: bad arguments");
  cudaSetDevice(ctx->device);
  Stager st(ctx, mem);
  SeedParams P;
  memset(&P, 0, sizeof(P));
  P.ref_pyr = makeView(ref_pyr);
  P.cur_pyr = makeView(cur_pyr);
  P.cam_ref = *cam_ref;
  P.cam_cur = *cam_cur;
  P.S = S; P.n_obs = n_obs;
  const size_t so = (size_t)S * n_obs;
  int n_T = 0;
  if (mem == SVO_MEM_HOST) {
    for (size_t i = 0; i < so; ++i) n_T = obs_T_idx[i] + 1 > n_T ? obs_T_idx[i] + 1 : n_T;
  }
  P.ref_frame_idx = st.in(ref_frame_idx, (size_t)S);
  P.ftrs = st.in(ftrs, (size_t)S);
  P.types = st.inout(types, (size_t)S);
  P.state = st.inout(state, (size_t)S * 4);
  P.seed_mu_range = st.in(seed_mu_range, (size_t)S);
  P.obs_frame_idx = st.in(obs_frame_idx, so);
  P.obs_T_idx = st.in(obs_T_idx, so);
  P.T_cur_ref = st.in(T_cur_ref, (size_t)n_T * 7);
  P.mopt = *mopt;
  P.dopt = *dopt;
  P.px_error_angle = dopt->px_error_angle > 0.0 ? dopt->px_error_angle : camAngleError(*cam_cur, 1.0);
  int* d_ns = st.out(n_success, 1);
  P.n_success = d_ns;
  P.match_results = st.out(match_results, so);
  if (st.failed()) return st.finish();
  SVO_CUDA_TRY(ctx, cudaMemsetAsync(d_ns, 0, sizeof(int), ctx->stream));
  if (S > 0 && n_obs > 0) {
    update_seeds_kernel<<<(S + kGroupsPerCta - 1) / kGroupsPerCta, kThreads, 0, ctx->stream>>>(P);
    SVO_LAUNCH_CHECK(ctx);
  }
  return st.finish();
}

}  // extern "C"
