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
#include "matcher_dev.cuh"

using namespace svo_dev;

namespace {

SVO_D double normPdf(double x, double mean, double sigma) {
  double exponent = x - mean;
  exponent *= -exponent;
  exponent /= 2 * sigma * sigma;
  double result = exp(exponent);
  result /= sigma * sqrt(2 * 3.14159265358979323846);
  return result;
}

// depth_filter.cpp:501-552; s = (mu, sigma2, a, b) in/out
SVO_D bool updateFilterVogiatzis(double z, double tau2, double mu_range, double s[4]) {
  double mu = s[0], sigma2 = s[1], a = s[2], b = s[3];
  const double norm_scale = sqrt(sigma2 + tau2);
  if (norm_scale != norm_scale) return false;
  const double oldsigma2 = sigma2;
  const double s2 = 1.0 / (1.0 / sigma2 + 1.0 / tau2);
  const double m = s2 * (mu / sigma2 + z / tau2);
  const double uniform_x = 1.0 / mu_range;
  double C1 = a / (a + b) * normPdf(z, mu, norm_scale);
  double C2 = b / (a + b) * uniform_x;
  const double normalization_constant = C1 + C2;
  C1 /= normalization_constant;
  C2 /= normalization_constant;
  const double f = C1 * (a + 1.0) / (a + b + 1.0) + C2 * a / (a + b + 1.0);
  const double e = C1 * (a + 1.0) * (a + 2.0) / ((a + b + 1.0) * (a + b + 2.0)) + C2 * a * (a + 1.0) / ((a + b + 1.0) * (a + b + 2.0));
  const double mu_new = C1 * m + C2 * mu;
  sigma2 = C1 * (s2 + m * m) + C2 * (sigma2 + mu * mu) - mu_new * mu_new;
  mu = mu_new;
  a = (e - f) / (f - e / f);
  b = a * (1.0 - f) / f;
  bool ok = true;
  if (sigma2 < 0.0) sigma2 = oldsigma2;
  if (mu < 0.0) { mu = 1.0; ok = false; }
  s[0] = mu; s[1] = sigma2; s[2] = a; s[3] = b;
  return ok;
}

// depth_filter.cpp:554-578
SVO_D bool updateFilterGaussian(double z, double tau2, double s[4]) {
  const double norm_scale = sqrt(s[1] + tau2);
  if (norm_scale != norm_scale) return false;
  const double denom = s[1] + tau2;
  s[0] = (s[1] * z + tau2 * s[0]) / denom;
  s[1] = s[1] * tau2 / denom;
  return true;
}

// depth_filter.cpp:580-596
SVO_D double computeTau(const V3d& t, const V3d& f, double z, double px_error_angle) {
  const V3d a = f * z - t;
  const double t_norm = norm3(t);
  const double a_norm = norm3(a);
  const double alpha = acos(dot3(f, t) / t_norm);
  const double beta = acos(dot3(a, -t) / (t_norm * a_norm));
  const double beta_plus = beta + px_error_angle;
  const double gamma_plus = 3.14159265358979323846 - alpha - beta_plus;
  const double z_plus = t_norm * sin(beta_plus) / sin(gamma_plus);
  return z_plus - z;
}

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
  const V3d f_ref{ft.f[0], ft.f[1], ft.f[2]};
  int n_ok = 0;
  for (int o = 0; o < P.n_obs; ++o) {
    const size_t oi = (size_t)o * P.S + s;
    int mr = -1;
    const int cf = P.obs_frame_idx[oi];
    bool done = cf < 0;  // depth_filter.cpp:377-381 (cur frame == ref frame): the caller marks such observations
    // :387-399
    if (!done && type == kOutlier) done = true;
    if (!done && (type == kCornerSeedConverged || type == kEdgeletSeedConverged || type == kMapPointSeedConverged) &&
        P.dopt.check_convergence)
      done = true;
    if (!done) {
      const SE3d T = se3Load(P.T_cur_ref + 7 * (size_t)P.obs_T_idx[oi]);
      bool visible = true;
      if (P.dopt.check_visibility) {  // :406-420
        const V3d xyz_f = se3Apply(T, f_ref * (1.0 / st[0]));
        const V2d px = camProject3(P.cam_cur, xyz_f);
        visible = px.x >= 0.0 && px.y >= 0.0 && px.x < (double)P.cam_cur.width && px.y < (double)P.cam_cur.height;
        if (visible) {
          const int pxi0 = (int)px.x, pxi1 = (int)px.y;
          const int boundary = 9;
          visible = pxi0 >= boundary && pxi1 >= boundary && pxi0 < P.cam_cur.width - boundary && pxi1 < P.cam_cur.height - boundary;
        }
      }
      if (visible) {
        const bool align_1d = (type == kEdgeletSeed || type == kEdgeletSeedConverged);  // :423-427
        ft.type = type;
        MatchState m;
        initMatchState(m);
        double depth = 0.0;
        // seed.h:115-128: d_estimate_inv = mu, d_min_inv = mu + sigma, d_max_inv = max(mu - sigma, 1e-8)
        const double sig = sqrt(st[1]);
        mr = findEpipolarMatchDirect(g, P.ref_pyr, rf, P.cur_pyr, cf, P.cam_ref, P.cam_cur, T, ft, st[0], st[0] + sig,
                                     fmax(st[0] - sig, 0.00000001), P.mopt, align_1d, pwb, m, &depth);
        if (mr != kSuccess) {
          if (!m.reject) st[3] += 1;  // seed::increaseOutlierProbability, :445-450
        } else {
          const SE3d T_ref_cur = se3Inv(T);
          const double depth_sigma = computeTau(T_ref_cur.t, f_ref, depth, P.px_error_angle);  // :459
          const double zi = 1.0 / depth;
          // seed::getSigma2FromDepthSigma (seed.h:155-160)
          const double sg = 0.5 * (1.0 / fmax(0.000000000001, depth - depth_sigma) - 1.0 / (depth + depth_sigma));
          const double tau2 = sg * sg;
          const bool ok = P.dopt.use_vogiatzis_update ? updateFilterVogiatzis(zi, tau2, mu_range, st) : updateFilterGaussian(zi, tau2, st);
          if (!ok) {
            type = kOutlier;  // :470-471, :481-482
          } else {
            // DepthFilter::updateSeeds picks the threshold by type (:214-221); isConverged: seed.h:145-153
            const double cur_thresh = (type == kMapPointSeed || type == kMapPointSeedConverged)
                                          ? P.dopt.mappoint_convergence_sigma2_thresh : P.dopt.seed_convergence_sigma2_thresh;
            const double thresh = mu_range / cur_thresh;
            if (st[1] < thresh * thresh) {
              if (type == kCornerSeed) type = kCornerSeedConverged;
              else if (type == kEdgeletSeed) type = kEdgeletSeedConverged;
              else if (type == kMapPointSeed) type = kMapPointSeedConverged;
            }
            ++n_ok;
          }
        }
      }
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
