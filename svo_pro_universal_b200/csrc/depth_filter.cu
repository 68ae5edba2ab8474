// (d) Depth-filter seed update: Vogiatzis Gaussian x Beta filter on inverse depth, driven by the epipolar matcher.
//
// ref: src/svo_direct/src/depth_filter.cpp:200-249 (updateSeeds), :367-499 (updateSeed), :501-552 (updateFilterVogiatzis),
//      :554-578 (updateFilterGaussian), :580-596 (computeTau)
//      src/svo_common/include/svo/common/seed.h:110-169 (inverse-depth parametrisation)
//      src/vikit/vikit_common/include/vikit/math_utils.h:186-194 (normPdf)
//
// Kernels:
//   vogiatzis_kernel      one thread per independent update; state is streamed as 2 x double2 (80 B per update in+out).
//   filter_seq_kernel     one thread per seed applies n_obs ORDERED updates with the state in registers (16 B streamed per update).
//   compute_tau_kernel    one thread per (T_ref_cur, f, z).
//   seed_step_kernel /    depth_filter_utils::updateSeed, phased. The filter is sequential per seed and seeds are independent, so
//   seed_match_kernel     observation o of ALL seeds forms one wave: a one-thread-per-seed step kernel finishes the seeds' previous
//                         observation (bearing, triangulation, tau, filter update, convergence flag) and prepares the next one
//                         (type / visibility gates, epipolar geometry, affine warp matrix, scan parameters) into a compact work
//                         list; a one-8-lane-group-per-work-item match kernel does what needs a patch (affine warp, ZMSSD scan,
//                         sub-pixel alignment). All warps of a launch run the same short code (the all-in-one kernel of round 1
//                         spent 71 % of its warp samples waiting for instructions) and the FP64 geometry is no longer computed
//                         eight times per seed.
#include "depth_filter_dev.cuh"
#include <cstdlib>

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
// Concurrent seed groups of one large svo_cuda_update_seeds call (env SVO_SEED_GROUPS overrides per call). B200, 50 k seeds x 64
// observations: 1 group 5.89 ms, 2: 5.68, 3: 5.29, 4: 4.51, 6: 4.53, 8: 5.1-5.4 (the host cannot issue 1024 launches fast enough).
constexpr int kSeedGroups = 4;
constexpr int kSeedGroupMin = 16384;          // calls with fewer seeds stay on one stream

struct SeedParams {
  PyrView ref_pyr, cur_pyr;
  svo_camera cam_ref, cam_cur;
  int S, n_obs;     // seeds of this group, observation waves
  int s0, S_all;    // first seed of the group, seeds of the call (stride of the observation-major arrays)
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

// One pending observation of a seed, handed from the step kernel to the match kernel and back.
struct SeedWork {
  EpiSetup e;
  int cur_frame, T_idx;
  int align_1d;
  int result;         // Matcher::MatchResult of the group part (match kernel)
  double px_x, px_y;  // px_cur_ (match kernel)
};

// Wave o (0 <= o <= n_obs), one thread per seed: finish observation o - 1 of the seeds that had a match pending, then prepare
// observation o. Between the two halves the seed's type and state are exactly what the sequential loop of the reference holds
// between two updateSeed calls.
#ifndef SVO_SEED_STEP_MINB
#define SVO_SEED_STEP_MINB 1   // 6 / 8 CTAs per SM (80 / 64 registers) measured: no gain (6.93 -> 7.09 / 7.12 ms per 3.2 M seed-observations)
#endif
__global__ void __launch_bounds__(kThreads, SVO_SEED_STEP_MINB) seed_step_kernel(const SeedParams P, int o, SeedWork* __restrict__ work, uint8_t* __restrict__ pending,
                                                             int* __restrict__ list, int* __restrict__ counts) {
  const int sl = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = P.s0 + sl;
  const int lane = threadIdx.x & 31;
  bool active = false, edge_item = false;
  int n_ok = 0;
  if (sl < P.S) {
    int type = P.types[s];
    double2* sp = reinterpret_cast<double2*>(P.state) + 2 * (size_t)s;
    const double2 s01 = sp[0], s23 = sp[1];
    double st[4] = {s01.x, s01.y, s23.x, s23.y};
    svo_feature ft = P.ftrs[s];
    const V3d f_ref{ft.f[0], ft.f[1], ft.f[2]};
    bool dirty = false;
    if (o > 0 && pending[s]) {  // depth_filter.cpp:441-498 for observation o - 1
      const SeedWork& w = work[s];
      const SE3d T = se3Load(P.T_cur_ref + 7 * (size_t)w.T_idx);
      int mr = w.result;
      double depth = 0.0;
      if (mr == kSuccess) {
        V3d f_cur;
        mr = epiFinish(P.cam_cur, T, f_ref, w.px_x, w.px_y, f_cur, &depth);
      }
      if (mr != kSuccess) {
        st[3] += 1;  // seed::increaseOutlierProbability (the matcher's reject_ flag is only set by the angle gate of the step kernel)
      } else {
        const double mu_range = P.seed_mu_range[s];
        const double cur_thresh = (type == kMapPointSeed || type == kMapPointSeedConverged) ? P.dopt.mappoint_convergence_sigma2_thresh
                                                                                               : P.dopt.seed_convergence_sigma2_thresh;
        const SE3d T_ref_cur = se3Inv(T);
        const double depth_sigma = computeTau(T_ref_cur.t, f_ref, depth, P.px_error_angle);  // :459
        const double zi = 1.0 / depth;
        const double sg = 0.5 * (1.0 / fmax(0.000000000001, depth - depth_sigma) - 1.0 / (depth + depth_sigma));  // seed.h:155-160
        const double tau2 = sg * sg;
        const bool ok = P.dopt.use_vogiatzis_update ? updateFilterVogiatzis(zi, tau2, mu_range, st) : updateFilterGaussian(zi, tau2, st);
        if (!ok) {
          type = kOutlier;  // :470-471, :481-482
        } else {
          const double thresh = mu_range / cur_thresh;  // seed::isConverged, seed.h:145-153
          if (st[1] < thresh * thresh) {
            if (type == kCornerSeed) type = kCornerSeedConverged;
            else if (type == kEdgeletSeed) type = kEdgeletSeedConverged;
            else if (type == kMapPointSeed) type = kMapPointSeedConverged;
          }
          n_ok = 1;
        }
      }
      if (P.match_results) P.match_results[(size_t)(o - 1) * P.S_all + s] = mr;
      dirty = true;
    }
    if (o < P.n_obs) {  // depth_filter.cpp:377-439 for observation o
      const size_t oi = (size_t)o * P.S_all + s;
      int mr = -1;
      const int cf = P.obs_frame_idx[oi];
      bool go = cf >= 0;  // :377-381 (cur frame == ref frame): the caller marks such observations with a negative index
      if (go && type == kOutlier) go = false;  // :387-392
      if (go && P.dopt.check_convergence && (type == kCornerSeedConverged || type == kEdgeletSeedConverged || type == kMapPointSeedConverged))
        go = false;  // :394-399
      if (go) {
        const int ti = P.obs_T_idx[oi];
        const SE3d T = se3Load(P.T_cur_ref + 7 * (size_t)ti);
        if (P.dopt.check_visibility) {  // :406-420
          const V3d xyz_f = se3Apply(T, f_ref * (1.0 / st[0]));
          const V2d px = camProject3(P.cam_cur, xyz_f);
          if (!(px.x >= 0.0 && px.y >= 0.0 && px.x < (double)P.cam_cur.width && px.y < (double)P.cam_cur.height)) go = false;
          const int pxi0 = (int)px.x, pxi1 = (int)px.y;
          const int boundary = 9;
          if (go && !(pxi0 >= boundary && pxi1 >= boundary && pxi0 < P.cam_cur.width - boundary && pxi1 < P.cam_cur.height - boundary)) go = false;
        }
        if (go) {
          ft.type = type;
          // seed.h:115-128: d_estimate_inv = mu, d_min_inv = mu + sigma, d_max_inv = max(mu - sigma, 1e-8)
          const double sig = sqrt(st[1]);
          SeedWork w;
          epiSetup(P.cam_ref, P.cam_cur, T, ft, st[0], st[0] + sig, fmax(st[0] - sig, 0.00000001), P.mopt, P.ref_pyr.n_levels - 1, w.e);
          if (w.e.early >= 0) {
            mr = w.e.early;  // the angle gate: reject_ is set, the outlier count stays (:445-450)
          } else {
            w.cur_frame = cf; w.T_idx = ti;
            w.align_1d = (type == kEdgeletSeed || type == kEdgeletSeedConverged) ? 1 : 0;  // :423-427
            w.result = -1; w.px_x = 0.0; w.px_y = 0.0;
            work[s] = w;
            active = true;
            edge_item = w.align_1d != 0;
          }
        }
      }
      if (!active && P.match_results) P.match_results[oi] = mr;
      pending[s] = active ? 1 : 0;
    }
    if (dirty) {
      P.types[s] = (uint8_t)type;
      sp[0] = make_double2(st[0], st[1]);
      sp[1] = make_double2(st[2], st[3]);
    }
  }
  // Compact work list of this wave, two-ended: edgelet seeds (1-D alignment) fill it from the front, corner seeds (2-D alignment) from
  // the back, so that the four items of a warp of the match kernel run the same alignment code (a warp with both kinds executes
  // both, one after the other). One atomic per warp and kind; the order inside a warp is the seed order.
  const bool edge = active && edge_item;
  const unsigned balA = __ballot_sync(0xffffffffu, active && edge), balB = __ballot_sync(0xffffffffu, active && !edge);
  if (balA) {
    int base = 0;
    const int leader = __ffs((int)balA) - 1;
    if (lane == leader) base = atomicAdd(&counts[2 * o], __popc(balA));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (active && edge) list[base + __popc(balA & ((1u << lane) - 1u))] = s;
  }
  if (balB) {
    int base = 0;
    const int leader = __ffs((int)balB) - 1;
    if (lane == leader) base = atomicAdd(&counts[2 * o + 1], __popc(balB));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (active && !edge) list[P.S - 1 - (base + __popc(balB & ((1u << lane) - 1u)))] = s;
  }
  const unsigned okb = __ballot_sync(0xffffffffu, n_ok != 0);
  if (okb && lane == 0) atomicAdd(P.n_success, __popc(okb));
}

// One 8-lane group per work item of wave o: affine warp of the reference patch, ZMSSD scan, sub-pixel alignment.
template <int SCAN>
#ifndef SVO_SEED_MATCH_MINB
#define SVO_SEED_MATCH_MINB 5   // 3.2 M seed-observations: 6.94 ms at 4 CTAs / SM, 6.40 at 5, 6.45 at 6
#endif
__global__ void __launch_bounds__(kThreads, SVO_SEED_MATCH_MINB) seed_match_kernel(const SeedParams P, int o, SeedWork* __restrict__ work, const int* __restrict__ list,
                                                                 const int* __restrict__ counts) {
  __shared__ __align__(16) uint8_t s_pwb[kGroupsPerCta * kPwbPitch];
  const Group g = makeGroup();
  const int gi = threadIdx.x / kGroup;
  const int k = blockIdx.x * kGroupsPerCta + gi;
  if (k >= P.S || (k >= counts[2 * o] && k < P.S - counts[2 * o + 1])) return;  // front: edgelet seeds, back: corner seeds
  const int s = list[k];
  uint8_t* pwb = s_pwb + gi * kPwbPitch;
  SeedWork& w = work[s];
  const EpiSetup e = w.e;
  const svo_feature ft = P.ftrs[s];
  const int rf = P.ref_frame_idx ? P.ref_frame_idx[s] : 0;
  double px_x = 0.0, px_y = 0.0, h_inv = 0.0;
  const int res = epiMatch<SCAN>(g, P.ref_pyr, rf, P.cur_pyr, w.cur_frame, P.cam_cur, ft, e, P.mopt, w.align_1d != 0, pwb, px_x, px_y, &h_inv);
  if (g.r == 0) { w.result = res; w.px_x = px_x; w.px_y = px_y; }
}

// n_obs ordered filter updates per seed with the state in registers: z / tau2 are [n_obs][n] (observation-major, coalesced).
// The reference's nineteen divisions per update are folded into two reciprocals (vogiatzisRational below): the kernel is bound by the
// FP64 pipe and by the length of one update's dependent chain, and the result differs from the reference's IEEE divisions by a few
// ulp per update (tests: rtol 1e-9 after 64 updates; the north-star tolerance for seed mean / variance is 1e-4).
SVO_D double rcpGuarded(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, fma(e, e, e), x);
  e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  return x == x ? x : 1.0 / d;  // zero, infinite and denormal divisors take the IEEE path
}
// updateFilterVogiatzis as ONE rational expression per output: with c1 = a N(z; mu, sigma2 + tau2), c2 = b / mu_range (the common
// factor 1 / (a + b) of C1, C2 cancels), f = Nf / Df and e = Ne / (Df (a + b + 2)),
//   a' = (Ne - Nf ab2) Nf / (Nf^2 ab2 - Ne Df),   b' = (Ne - Nf ab2) (Df - Nf) / (Nf^2 ab2 - Ne Df),
//   s2 = sigma2 tau2 / (sigma2 + tau2),   m = (mu tau2 + z sigma2) / (sigma2 + tau2),
// so an update costs one rsqrt, one exp and two reciprocals on its dependent chain instead of nineteen divisions.
SVO_D bool vogiatzisRational(double z, double tau2, double inv_range, double s[4]) {
  const double mu = s[0], sigma2 = s[1], a = s[2], b = s[3];
  const double v = sigma2 + tau2;
  if (!(v >= 0.0)) return false;  // sqrt(sigma2 + tau2) is NaN (depth_filter.cpp:509-512)
  const double rs = rsqrt(v), rs2 = rs * rs;  // 1 / norm_scale, 1 / (sigma2 + tau2)
  const double s2 = sigma2 * tau2 * rs2;
  const double m = (mu * tau2 + z * sigma2) * rs2;
  const double d = z - mu;
  const double pdf = exp(-(d * d) * (0.5 * rs2)) * (rs * 0.3989422804014326779);  // vk::normPdf(z, mu, norm_scale)
  const double c1 = a * pdf, c2 = b * inv_range;
  const double ic = rcpGuarded(c1 + c2);
  const double ab1 = a + b + 1.0, ab2 = a + b + 2.0;
  const double Nf = c1 * (a + 1.0) + c2 * a;
  const double Ne = (c1 * (a + 2.0) + c2 * a) * (a + 1.0);
  const double Df = (c1 + c2) * ab1;
  const double iden = rcpGuarded(Nf * Nf * ab2 - Ne * Df);
  const double t = Ne - Nf * ab2;
  const double mu_new = (c1 * m + c2 * mu) * ic;
  double sigma2_new = (c1 * (s2 + m * m) + c2 * (sigma2 + mu * mu)) * ic - mu_new * mu_new;
  bool ok = true;
  double mu_out = mu_new;
  if (sigma2_new < 0.0) sigma2_new = sigma2;
  if (mu_out < 0.0) { mu_out = 1.0; ok = false; }
  s[0] = mu_out; s[1] = sigma2_new; s[2] = t * Nf * iden; s[3] = t * (Df - Nf) * iden;
  return ok;
}

template <bool GAUSS>
__global__ void __launch_bounds__(64) filter_seq_kernel(int n, int n_obs, const double* __restrict__ z, const double* __restrict__ tau2,
                                                        const double* __restrict__ mu_range, double* __restrict__ state,
                                                        uint8_t* __restrict__ ok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2* sp = reinterpret_cast<double2*>(state) + 2 * (size_t)i;
  const double2 s01 = sp[0], s23 = sp[1];
  double s[4] = {s01.x, s01.y, s23.x, s23.y};
  const double inv_range = GAUSS ? 0.0 : rcpGuarded(mu_range[i]);
  double zn = z[i], tn = tau2[i];
  for (int o = 0; o < n_obs; ++o) {
    const double zc = zn, tc = tn;
    if (o + 1 < n_obs) { zn = z[(size_t)(o + 1) * n + i]; tn = tau2[(size_t)(o + 1) * n + i]; }  // next observation in flight
    const bool r = GAUSS ? updateFilterGaussian(zc, tc, s) : vogiatzisRational(zc, tc, inv_range, s);
    if (ok) ok[(size_t)o * n + i] = r ? 1 : 0;
  }
  sp[0] = make_double2(s[0], s[1]);
  sp[1] = make_double2(s[2], s[3]);
}

}  // namespace

extern "C" {

int svo_cuda_update_filter_vogiatzis(svo_cuda_ctx* ctx, int n, const double* z, const double* tau2, const double* mu_range,
                                     double* state, uint8_t* ok, svo_mem mem) {
  if (!ctx || n < 0 || !z || !tau2 || !mu_range || !state)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_update_filter_vogiatzis: bad arguments");
  if (n == 0) return SVO_OK;
  SVO_BIND(ctx);
  Stager st(ctx, mem);
  const double* dz = st.in(z, (size_t)n);
  const double* dt = st.in(tau2, (size_t)n);
  const double* dm = st.in(mu_range, (size_t)n);
  double* ds = st.inout(state, (size_t)n * 4);
  uint8_t* dok = st.out(ok, (size_t)n);
  if (!st.send()) return st.finish();
  vogiatzis_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, dz, dt, dm, ds, dok);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

int svo_cuda_update_filter_seq(svo_cuda_ctx* ctx, int n, int n_obs, const double* z, const double* tau2, const double* mu_range,
                               double* state, uint8_t* ok, int gaussian, svo_mem mem) {
  if (!ctx || n < 0 || n_obs < 0 || !z || !tau2 || (!gaussian && !mu_range) || !state)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_update_filter_seq: bad arguments");
  if (n == 0 || n_obs == 0) return SVO_OK;
  SVO_BIND(ctx);
  Stager st(ctx, mem);
  const size_t no = (size_t)n * n_obs;
  const double* dz = st.in(z, no);
  const double* dt = st.in(tau2, no);
  const double* dm = st.in(mu_range, (size_t)n);
  double* ds = st.inout(state, (size_t)n * 4);
  uint8_t* dok = st.out(ok, no);
  if (!st.send()) return st.finish();
  if (gaussian) filter_seq_kernel<true><<<(n + 63) / 64, 64, 0, ctx->stream>>>(n, n_obs, dz, dt, dm, ds, dok);
  else filter_seq_kernel<false><<<(n + 63) / 64, 64, 0, ctx->stream>>>(n, n_obs, dz, dt, dm, ds, dok);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

int svo_cuda_compute_tau(svo_cuda_ctx* ctx, int n, const double* T_ref_cur, const double* f, const double* z, double px_error_angle,
                         double* tau, svo_mem mem) {
  if (!ctx || n < 0 || !T_ref_cur || !f || !z || !tau) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_compute_tau: bad arguments");
  if (n == 0) return SVO_OK;
  SVO_BIND(ctx);
  Stager st(ctx, mem);
  const double* dT = st.in(T_ref_cur, (size_t)n * 7);
  const double* df = st.in(f, (size_t)n * 3);
  const double* dz = st.in(z, (size_t)n);
  double* dtau = st.out(tau, (size_t)n);
  if (!st.send()) return st.finish();
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
  SVO_BIND(ctx);
  Stager st(ctx, mem);
  SeedParams P;
  memset(&P, 0, sizeof(P));
  P.ref_pyr = makeView(ref_pyr);
  P.cur_pyr = makeView(cur_pyr);
  P.cam_ref = *cam_ref;
  P.cam_cur = *cam_cur;
  P.S = S; P.n_obs = n_obs; P.s0 = 0; P.S_all = S;
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
  // wave scratch: one work item / pending flag / list slot per seed, one counter per wave
  SeedWork* d_work = (SeedWork*)st.scratch(sizeof(SeedWork) * (size_t)(S > 0 ? S : 1));
  uint8_t* d_pending = (uint8_t*)st.scratch((size_t)(S > 0 ? S : 1));
  int* d_list = (int*)st.scratch(sizeof(int) * (size_t)(S > 0 ? S : 1));
  // Seeds are independent, only the observations of ONE seed are ordered: large calls are cut into groups of seeds whose wave chains run
  // on the context's side streams, so that the step kernel of one group (one thread per seed, long dependent FP64 chains, few warps) and
  // the tail of its match kernel overlap the match kernels of the other groups.
  int G = 1;
  if (S >= kSeedGroupMin) G = kSeedGroups;
  if (const char* e = getenv("SVO_SEED_GROUPS")) G = atoi(e);
  G = G < 1 ? 1 : (G > 1 + svo_cuda_ctx::kSideStreams ? 1 + svo_cuda_ctx::kSideStreams : G);
  if (S < G * kThreads) G = 1;
  const size_t counts_per_group = 2 * (size_t)(n_obs + 1);
  int* d_counts = (int*)st.scratch(sizeof(int) * counts_per_group * G);  // per group and wave: edgelet items (front of the list), corner items (back)
  if (!st.send() || !d_work || !d_pending || !d_list || !d_counts) return st.finish();
  SVO_CUDA_TRY(ctx, cudaMemsetAsync(d_ns, 0, sizeof(int), ctx->stream));
  if (S > 0 && n_obs > 0) {
    SVO_CUDA_TRY(ctx, cudaMemsetAsync(d_counts, 0, sizeof(int) * counts_per_group * G, ctx->stream));
    if (G > 1) {
      const int rc = svoEnsureSideStreams(ctx);
      if (rc != SVO_OK) { st.finish(); return rc; }
      SVO_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
      for (int g = 1; g < G; ++g) SVO_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->side_stream[g - 1], ctx->ev_fork, 0));
    }
    const int chunk = ((S + G - 1) / G + kThreads - 1) / kThreads * kThreads;  // whole step CTAs per group
    for (int o = 0; o <= n_obs; ++o) {
      for (int g = 0; g < G; ++g) {
        SeedParams Q = P;
        Q.s0 = g * chunk;
        Q.S = S - Q.s0 < chunk ? S - Q.s0 : chunk;
        if (Q.S <= 0) continue;
        cudaStream_t stream = g == 0 ? ctx->stream : ctx->side_stream[g - 1];
        int* list = d_list + Q.s0;
        int* counts = d_counts + counts_per_group * g;
        const int step_grid = (Q.S + kThreads - 1) / kThreads, match_grid = (Q.S + kGroupsPerCta - 1) / kGroupsPerCta;
        seed_step_kernel<<<step_grid, kThreads, 0, stream>>>(Q, o, d_work, d_pending, list, counts);
        SVO_LAUNCH_CHECK(ctx);
        if (o < n_obs) {
          if (mopt->scan_on_unit_sphere) seed_match_kernel<1><<<match_grid, kThreads, 0, stream>>>(Q, o, d_work, list, counts);
          else seed_match_kernel<0><<<match_grid, kThreads, 0, stream>>>(Q, o, d_work, list, counts);
          SVO_LAUNCH_CHECK(ctx);
        }
      }
    }
    for (int g = 1; g < G; ++g) {
      SVO_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join[g - 1], ctx->side_stream[g - 1]));
      SVO_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[g - 1], 0));
    }
  }
  return st.finish();
}

}  // extern "C"
