// (b) svo::SparseImgAlign::run — inverse-compositional sparse image alignment, whole run() in ONE launch.
//
// ref: src/svo_img_align/src/sparse_img_align.cpp:34-156 (run / evaluateError), :209-541 (utils)
//      src/svo_img_align/src/sparse_img_align_base.cpp:35-107 (defaults, update, applyPrior)
//      src/vikit/vikit_solver/include/vikit/solver/implementation/mini_least_squares_solver.hpp:42-107 (Gauss-Newton),
//      :253-262 (H.ldlt().solve(g)); src/vikit/vikit_solver/src/robust_cost.cpp:48-60 (Tukey)
//      src/svo_common/include/svo/common/frame.h:342-357, src/svo_common/src/frame.cpp:274-290 (projection Jacobians)
//
// Mapping: one CTA per frame-bundle pair; the CTA walks all pyramid levels and all Gauss-Newton iterations on the
// device (no host round trips). One thread per feature patch:
//   setup      feature subset (b2) + base caches (b3): xyz_ref and 1/z (pinhole Jacobian) or the 2x3 projection Jacobian
//              (use_distortion_jacobian) -> shared memory (k-major, no bank conflicts), once per run;
//   per level  6x6 bilinear reference patch (b4) -> 32 doubles per feature in shared memory (centre + the 4-neighbour
//              values the central differences need);
//   per iter   project, visibility test, 5x5 cur-image taps via aligned 32-bit loads, 16 residuals (b5);
//              the per-pixel Jacobian factorises as J = [ (dx*Jp0 + dy*Jp1)*scale ; a6 ; a7 ] with
//              Jp = Jproj * R_cam_imu * [I | -skew(p_imu)], so the gradient of a patch is
//                  g_trans = -scale * R_imu_cam * c,   g_rot = -scale * (R_imu_cam * (xyz_ref x c) + t_imu_cam x R_imu_cam c),
//              c = Jproj^T [sum dx*res ; sum dy*res]: every patch contributes 6 numbers (c, xyz_ref x c) that are summed per
//              camera and rotated ONCE after the reduction — no per-patch 2x6 Jacobian is stored (b6). H is a rank-2
//              expansion of a few patch sums with Jp rebuilt on the fly; without robust weights H depends only on WHICH
//              patches are visible, so it is reduced and factorised (LDL^T) once per level and again only when the
//              visibility pattern changes; every other iteration reduces 7 numbers and runs two triangular solves;
//              warp 0 then applies the prior and the SE3/illumination update (b7, b8) and rebuilds the cameras'
//              T_cur_ref lane-parallel — two block barriers per iteration.
// All arithmetic is FP64 like the reference (FloatType = double, src/svo_common/include/svo/common/types.h:16);
// Tukey weights and the alpha/beta handed to the residual are float, as in the reference signatures.
// Shared memory: 36 doubles per feature -> 4 CTAs per SM for <= 180 features (3 with the distortion Jacobian).
#include "sparse_align_dev.cuh"

using namespace svo_align;

namespace {

// Threads per CTA (= per pair): a thread owns feature slots tid, tid + kThreads, ... Measured on the B200 (4096 pairs, FP32 patch cache):
// 192 threads / 4 pairs per SM 0.821 ms, 128 threads / 6 pairs 0.928 ms, 96 threads / 7 pairs 0.854 ms: the register file holds about 24
// warps of this kernel whatever the CTA size, so smaller CTAs only lengthen a pair's own path. A/B builds: -DSVO_ALIGN_THREADS=.
#ifndef SVO_ALIGN_THREADS
#define SVO_ALIGN_THREADS 192
#endif
constexpr int kThreads = SVO_ALIGN_THREADS;
constexpr int kWarps = kThreads / 32;
static_assert(kThreads % 32 == 0 && kWarps >= 1 && kWarps <= 8, "kThreads");
// Bundles with more than kFixedSlots features (stereo rigs: 2 x 150-180) get twice the threads, so that their slots are still walked in
// one pass and an SM still holds 24 warps (2 CTAs x 12): 9.5 -> see profiles/ ms for the 8192 stereo pairs of the front-end chain.
#ifndef SVO_ALIGN_WIDE_ILLUM_MINB
#define SVO_ALIGN_WIDE_ILLUM_MINB 2   // 8-DoF stereo bundles of the front-end chain: 1 CTA / SM (162 registers) 7.9 ms, 2 CTAs / SM (80) 5.7 ms per 8192 pairs
#endif
#ifndef SVO_ALIGN_ILLUM_MINB
#define SVO_ALIGN_ILLUM_MINB kMinBlocks   // 8-DoF mono, 4096 pairs: 3 CTAs / SM 1.547 ms, 4 CTAs / SM (80 registers) 1.458 ms
#endif
constexpr int kThreadsWide = 2 * kThreads;
constexpr int kMaxWarps = kThreadsWide / 32;
// resident CTAs per SM the register allocation is held to: (common case, variants with more per-thread state)
#ifdef SVO_ALIGN_MINB
constexpr int kMinBlocks = SVO_ALIGN_MINB;
#else
constexpr int kMinBlocks = kThreads <= 96 ? 7 : (kThreads <= 128 ? 6 : 4);
#endif
#ifdef SVO_ALIGN_HEAVY_MINB
constexpr int kMinBlocksHeavy = SVO_ALIGN_HEAVY_MINB;
#else
constexpr int kMinBlocksHeavy = kThreads <= 96 ? 6 : (kThreads <= 128 ? 4 : 3);
#endif
// The warp that runs the serial part of an iteration. (Warps are dealt to the four SM sub-partitions round robin, so sub-partitions
// 2 and 3 hold one warp of every resident CTA instead of two; moving the serial work there was measured on the B200: 0.888 ms
// instead of 0.869 ms per 4096 pairs, so it stays on warp 0.)
constexpr int kSerialWarp = 0;
[[maybe_unused]] constexpr int kTimerTid = 32 * kSerialWarp;  // used by the profiling build (make timing)

struct Ctl {
  SE3d T, T_old;
  double alpha, beta, alpha_old, beta_old;
  double I_prior[8];            // diagonal of I_prior_
  double L[28], rd[8];          // cached LDL^T factor of H (+ prior) and the pivot reciprocals
  double chi_num, chi_den;      // chi2 = float(chi_num / chi_den) of the last accepted iteration, divided once at the end
  int chi_set;                  // 0 until an iteration was accepted (chi2 = 1e10, the solver's reset value)
  float alpha_f, beta_f;
  double dx[8];                 // serial-solve path: lane 0 hands dx to the warp
  int stop, brk;
  int h_dirty;                  // a patch entered or left the image in this iteration: H must be re-reduced
  int warp_cnt[kMaxWarps];
  int n_total;                  // features in the run (read from here inside the loops: a register copy would be spilled)
  int iters[SVO_MAX_LEVELS];
#ifdef SVO_ALIGN_TIMING
  // phase clocks of thread 0 (profiling builds only; reported in the unused rows 6-7 of the result's H for 6-DoF runs):
  // 0 total, 1 setup, 2 reference patches, 3 residual pass (thread 0's own), 4 wait at the first barrier, 5 H rebuild,
  // 6 serial solve + update + camera refresh, 7 wait at the second barrier, 8 iterations, 9 H rebuilds, 10-15 residual pass of warps
  // 0-5 (each warp's own clock), 16-20 the serial phase split: cross-warp totals, gradient (+ prior), solve, state update, camera refresh
  long long tm[28];
  long long t_mark;
#endif
};
#ifdef SVO_ALIGN_TIMING
#define SVO_TM_MARK() do { if (tid == kTimerTid) ctl.t_mark = clock64(); } while (0)
#define SVO_TM_ADD(k) do { if (tid == kTimerTid) { const long long _t = clock64(); ctl.tm[k] += _t - ctl.t_mark; ctl.t_mark = _t; } } while (0)
#else
#define SVO_TM_MARK() do { } while (0)
#define SVO_TM_ADD(k) do { } while (0)
#endif

// ILL: 0 = no illumination parameters and alpha = beta = 0 (the subtraction of the reference pixel rides in the
// interpolation's FMA chain), 1 = no illumination parameters but non-zero initial alpha/beta, 2 = gain and/or offset estimated.
// SPLIT2: two threads per patch (two of its four pixel rows each) on kThreadsWide threads — the latency variant for batches that leave
// SMs idle anyway (B <= number of SMs): the residual pass of an iteration is a dependent chain per thread, half as long with half the
// pixels. The gradient contributions are linear in the per-thread sums, so the two halves need no exchange: the warp reduction adds
// them. Instantiated for the fixed-slot variants without robust weights and without the distortion Jacobian.
// (register caps measured per variant on the B200, 4096 pairs: robust weights alone 1.473 ms at 3 CTAs / SM, 1.410 at 4; robust + illumination 3.19 at 3, 3.70 at 4)
template <int ILL, bool ROBUST, bool DJ, int SLOTS, bool SPLIT2 = false>
__global__ void __launch_bounds__((SLOTS && !SPLIT2) ? kThreads : kThreadsWide,
                                  SPLIT2 ? 1 : (SLOTS ? ((DJ || (ROBUST && ILL == 2)) ? kMinBlocksHeavy : (ILL == 2 ? SVO_ALIGN_ILLUM_MINB : kMinBlocks))
                                                      : ((DJ || ROBUST) ? 1 : (ILL == 2 ? SVO_ALIGN_WIDE_ILLUM_MINB : 2))))
sparse_align_kernel(const AlignParams P) {
  static_assert(!SPLIT2 || !ROBUST, "SPLIT2 is instantiated for the variants without robust weights (every per-pixel sum is linear, the split would be valid there too)");
  constexpr int TH = (SLOTS && !SPLIT2) ? kThreads : kThreadsWide;  // threads of this variant
  constexpr int NY = SPLIT2 ? 2 : 4;                                   // pixel rows of a patch per thread
  constexpr int NW = TH / 32;
  constexpr bool ILLUM = ILL == 2;
  constexpr bool unit_gain = ILL == 0;
  constexpr int D = ILLUM ? 8 : 6;
  constexpr int NH = D * (D + 1) / 2;
  constexpr int NAUX = DJ ? 6 : 1;
  extern __shared__ __align__(16) double smem[];
  const int stride = SLOTS ? SLOTS : P.slots;
  const int n_cams = P.n_cams;
  // per-warp accumulators: H | (c, xyz x c) per camera | g6 g7 | chi2 | n_meas | changed
  const int iCM = NH, iG6 = NH + 6 * n_cams, iChi = iG6 + (ILLUM ? 2 : 0), iN = iChi + 1, NV = iChi + 3;
  double* s_xyz = smem;                          // [3][stride]
  double* s_aux = s_xyz + 3 * stride;            // [NAUX][stride]  1/z, or the 2x3 projection Jacobian
  double* s_red = s_aux + NAUX * stride;         // [NW][NV]
  double* s_tot = s_red + NW * NV;           // [NV]
  double* s_camblk = s_tot + NV;                 // [n_cams][kCamBlk]
  const int NG = 6 * n_cams + (ILLUM ? 2 : 0);   // gradient-related totals: per-camera (c, xyz x c), then g6 g7
  double* s_Hinv = s_camblk + kCamBlk * n_cams;  // [D][D]   H^-1 (rebuilt with H)
  double* s_P = s_Hinv + D * D;                  // [D][NG]  H^-1 M: dx = P * totals
  Ctl& ctl = *reinterpret_cast<Ctl*>(s_P + D * NG + (NG & 1));
  PatchT* s_patch = reinterpret_cast<PatchT*>(&ctl + 1);       // [32][stride] reference patch cache of the current level
  int* s_src = reinterpret_cast<int*>(s_patch + 32 * stride);   // [stride] feature index inside its camera's array
  uint8_t* s_cam = reinterpret_cast<uint8_t*>(s_src + stride);  // [stride]

  const int pair = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const svo_sparse_align_options& opt = P.opt;

  // ---- setup -----------------------------------------------------------------------------------------------
#ifdef SVO_ALIGN_TIMING
  const long long t_start = clock64();
  if (tid == kTimerTid) { for (int i = 0; i < 28; ++i) ctl.tm[i] = 0; ctl.t_mark = t_start; }
#endif
  if (tid < n_cams) {
    const SE3d Tci = se3Load(P.T_cam_imu[tid]);
    const SE3d Tic = se3Inv(Tci);
    const M3d Rci = quatToMatrix(Tci.q), Ric = quatToMatrix(Tic.q);
    double* cb = s_camblk + kCamBlk * tid;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) { cb[3 * r + c] = Rci.m[r][c]; cb[9 + 3 * r + c] = Ric.m[r][c]; }
    cb[18] = Tci.t.x; cb[19] = Tci.t.y; cb[20] = Tci.t.z;
    cb[21] = Tic.t.x; cb[22] = Tic.t.y; cb[23] = Tic.t.z;
  }
  if (tid == 32) {
    const SE3d T_iref_world = se3Load(P.T_imu_world_ref + 7 * (size_t)pair);
    const SE3d T_icur_world = se3Load(P.T_imu_world_cur + 7 * (size_t)pair);
    ctl.T = se3Mul(T_icur_world, se3Inv(T_iref_world));  // sparse_img_align.cpp:74-75
    ctl.T_old = ctl.T;
    ctl.alpha = ctl.alpha_old = opt.alpha_init;
    ctl.beta = ctl.beta_old = opt.beta_init;
    ctl.alpha_f = (float)opt.alpha_init;
    ctl.beta_f = (float)opt.beta_init;
    ctl.stop = 0; ctl.brk = 0; ctl.h_dirty = 0;
    ctl.chi_num = 0.0; ctl.chi_den = 0.0; ctl.chi_set = 0;  // no accepted iteration yet: chi2 = 1e10 (reset(): mini_least_squares_solver.hpp:243)
    for (int i = 0; i < SVO_MAX_LEVELS; ++i) ctl.iters[i] = 0;
    for (int i = 0; i < 8; ++i) { ctl.I_prior[i] = 0.0; ctl.rd[i] = 0.0; }
    for (int i = 0; i < 28; ++i) ctl.L[i] = 0.0;
  }
  for (int i = tid; i < NV; i += TH) s_tot[i] = 0.0;
  __syncthreads();

  // b2 + b3: ordered compaction of the eligible, in-bounds features of every ref camera, base caches
  int n_total = 0;
  for (int c = 0; c < n_cams; ++c) {
    const size_t fbase = ((size_t)pair * n_cams + c) * P.max_features;
    const int n = min(P.n_features[(size_t)pair * n_cams + c], P.max_features);
    const PyrView& rp = P.ref_pyr[c];
    const int ml = opt.max_level;
    const double scale_max = 1.0f / (1 << ml);
    const int rows_m2 = rp.rows[ml] - 2, cols_m2 = rp.cols[ml] - 2;
    for (int base = 0; base < n; base += TH) {
      const int i = base + tid;
      bool ok = false;
      if (i < n && P.eligible[fbase + i]) {
        const double pu = P.px[2 * (fbase + i)], pv = P.px[2 * (fbase + i) + 1];
        // sparse_img_align.cpp:249-257 with patch_size_wb = 6, patch_center_wb = 2.5
        const int u_tl_i = (int)floor(pu * scale_max - 2.5);
        const int v_tl_i = (int)floor(pv * scale_max - 2.5);
        ok = !(u_tl_i < 0 || v_tl_i < 0 || u_tl_i + 6 >= cols_m2 || v_tl_i + 6 >= rows_m2);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) ctl.warp_cnt[warp] = __popc(bal);
      __syncthreads();
      int off = n_total, chunk = 0;
      for (int w = 0; w < NW; ++w) {
        if (w < warp) off += ctl.warp_cnt[w];
        chunk += ctl.warp_cnt[w];
      }
      if (ok) {
        const int s = off + __popc(bal & ((1u << lane) - 1u));
        if (s < stride) {
          // sparse_img_align.cpp:262-317: xyz_ref = f * depth; the point in the ref camera frame is xyz_ref itself
          const double depth = P.depth[fbase + i];
          const V3d xyz_ref = V3d{P.f[3 * (fbase + i)], P.f[3 * (fbase + i) + 1], P.f[3 * (fbase + i) + 2]} * depth;
          if (DJ) {  // Frame::jacobian_xyz2image_imu, frame.cpp:274-290, times -1
            double Jp[2][3];
            camProject3Jac(P.cams[c], xyz_ref, Jp);
            for (int r = 0; r < 2; ++r)
              for (int k = 0; k < 3; ++k) s_aux[(3 * r + k) * stride + s] = -Jp[r][k];
          } else {
            s_aux[s] = 1.0 / xyz_ref.z;
          }
          s_xyz[0 * stride + s] = xyz_ref.x; s_xyz[1 * stride + s] = xyz_ref.y; s_xyz[2 * stride + s] = xyz_ref.z;
          s_src[s] = i;
          s_cam[s] = (uint8_t)c;
        }
      }
      n_total += chunk;
      __syncthreads();
    }
  }
  n_total = min(n_total, stride);
  if (tid == 0) ctl.n_total = n_total;
  __syncthreads();
  SVO_TM_ADD(1);

  if (n_total > 0) {
    if (warp == 0) refreshCameraTransforms(ctl.T, s_camblk, n_cams, lane);
    const bool est_gain = ILLUM && opt.estimate_illumination_gain;
    const bool est_off = ILLUM && opt.estimate_illumination_offset;
    const float wscale_f = (float)opt.weight_scale;
    // Loop-carried scalars of the thread live in shared memory (the feature count in ctl, the visibility bits in the upper bits
    // of s_cam): the residual pass needs every one of its 80 registers, and a spilled value is a local-memory load that misses
    // the little L1 left beside the patch store (~260 cycles from L2, on the critical path of every iteration).
    const volatile int* n_total_s = &ctl.n_total;

    for (int level = opt.max_level; level >= opt.min_level; --level) {
      const double scale = 1.0f / (1 << level);
      // ---- b4: reference patches of this level (sparse_img_align.cpp:319-403) ----
      for (int s = tid; s < *n_total_s; s += TH) {
        const int c = s_cam[s] & 3;
        const PyrView& rp = P.ref_pyr[c];
        const int rf = P.ref_frame_idx ? P.ref_frame_idx[(size_t)pair * n_cams + c] : pair;
        const uint8_t* img = rp.level(rf, level);
        const int pitch = rp.pitch[level];
        const size_t fi = ((size_t)pair * n_cams + c) * P.max_features + s_src[s];
        const double u_tl = P.px[2 * fi] * scale - 2.5, v_tl = P.px[2 * fi + 1] * scale - 2.5;
        const int ui = (int)floor(u_tl), vi = (int)floor(v_tl);
        const double su = u_tl - ui, sv = v_tl - vi;
        const double wtl = (1.0 - su) * (1.0 - sv), wtr = su * (1.0 - sv), wbl = (1.0 - su) * sv, wbr = su * sv;
        // 7x7 taps -> the 32 used values of the 6x6 interpolated patch; every tap is converted to double once
        double tp[7], tn[7];
        {
          unsigned ra, rb;
          loadRow8(img + (size_t)vi * pitch, ui, ra, rb);
#pragma unroll
          for (int x = 0; x < 7; ++x) tp[x] = u8ToDouble(x < 4 ? byteAt(ra, x) : byteAt(rb, x - 4));
        }
#pragma unroll
        for (int y = 0; y < 6; ++y) {
          unsigned na, nb;
          loadRow8(img + (size_t)(vi + y + 1) * pitch, ui, na, nb);
#pragma unroll
          for (int x = 0; x < 7; ++x) tn[x] = u8ToDouble(x < 4 ? byteAt(na, x) : byteAt(nb, x - 4));
#pragma unroll
          for (int x = 0; x < 6; ++x) {
            const bool used = (y >= 1 && y <= 4) || (x >= 1 && x <= 4);
            if (!used) continue;
            s_patch[patchIdx(x, y) * stride + s] = (PatchT)(wtl * tp[x] + wtr * tp[x + 1] + wbl * tn[x] + wbr * tn[x + 1]);
          }
#pragma unroll
          for (int x = 0; x < 7; ++x) tp[x] = tn[x];
        }
      }
      __syncthreads();
      SVO_TM_ADD(2);

      // ---- Gauss-Newton iterations of this level (mini_least_squares_solver.hpp:42-107) ----
      const int max_iter = opt.max_iter;
      // optimizeGaussNewton starts with `old_state = state` (mini_least_squares_solver.hpp:45): a NaN step in the first
      // iteration of a level leaves the state as the previous level left it. Thread 0 is the only reader / writer of these.
      if (tid == 0) { ctl.T_old = ctl.T; ctl.alpha_old = ctl.alpha; ctl.beta_old = ctl.beta; }
      for (int iter = 0; iter < max_iter; ++iter) {
        double* red = s_red + warp * NV;
#ifdef SVO_ALIGN_TIMING
        const long long t_warp0 = clock64();
#endif
        for (int k = lane; k < NV; k += 32) red[k] = 0.0;
        __syncwarp();
        const float alpha_f = ctl.alpha_f, beta_f = ctl.beta_f;
        const int n_now = *n_total_s, n_round = (((SPLIT2 ? 2 : 1) * n_now + TH - 1) / TH) * TH;
        for (int st = tid; st < n_round; st += TH) {
          const int s = SPLIT2 ? (st >> 1) : st;   // patch slot
          const int y0 = SPLIT2 ? 2 * (st & 1) : 0;  // first of this thread's pixel rows
          bool vis = false;
          int c = 0;
          double gx = 0, gy = 0, chi = 0, g6 = 0, g7 = 0;
          PatchSums ps = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
          bool vis_before = false;  // bit 7 of s_cam: the slot was visible in the previous iteration of this level
          if (s < n_now) {
            const unsigned cs = s_cam[s];
            c = cs & 3u;
            vis_before = (cs & 0x80u) != 0u;
            const double* Rt = s_camblk + kCamBlk * c + 24;
            const double X = s_xyz[s], Y = s_xyz[stride + s], Z = s_xyz[2 * stride + s];
            const V3d pc{Rt[0] * X + Rt[1] * Y + Rt[2] * Z + Rt[9], Rt[3] * X + Rt[4] * Y + Rt[5] * Z + Rt[10],
                         Rt[6] * X + Rt[7] * Y + Rt[8] * Z + Rt[11]};
            if (!(pc.z < 0.0)) {  // sparse_img_align.cpp:432-438
              // PinholeProjection::project3 (pinhole_projection.hpp:30-44) with a Newton reciprocal for 1/z
              const svo_camera& cm = P.cams[c];
              const double z_inv = fastRcp(pc.z);
              double ud, vd;
              camDistort(cm, pc.x * z_inv, pc.y * z_inv, ud, vd);
              const V2d uvc{cm.fx * ud + cm.cx, cm.fy * vd + cm.cy};
              const PyrView& cp = P.cur_pyr[c];
              const double u_tl = uvc.x * scale - 1.5, v_tl = uvc.y * scale - 1.5;
              // sparse_img_align.cpp:449-456 (NaN coordinates fall through as "visible" in the reference; they cannot
              // be sampled, so they are dropped here)
              if (!(u_tl < 0.0 || v_tl < 0.0 || u_tl + 4 + 2.0 >= cp.cols[level] || v_tl + 4 + 2.0 >= cp.rows[level]) &&
                  u_tl == u_tl && v_tl == v_tl) {
                vis = true;
                const int cf = P.cur_frame_idx ? P.cur_frame_idx[(size_t)pair * n_cams + c] : pair;
                const uint8_t* img = cp.level(cf, level);
                const int pitch = cp.pitch[level];
                const int ui = (int)floor(u_tl), vi = (int)floor(v_tl);
                const double su = u_tl - ui, sv = v_tl - vi;
                const double wtl = (1.0 - su) * (1.0 - sv), wtr = su * (1.0 - sv), wbl = (1.0 - su) * sv, wbr = su * sv;
                const double gain = 1.0 + alpha_f;
                const PatchT* patch = s_patch + s;
                // 5x5 taps, each converted once; the 32 stored patch values are read exactly once (rolling rows)
                double tp[5], tn[5];
                {
                  unsigned ra, rb;
                  loadRow5(img + (size_t)(vi + y0) * pitch, ui, ra, rb);
                  tp[0] = tapToDouble<0>(ra, 0); tp[1] = tapToDouble<1>(ra, 1); tp[2] = tapToDouble<2>(ra, 2); tp[3] = tapToDouble<3>(ra, 3); tp[4] = tapToDouble<4>(rb, 0);
                }
                double up[4], mid[6], low[6];
#pragma unroll
                for (int x = 0; x < 4; ++x) up[x] = patchLoad(patch + (SPLIT2 ? patchIdxRt(x + 1, y0) : patchIdx(x + 1, 0)) * stride);
#pragma unroll
                for (int x = 0; x < 6; ++x) mid[x] = patchLoad(patch + (SPLIT2 ? patchIdxRt(x, y0 + 1) : patchIdx(x, 1)) * stride);
#pragma unroll
                for (int y = 0; y < NY; ++y) {
                  unsigned na, nb;
                  loadRow5(img + (size_t)(vi + y0 + y + 1) * pitch, ui, na, nb);
                  tn[0] = tapToDouble<0>(na, 0); tn[1] = tapToDouble<1>(na, 1); tn[2] = tapToDouble<2>(na, 2); tn[3] = tapToDouble<3>(na, 3); tn[4] = tapToDouble<4>(nb, 0);
#pragma unroll
                  for (int x = (y < NY - 1 ? 0 : 1); x < (y < NY - 1 ? 6 : 5); ++x)  // the thread's last row is only a "below" row: x = 1..4
                    low[x] = patchLoad(patch + (SPLIT2 ? patchIdxRt(x, y0 + y + 2) : patchIdx(x, y + 2)) * stride);
#pragma unroll
                  for (int x = 0; x < 4; ++x) {
                    const double ref = mid[x + 1];
                    // twice the central differences; the exact factor 0.5 is applied to the sums afterwards
                    const double dx2 = mid[x + 2] - mid[x];
                    const double dy2 = low[x + 1] - up[x];
                    double res;  // I_cur * (1 + alpha) + beta - I_ref, sparse_img_align.cpp:488-489
                    if (unit_gain) {  // alpha == beta == 0: the subtraction rides in the interpolation's FMA chain
                      res = fma(wbr, tn[x + 1], fma(wbl, tn[x], fma(wtr, tp[x + 1], fma(wtl, tp[x], -ref))));
                    } else {
                      const double I = wtl * tp[x] + wtr * tp[x + 1] + wbl * tn[x] + wbr * tn[x + 1];
                      res = (I * gain + beta_f) - ref;
                    }
                    if (ROBUST) {
                      const double w = (double)tukeyWeight((float)(res / wscale_f));
                      const double wdx = w * dx2, wdy = w * dy2;
                      ps.sxx += wdx * dx2; ps.sxy += wdx * dy2; ps.syy += wdy * dy2;
                      gx += wdx * res; gy += wdy * res;
                      chi += res * res * w;
                      if (ILLUM) {
                        const double a6 = est_gain ? -ref : 0.0, a7 = est_off ? -1.0 : 0.0;
                        ps.sx6 += wdx * a6; ps.sy6 += wdy * a6; ps.sx7 += wdx * a7; ps.sy7 += wdy * a7;
                        ps.s66 += w * a6 * a6; ps.s67 += w * a6 * a7; ps.s77 += w * a7 * a7;
                        g6 += w * a6 * res; g7 += w * a7 * res;
                      }
                    } else {
                      gx += dx2 * res; gy += dy2 * res;
                      chi += res * res;
                      if (ILLUM) {
                        if (est_gain) g6 -= ref * res;
                        if (est_off) g7 -= res;
                      }
                    }
                  }
#pragma unroll
                  for (int x = 0; x < 5; ++x) tp[x] = tn[x];
#pragma unroll
                  for (int x = 0; x < 4; ++x) up[x] = mid[x + 1];
#pragma unroll
                  for (int x = 0; x < 6; ++x) mid[x] = low[x];
                }
                gx *= 0.5; gy *= 0.5;
                if (ROBUST) {
                  ps.sxx *= 0.25; ps.sxy *= 0.25; ps.syy *= 0.25;
                  ps.sx6 *= 0.5; ps.sy6 *= 0.5; ps.sx7 *= 0.5; ps.sy7 *= 0.5;
                }
              }
            }
          }
          if (SPLIT2) __syncwarp();  // both threads of a patch have read the slot's flags
          if (s < n_now && y0 == 0) s_cam[s] = (uint8_t)(c | (vis ? 0x80 : 0));  // only the slot's own (first) thread writes it
          const bool changed = (iter == 0) || (vis != vis_before);
          const unsigned visb = __ballot_sync(0xffffffffu, vis);
          const unsigned chb = __ballot_sync(0xffffffffu, changed);
          if ((visb | chb) == 0u) continue;
          const int sl = (s < n_now) ? s : 0;
          if (ROBUST && visb) {
            double jp0[6], jp1[6];
            patchJacobian<DJ>(s_xyz + sl, s_aux + sl, stride, s_camblk + kCamBlk * c, fabs(P.cams[c].fx), scale, jp0, jp1);
            reduceH<D>(ps, jp0, jp1, red, lane);
          }
          if (visb) {
            // c = (mult * Jproj)^T [gx ; gy] in the camera frame and xyz_ref x c; rotated into the IMU frame after the reduction
            double gv[8];
            {
              const double X = s_xyz[sl], Y = s_xyz[stride + sl], Z = s_xyz[2 * stride + sl];
              double cx, cy, cz;
              if (DJ) {
                cx = s_aux[sl] * gx + s_aux[3 * stride + sl] * gy;
                cy = s_aux[stride + sl] * gx + s_aux[4 * stride + sl] * gy;
                cz = s_aux[2 * stride + sl] * gx + s_aux[5 * stride + sl] * gy;
              } else {
                const double iz = s_aux[sl], sI = -fabs(P.cams[c].fx) * iz;
                cx = sI * gx; cy = sI * gy;
                cz = -(X * cx + Y * cy) * iz;
              }
              gv[0] = cx; gv[1] = cy; gv[2] = cz;
              gv[3] = Y * cz - Z * cy; gv[4] = Z * cx - X * cz; gv[5] = X * cy - Y * cx;
              gv[6] = ILLUM ? -g6 : chi;
              gv[7] = ILLUM ? -g7 : 0.0;
            }
            for (int cc = 0; cc < n_cams; ++cc) {  // warps hold one camera except at a camera boundary
              const unsigned mine = __ballot_sync(0xffffffffu, vis && c == cc);
              if (mine == 0u) continue;
              double v8[8];
              const bool me = (mine >> lane) & 1u;
#pragma unroll
              for (int k = 0; k < 8; ++k) v8[k] = me ? gv[k] : 0.0;
              const double t = warpSumMulti<8>(v8, lane);  // lanes 4k hold value k
              const int k = lane >> 2;
              if ((lane & 3) == 0) {
                if (k < 6) red[iCM + 6 * cc + k] += t;
                else if (ILLUM) red[iG6 + (k - 6)] += t;
                else if (k == 6) red[iChi] += t;
              }
            }
            if (ILLUM) {
              const double cs = warpSum(chi);
              if (lane == 0) red[iChi] += cs;
            }
          }
          if (lane == 0) {
            red[iN] += (SPLIT2 ? 8.0 : 16.0) * __popc(visb);
            if (!ROBUST && chb) ctl.h_dirty = 1;  // benign race: every writer stores 1; read after the barrier, cleared by thread 0
          }
        }
        SVO_TM_ADD(3);
#ifdef SVO_ALIGN_TIMING
        if (lane == 0) ctl.tm[10 + warp] += clock64() - t_warp0;  // every warp's own residual pass
#endif
        __syncthreads();
        SVO_TM_ADD(4);
        bool h_fresh = ROBUST;
        if (!ROBUST) {
          // H depends only on the visible set: rebuild it when some patch entered or left the image (flag raised by the residual pass)
          h_fresh = ctl.h_dirty != 0;
          if (h_fresh) {
            for (int s = tid; s < n_round; s += TH) {
              const int sl = (s < n_now) ? s : 0;
              const unsigned cs = s_cam[sl];
              const bool vis = s < n_now && (cs & 0x80u) != 0u;
              if (__ballot_sync(0xffffffffu, vis) == 0u) continue;
              const int c = cs & 3u;
              PatchSums ps = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
              if (vis) ps = unitWeightSums<ILLUM>(s_patch + sl, stride, est_gain, est_off);
              double jp0[6], jp1[6];
              patchJacobian<DJ>(s_xyz + sl, s_aux + sl, stride, s_camblk + kCamBlk * c, fabs(P.cams[c].fx), scale, jp0, jp1);
              reduceH<D>(ps, jp0, jp1, red, lane);
            }
            __syncthreads();
            if (tid == 0) ctl.h_dirty = 0;  // every thread has read the flag (they all passed the barrier above)
#ifdef SVO_ALIGN_TIMING
            if (tid == kTimerTid) ctl.tm[9] += 1;
#endif
          }
        }
        SVO_TM_ADD(5);
        if (warp == kSerialWarp) {
          // ---- the serial part of an iteration, kept as short as its data dependences allow (one warp, 8-cycle FP64 latency,
          // ~30 cycles per shared-memory or shuffle hop): totals -> dx -> state update -> camera transforms.
          // cross-warp totals: lane k owns accumulator k; the H part is refreshed only when it was re-reduced
          for (int k = lane; k < NV; k += 32) {
            if (k < NH && !h_fresh) continue;
            const double* r0 = s_red + k;
            double t = r0[0];
#pragma unroll
            for (int w = 1; w < NW; ++w) t += r0[w * NV];
            s_tot[k] = t;
          }
          __syncwarp();
#ifdef SVO_ALIGN_TIMING
          long long t_s = 0;
          if (tid == kTimerTid) { t_s = clock64(); ctl.tm[16] += t_s - ctl.t_mark; }
#define SVO_TS(k) do { if (tid == kTimerTid) { const long long _t = clock64(); ctl.tm[k] += _t - t_s; t_s = _t; } } while (0)
#else
#define SVO_TS(k) do { } while (0)
#endif
#ifdef SVO_ALIGN_LEAN  // experiment: no prior path compiled in
          const bool serial_solve = ROBUST;
#else
          const bool serial_solve = ROBUST || P.priors != nullptr;
#endif
          double dxk = 0.0;  // lane k < D: dx[k]
          if (serial_solve) {
            // robust weights (H changes every iteration) or a prior (its gradient needs log(T_prior^-1 T)): lane 0 forms g and
            // solves with the LDL^T factor, as round 1 did for every case
            if (lane == 0) {
              double g[D], dx[8], padd[D];
#pragma unroll
              for (int a2 = 0; a2 < D; ++a2) { g[a2] = 0.0; padd[a2] = 0.0; }
              for (int cc = 0; cc < n_cams; ++cc) {  // gradient of the pose block from the per-camera sums
                const double* cb = s_camblk + kCamBlk * cc;
                const double* Ric = cb + 9;
                const double* cm = s_tot + iCM + 6 * cc;
                double bv[3], mv[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                  bv[r] = Ric[3 * r] * cm[0] + Ric[3 * r + 1] * cm[1] + Ric[3 * r + 2] * cm[2];
                  mv[r] = Ric[3 * r] * cm[3] + Ric[3 * r + 1] * cm[4] + Ric[3 * r + 2] * cm[5];
                }
                const double tx = cb[21], ty = cb[22], tz = cb[23];
                g[0] -= scale * bv[0]; g[1] -= scale * bv[1]; g[2] -= scale * bv[2];
                g[3] -= scale * (mv[0] + (ty * bv[2] - tz * bv[1]));
                g[4] -= scale * (mv[1] + (tz * bv[0] - tx * bv[2]));
                g[5] -= scale * (mv[2] + (tx * bv[1] - ty * bv[0]));
              }
              if (ILLUM) { g[D - 2] = s_tot[iG6]; g[D - 1] = s_tot[iG6 + 1]; }
              if (P.priors) {  // applyPrior, sparse_img_align_base.cpp:77-107
                const svo_align_prior& pr = P.priors[pair];
                if (iter == 0) {
                  double mt = 0, mr = 0;
#pragma unroll
                  for (int j = 0; j < 3; ++j) mt = fmax(mt, fabs(s_tot[j * D - (j * (j - 1)) / 2]));
#pragma unroll
                  for (int j = 3; j < 6; ++j) mr = fmax(mr, fabs(s_tot[j * D - (j * (j - 1)) / 2]));
                  for (int j = 0; j < 3; ++j) ctl.I_prior[j] = 1.0 * opt.lambda_trans * mt;
                  for (int j = 3; j < 6; ++j) ctl.I_prior[j] = 1.0 * opt.lambda_rot * mr;
                  ctl.I_prior[6] = (D == 8) ? opt.lambda_alpha * s_tot[6 * D - 15] : 0.0;
                  ctl.I_prior[7] = (D == 8) ? opt.lambda_beta * s_tot[7 * D - 21] : 0.0;
                }
                const SE3d Tp = se3Load(pr.T);
                const SE3d E = se3Mul(se3Inv(Tp), ctl.T);
                const V3d lr = quatLog(E.q);
                const double l[6] = {E.t.x, E.t.y, E.t.z, lr.x, lr.y, lr.z};
#pragma unroll
                for (int j = 0; j < 6; ++j) { padd[j] = ctl.I_prior[j]; g[j] += ctl.I_prior[j] * l[j]; }
                if (D == 8) {
                  padd[D - 2] = ctl.I_prior[6]; padd[D - 1] = ctl.I_prior[7];
                  g[D - 2] += ctl.I_prior[6] * (pr.alpha - ctl.alpha);
                  g[D - 1] += ctl.I_prior[7] * (pr.beta - ctl.beta);
                }
              }
              dx[6] = 0.0; dx[7] = 0.0;
              if (h_fresh) ldltFactor<D>(s_tot, padd, ctl.L, ctl.rd);  // H (+ prior) is constant until the visible set changes
              ldltSolve<D>(ctl.L, ctl.rd, g, dx);
#pragma unroll
              for (int i = 0; i < 8; ++i) ctl.dx[i] = dx[i];
            }
          } else {
            // H is constant between changes of the visible set, so dx = H^-1 g = (H^-1 M) * totals with g = M * totals linear in
            // the reduced sums (M: the rotation of the per-camera sums into the IMU frame, see the header comment). P = H^-1 M is
            // rebuilt with H; an iteration then costs one short dot product per lane instead of the gradient assembly and two
            // triangular solves on one lane.
            if (h_fresh) {
              if (lane == 0) {
                double padd[D];
#pragma unroll
                for (int a2 = 0; a2 < D; ++a2) padd[a2] = 0.0;
                ldltFactor<D>(s_tot, padd, ctl.L, ctl.rd);
              }
              __syncwarp();
              if (lane < D) {  // lane i: column i (= row i) of H^-1
                double e[D], x[D];
#pragma unroll
                for (int i = 0; i < D; ++i) e[i] = (i == lane) ? 1.0 : 0.0;
                ldltSolve<D>(ctl.L, ctl.rd, e, x);
#pragma unroll
                for (int i = 0; i < D; ++i) s_Hinv[lane * D + i] = x[i];
              }
              __syncwarp();
              for (int idx = lane; idx < D * NG; idx += 32) {
                const int k = idx / NG, j = idx - k * NG;
                const double* hk = s_Hinv + k * D;
                double v;
                if (j < 6 * n_cams) {
                  const int cc = j / 6, jj = j - 6 * cc;
                  const double* cb = s_camblk + kCamBlk * cc;
                  const double* Ric = cb + 9;
                  const double tx = cb[21], ty = cb[22], tz = cb[23];
                  const int col = jj < 3 ? jj : jj - 3;
                  const double r0c = Ric[col], r1c = Ric[3 + col], r2c = Ric[6 + col];
                  if (jj < 3) {  // translation sums: rows 0-2 through R_imu_cam, rows 3-5 through [t_imu_cam]x R_imu_cam
                    v = (hk[0] * r0c + hk[1] * r1c + hk[2] * r2c) +
                        (hk[3] * (ty * r2c - tz * r1c) + hk[4] * (tz * r0c - tx * r2c) + hk[5] * (tx * r1c - ty * r0c));
                  } else {       // moment sums: rows 3-5 through R_imu_cam
                    v = hk[3] * r0c + hk[4] * r1c + hk[5] * r2c;
                  }
                  v = -scale * v;
                } else {         // illumination gradient entries pass straight through
                  v = hk[6 + (j - 6 * n_cams)];
                }
                s_P[idx] = v;
              }
              __syncwarp();
            }
            if (lane < D) {
              const double* pk = s_P + lane * NG;
              const double* tg = s_tot + iCM;  // per-camera sums, then g6 g7: contiguous
              double acc0 = 0.0, acc1 = 0.0;
              for (int j = 0; j + 1 < NG; j += 2) { acc0 = fma(pk[j], tg[j], acc0); acc1 = fma(pk[j + 1], tg[j + 1], acc1); }
              dxk = acc0 + acc1;
            }
          }
          SVO_TS(17);
          if (!serial_solve && lane < 8) ctl.dx[lane] = dxk;  // lanes D..7 hold 0
          __syncwarp();
          SVO_TS(18);
          if (lane == 0) {
            // One lane applies the update out of shared memory: the residual pass owns the register file (80 registers at 4 CTAs / SM),
            // and anything spilled here would go to local memory, which misses the 28 KB of L1 left beside the patch store.
            double dx[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) dx[i] = ctl.dx[i];
            int stop = ctl.stop;
            if (dx[0] != dx[0]) stop = 1;  // solveDefaultImpl: isnan(dx[0]) -> stop_
            int brk = 0;
            ctl.iters[opt.max_level - level] = iter + 1;
            if (stop) {
              ctl.T = ctl.T_old; ctl.alpha = ctl.alpha_old; ctl.beta = ctl.beta_old;  // rollback (mini_least_squares_solver.hpp:76-84)
              ctl.stop = 1;
              brk = 1;
            } else {
              // update, sparse_img_align_base.cpp:64-75
              const SE3d Tc = ctl.T;
              SE3d inc;
              inc.q = quatExpFast(V3d{-dx[3], -dx[4], -dx[5]});
              inc.t = V3d{-dx[0], -dx[1], -dx[2]};
              SE3d Tn = se3Mul(Tc, inc);
              quatNormalizeFast(Tn.q);
              ctl.T_old = Tc;
              ctl.T = Tn;
              const double alpha_c = ctl.alpha, beta_c = ctl.beta;
              ctl.alpha_old = alpha_c; ctl.beta_old = beta_c;
              if (ILLUM) {
                ctl.alpha = (alpha_c - dx[6]) / (1.0 + dx[6]);
                ctl.beta = (beta_c - dx[7]) / (1.0 + dx[6]);
              }
              ctl.chi_num = s_tot[iChi]; ctl.chi_den = s_tot[iN]; ctl.chi_set = 1;  // chi2_ = new_chi2 = float chi2 / n_meas (:540)
              double x_norm = -1.0;
#pragma unroll
              for (int i = 0; i < 8; ++i) { const double a2 = fabs(dx[i]); if (a2 > x_norm) x_norm = a2; }
              if (x_norm < opt.eps) brk = 1;
            }
            ctl.alpha_f = (float)ctl.alpha;
            ctl.beta_f = (float)ctl.beta;
            ctl.brk = brk;
          }
          __syncwarp();
          SVO_TS(19);
          refreshCameraTransforms(ctl.T, s_camblk, n_cams, lane);
          SVO_TS(20);
        }
        SVO_TM_ADD(6);
        __syncthreads();
        SVO_TM_ADD(7);
#ifdef SVO_ALIGN_TIMING
        if (tid == kTimerTid) ctl.tm[8] += 1;
#endif
        if (ctl.brk) break;
      }
      __syncthreads();
    }
  }

  // ---- outputs (sparse_img_align.cpp:102-112) ----
  if (tid == 0) {
    svo_align_result& r = P.results[pair];
    const SE3d T_iref_world = se3Load(P.T_imu_world_ref + 7 * (size_t)pair);
    se3Store(ctl.T, r.T_icur_iref);
    for (int c = 0; c < SVO_MAX_CAMS; ++c) {
      if (c < n_cams) se3Store(se3Mul(se3Mul(se3Load(P.T_cam_imu[c]), ctl.T), T_iref_world), r.T_f_w[c]);
      else for (int k = 0; k < 7; ++k) r.T_f_w[c][k] = 0.0;
    }
    r.alpha = ctl.alpha;
    r.beta = ctl.beta;
    r.chi2 = ctl.chi_set ? (double)(float)(ctl.chi_num / ctl.chi_den) : 1e10;  // 0 / 0 = NaN when no patch was visible, as the reference
    // getHessian(): the last evaluated H_ (incl. the prior information) rebuilt from the reduced upper triangle
    for (int i = 0; i < 64; ++i) r.H[i] = 0.0;
    if (ctl.n_total > 0) {
      int idx = 0;
      for (int a = 0; a < D; ++a)
        for (int b = a; b < D; ++b) {  // (the result may live in mapped host memory: every element is written, none is read back)
          const double h = s_tot[idx] + ((a == b && P.priors) ? ctl.I_prior[a] : 0.0);
          r.H[a * 8 + b] = h; r.H[b * 8 + a] = h; ++idx;
        }
      if (P.priors) for (int j = D; j < 8; ++j) r.H[j * 8 + j] = ctl.I_prior[j];
    }
    r.n_tracked = ctl.n_total;
#ifdef SVO_ALIGN_TIMING
    ctl.tm[0] = clock64() - t_start;
    if (D == 6) {  // rows 6-7, then columns 6-7 of rows 0-5: unused by a 6-DoF result
      for (int i = 0; i < 16; ++i) r.H[48 + i] = (double)ctl.tm[i];
      for (int i = 16; i < 28; ++i) r.H[((i - 16) >> 1) * 8 + 6 + ((i - 16) & 1)] = (double)ctl.tm[i];
    }
#endif
    for (int i = 0; i < SVO_MAX_LEVELS; ++i) r.iters[i] = ctl.iters[i];
    r.stop = ctl.stop;
  }
}

inline size_t alignSmemBytes(int slots, int n_cams, bool illum, bool dj, int n_warps) {
  const int D = illum ? 8 : 6, NH = D * (D + 1) / 2;
  const int NV = NH + 6 * n_cams + (illum ? 2 : 0) + 3;
  const int NG = 6 * n_cams + (illum ? 2 : 0);
  const size_t doubles = (size_t)slots * (3 + (dj ? 6 : 1)) + (size_t)(n_warps + 1) * NV + (size_t)kCamBlk * n_cams + (size_t)D * D +
                         (size_t)D * NG + (NG & 1);
  return doubles * 8 + sizeof(Ctl) + (size_t)slots * 32 * sizeof(PatchT) + (size_t)slots * 5 + 16;
}

template <int ILL, bool ROBUST, bool DJ, int SLOTS, bool SPLIT2 = false>
cudaError_t launchAlign(const AlignParams& P, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(sparse_align_kernel<ILL, ROBUST, DJ, SLOTS, SPLIT2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  sparse_align_kernel<ILL, ROBUST, DJ, SLOTS, SPLIT2><<<P.B, (SLOTS && !SPLIT2) ? kThreads : kThreadsWide, smem, stream>>>(P);
  return cudaGetLastError();
}
template <int ILL, bool ROBUST>
cudaError_t launchAlignSel(const AlignParams& P, size_t smem, cudaStream_t stream, bool dj, bool fixed) {
  if (dj) return fixed ? launchAlign<ILL, ROBUST, true, kFixedSlots>(P, smem, stream) : launchAlign<ILL, ROBUST, true, 0>(P, smem, stream);
  return fixed ? launchAlign<ILL, ROBUST, false, kFixedSlots>(P, smem, stream) : launchAlign<ILL, ROBUST, false, 0>(P, smem, stream);
}

}  // namespace

extern "C" int svo_cuda_sparse_align(svo_cuda_ctx* ctx, int n_cams, const svo_cuda_pyr* const* ref_pyr,
                                     const svo_cuda_pyr* const* cur_pyr, const int* ref_frame_idx, const int* cur_frame_idx,
                                     const svo_camera* cams, const double* T_cam_imu, int B, const double* T_imu_world_ref,
                                     const double* T_imu_world_cur, const int* n_features, int max_features, const double* px,
                                     const double* f, const double* depth, const uint8_t* eligible,
                                     const svo_sparse_align_options* opt, const svo_align_prior* priors,
                                     svo_align_result* results, svo_mem mem) {
  if (!ctx || n_cams < 1 || n_cams > SVO_MAX_CAMS || !ref_pyr || !cur_pyr || !cams || !T_cam_imu || B < 0 || !T_imu_world_ref ||
      !T_imu_world_cur || !n_features || max_features < 1 || !px || !f || !depth || !eligible || !opt || !results)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_sparse_align: bad arguments");
  if (opt->min_level < 0 || opt->max_level < opt->min_level || opt->max_iter < 0 || opt->max_level - opt->min_level >= SVO_MAX_LEVELS)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_sparse_align: bad level range / max_iter");
  for (int c = 0; c < n_cams; ++c) {
    if (!ref_pyr[c] || !cur_pyr[c] || opt->max_level >= ref_pyr[c]->n_levels || opt->max_level >= cur_pyr[c]->n_levels)
      return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_sparse_align: pyramid has fewer levels than max_level+1");
    if (!ref_frame_idx && ref_pyr[c]->n_frames < B) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_sparse_align: ref batch < B");
    if (!cur_frame_idx && cur_pyr[c]->n_frames < B) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_sparse_align: cur batch < B");
  }
  if (B == 0) return SVO_OK;
  cudaSetDevice(ctx->device);

  AlignParams P;
  memset(&P, 0, sizeof(P));
  P.n_cams = n_cams; P.B = B; P.max_features = max_features;
  const int need = n_cams * max_features;
  const bool fixed = need <= kFixedSlots;
  const int slots = fixed ? kFixedSlots : ((need + 7) / 8) * 8;
  if (slots > 32 * kThreads) return SVO_FAIL(ctx, SVO_ERR_TOO_MANY_FEATURES, "svo_cuda_sparse_align: too many features per bundle");
  P.slots = slots;
  const bool illum = opt->estimate_illumination_gain || opt->estimate_illumination_offset;
  const bool robust = opt->robustification != 0;
  const bool dj = opt->use_distortion_jacobian != 0;
  const bool unit = (float)opt->alpha_init == 0.0f && (float)opt->beta_init == 0.0f;  // residual uses float alpha/beta
  // few pairs (SMs idle anyway): the variants without robust weights / distortion Jacobian run with two threads per patch (SVO_ALIGN_SPLIT=0 / 1 overrides: tests, A/B)
  bool split = fixed && !robust && !dj && B <= ctx->sm_count;
  if (const char* e = getenv("SVO_ALIGN_SPLIT")) split = fixed && !robust && !dj && atoi(e) != 0;
  size_t smem = alignSmemBytes(slots, n_cams, illum, dj, ((fixed && !split) ? kThreads : kThreadsWide) / 32);
  if (const char* pad = getenv("SVO_ALIGN_PAD_SMEM")) smem += (size_t)atoi(pad);  // occupancy experiments only
  if (smem > 227 * 1024) return SVO_FAIL(ctx, SVO_ERR_TOO_MANY_FEATURES, "svo_cuda_sparse_align: n_cams*max_features exceeds the shared-memory capacity (~1380 features per bundle)");
  for (int c = 0; c < n_cams; ++c) {
    P.ref_pyr[c] = makeView(ref_pyr[c]);
    P.cur_pyr[c] = makeView(cur_pyr[c]);
    P.cams[c] = cams[c];
    for (int k = 0; k < 7; ++k) P.T_cam_imu[c][k] = T_cam_imu[7 * c + k];
  }
  P.opt = *opt;
  Stager st(ctx, mem);
  const size_t nf = (size_t)B * n_cams * max_features;
  P.ref_frame_idx = st.in(ref_frame_idx, (size_t)B * n_cams);
  P.cur_frame_idx = st.in(cur_frame_idx, (size_t)B * n_cams);
  P.T_imu_world_ref = st.in(T_imu_world_ref, (size_t)B * 7);
  P.T_imu_world_cur = st.in(T_imu_world_cur, (size_t)B * 7);
  P.n_features = st.in(n_features, (size_t)B * n_cams);
  P.px = st.in(px, nf * 2);
  P.f = st.in(f, nf * 3);
  P.depth = st.in(depth, nf);
  P.eligible = st.in(eligible, nf);
  P.priors = st.in(priors, (size_t)B);
  P.results = st.outWriteOnly(results, (size_t)B);
  if (!st.send()) return st.finish();

  cudaError_t e;
  if (split) e = illum ? launchAlign<2, false, false, kFixedSlots, true>(P, smem, ctx->stream)
                       : (unit ? launchAlign<0, false, false, kFixedSlots, true>(P, smem, ctx->stream)
                               : launchAlign<1, false, false, kFixedSlots, true>(P, smem, ctx->stream));
  else if (illum) e = robust ? launchAlignSel<2, true>(P, smem, ctx->stream, dj, fixed) : launchAlignSel<2, false>(P, smem, ctx->stream, dj, fixed);
  else if (unit) e = robust ? launchAlignSel<0, true>(P, smem, ctx->stream, dj, fixed) : launchAlignSel<0, false>(P, smem, ctx->stream, dj, fixed);
  else e = robust ? launchAlignSel<1, true>(P, smem, ctx->stream, dj, fixed) : launchAlignSel<1, false>(P, smem, ctx->stream, dj, fixed);
  ctx->launches++;
  SVO_CUDA_TRY(ctx, e);
  return st.finish();
}
