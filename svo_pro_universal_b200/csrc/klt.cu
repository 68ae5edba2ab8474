// (f3, first "next" row of SURVEY §8f) feature_alignment::alignPyr2D / alignPyr2DVec — pyramidal inverse-compositional KLT
// with integer gradients and 7-bit fixed-point bilinear weights, the tracker FeatureTracker::trackFrameBundle runs.
// ref: src/svo_direct/src/feature_alignment.cpp:731-759 (alignPyr2DVec), :761-973 (alignPyr2D, non-NEON path)
//
// One warp per feature; lane l owns pixels l, l+32, ... of the patch in raster order. Template (u8) and its int16
// gradients live in shared memory. The reference accumulates H and Jres as float sums in raster order; every term is an
// integer (|dx|, |dy|, |res| <= 255) and a patch has at most 16x16 pixels, so every partial sum stays below
// 255*255*256 = 16 646 400 < 2^24: float accumulation is exact in ANY order and equals the int32 sum, which is what the
// warp reduces. Everything else (float position update, the 2x2 inverse, double <-> float conversions) follows the
// reference operation by operation; this translation unit is compiled with -fmad=false.
#include "common.cuh"

namespace {

constexpr int kKltThreads = 128;
constexpr int kKltWarps = kKltThreads / 32;
constexpr int kMaxPatch = 16;
constexpr int kMaxArea = kMaxPatch * kMaxPatch;

struct KltParams {
  PyrView ref_pyr, cur_pyr;
  const int* ref_frame_idx;
  const int* cur_frame_idx;
  int M, max_level, min_level, n_iter;
  int patch_sizes[SVO_MAX_LEVELS];
  float min_update_squared;
  const int* px_ref;  // [M][2] level-0 integer pixel of the reference feature
  double* px_cur;     // [M][2] in/out
  uint8_t* status;    // [M] out
};

SVO_D int warpSumInt(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kKltThreads) klt_pyr2d_kernel(const KltParams P) {
  __shared__ uint8_t s_patch[kKltWarps][kMaxArea];
  __shared__ short s_dx[kKltWarps][kMaxArea];
  __shared__ short s_dy[kKltWarps][kMaxArea];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kKltWarps + warp;
  if (i >= P.M) return;
  uint8_t* patch = s_patch[warp];
  short* pdx = s_dx[warp];
  short* pdy = s_dy[warp];
  const int rf = P.ref_frame_idx ? P.ref_frame_idx[i] : 0;
  const int cf = P.cur_frame_idx ? P.cur_frame_idx[i] : 0;
  const int px0 = P.px_ref[2 * i], py0 = P.px_ref[2 * i + 1];
  double cur_x = P.px_cur[2 * i], cur_y = P.px_cur[2 * i + 1];
  bool converged = false, failed = false;

  for (int level = P.max_level; level >= P.min_level && !failed; --level) {
    const int patch_size = P.patch_sizes[level];
    const int halfpatch_size = patch_size / 2;
    const int shift = patch_size == 16 ? 4 : 3;
    const int area = patch_size * patch_size;
    const int scale = (1 << level);
    const int width = P.ref_pyr.cols[level], height = P.ref_pyr.rows[level];
    const uint8_t* img_ref = P.ref_pyr.level(rf, level);
    const uint8_t* img_cur = P.cur_pyr.level(cf, level);
    const int step_ref = P.ref_pyr.pitch[level], step_cur = P.cur_pyr.pitch[level];
    // feature_alignment.cpp:797-800
    const float prx = (float)px0 / (float)scale - (float)halfpatch_size, pry = (float)py0 / (float)scale - (float)halfpatch_size;
    const int rx = (int)prx, ry = (int)pry;
    const float offx = prx - (float)rx, offy = pry - (float)ry;
    if (rx < 1 || ry < 1 || rx >= width - patch_size - 1 || ry >= height - patch_size - 1) continue;  // :801-808

    // template, gradients (twice the central difference) and H = sum J J^T (:810-827)
    __syncwarp();
    int sxx = 0, sxy = 0, syy = 0;
    for (int p = lane; p < area; p += 32) {
      const int y = p >> shift, x = p & (patch_size - 1);
      const uint8_t* it = img_ref + (size_t)(ry + y) * step_ref + (rx + x);
      const int dx = (int)it[1] - (int)it[-1];
      const int dy = (int)it[step_ref] - (int)it[-step_ref];
      patch[p] = it[0];
      pdx[p] = (short)dx;
      pdy[p] = (short)dy;
      sxx += dx * dx; sxy += dx * dy; syy += dy * dy;
    }
    __syncwarp();
    const float H00 = (float)warpSumInt(sxx), H01 = (float)warpSumInt(sxy), H11 = (float)warpSumInt(syy);
    // Eigen::Matrix2f::inverse(): compute_inverse_size2_helper
    const float invdet = 1.0f / (H00 * H11 - H01 * H01);
    const float Hi00 = H11 * invdet, Hi01 = -H01 * invdet, Hi10 = -H01 * invdet, Hi11 = H00 * invdet;

    // :830-832 (double arithmetic, then narrowed)
    float u = (float)(cur_x / scale - halfpatch_size - offx);
    float v = (float)(cur_y / scale - halfpatch_size - offy);
    bool go_to_next_level = false;
    converged = false;
    for (int iter = 0; iter < P.n_iter; ++iter) {
      if (u != u || v != v) { failed = true; converged = false; break; }  // :841-847 (returns false)
      go_to_next_level = false;
      const int u_r = (int)floorf(u), v_r = (int)floorf(v);
      if (u_r < 0 || v_r < 0 || u_r >= width - patch_size || v_r >= height - patch_size) {  // :851-861
        go_to_next_level = true;
        break;
      }
      const float subpix_x = u - (float)u_r, subpix_y = v - (float)v_r;
      const int wTL = (int)(unsigned short)((1.0f - subpix_x) * (1.0f - subpix_y) * 128.0f);
      const int wTR = (int)(unsigned short)(subpix_x * (1.0f - subpix_y) * 128.0f);
      const int wBL = (int)(unsigned short)((1.0f - subpix_x) * subpix_y * 128.0f);
      const int wBR = (int)(unsigned short)(128 - wTL - wTR - wBL);
      int j0 = 0, j1 = 0;
      for (int p = lane; p < area; p += 32) {
        const int y = p >> shift, x = p & (patch_size - 1);
        const uint8_t* it = img_cur + (size_t)(v_r + y) * step_cur + (u_r + x);
        const int cur = (int)(unsigned short)((wTL * it[0] + wTR * it[1] + wBL * it[step_cur] + wBR * it[step_cur + 1] + 64) >> 7);
        const int res = cur - (int)patch[p];
        j0 += res * (int)pdx[p];
        j1 += res * (int)pdy[p];
      }
      // Jres -= res * grad in floats, exact (see the header), starting from +0.0f
      const float Jres0 = 0.0f - (float)warpSumInt(j0), Jres1 = 0.0f - (float)warpSumInt(j1);
      const float up0 = (Hi00 * Jres0 + Hi01 * Jres1) * 2.0f, up1 = (Hi10 * Jres0 + Hi11 * Jres1) * 2.0f;  // :949
      u += up0;
      v += up1;
      if (up0 * up0 + up1 * up1 < P.min_update_squared) { converged = true; break; }
    }
    if (failed) break;
    cur_x = (double)((u + (float)halfpatch_size + offx) * (float)scale);  // :967-968
    cur_y = (double)((v + (float)halfpatch_size + offy) * (float)scale);
    if (!converged && !go_to_next_level) failed = true;  // :969-970
  }
  if (lane == 0) {
    P.px_cur[2 * i] = cur_x;
    P.px_cur[2 * i + 1] = cur_y;
    P.status[i] = (converged && !failed) ? 1 : 0;
  }
}

}  // namespace

extern "C" int svo_cuda_align_pyr2d(svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const svo_cuda_pyr* cur_pyr, const int* ref_frame_idx,
                                    const int* cur_frame_idx, int M, const int* px_ref_level_0, double* px_cur, int max_level,
                                    int min_level, const int* patch_sizes, int n_iter, float min_update_squared, uint8_t* status,
                                    svo_mem mem) {
  if (!ctx || !ref_pyr || !cur_pyr || M < 0 || !px_ref_level_0 || !px_cur || !patch_sizes || !status)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_align_pyr2d: bad arguments");
  if (min_level < 0 || max_level < min_level || max_level >= ref_pyr->n_levels || max_level >= cur_pyr->n_levels || n_iter < 0)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_align_pyr2d: bad level range / n_iter");
  for (int l = min_level; l <= max_level; ++l) {
    if (patch_sizes[l] != 8 && patch_sizes[l] != 16)  // the reference CHECKs patch_size % 8 == 0 (:789); 8 and 16 are what it ships
      return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_align_pyr2d: patch sizes must be 8 or 16");
    if (ref_pyr->cols[l] != cur_pyr->cols[l] || ref_pyr->rows[l] != cur_pyr->rows[l])
      return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_align_pyr2d: ref and cur pyramids differ in size");
  }
  if (M == 0) return SVO_OK;
  cudaSetDevice(ctx->device);
  Stager st(ctx, mem);
  KltParams P;
  memset(&P, 0, sizeof(P));
  P.ref_pyr = makeView(ref_pyr);
  P.cur_pyr = makeView(cur_pyr);
  P.ref_frame_idx = st.in(ref_frame_idx, (size_t)M);
  P.cur_frame_idx = st.in(cur_frame_idx, (size_t)M);
  P.M = M; P.max_level = max_level; P.min_level = min_level; P.n_iter = n_iter;
  for (int l = 0; l < SVO_MAX_LEVELS; ++l) P.patch_sizes[l] = (l >= min_level && l <= max_level) ? patch_sizes[l] : 8;
  P.min_update_squared = min_update_squared;
  P.px_ref = st.in(px_ref_level_0, (size_t)M * 2);
  P.px_cur = st.inout(px_cur, (size_t)M * 2);
  P.status = st.out(status, (size_t)M);
  if (!st.send()) return st.finish();
  klt_pyr2d_kernel<<<(M + kKltWarps - 1) / kKltWarps, kKltThreads, 0, ctx->stream>>>(P);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}
