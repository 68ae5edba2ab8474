// (f2) Edgelet detector: Gaussian 3x3 -> Scharr -> gradient-magnitude score -> the reference's 8-neighbour test ->
// grid-cell arg-max -> gradient-orientation histogram of the winners; and the FastGrad combination (FAST corners first,
// edgelets in the cells FAST left empty).
//
// ref: src/svo_direct/src/feature_detection_utils.cpp:313-385 (edgeletDetector_V2), :831-839 + :945-1009 (angle histogram),
//      src/svo_direct/src/feature_detection.cpp:130-194 (GradientDetectorGrid::detect, FastGradDetector::detect),
//      OpenCV 8-bit GaussianBlur(3x3, sigma 0) = (1-2-1 x 1-2-1 window sum + 8) >> 4 and Scharr 8U -> 16S = exact 3-10-3
//      differences (both pinned against cv2 4.13, tests/golden/cv_imgproc_golden.npz).
//
// The detector works on pyramid level 1 only and reports level 0 with px = 2 * level-1 pixel. Quirk kept: the reference
// offsets a float pointer by `score.step` (bytes), so its "vertical" neighbours are FOUR rows away: a pixel survives when
// score(x+1,y) < s, score(x-1,y) <= s, score(x,y+4) < s, score(x,y-4) <= s, score(x+-1,y+4) < / <= s, score(x+-1,y-4) < / <= s.
// Scores exist on [border, cols-border) x [border, rows-border) and are 0 elsewhere; border >= 4 keeps the reference inside
// its score map and is required here (it also means no BORDER_REFLECT_101 pixel is ever read).
//
// Kernel layout: one CTA per 96x48 tile of the level-1 image of one frame. The u8 tile with a 3 (x) / 6 (y) pixel halo is staged
// in shared memory with aligned 32-bit loads, blurred into a second u8 tile (halo 2 / 5), turned into a float score tile
// (halo 1 / 4; the float is the reference's float(std::sqrt(double(int)))), and the survivors of the neighbour test go to a 64-bit
// atomicMax per grid cell (score bits << 32 | ~raster order: strict `>` keeps the first of equal scores in raster order).
// A second kernel, one warp per cell, builds the 9x9 orientation histogram of each winner: lanes evaluate atan2 / sqrt for the
// 81 pixels, the per-bin sums are then formed in the reference's raster order (lane b walks the 81 terms of bin b), smoothed
// and arg-maxed.
#include "common.cuh"

namespace {

constexpr int kETW = 96, kETH = 48;
constexpr int kIX = 4, kIY = 6;                 // staged image halo (x halo 3, rounded up to a word)
constexpr int kIPitch = kETW + 2 * kIX;         // 104 bytes
constexpr int kIRows = kETH + 2 * kIY;          // 60
constexpr int kBX = 2, kBY = 5;                 // blur halo
constexpr int kBPitch = kETW + 2 * kBX + 4;     // 104 (100 used)
constexpr int kBRows = kETH + 2 * kBY;          // 58
constexpr int kSX = 1, kSY = 4;                 // score halo
constexpr int kSPitchE = kETW + 2 * kSX + 1;    // 99 floats (odd pitch: the +-4-row reads of a warp hit different banks)
constexpr int kSRowsE = kETH + 2 * kSY;         // 56
constexpr int kThreadsE = 256;

inline unsigned divMagicE(int d) { return d <= 1 ? 0u : (unsigned)(0x100000000ull / (unsigned)d + 1ull); }
SVO_D int divFastE(int x, unsigned magic) { return magic ? (int)__umulhi((unsigned)x, magic) : x; }

struct EdgeletParams {
  int first, threshold, border, n_cols, n_cells, tiles_x;
  unsigned cell_magic, tiles_x_magic;
  unsigned long long* keys;   // [count][n_cells]
  const uint8_t* occupancy;   // [count][n_cells] or nullptr
};

// float(std::sqrt(double(n))) for 0 <= n < 2^31. Below 2^24 the int -> float conversion is exact and the correctly rounded float
// square root equals the double square root rounded to float (sqrt of an integer < 2^24 is never within double precision of a
// float midpoint unless it is an integer); above, take the double path.
SVO_D float sqrtIntAsFloat(int n) {
  if (n < (1 << 24)) return __fsqrt_rn((float)n);
  return (float)sqrt((double)n);
}

__global__ void __launch_bounds__(kThreadsE) edgelet_score_kernel(PyrView v, EdgeletParams P) {
  __shared__ __align__(16) uint8_t s_img[kIRows * kIPitch];
  __shared__ __align__(16) uint8_t s_blur[kBRows * kBPitch];
  __shared__ float s_score[kSRowsE * kSPitchE];
  const int tid = threadIdx.x;
  const int ty = divFastE(blockIdx.x, P.tiles_x_magic), tx = blockIdx.x - ty * P.tiles_x;
  const int cols = v.cols[1], rows = v.rows[1], pitch = v.pitch[1];
  const int frame_local = blockIdx.y;
  const uint8_t* img = v.level(P.first + frame_local, 1);
  const int x0 = tx * kETW, y0 = ty * kETH;

  // stage: words [x0 - 4, x0 + 100) of rows [y0 - 6, y0 + 54), rows clamped, words outside the row read as 0 (never used:
  // scores exist only `border` >= 4 pixels inside the image)
  for (int i = tid; i < kIRows * (kIPitch / 4); i += kThreadsE) {
    const int r = i / (kIPitch / 4), w = i - r * (kIPitch / 4);
    const int gy = min(max(y0 - kIY + r, 0), rows - 1), gx = x0 - kIX + 4 * w;
    unsigned val = 0;
    if (gx >= 0 && gx < pitch) val = __ldg(reinterpret_cast<const unsigned*>(img + (size_t)gy * pitch + gx));
    reinterpret_cast<unsigned*>(s_img)[i] = val;
  }
  __syncthreads();

  // blur tile: pixel (bx, by) of the blur tile is image pixel (x0 - 2 + bx, y0 - 5 + by) = s_img[(by + 1) * pitch + bx + 2]
  for (int i = tid; i < kBRows * (kETW + 2 * kBX); i += kThreadsE) {
    const int by = i / (kETW + 2 * kBX), bx = i - by * (kETW + 2 * kBX);
    const uint8_t* p = &s_img[(by + 1) * kIPitch + bx + 2];
    const int sum = (p[-kIPitch - 1] + p[-kIPitch + 1] + p[kIPitch - 1] + p[kIPitch + 1]) +
                    2 * (p[-kIPitch] + p[kIPitch] + p[-1] + p[1]) + 4 * p[0];
    s_blur[by * kBPitch + bx] = (uint8_t)((sum + 8) >> 4);
  }
  __syncthreads();

  // score tile: pixel (sx, sy) is image pixel (x0 - 1 + sx, y0 - 4 + sy) = s_blur[(sy + 1) * pitch + sx + 1]
  const int thr = P.threshold;
  const long long thr2 = (long long)thr * thr;
  for (int i = tid; i < kSRowsE * (kETW + 2 * kSX); i += kThreadsE) {
    const int sy = i / (kETW + 2 * kSX), sx = i - sy * (kETW + 2 * kSX);
    const int gx = x0 - kSX + sx, gy = y0 - kSY + sy;
    float sc = 0.0f;
    if (gx >= P.border && gy >= P.border && gx < cols - P.border && gy < rows - P.border) {
      const uint8_t* p = &s_blur[(sy + 1) * kBPitch + sx + 1];
      const int a = p[-kBPitch - 1], b = p[-kBPitch], c = p[-kBPitch + 1];
      const int d = p[-1], e = p[1];
      const int f = p[kBPitch - 1], g = p[kBPitch], h = p[kBPitch + 1];
      const int dx = 3 * ((c - a) + (h - f)) + 10 * (e - d);
      const int dy = 3 * ((f - a) + (h - c)) + 10 * (g - b);
      const int n = dx * dx + dy * dy;
      if ((long long)n > thr2) {  // necessary for mag > threshold; the float comparison below is the reference's
        const float mag = sqrtIntAsFloat(n);
        if (mag > (float)thr) sc = mag;
      }
    }
    s_score[sy * kSPitchE + sx] = sc;
  }
  __syncthreads();

  // neighbour test + cell arg-max
  for (int i = tid; i < kETH * kETW; i += kThreadsE) {
    const int r = i / kETW, c = i - r * kETW;
    const float* q = &s_score[(r + kSY) * kSPitchE + c + kSX];
    const float s = q[0];
    if (s == 0.0f) continue;  // 0 = below the threshold, outside the scored region or outside the image
    if (q[1] >= s || q[-1] > s) continue;
    if (q[4 * kSPitchE] >= s || q[-4 * kSPitchE] > s) continue;
    if (q[4 * kSPitchE + 1] >= s || q[4 * kSPitchE - 1] > s) continue;
    if (q[-4 * kSPitchE + 1] >= s || q[-4 * kSPitchE - 1] > s) continue;
    const int gx = x0 + c, gy = y0 + r;
    const int k = divFastE(2 * gy, P.cell_magic) * P.n_cols + divFastE(2 * gx, P.cell_magic);
    if (P.occupancy && P.occupancy[(size_t)frame_local * P.n_cells + k]) continue;
    const unsigned order = ((unsigned)gy << 14) | (unsigned)gx;
    const unsigned long long key = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - order);
    atomicMax(&P.keys[(size_t)frame_local * P.n_cells + k], key);
  }
}

__global__ void edgelet_keys_init_kernel(unsigned long long* keys, size_t n, int threshold) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((unsigned long long)__float_as_uint((float)threshold) << 32) | 0xFFFFFFFFull;
}

// angle_hist::angleHistogram bin of a central-difference gradient: round(36 * (atan2(gy, gx) + pi) / (2 pi)), 36 -> 0.
// The only gradients whose angle sits on a bin boundary are the diagonals (|gx| == |gy|: 45 / 135 degrees = 22.5, 31.5, 4.5,
// 13.5 bins); there the result depends on the last bit of atan2, so those angles are the correctly rounded constants glibc
// returns; everywhere else the distance to a boundary (>= 1e-6 rad) dwarfs any 2-ulp difference between libm implementations.
SVO_D int angleBin(int gx, int gy) {
  const double kPi = 3.14159265358979323846;
  double angle;
  if (gx != 0 && (gx == gy || gx == -gy)) {
    const double a = gx > 0 ? 0.78539816339744830962 : 2.35619449019234492885;
    angle = gy > 0 ? a : -a;
  } else {
    angle = atan2((double)gy, (double)gx);
  }
  const double t = __ddiv_rn(__dmul_rn(36.0, __dadd_rn(angle, kPi)), __dmul_rn(2.0, kPi));
  const unsigned long long bin = (unsigned long long)round(t);
  return bin < 36ull ? (int)bin : 0;
}

constexpr int kHistWarps = 8;

// One warp per (frame, cell): decode the winner and compute its gradient-orientation histogram angle (half patch 4).
__global__ void __launch_bounds__(kHistWarps * 32) edgelet_decode_kernel(PyrView v, int first, const unsigned long long* keys,
                                                                       int n_cells, size_t n, int threshold, svo_corner* out) {
  __shared__ double s_mag[kHistWarps][81];
  __shared__ int8_t s_bin[kHistWarps][84];
  __shared__ double s_hist[kHistWarps][36];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t i = (size_t)blockIdx.x * kHistWarps + warp;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  const float sc = __uint_as_float((unsigned)(key >> 32));
  svo_corner c;
  if (!(sc > (float)threshold)) {
    if (lane == 0) { c.x = 0; c.y = 0; c.level = 0; c.score = (float)threshold; c.angle = 0.0f; out[i] = c; }
    return;
  }
  const unsigned order = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
  const int py = (order >> 14) & 0x3FFF, px = order & 0x3FFF;
  const int frame = first + (int)(i / (size_t)n_cells);
  const uint8_t* img = v.level(frame, 1);
  const int cols = v.cols[1], rows = v.rows[1], pitch = v.pitch[1];
  for (int t = lane; t < 81; t += 32) {
    const int dv = t / 9, du = t - dv * 9;
    const int u = px + du - 4, w = py + dv - 4;
    int bin = -1;
    double mag = 0.0;
    if (w > 0 && w < rows - 1 && u > 0 && u < cols - 1) {
      const int gx = (int)img[(size_t)w * pitch + u + 1] - (int)img[(size_t)w * pitch + u - 1];
      const int gy = (int)img[(size_t)(w + 1) * pitch + u] - (int)img[(size_t)(w - 1) * pitch + u];
      mag = sqrt((double)(gx * gx + gy * gy));
      bin = angleBin(gx, gy);
    }
    s_mag[warp][t] = mag;
    s_bin[warp][t] = (int8_t)bin;
  }
  __syncwarp();
  // ordered per-bin sums: lane b owns bins b and b + 32 and walks the 81 terms in raster order
  double h0 = 0.0, h1 = 0.0;
  for (int t = 0; t < 81; ++t) {
    const int b = s_bin[warp][t];
    const double m = s_mag[warp][t];
    if (b == lane) h0 = __dadd_rn(h0, m);
    if (b == lane + 32) h1 = __dadd_rn(h1, m);
  }
  s_hist[warp][lane] = h0;
  if (lane < 4) s_hist[warp][lane + 32] = h1;
  __syncwarp();
  // circular 1-2-1 smoothing on the un-smoothed neighbours, then the first maximum
  double sm0, sm1 = 0.0;
  {
    const double* h = s_hist[warp];
    const int b = lane;
    sm0 = __dadd_rn(__dadd_rn(__dmul_rn(0.25, h[(b + 35) % 36]), __dmul_rn(0.5, h[b])), __dmul_rn(0.25, h[(b + 1) % 36]));
    if (lane < 4) {
      const int b1 = lane + 32;
      sm1 = __dadd_rn(__dadd_rn(__dmul_rn(0.25, h[b1 - 1]), __dmul_rn(0.5, h[b1])), __dmul_rn(0.25, h[(b1 + 1) % 36]));
    }
  }
  __syncwarp();
  s_hist[warp][lane] = sm0;
  if (lane < 4) s_hist[warp][lane + 32] = sm1;
  __syncwarp();
  if (lane == 0) {
    int best = 0;
    double bv = s_hist[warp][0];
    for (int b = 1; b < 36; ++b)
      if (s_hist[warp][b] > bv) { bv = s_hist[warp][b]; best = b; }
    const double kPi = 3.14159265358979323846;
    const double angle = __ddiv_rn(__dmul_rn(__dmul_rn((double)best, 2.0), kPi), 36.0);
    c.x = 2 * px; c.y = 2 * py; c.level = 0; c.score = sc; c.angle = (float)angle;
    out[i] = c;
  }
}

__global__ void angle_bin_table_kernel(int8_t* out) {  // bins of every gradient (gx, gy) in [-255, 255]^2, row = gy + 255
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 511 * 511) out[i] = (int8_t)angleBin(i % 511 - 255, i / 511 - 255);
}

// FastGradDetector::detect between its two stages (feature_detection.cpp:166-179): cells that received a FAST corner become
// occupied (fillFeatures, mask empty), and when the corners already fill max_n_features the edgelet stage is skipped.
__global__ void fastgrad_merge_kernel(const svo_corner* fast_corners, const uint8_t* occ_in, int n_cells, int threshold_primary,
                                      int max_n_features, uint8_t* occ_out) {
  __shared__ int s_count;
  const int frame = blockIdx.x;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  int local = 0;
  for (int k = threadIdx.x; k < n_cells; k += blockDim.x) {
    const bool found = fast_corners[(size_t)frame * n_cells + k].score > (float)threshold_primary;
    local += found;
    occ_out[(size_t)frame * n_cells + k] = (found || (occ_in && occ_in[(size_t)frame * n_cells + k])) ? 1 : 0;
  }
  atomicAdd(&s_count, local);
  __syncthreads();
  if (max_n_features - min(s_count, max_n_features) <= 0)
    for (int k = threadIdx.x; k < n_cells; k += blockDim.x) occ_out[(size_t)frame * n_cells + k] = 1;
}

int edgeletDeviceImpl(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, int threshold, int border, int cell_size,
                      const uint8_t* d_occ, svo_corner* d_out, unsigned long long* keys) {
  int n_cols, n_rows;
  const int n_cells = svo_cuda_grid_cells(pyr->cols[0], pyr->rows[0], cell_size, &n_cols, &n_rows);
  const size_t n = (size_t)n_cells * count;
  edgelet_keys_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(keys, n, threshold);
  SVO_LAUNCH_CHECK(ctx);
  const PyrView v = makeView(pyr);
  EdgeletParams P;
  P.threshold = threshold; P.border = border; P.n_cols = n_cols; P.n_cells = n_cells;
  P.cell_magic = divMagicE(cell_size);
  P.tiles_x = (v.cols[1] + kETW - 1) / kETW;
  P.tiles_x_magic = divMagicE(P.tiles_x);
  const int n_tiles = P.tiles_x * ((v.rows[1] + kETH - 1) / kETH);
  for (int f0 = 0; f0 < count; f0 += 65535) {  // grid.y limit
    const int nf = min(65535, count - f0);
    P.first = first + f0;
    P.keys = keys + (size_t)f0 * n_cells;
    P.occupancy = d_occ ? d_occ + (size_t)f0 * n_cells : nullptr;
    edgelet_score_kernel<<<dim3(n_tiles, nf, 1), kThreadsE, 0, ctx->stream>>>(v, P);
    SVO_LAUNCH_CHECK(ctx);
  }
  edgelet_decode_kernel<<<(unsigned)((n + kHistWarps - 1) / kHistWarps), kHistWarps * 32, 0, ctx->stream>>>(v, first, keys, n_cells, n,
                                                                                                        threshold, d_out);
  SVO_LAUNCH_CHECK(ctx);
  return SVO_OK;
}

int checkEdgeletArgs(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, int threshold, int border, int cell_size,
                     const char* who) {
  if (!ctx || !pyr || first < 0 || count < 0 || first + count > pyr->n_frames) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, who);
  if (pyr->n_levels < 2 || cell_size <= 0 || threshold < 0 || threshold > 46340 || border < 4 || pyr->cols[0] >= 16384 ||
      pyr->rows[0] >= 16384)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, who);
  return SVO_OK;
}

}  // namespace

int svoFastDetectImpl(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, const svo_detector_options* opt,
                      const uint8_t* occupancy_in, svo_corner* corners_out, svo_mem mem);

extern "C" {

int svo_cuda_edgelet_detect(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, int threshold, int border,
                            int cell_size, const uint8_t* occupancy_in, svo_corner* corners_out, svo_mem mem) {
  const int rc = checkEdgeletArgs(ctx, pyr, first, count, threshold, border, cell_size, "svo_cuda_edgelet_detect: bad arguments");
  if (rc != SVO_OK) return rc;
  if (!corners_out) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_edgelet_detect: corners_out is NULL");
  if (count == 0) return SVO_OK;
  const size_t n = (size_t)svo_cuda_grid_cells(pyr->cols[0], pyr->rows[0], cell_size, nullptr, nullptr) * count;
  Stager st(ctx, mem);
  const uint8_t* d_occ = st.in(occupancy_in, n);
  svo_corner* d_out = st.out(corners_out, n);
  unsigned long long* keys = (unsigned long long*)st.scratch(n * sizeof(unsigned long long));
  if (st.failed() || !keys) return st.finish();
  const int rc2 = edgeletDeviceImpl(ctx, pyr, first, count, threshold, border, cell_size, d_occ, d_out, keys);
  if (rc2 != SVO_OK) return rc2;
  return st.finish();
}

int svo_cuda_angle_histogram_bins(svo_cuda_ctx* ctx, int8_t* bins_out, svo_mem mem) {
  if (!ctx || !bins_out) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_angle_histogram_bins: bad arguments");
  Stager st(ctx, mem);
  int8_t* d = st.out(bins_out, (size_t)511 * 511);
  if (st.failed()) return st.finish();
  angle_bin_table_kernel<<<(511 * 511 + 255) / 256, 256, 0, ctx->stream>>>(d);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

int svo_cuda_fastgrad_detect(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, const svo_detector_options* opt,
                             int threshold_secondary, int max_n_features, const uint8_t* occupancy_in, svo_corner* corners_out,
                             svo_corner* edgelets_out, svo_mem mem) {
  if (!opt) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_fastgrad_detect: options are NULL");
  const int rc = checkEdgeletArgs(ctx, pyr, first, count, threshold_secondary, opt->border, opt->cell_size,
                                  "svo_cuda_fastgrad_detect: bad arguments");
  if (rc != SVO_OK) return rc;
  if (!corners_out || !edgelets_out || max_n_features < 0)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_fastgrad_detect: bad arguments");
  if (count == 0) return SVO_OK;
  const int n_cells = svo_cuda_grid_cells(pyr->cols[0], pyr->rows[0], opt->cell_size, nullptr, nullptr);
  const size_t n = (size_t)n_cells * count;
  Stager st(ctx, mem);
  const uint8_t* d_occ = st.in(occupancy_in, n);
  svo_corner* d_fast = st.out(corners_out, n);
  svo_corner* d_edge = st.out(edgelets_out, n);
  unsigned long long* keys = (unsigned long long*)st.scratch(n * sizeof(unsigned long long));
  uint8_t* d_occ2 = (uint8_t*)st.scratch(n);
  if (st.failed() || !keys || !d_occ2) return st.finish();
  int rc2 = svoFastDetectImpl(ctx, pyr, first, count, opt, d_occ, d_fast, SVO_MEM_DEVICE);
  if (rc2 != SVO_OK) return rc2;
  fastgrad_merge_kernel<<<count, 128, 0, ctx->stream>>>(d_fast, d_occ, n_cells, opt->threshold, max_n_features, d_occ2);
  SVO_LAUNCH_CHECK(ctx);
  rc2 = edgeletDeviceImpl(ctx, pyr, first, count, threshold_secondary, opt->border, opt->cell_size, d_occ2, d_edge, keys);
  if (rc2 != SVO_OK) return rc2;
  return st.finish();
}

}  // extern "C"
