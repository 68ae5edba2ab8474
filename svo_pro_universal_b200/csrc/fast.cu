// (a2-a5) Pyramidal FAST detector with 3x3 non-max suppression and grid-cell arg-max.
//
// ref: src/svo_direct/src/feature_detection_utils.cpp:145-194 (fastDetector)
//      src/fast_neon/src/faster_corner_10_sse.cpp:15-202, fast_10.cpp (segment test, region [3,w-3)x[3,h-3))
//      src/fast_neon/src/fast_10_score.cpp:21-3178 (score = largest barrier at which the pixel is still a corner)
//      src/fast_neon/src/nonmax_3x3.cpp:17-112 (suppress when any 8-neighbour corner has score >= own)
//      src/svo_common/include/svo/common/occupancy_grid_2d.h:82-95 (cell index)
//
// Closed form used instead of the generated decision trees: with d_i = I_i - p on the 16-pixel circle,
//   margin = max over the 16 arcs of ARC contiguous pixels of max(min_arc d_i, -max_arc d_i) - 1
// the pixel is a corner at barrier b iff margin >= b, and fast_corner_score_10 = max(b, margin).
//
// Kernel layout: a CTA owns a 64x16 interior tile; the u8 tile with a 4-pixel halo is staged in shared memory
// with aligned 32-bit loads; phase A writes scores for the interior + 1 ring; phase B does the 3x3 non-max,
// border / occupancy tests and a 64-bit atomicMax per grid cell whose key reproduces the reference's visiting
// order (score desc, then level asc, y asc, x asc — `score > corners[k].score` is strict).
#include "common.cuh"

namespace {

constexpr int kTW = 64, kTH = 16, kHalo = 4;
constexpr int kSW = kTW + 2 * kHalo;          // 72 staged columns
constexpr int kSH = kTH + 2 * kHalo;          // 24 staged rows
constexpr int kScW = kTW + 2, kScH = kTH + 2; // score region (interior + 1 ring)
constexpr int kScPitch = kScW + 2;            // 68

template <int ARC>
SVO_D int fastMargin(const uint8_t* p, int pitch) {
  const int c = p[0];
  int d[16];
  d[0] = p[3 * pitch] - c;       d[1] = p[3 * pitch + 1] - c;   d[2] = p[2 * pitch + 2] - c;   d[3] = p[pitch + 3] - c;
  d[4] = p[3] - c;               d[5] = p[-pitch + 3] - c;      d[6] = p[-2 * pitch + 2] - c;  d[7] = p[-3 * pitch + 1] - c;
  d[8] = p[-3 * pitch] - c;      d[9] = p[-3 * pitch - 1] - c;  d[10] = p[-2 * pitch - 2] - c; d[11] = p[-pitch - 3] - c;
  d[12] = p[-3] - c;             d[13] = p[pitch - 3] - c;      d[14] = p[2 * pitch - 2] - c;  d[15] = p[3 * pitch - 1] - c;
  // sliding min / max over ARC contiguous entries, log-step
  int mn2[16], mx2[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { mn2[k] = min(d[k], d[(k + 1) & 15]); mx2[k] = max(d[k], d[(k + 1) & 15]); }
  int mn4[16], mx4[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { mn4[k] = min(mn2[k], mn2[(k + 2) & 15]); mx4[k] = max(mx2[k], mx2[(k + 2) & 15]); }
  // bright = max over arcs of the arc minimum, dark = min over arcs of the arc maximum; margin = max(bright, -dark) - 1.
  // NB: nvcc 12.9 / sm_100a folds a negated operand into VIMNMX3 incorrectly (max(a, max(b, -c)) returns wrong values,
  // see tools/mm_test.cu), so the negation is kept out of every min/max chain and hidden behind an asm barrier.
  int bright = -256, dark = 256;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int mn8 = min(mn4[k], mn4[(k + 4) & 15]);
    const int mx8 = max(mx4[k], mx4[(k + 4) & 15]);
    int mn, mx;
    if (ARC == 10) { mn = min(mn8, mn2[(k + 8) & 15]); mx = max(mx8, mx2[(k + 8) & 15]); }
    else           { mn = min(mn8, d[(k + 8) & 15]);   mx = max(mx8, d[(k + 8) & 15]); }
    bright = max(bright, mn);
    dark = min(dark, mx);
  }
  int ndark = -dark;
  asm volatile("" : "+r"(ndark));
  return max(bright, ndark) - 1;
}

struct FastParams {
  int level, threshold, border, cell_size, n_cols, n_cells, first;
  unsigned long long* keys;        // [count][n_cells] or nullptr
  const uint8_t* occupancy;        // [count][n_cells] or nullptr
  short* score_map;                // dense debug maps for one frame (level coords) or nullptr
  uint8_t* nonmax_map;
};

template <int ARC>
__global__ void __launch_bounds__(256) fast_level_kernel(PyrView v, FastParams P) {
  __shared__ __align__(16) uint8_t s_img[kSH * kSW];
  __shared__ short s_score[kScH * kScPitch];
  const int L = P.level;
  const int cols = v.cols[L], rows = v.rows[L], pitch = v.pitch[L];
  const int frame_local = blockIdx.z;
  const uint8_t* img = v.level(P.first + frame_local, L);
  const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
  const int tid = threadIdx.x;

  // stage tile + halo: words of 4 px, aligned because x0 - 4 is a multiple of 4 and rows are 16-B aligned
  for (int i = tid; i < kSH * (kSW / 4); i += 256) {
    const int r = i / (kSW / 4), cw = i - r * (kSW / 4);
    const int gy = y0 - kHalo + r, gx = x0 - kHalo + cw * 4;
    unsigned w = 0;
    if (gy >= 0 && gy < rows && gx >= 0 && gx < pitch) w = __ldg(reinterpret_cast<const unsigned*>(img + (size_t)gy * pitch + gx));
    *reinterpret_cast<unsigned*>(&s_img[r * kSW + cw * 4]) = w;
  }
  __syncthreads();

  // phase A: scores on interior + 1 ring
  for (int i = tid; i < kScH * kScW; i += 256) {
    const int r = i / kScW, c = i - r * kScW;
    const int gy = y0 - 1 + r, gx = x0 - 1 + c;
    short sc = 0;
    if (gx >= 3 && gy >= 3 && gx < cols - 3 && gy < rows - 3) {
      const uint8_t* p = &s_img[(r + kHalo - 1) * kSW + (c + kHalo - 1)];
      // cheap necessary condition: an arc of >= 9 contiguous circle pixels contains 2 adjacent compass points
      const int cpx = p[0];
      const int hi = cpx + P.threshold, lo = cpx - P.threshold;
      const int n = p[-3 * kSW], s = p[3 * kSW], e = p[3], w = p[-3];
      const bool bright = ((n > hi) + (e > hi) + (s > hi) + (w > hi)) >= 2;
      const bool dark = ((n < lo) + (e < lo) + (s < lo) + (w < lo)) >= 2;
      if (bright || dark) {
        const int m = fastMargin<ARC>(p, kSW);
        if (m >= P.threshold) sc = (short)m;  // score = max(threshold, margin) = margin; threshold >= 1 so 0 means "no corner"
      }
    }
    s_score[r * kScPitch + c] = sc;
  }
  __syncthreads();

  // phase B: non-max, border, cell arg-max
  const int scale = 1 << L;
  for (int i = tid; i < kTH * kTW; i += 256) {
    const int r = i / kTW, c = i - r * kTW;
    const int gy = y0 + r, gx = x0 + c;
    if (gx >= cols || gy >= rows) continue;
    const short* sp = &s_score[(r + 1) * kScPitch + (c + 1)];
    const int sc = sp[0];
    if (P.score_map) P.score_map[(size_t)gy * cols + gx] = (short)sc;
    bool keep = sc > 0;
    if (keep) {
      keep = !(sp[-kScPitch - 1] >= sc || sp[-kScPitch] >= sc || sp[-kScPitch + 1] >= sc || sp[-1] >= sc || sp[1] >= sc ||
               sp[kScPitch - 1] >= sc || sp[kScPitch] >= sc || sp[kScPitch + 1] >= sc);
    }
    if (P.nonmax_map) P.nonmax_map[(size_t)gy * cols + gx] = keep ? 1 : 0;
    if (!keep || !P.keys) continue;
    if (gx < P.border || gy < P.border || gx >= cols - P.border || gy >= rows - P.border) continue;
    const int k = ((gy * scale) / P.cell_size) * P.n_cols + (gx * scale) / P.cell_size;
    if (P.occupancy && P.occupancy[(size_t)frame_local * P.n_cells + k]) continue;
    const unsigned order = ((unsigned)L << 28) | ((unsigned)gy << 14) | (unsigned)gx;
    const unsigned long long key = ((unsigned long long)(unsigned)sc << 32) | (unsigned long long)(0xFFFFFFFFu - order);
    atomicMax(&P.keys[(size_t)frame_local * P.n_cells + k], key);
  }
}

__global__ void fast_keys_init_kernel(unsigned long long* keys, size_t n, int threshold) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((unsigned long long)(unsigned)threshold << 32) | 0xFFFFFFFFull;
}

__global__ void fast_keys_decode_kernel(const unsigned long long* keys, size_t n, int threshold, svo_corner* out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  const int sc = (int)(key >> 32);
  svo_corner c;
  if (sc > threshold) {
    const unsigned order = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
    const int L = order >> 28, y = (order >> 14) & 0x3FFF, x = order & 0x3FFF;
    c.x = x << L; c.y = y << L; c.level = L; c.score = (float)sc; c.angle = 0.0f;
  } else {
    c.x = 0; c.y = 0; c.level = 0; c.score = (float)threshold; c.angle = 0.0f;
  }
  out[i] = c;
}

int launchLevel(svo_cuda_ctx* ctx, const PyrView& v, const FastParams& P, int arc, int count) {
  dim3 grid((v.cols[P.level] + kTW - 1) / kTW, (v.rows[P.level] + kTH - 1) / kTH, count);
  if (arc == 9) fast_level_kernel<9><<<grid, 256, 0, ctx->stream>>>(v, P);
  else fast_level_kernel<10><<<grid, 256, 0, ctx->stream>>>(v, P);
  SVO_LAUNCH_CHECK(ctx);
  return SVO_OK;
}

}  // namespace

int svoPyrBuildLaunch(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr, int first, int count);

static int fastDetectImpl(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, const svo_detector_options* opt,
                          const uint8_t* occupancy_in, svo_corner* corners_out, svo_mem mem) {
  if (!ctx || !pyr || !opt || !corners_out || first < 0 || count < 0 || first + count > pyr->n_frames)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_fast_detect: bad arguments");
  if (opt->min_level < 0 || opt->max_level < opt->min_level || opt->max_level >= pyr->n_levels || opt->cell_size <= 0 ||
      opt->threshold < 1 || opt->threshold > 254 || opt->max_level > 15 || pyr->cols[0] >= 16384 || pyr->rows[0] >= 16384)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_fast_detect: bad detector options");
  const int arc = opt->arc_length == 9 ? 9 : 10;
  if (count == 0) return SVO_OK;
  int n_cols, n_rows;
  const int n_cells = svo_cuda_grid_cells(pyr->cols[0], pyr->rows[0], opt->cell_size, &n_cols, &n_rows);
  const size_t n = (size_t)n_cells * count;
  Stager st(ctx, mem);
  const uint8_t* d_occ = st.in(occupancy_in, n);
  svo_corner* d_out = st.out(corners_out, n);
  unsigned long long* keys = (unsigned long long*)st.scratch(n * sizeof(unsigned long long));
  if (st.failed()) return st.finish();
  fast_keys_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(keys, n, opt->threshold);
  SVO_LAUNCH_CHECK(ctx);
  const PyrView v = makeView(pyr);
  for (int L = opt->min_level; L <= opt->max_level; ++L) {
    FastParams P;
    P.level = L; P.threshold = opt->threshold; P.border = opt->border; P.cell_size = opt->cell_size;
    P.n_cols = n_cols; P.n_cells = n_cells; P.first = first;
    P.keys = keys; P.occupancy = d_occ; P.score_map = nullptr; P.nonmax_map = nullptr;
    const int rc = launchLevel(ctx, v, P, arc, count);
    if (rc != SVO_OK) return rc;
  }
  fast_keys_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(keys, n, opt->threshold, d_out);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

extern "C" {

int svo_cuda_fast_detect(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, const svo_detector_options* opt,
                         const uint8_t* occupancy_in, svo_corner* corners_out, svo_mem mem) {
  return fastDetectImpl(ctx, pyr, first, count, opt, occupancy_in, corners_out, mem);
}

int svo_cuda_pyramid_fast_detect(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr, int first, int count, const svo_detector_options* opt,
                                 const uint8_t* occupancy_in, svo_corner* corners_out, svo_mem mem) {
  if (!ctx || !pyr || first < 0 || count < 0 || first + count > pyr->n_frames)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_pyramid_fast_detect: bad arguments");
  const int rc = svoPyrBuildLaunch(ctx, pyr, first, count);
  if (rc != SVO_OK) return rc;
  return fastDetectImpl(ctx, pyr, first, count, opt, occupancy_in, corners_out, mem);
}

int svo_cuda_fast_level_maps(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int frame, int level, int threshold, int arc_length,
                             int16_t* score_map, uint8_t* nonmax_map, svo_mem mem) {
  if (!ctx || !pyr || frame < 0 || frame >= pyr->n_frames || level < 0 || level >= pyr->n_levels || threshold < 1 || threshold > 254)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_fast_level_maps: bad arguments");
  const size_t n = (size_t)pyr->cols[level] * pyr->rows[level];
  Stager st(ctx, mem);
  int16_t* d_sc = st.out(score_map, n);
  uint8_t* d_nm = st.out(nonmax_map, n);
  if (st.failed()) return st.finish();
  FastParams P;
  P.level = level; P.threshold = threshold; P.border = 0; P.cell_size = 1; P.n_cols = 1; P.n_cells = 1; P.first = frame;
  P.keys = nullptr; P.occupancy = nullptr; P.score_map = d_sc; P.nonmax_map = d_nm;
  const int rc = launchLevel(ctx, makeView(pyr), P, arc_length == 9 ? 9 : 10, 1);
  if (rc != SVO_OK) return rc;
  return st.finish();
}

}  // extern "C"
