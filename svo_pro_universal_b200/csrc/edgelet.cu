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
// Kernel layout: one CTA per 96x48 tile of the level-1 image of one frame. The u8 tile with a 4 (x) / 6 (y) pixel halo is staged
// in shared memory with aligned 32-bit loads, blurred into a second u8 tile (y halo 5), turned into a float score tile
// (y halo 4; the float is the reference's float(std::sqrt(double(int)))), and the survivors of the neighbour test go to a 64-bit
// atomicMax per grid cell (score bits << 32 | ~raster order: strict `>` keeps the first of equal scores in raster order).
// Blur and Scharr run on 2 x 4 pixels per thread in 16-bit SIMD lanes of ordinary 32-bit integer instructions (byte-permute to
// split a word into even / odd pixels, biased lanes so no borrow crosses a lane), and the score tile holds the squared magnitude
// as an integer (see suppressedSlow) so that the float square root is taken for the few survivors only.
// A second kernel, one warp per cell, builds the 9x9 orientation histogram of each winner: lanes evaluate atan2 / sqrt for the
// 81 pixels, the per-bin sums are then formed in the reference's raster order (lane b walks the 81 terms of bin b), smoothed
// and arg-maxed.
#include "common.cuh"

namespace {

constexpr int kETW = 96, kETH = 48;
// Shared-memory tiles. Image and blur rows are 32 words = 128 bytes starting at the 16-byte aligned pixel x0 - 16 (word w covers
// x0 - 16 + 4 w ... + 3), so rows are staged with 16-byte cp.async; rows: image y0 - 6 ..., blur y0 - 5 ..., score y0 - 4 ...
// Blur is computed for words 2 .. 29, scores for words 3 .. 28 (pixels x0 - 4 .. x0 + 99; score column c = pixel x0 - 4 + c).
constexpr int kPitchW = 32;
constexpr int kBlurW0 = 2, kBlurWords = 28;
constexpr int kScoreW0 = 3, kWords = 26;
constexpr int kIRows = kETH + 12;               // 60 image rows
constexpr int kBRows = kETH + 10;               // 58 blur rows
constexpr int kSRowsE = kETH + 8;               // 56 score rows
constexpr int kSPitchE = 4 * kWords;            // 104 ints
constexpr int kThreadsE = 256;
constexpr int kExactBelow = 1 << 22;            // squared magnitudes below this map to distinct floats (see suppressedSlow)

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
__device__ __noinline__ float sqrtIntAsFloat(int n) {  // out of line: rare (survivors of the neighbour test, huge gradients)
  if (n < (1 << 24)) return __fsqrt_rn((float)n);
  return (float)sqrt((double)n);
}

// The score tile holds the SQUARED gradient magnitude n = dx^2 + dy^2 (0 = no score) instead of the reference's float
// mag(n) = float(sqrt(n)). mag is monotone, strictly so below 2^22 (sqrt(n+1) - sqrt(n) > one float ulp of sqrt(n) there); above,
// up to 5 consecutive integers collapse onto one float (n < 2^25.1: ulp <= 2^-11, sqrt(b) - sqrt(a) < ulp needs b - a <= 5).
// Hence every comparison of the reference is decided by the integers except in a window of 8 around equality at n >= 2^22,
// where the floats are formed and compared — out of line: the straight-line code of the common path must stay small (a version
// with the float path inlined at every comparison was instruction-fetch bound and 5x slower), and rare (a version that took
// the float path for every n >= 2^22 spent a quarter of its instructions there on sharp synthetic edges).
// Only the survivors of the neighbour test need their float score (arg-max key, output).
constexpr int kCollapse = 8;
// ge / gt = the largest neighbour compared with >= / > (the test "any neighbour's mag >= mag(s)" is the test on the largest one)
__device__ __noinline__ bool suppressedSlow(int ge, int gt, int s) {
  const float fs = sqrtIntAsFloat(s);
  return sqrtIntAsFloat(ge) >= fs || sqrtIntAsFloat(gt) > fs;
}
__device__ __noinline__ int scoreSlow(int n, int thr) { return sqrtIntAsFloat(n) > (float)thr ? n : 0; }
// High word of the per-cell arg-max key: ordered like mag(n) and equal exactly when the floats are equal. Below 2^22 that is n
// itself (the float is formed once per cell by the decode kernel); from 2^22 on it is the bit pattern of the float, which is
// >= 0x45000000 (2048.0f) and so stays above every small key.
__device__ __noinline__ unsigned scoreKeySlow(int n) { return __float_as_uint(sqrtIntAsFloat(n)); }
SVO_D unsigned scoreKey(int n) { return n < kExactBelow ? (unsigned)n : scoreKeySlow(n); }
SVO_D float scoreOfKey(unsigned k) { return k < (unsigned)kExactBelow ? __fsqrt_rn((float)k) : __uint_as_float(k); }

// Four neighbouring pixels x .. x+3 of one row as 16-bit lanes: "even" registers hold (x, x+2), "odd" ones (x+1, x+3).
//   le = (x-1, x+1), ce = (x, x+2), co = (x+1, x+3), ro = (x+2, x+4)
struct Lanes { unsigned le, ce, co, ro; };
SVO_D Lanes splitRow(const unsigned* r) {  // r points at the word of x .. x+3; r[-1] and r[1] are staged too
  const unsigned w0 = r[-1], w1 = r[0], w2 = r[1];
  const unsigned l = __byte_perm(w0, w1, 0x6543);  // bytes x-1, x, x+1, x+2
  const unsigned rr = __byte_perm(w1, w2, 0x4321); // bytes x+1, x+2, x+3, x+4
  Lanes o;
  o.le = l & 0x00FF00FFu;
  o.ce = w1 & 0x00FF00FFu;
  o.co = __byte_perm(w1, 0u, 0x4341);
  o.ro = __byte_perm(rr, 0u, 0x4341);
  return o;
}

// squared magnitude of one pixel from its biased Scharr responses (dx + 4096, dy + 4096); 0 unless mag > threshold
SVO_D int edgeletScore(unsigned bdx, unsigned bdy, int thr, int thr2) {
  const int dx = (int)bdx - 4096, dy = (int)bdy - 4096;
  const int n = dx * dx + dy * dy;
  if (n <= thr2) return 0;  // mag(thr^2) = thr exactly, so n <= thr^2 means mag <= threshold
  // thresholds >= 2048 only: within the collapse window above thr^2 the float comparison decides
  if (n >= kExactBelow && n <= thr2 + kCollapse) return scoreSlow(n, thr);
  return n;
}

#ifndef SVO_EDGELET_MIN_CTAS
#define SVO_EDGELET_MIN_CTAS 5  // A/B on the B200: 4 -> 1.567 ms, 5 -> 1.548 ms per 1024 frames of FastGrad (shared memory allows 5)
#endif
__global__ void __launch_bounds__(kThreadsE, SVO_EDGELET_MIN_CTAS) edgelet_score_kernel(PyrView v, EdgeletParams P) {
  __shared__ __align__(16) unsigned s_img[kIRows * kPitchW];
  __shared__ __align__(16) unsigned s_blur[kBRows * kPitchW];
  __shared__ __align__(16) int s_score[kSRowsE * kSPitchE];
  const int tid = threadIdx.x;
  const int ty = divFastE(blockIdx.x, P.tiles_x_magic), tx = blockIdx.x - ty * P.tiles_x;
  const int cols = v.cols[1], rows = v.rows[1], pitch = v.pitch[1];
  const int frame_local = blockIdx.y;
  const uint8_t* img = v.level(P.first + frame_local, 1);
  const int x0 = tx * kETW, y0 = ty * kETH;

  // stage 60 rows x 128 bytes with 16-byte cp.async (x0 - 16 and the row pitch are multiples of 16); rows are clamped, chunks
  // outside the row are zero-filled (src-size 0) — never used for a score: scores exist only `border` >= 4 pixels inside the image
  for (int i = tid; i < kIRows * (kPitchW / 4); i += kThreadsE) {
    const int r = i >> 3, c = i & 7;
    const int gy = min(max(y0 - 6 + r, 0), rows - 1), gx = x0 - 16 + 16 * c;
    const bool in = gx >= 0 && gx < pitch;
    const uint8_t* src = img + (size_t)gy * pitch + (in ? gx : 0);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_img[r * kPitchW + 4 * c]);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  // blur: an item is (row pair, word) = 2 x 4 pixels from 4 image rows. Horizontal 1-2-1 sums in 16-bit lanes (<= 1020), vertical
  // 1-2-1 (<= 4080), + 8, >> 4.
  // (a warp owns a row pair, lane = word: 28 of 32 lanes busy, no two lanes of a warp on one shared-memory bank)
  const int warp = tid >> 5, lane = tid & 31;
  for (int rp = warp; rp < kBRows / 2; rp += kThreadsE / 32) {
    if (lane >= kBlurWords) continue;
    const int w = lane + kBlurW0;
    const unsigned* r = &s_img[(2 * rp) * kPitchW + w];  // blur row b reads image rows b, b+1, b+2
    unsigned he[4], ho[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const Lanes a = splitRow(r + j * kPitchW);
      he[j] = a.le + 2u * a.ce + a.co;
      ho[j] = a.ce + 2u * a.co + a.ro;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const unsigned be = ((he[j] + 2u * he[j + 1] + he[j + 2] + 0x00080008u) >> 4) & 0x00FF00FFu;
      const unsigned bo = ((ho[j] + 2u * ho[j + 1] + ho[j + 2] + 0x00080008u) >> 4) & 0x00FF00FFu;
      s_blur[(2 * rp + j) * kPitchW + w] = __byte_perm(be, bo, 0x6240);
    }
  }
  __syncthreads();

  // score: an item is (row pair, word) = 2 x 4 pixels from 4 blur rows. Scharr in biased 16-bit lanes:
  //   D = b(x+1) - b(x-1) + 256, H = 3 b(x-1) + 10 b(x) + 3 b(x+1);  dx + 4096 = 3 D(y-1) + 10 D(y) + 3 D(y+1),
  //   dy + 4096 = H(y+1) - H(y-1) + 4096.
  const int thr = P.threshold, thr2 = thr * thr;
  for (int rp = warp; rp < kSRowsE / 2; rp += kThreadsE / 32) {
    if (lane >= kWords) continue;
    const int wi = lane;
    const unsigned* r = &s_blur[(2 * rp) * kPitchW + wi + kScoreW0];  // score row s reads blur rows s, s+1, s+2
    unsigned de[4], dod[4], he[4], ho[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const Lanes a = splitRow(r + j * kPitchW);
      de[j] = a.co + 0x01000100u - a.le;
      dod[j] = a.ro + 0x01000100u - a.ce;
      he[j] = 3u * (a.le + a.co) + 10u * a.ce;
      ho[j] = 3u * (a.ce + a.ro) + 10u * a.co;
    }
    const int gx = x0 - 4 + 4 * wi;
    const int lo = P.border - gx, hi = cols - P.border - gx;  // pixel k of the word is scored iff lo <= k < hi
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int gy = y0 - 4 + 2 * rp + j;
      int4 sc = make_int4(0, 0, 0, 0);
      if (gy >= P.border && gy < rows - P.border && hi > 0 && lo < 4) {
        const unsigned dxe = 3u * (de[j] + de[j + 2]) + 10u * de[j + 1];
        const unsigned dxo = 3u * (dod[j] + dod[j + 2]) + 10u * dod[j + 1];
        const unsigned dye = he[j + 2] + 0x10001000u - he[j];
        const unsigned dyo = ho[j + 2] + 0x10001000u - ho[j];
        sc.x = edgeletScore(dxe & 0xFFFFu, dye & 0xFFFFu, thr, thr2);
        sc.y = edgeletScore(dxo & 0xFFFFu, dyo & 0xFFFFu, thr, thr2);
        sc.z = edgeletScore(dxe >> 16, dye >> 16, thr, thr2);
        sc.w = edgeletScore(dxo >> 16, dyo >> 16, thr, thr2);
        if (lo > 0 || hi < 4) {  // the word straddles the border of the scored region
          if (!(0 >= lo && 0 < hi)) sc.x = 0;
          if (!(1 >= lo && 1 < hi)) sc.y = 0;
          if (!(2 >= lo && 2 < hi)) sc.z = 0;
          if (!(3 >= lo && 3 < hi)) sc.w = 0;
        }
      }
      *reinterpret_cast<int4*>(&s_score[(2 * rp + j) * kSPitchE + 4 * wi]) = sc;
    }
  }
  __syncthreads();

  // neighbour test + cell arg-max: an item is 4 interior pixels; their 3 x 6 neighbourhood (rows y-4, y, y+4) is read with three
  // 128-bit and six 32-bit loads, and the four tests run in registers
  for (int i = tid; i < kETH * (kETW / 4); i += kThreadsE) {
    const int r = i / (kETW / 4), w = i - r * (kETW / 4);
    const int* q4 = &s_score[(r + 4) * kSPitchE + 4 + 4 * w];
    const int4 s4 = *reinterpret_cast<const int4*>(q4);
    if ((s4.x | s4.y | s4.z | s4.w) == 0) continue;  // 0 = below threshold / not scored / outside
    const int4 a4 = *reinterpret_cast<const int4*>(q4 + 4 * kSPitchE), b4 = *reinterpret_cast<const int4*>(q4 - 4 * kSPitchE);
    const int c[6] = {q4[-1], s4.x, s4.y, s4.z, s4.w, q4[4]};
    const int a[6] = {q4[4 * kSPitchE - 1], a4.x, a4.y, a4.z, a4.w, q4[4 * kSPitchE + 4]};   // row y + 4
    const int b[6] = {q4[-4 * kSPitchE - 1], b4.x, b4.y, b4.z, b4.w, q4[-4 * kSPitchE + 4]}; // row y - 4
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int s = c[k + 1];
      if (s == 0) continue;
      const int ge = max(max(c[k + 2], a[k + 1]), max(a[k + 2], b[k + 2]));  // (+1,0) (0,+4) (+1,+4) (+1,-4): compared with >=
      const int gt = max(max(c[k], b[k + 1]), max(a[k], b[k]));              // (-1,0) (0,-4) (-1,+4) (-1,-4): compared with >
      if (ge >= s || gt > s + kCollapse) continue;              // suppressed for certain (mag is monotone)
      if (gt > s || ge >= s - kCollapse) {                      // inside the window around equality
        if (max(gt, s) < kExactBelow) { if (gt > s) continue; } // exact integer domain: ge < s, so only gt decides
        else if (suppressedSlow(ge, gt, s)) continue;
      }
      const int gx = x0 + 4 * w + k, gy = y0 + r;
      const int cell = divFastE(2 * gy, P.cell_magic) * P.n_cols + divFastE(2 * gx, P.cell_magic);
      if (P.occupancy && P.occupancy[(size_t)frame_local * P.n_cells + cell]) continue;
      const unsigned order = ((unsigned)gy << 14) | (unsigned)gx;
      atomicMax(&P.keys[(size_t)frame_local * P.n_cells + cell], ((unsigned long long)scoreKey(s) << 32) | (unsigned long long)(0xFFFFFFFFu - order));
    }
  }
}

__global__ void edgelet_keys_init_kernel(unsigned long long* keys, size_t n, int threshold) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const int t2 = threshold * threshold;  // mag(threshold^2) == threshold: an empty cell holds the key of score = threshold
  if (i < n) keys[i] = ((unsigned long long)(t2 < kExactBelow ? (unsigned)t2 : __float_as_uint((float)threshold)) << 32) | 0xFFFFFFFFull;
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

// A CTA decodes 256 cells: empty ones at once, then one warp per winner computes the gradient-orientation histogram angle (half patch 4).
__global__ void __launch_bounds__(kHistWarps * 32) edgelet_decode_kernel(PyrView v, int first, const unsigned long long* keys,
                                                                       int n_cells, size_t n, int threshold,
                                                                       const int8_t* __restrict__ bins, svo_corner* out) {
  __shared__ double s_mag[kHistWarps][81];
  __shared__ int8_t s_bin[kHistWarps][84];
  __shared__ double s_hist[kHistWarps][36];
  __shared__ int s_list[kHistWarps * 32];
  __shared__ int s_count;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // pass 1, one thread per cell: empty cells are written at once, winners are compacted (most cells are empty when FAST corners
  // already occupy the grid: ~17 winners per frame of 416 cells)
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  {
    const size_t ci = (size_t)blockIdx.x * (kHistWarps * 32) + threadIdx.x;
    if (ci < n) {
      const float sc0 = scoreOfKey((unsigned)(keys[ci] >> 32));
      if (sc0 > (float)threshold) {
        s_list[atomicAdd(&s_count, 1)] = threadIdx.x;
      } else {
        svo_corner e;
        e.x = 0; e.y = 0; e.level = 0; e.score = (float)threshold; e.angle = 0.0f;
        out[ci] = e;
      }
    }
  }
  __syncthreads();
  // pass 2, one warp per winner
  for (int wi = warp; wi < s_count; wi += kHistWarps) {
  const size_t i = (size_t)blockIdx.x * (kHistWarps * 32) + s_list[wi];
  const unsigned long long key = keys[i];
  const float sc = scoreOfKey((unsigned)(key >> 32));
  svo_corner c;
  const unsigned order = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
  const int py = (order >> 14) & 0x3FFF, px = order & 0x3FFF;
  const int frame = first + (int)(i / (size_t)n_cells);
  const uint8_t* img = v.level(frame, 1);
  const int cols = v.cols[1], rows = v.rows[1], pitch = v.pitch[1];
  bool any_high = false;
  for (int t = lane; t < 81; t += 32) {
    const int dv = t / 9, du = t - dv * 9;
    const int u = px + du - 4, w = py + dv - 4;
    int bin = -1;
    double mag = 0.0;
    if (w > 0 && w < rows - 1 && u > 0 && u < cols - 1) {
      const int gx = (int)img[(size_t)w * pitch + u + 1] - (int)img[(size_t)w * pitch + u - 1];
      const int gy = (int)img[(size_t)(w + 1) * pitch + u] - (int)img[(size_t)(w - 1) * pitch + u];
      mag = sqrt((double)(gx * gx + gy * gy));
      bin = bins[(gy + 255) * 511 + gx + 255];  // angleBin(gx, gy), tabulated once per context (FP64 atan2 was 85 % of this kernel)
    }
    s_mag[warp][t] = mag;
    s_bin[warp][t] = (int8_t)bin;
    any_high |= bin >= 32;
  }
  const bool high_bins = __any_sync(0xFFFFFFFFu, any_high);
  __syncwarp();
  // ordered per-bin sums: lane b owns bins b and b + 32 and walks the 81 terms in raster order
  // (branch-free: adding +0.0 leaves a non-negative partial sum unchanged; bins 32 .. 35 get their own pass only when one occurs)
  double h0 = 0.0, h1 = 0.0;
#pragma unroll 9
  for (int t = 0; t < 81; ++t) h0 = __dadd_rn(h0, s_bin[warp][t] == lane ? s_mag[warp][t] : 0.0);
  if (high_bins) {
    for (int t = 0; t < 81; ++t) h1 = __dadd_rn(h1, s_bin[warp][t] == lane + 32 ? s_mag[warp][t] : 0.0);
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
  __syncwarp();
  }  // winners of this CTA
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

// The bin table lives as long as the context; the one-time build is synchronised so that later calls may use any stream.
int ensureAngleBins(svo_cuda_ctx* ctx) {
  if (ctx->angle_bins) return SVO_OK;
  int8_t* p = nullptr;
  SVO_CUDA_TRY(ctx, cudaMalloc(&p, (size_t)511 * 511));
  angle_bin_table_kernel<<<(511 * 511 + 255) / 256, 256, 0, ctx->stream>>>(p);
  ctx->launches++;
  const cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { cudaFree(p); return SVO_FAIL(ctx, SVO_ERR_CUDA, cudaGetErrorString(e)); }
  ctx->angle_bins = p;
  return SVO_OK;
}

int edgeletDeviceImpl(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, int threshold, int border, int cell_size,
                      const uint8_t* d_occ, svo_corner* d_out, unsigned long long* keys) {
  int n_cols, n_rows;
  const int n_cells = svo_cuda_grid_cells(pyr->cols[0], pyr->rows[0], cell_size, &n_cols, &n_rows);
  const size_t n = (size_t)n_cells * count;
  const int rcb = ensureAngleBins(ctx);
  if (rcb != SVO_OK) return rcb;
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
  edgelet_decode_kernel<<<(unsigned)((n + kHistWarps * 32 - 1) / (kHistWarps * 32)), kHistWarps * 32, 0, ctx->stream>>>(v, first, keys, n_cells, n,
                                                                                                        threshold, ctx->angle_bins, d_out);
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
  SVO_BIND(ctx);
  const size_t n = (size_t)svo_cuda_grid_cells(pyr->cols[0], pyr->rows[0], cell_size, nullptr, nullptr) * count;
  Stager st(ctx, mem);
  const uint8_t* d_occ = st.in(occupancy_in, n);
  svo_corner* d_out = st.out(corners_out, n);
  unsigned long long* keys = (unsigned long long*)st.scratch(n * sizeof(unsigned long long));
  if (!st.send() || !keys) return st.finish();
  const int rc2 = edgeletDeviceImpl(ctx, pyr, first, count, threshold, border, cell_size, d_occ, d_out, keys);
  if (rc2 != SVO_OK) return rc2;
  return st.finish();
}

int svo_cuda_angle_histogram_bins(svo_cuda_ctx* ctx, int8_t* bins_out, svo_mem mem) {
  if (!ctx || !bins_out) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_angle_histogram_bins: bad arguments");
  SVO_BIND(ctx);
  Stager st(ctx, mem);
  int8_t* d = st.out(bins_out, (size_t)511 * 511);
  if (!st.send()) return st.finish();
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
  SVO_BIND(ctx);
  const int n_cells = svo_cuda_grid_cells(pyr->cols[0], pyr->rows[0], opt->cell_size, nullptr, nullptr);
  const size_t n = (size_t)n_cells * count;
  Stager st(ctx, mem);
  const uint8_t* d_occ = st.in(occupancy_in, n);
  svo_corner* d_fast = st.out(corners_out, n);
  svo_corner* d_edge = st.out(edgelets_out, n);
  unsigned long long* keys = (unsigned long long*)st.scratch(n * sizeof(unsigned long long));
  uint8_t* d_occ2 = (uint8_t*)st.scratch(n);
  if (!st.send() || !keys || !d_occ2) return st.finish();
  int rc2 = svoFastDetectImpl(ctx, pyr, first, count, opt, d_occ, d_fast, SVO_MEM_DEVICE);
  if (rc2 != SVO_OK) return rc2;
  fastgrad_merge_kernel<<<count, 128, 0, ctx->stream>>>(d_fast, d_occ, n_cells, opt->threshold, max_n_features, d_occ2);
  SVO_LAUNCH_CHECK(ctx);
  rc2 = edgeletDeviceImpl(ctx, pyr, first, count, threshold_secondary, opt->border, opt->cell_size, d_occ2, d_edge, keys);
  if (rc2 != SVO_OK) return rc2;
  return st.finish();
}

}  // extern "C"
