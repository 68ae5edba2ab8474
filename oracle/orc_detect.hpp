// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header for the rules).
//
// Rows a1-a6 and f2 of SURVEY.md §8: image pyramid, FAST-10 detect / score / 3x3 non-max, grid-cell arg-max (fastDetector),
// fillFeatures, the edgelet detector (Gaussian 3x3, Scharr, the reference's neighbour test, angle histogram) and the three grid
// detector classes (FastDetector, GradientDetectorGrid, FastGradDetector).
//
// Parity status: PINNED. a2-a4 against the reference's own fast_neon sources (oracle/_ref/libfast_ref.so), a1 against its
// vision.cpp (libdirect_ref.so), a5 / a6 / f2 against its feature_detection.cpp + feature_detection_utils.cpp (libdetect_ref.so);
// the two OpenCV imgproc functions the edgelet detector executes are restated here and pinned bit-for-bit against the real
// OpenCV (cv2 4.13) through tests/golden/cv_imgproc_golden.npz. Tests: tests/test_reference_pins_cpu.py, tests/test_edgelet_cpu.py.
#pragma once
#include <vector>
#include <numeric>
#include <cmath>
#include <cstring>
#include "orc_math.hpp"

namespace orc {

// ---------------------------------------------------------------------------
// a1. vk::halfSample
// ref: src/vikit/vikit_common/src/vision.cpp:19-44 (halfSampleSSE2), :72-111 (halfSample)
//
// SSE2 branch arithmetic, restated per output pixel:
//   here = _mm_avg_epu8(row0, row1)            -> v(x)  = (a + c + 1) >> 1   (vertical, rounds up)
//   _mm_avg_epu16(even bytes, odd bytes)       -> out   = (v(2x) + v(2x+1) + 1) >> 1
// The SSE2 loop handles sw = w>>4 blocks of 16 input px and sh = h>>1 row pairs and
// treats the input as contiguous with stride == w (vision.cpp:22,25-26,40-41).
inline void halfSampleSSE2Formula(const uint8_t* in, uint8_t* out, int w, int h) {
  const int sw = w >> 4;
  const int sh = h >> 1;
  const uint8_t* row0 = in;
  const uint8_t* row1 = in + w;
  for (int i = 0; i < sh; ++i) {
    for (int j = 0; j < sw; ++j) {
      for (int k = 0; k < 8; ++k) {
        const int a = row0[16 * j + 2 * k], b = row0[16 * j + 2 * k + 1];
        const int c = row1[16 * j + 2 * k], d = row1[16 * j + 2 * k + 1];
        const int v0 = (a + c + 1) >> 1;
        const int v1 = (b + d + 1) >> 1;
        out[8 * j + k] = static_cast<uint8_t>((v0 + v1 + 1) >> 1);
      }
    }
    out += 8 * sw;
    row0 += 2 * w;
    row1 += 2 * w;
  }
}

// vision.cpp:98-110 (scalar fallback): truncating mean of the 2x2 block.
inline void halfSampleScalar(const uint8_t* in, int in_stride, int in_rows,
                             uint8_t* out, int out_stride, int out_cols, int out_rows) {
  const uint8_t* top = in;
  const uint8_t* bottom = top + in_stride;
  const uint8_t* end = top + static_cast<ptrdiff_t>(in_stride) * in_rows;
  uint8_t* p = out;
  for (int y = 0; y < out_rows && bottom < end; y++, top += in_stride * 2, bottom += in_stride * 2, p += out_stride) {
    for (int x = 0; x < out_cols; x++) {
      p[x] = static_cast<uint8_t>((uint16_t(top[x * 2]) + top[x * 2 + 1] + bottom[x * 2] + bottom[x * 2 + 1]) / 4);
    }
  }
}

// vision.cpp:72-111. `mode`: -1 = the reference's x86 predicate (cols%16==0 && contiguous; the
// 16-byte alignment of cv::Mat buffers is taken as given, cf. cv::Mat::create 64-B alignment),
// 0 = force scalar formula, 1 = force SSE2 formula (requires in_stride == in_cols, cols%16==0).
inline void halfSample(const uint8_t* in, int in_cols, int in_rows, int in_stride,
                       uint8_t* out, int out_stride, int mode = -1) {
  const int out_cols = in_cols / 2, out_rows = in_rows / 2;
  bool sse = (in_cols % 16 == 0) && (in_cols == in_stride) && (out_stride == out_cols);
  if (mode == 0) sse = false;
  if (sse) { halfSampleSSE2Formula(in, out, in_cols, in_rows); return; }
  halfSampleScalar(in, in_stride, in_rows, out, out_stride, out_cols, out_rows);
}

// a1. frame_utils::createImgPyramid
// ref: src/svo_common/src/frame.cpp:372-386
// Level buffers are tightly packed (step == cols), as cv::Mat(rows, cols, CV_8U) allocates them.
struct Pyramid {
  std::vector<std::vector<uint8_t>> store;  // levels 1.. (level 0 aliases the caller's image)
  std::vector<Img> lv;
};
inline void createImgPyramid(const uint8_t* img0, int cols, int rows, int step, int n_levels, Pyramid& pyr, int mode = -1) {
  pyr.lv.resize(n_levels);
  pyr.store.resize(n_levels);
  pyr.lv[0] = Img{img0, cols, rows, step};
  for (int i = 1; i < n_levels; ++i) {
    const Img& p = pyr.lv[i - 1];
    const int c = p.cols / 2, r = p.rows / 2;
    pyr.store[i].assign(static_cast<size_t>(c) * r, 0);
    halfSample(p.data, p.cols, p.rows, p.step, pyr.store[i].data(), c, mode);
    pyr.lv[i] = Img{pyr.store[i].data(), c, r, c};
  }
}

// ---------------------------------------------------------------------------
// a2-a4. FAST-10 (libCVD-derived) — closed-form restatement of the generated trees.
// Circle offsets: src/fast_neon/src/fast_10_score.cpp:3158-3175 (== fast_10.cpp:16-33).
static const int kFastDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kFastDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

struct FastXY { short x, y; };  // ref: src/fast_neon/include/fast/fast.h:11-15

// Largest b such that >= ARC contiguous circle pixels are all > p+b, or all < p-b; -1 if none
// even for b = 0... (returns S-1 with S = max over arcs of the min margin; may be negative).
template <int ARC>
inline int fastMargin(const uint8_t* p, int stride) {
  int d[16 + ARC];
  const int c = *p;
  for (int i = 0; i < 16; ++i) d[i] = int(p[kFastDy[i] * stride + kFastDx[i]]) - c;
  for (int i = 0; i < ARC; ++i) d[16 + i] = d[i];
  int best = -256;
  for (int k = 0; k < 16; ++k) {
    int mn = 255, mx = -255;
    for (int j = 0; j < ARC; ++j) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
    best = std::max(best, std::max(mn, -mx));  // bright arc margin, dark arc margin
  }
  return best - 1;
}

// a2. fast_corner_detect_10_sse2 / fast_corner_detect_10 — segment test, raster order.
// ref: src/fast_neon/src/faster_corner_10_sse.cpp:15-202; src/fast_neon/src/fast_10.cpp:9-
// A pixel is a corner at barrier b iff fastMargin >= b (all ARC px strictly > p+b or < p-b).
// Region: y in [3,h-3), x in [3,w-3) (faster_corner_10_sse.cpp:27-32,180-185; fast_10.cpp:35-48).
template <int ARC>
inline void fastCornerDetect(const uint8_t* img, int w, int h, int stride, int barrier, std::vector<FastXY>& corners) {
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x)
      if (fastMargin<ARC>(img + y * stride + x, stride) >= barrier)
        corners.push_back(FastXY{short(x), short(y)});
}

// a3. fast_corner_score_10 — ref: src/fast_neon/src/fast_10_score.cpp:21-3148 (tree), :3150-3178 (wrapper)
// The tree starts at b = threshold+1 and raises b while the pixel remains a corner; it returns b-1.
inline void fastCornerScore10(const uint8_t* img, int stride, const std::vector<FastXY>& corners, int threshold,
                              std::vector<int>& scores) {
  scores.resize(corners.size());
  for (size_t n = 0; n < corners.size(); ++n)
    scores[n] = std::max(threshold, fastMargin<10>(img + corners[n].y * stride + corners[n].x, stride));
}

// a4. fast_nonmax_3x3 — ref: src/fast_neon/src/nonmax_3x3.cpp:17-112
// Keeps corner i iff no 8-neighbour that is also in the list has score >= score_i.
// Restated with a dense lookup instead of the row_start / point_above / point_below cursors.
inline void fastNonmax3x3(const std::vector<FastXY>& corners, const std::vector<int>& scores, std::vector<int>& nonmax) {
  nonmax.clear();
  if (corners.empty()) return;
  int maxx = 0, maxy = 0;
  for (const auto& c : corners) { maxx = std::max<int>(maxx, c.x); maxy = std::max<int>(maxy, c.y); }
  const int W = maxx + 3, H = maxy + 3;
  std::vector<int> grid(static_cast<size_t>(W) * H, -1);  // score at (x+1, y+1), -1 = no corner
  for (size_t i = 0; i < corners.size(); ++i) grid[(corners[i].y + 1) * W + corners[i].x + 1] = scores[i];
  for (size_t i = 0; i < corners.size(); ++i) {
    const int x = corners[i].x + 1, y = corners[i].y + 1, s = scores[i];
    bool keep = true;
    for (int dy = -1; dy <= 1 && keep; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        if (!dx && !dy) continue;
        const int g = grid[(y + dy) * W + x + dx];
        if (g >= 0 && g >= s) { keep = false; break; }
      }
    if (keep) nonmax.push_back(int(i));
  }
}

// ---------------------------------------------------------------------------
// a5. feature_detection_utils::fastDetector + OccupandyGrid2D::getCellIndex
// ref: src/svo_direct/include/svo/direct/feature_detection_types.h:17-29 (Corner)
struct Corner {
  int x, y, level;
  float score, angle;
};

// ref: src/svo_common/include/svo/common/occupancy_grid_2d.h:82-95
inline size_t gridCellIndex(int x, int y, int scale, int cell_size, int n_cols) {
  const double px0 = double(scale * x), px1 = double(scale * y);
  return static_cast<size_t>(std::floor(px1 / cell_size) * n_cols + std::floor(px0 / cell_size));
}

// ref: src/svo_direct/src/feature_detection_utils.cpp:145-194
// `corners` is in/out (pre-filled by the caller with score = threshold, feature_detection.cpp:63-65);
// `occupancy` is an input (cells already holding a feature are skipped).
inline void fastDetector(const std::vector<Img>& img_pyr, int threshold, int border, size_t min_level, size_t max_level,
                         std::vector<Corner>& corners, const std::vector<uint8_t>& occupancy, int cell_size, int n_cols) {
  for (size_t level = min_level; level <= max_level; ++level) {
    const int scale = (1 << level);
    const Img& im = img_pyr[level];
    std::vector<FastXY> fast_corners;
    fastCornerDetect<10>(im.data, im.cols, im.rows, im.step, threshold, fast_corners);
    std::vector<int> scores, nm_corners;
    fastCornerScore10(im.data, im.step, fast_corners, threshold, scores);
    fastNonmax3x3(fast_corners, scores, nm_corners);

    const int maxw = im.cols - border;
    const int maxh = im.rows - border;
    for (const int& i : nm_corners) {
      const FastXY& xy = fast_corners.at(i);
      if (xy.x < border || xy.y < border || xy.x >= maxw || xy.y >= maxh) continue;
      const size_t k = gridCellIndex(xy.x, xy.y, scale, cell_size, n_cols);
      if (occupancy.at(k)) continue;
      const float score = scores.at(i);
      if (score > corners.at(k).score) corners.at(k) = Corner{xy.x * scale, xy.y * scale, int(level), score, 0.0f};
    }
  }
}

// a6. fillFeatures — ref: src/svo_direct/src/feature_detection_utils.cpp:72-142
// Outputs appended SoA entries (px, score, level, gradient) and marks occupancy.
// std::sort in the reference is unstable: callers must compare tied scores as sets.
struct FeatureSoA {
  std::vector<double> px;     // 2 x N (x0,y0,x1,y1,...)
  std::vector<double> grad;   // 2 x N
  std::vector<double> score;  // N
  std::vector<int> level;     // N
};
inline void fillFeatures(const std::vector<Corner>& corners, const uint8_t* mask, int mask_step, double threshold,
                         size_t max_n_features, FeatureSoA& out, std::vector<uint8_t>& occupancy, int cell_size, int n_cols) {
  std::vector<double> kx, ky, gx, gy, sc;
  std::vector<int> lv;
  for (const Corner& c : corners) {
    if (c.score > threshold) {
      if (mask && mask[c.y * mask_step + c.x] == 0) continue;
      kx.push_back(c.x); ky.push_back(c.y);
      lv.push_back(c.level);
      sc.push_back(c.score);
      gx.push_back(std::cos(c.angle)); gy.push_back(std::sin(c.angle));
      occupancy[gridCellIndex(c.x, c.y, 1, cell_size, n_cols)] = 1;
    }
  }
  std::vector<size_t> idx(sc.size());
  std::iota(idx.begin(), idx.end(), 0u);
  std::sort(idx.begin(), idx.end(), [&sc](size_t i1, size_t i2) { return sc[i1] > sc[i2]; });
  const size_t n_new = std::min(max_n_features, kx.size());
  for (size_t i = 0; i < n_new; ++i) {
    const size_t j = idx[i];
    out.px.push_back(kx[j]); out.px.push_back(ky[j]);
    out.grad.push_back(gx[j]); out.grad.push_back(gy[j]);
    out.score.push_back(sc[j]);
    out.level.push_back(lv[j]);
  }
}

// ---------------------------------------------------------------------------
// f2. Edgelet detector: edgeletDetector_V2 + the gradient-orientation histogram.
// OpenCV pieces, restated from OpenCV's documented 8-bit behaviour and PINNED against cv2 4.13 (tests/golden/cv_imgproc_golden.npz):
//   cv::GaussianBlur(u8, Size(3,3), 0)  = one rounding of the 1-2-1 x 1-2-1 window: (sum + 8) >> 4, BORDER_REFLECT_101
//   cv::Scharr(u8 -> CV_16S)            = exact 3-10-3 differences, BORDER_REFLECT_101
inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
  return i;
}
inline void gaussianBlur3x3(const Img& im, std::vector<uint8_t>& out) {
  out.resize(size_t(im.cols) * im.rows);
  for (int y = 0; y < im.rows; ++y) {
    const uint8_t* ra = im.data + reflect101(y - 1, im.rows) * im.step;
    const uint8_t* rb = im.data + y * im.step;
    const uint8_t* rc = im.data + reflect101(y + 1, im.rows) * im.step;
    for (int x = 0; x < im.cols; ++x) {
      const int l = reflect101(x - 1, im.cols), r = reflect101(x + 1, im.cols);
      const int s = (ra[l] + rc[l] + ra[r] + rc[r]) + 2 * (ra[x] + rc[x] + rb[l] + rb[r]) + 4 * rb[x];
      out[size_t(y) * im.cols + x] = uint8_t((s + 8) >> 4);
    }
  }
}
inline void scharr3x3(const uint8_t* g, int cols, int rows, std::vector<int16_t>& dx, std::vector<int16_t>& dy) {
  dx.resize(size_t(cols) * rows); dy.resize(size_t(cols) * rows);
  for (int y = 0; y < rows; ++y) {
    const uint8_t* ra = g + size_t(reflect101(y - 1, rows)) * cols;
    const uint8_t* rb = g + size_t(y) * cols;
    const uint8_t* rc = g + size_t(reflect101(y + 1, rows)) * cols;
    for (int x = 0; x < cols; ++x) {
      const int l = reflect101(x - 1, cols), r = reflect101(x + 1, cols);
      dx[size_t(y) * cols + x] = int16_t(3 * ((ra[r] - ra[l]) + (rc[r] - rc[l])) + 10 * (rb[r] - rb[l]));
      dy[size_t(y) * cols + x] = int16_t(3 * ((rc[l] - ra[l]) + (rc[r] - ra[r])) + 10 * (rc[x] - ra[x]));
    }
  }
}

// angle_hist::{angleHistogram, gradientAndMagnitudeAtPixel, smoothOrientationHistogram, getDominantAngle} and
// getAngleAtPixelUsingHistogram — ref: src/svo_direct/src/feature_detection_utils.cpp:831-839, 945-1009;
// n_bins = 36 (include/svo/direct/feature_detection_utils.h:168).
inline int angleHistogramBin(int gx, int gy) {  // bin of a central-difference gradient (gx, gy)
  const double angle = std::atan2(double(gy), double(gx));
  size_t bin = size_t(std::round(36 * (angle + M_PI) / (2.0 * M_PI)));
  return int(bin < 36 ? bin : 0u);
}
inline double angleAtPixelUsingHistogram(const Img& im, int px, int py, int halfpatch) {
  double hist[36];
  for (double& h : hist) h = 0.0;
  for (int v = py - halfpatch; v <= py + halfpatch; ++v)
    for (int u = px - halfpatch; u <= px + halfpatch; ++u) {
      if (!(v > 0 && v < im.rows - 1 && u > 0 && u < im.cols - 1)) continue;
      const int gx = int(im.data[v * im.step + u + 1]) - int(im.data[v * im.step + u - 1]);
      const int gy = int(im.data[(v + 1) * im.step + u]) - int(im.data[(v - 1) * im.step + u]);
      hist[angleHistogramBin(gx, gy)] += std::sqrt(double(gx) * gx + double(gy) * gy);
    }
  // circular 1-2-1 smoothing, in place, each bin reading its un-smoothed neighbours
  double prev = hist[35];
  const double first = hist[0];
  for (int i = 0; i < 36; ++i) {
    const double here = hist[i];
    hist[i] = 0.25 * prev + 0.5 * here + 0.25 * (i == 35 ? first : hist[i + 1]);
    prev = here;
  }
  int best = 0;
  for (int i = 1; i < 36; ++i)
    if (hist[i] > hist[best]) best = i;
  return best * 2.0 * M_PI / 36;
}

// edgeletDetector_V2 — ref: src/svo_direct/src/feature_detection_utils.cpp:313-385. Works on pyramid level 1 only
// (:324-325) and reports level 0 (`level-1`, :379) with px = 2 * level-1 pixel.
// Quirk kept: the 8-neighbour test uses `stride = score.step` — a BYTE step — as a float-pointer offset (:352, :363-364), so
// the "vertical" neighbours it compares are 4 rows away: (x, y+-4) and (x+-1, y+-4). The loops need border >= 4 to stay inside
// the score map.
inline void edgeletDetectorV2(const std::vector<Img>& img_pyr, int threshold, int border, std::vector<Corner>& corners,
                              const std::vector<uint8_t>& occupancy, int cell_size, int n_cols) {
  const Img& im = img_pyr[1];
  const int W = im.cols, H = im.rows;
  std::vector<uint8_t> blur;
  std::vector<int16_t> gx, gy;
  gaussianBlur3x3(im, blur);
  scharr3x3(blur.data(), W, H, gx, gy);
  std::vector<float> score(size_t(W) * H, 0.0f);
  for (int y = border; y < H - border; ++y)
    for (int x = border; x < W - border; ++x) {
      const int a = gx[size_t(y) * W + x], b = gy[size_t(y) * W + x];
      const float mag = float(std::sqrt(double(a * a + b * b)));  // std::sqrt(int) is the double overload (:343)
      score[size_t(y) * W + x] = (mag > threshold) ? mag : 0.0f;
    }
  const int vstep = 4 * W;  // see the quirk above
  for (int y = border; y < H - border; ++y)
    for (int x = border; x < W - border; ++x) {
      const size_t k = gridCellIndex(x, y, 2, cell_size, n_cols);
      if (occupancy.at(k)) continue;
      const float* c = &score[size_t(y) * W + x];
      const float s = *c;
      if (s < threshold) continue;
      if (c[1] >= s || c[-1] > s) continue;
      if (c[vstep] >= s || c[-vstep] > s) continue;
      if (c[vstep + 1] >= s || c[vstep - 1] > s) continue;
      if (c[-vstep + 1] >= s || c[-vstep - 1] > s) continue;
      if (s > corners.at(k).score)
        corners.at(k) = Corner{2 * x, 2 * y, 0, s, float(angleAtPixelUsingHistogram(im, x, y, 4))};
    }
}

// AbstractDetector::detect(img_pyr, mask = empty, ...) for the three grid detectors of this path —
// ref: src/svo_direct/src/feature_detection.cpp:53-74 (FastDetector), :130-151 (GradientDetectorGrid), :154-194 (FastGradDetector).
// detector_type follows svo::DetectorType (feature_detection_types.h:33-45): 0 kFast, 2 kFastGrad, 5 kGridGrad.
// fillFeatures turns Corner::angle (float) into the gradient with the float overloads of cos / sin (:101).
struct DetectedFeatures {
  std::vector<double> px, grad, score;
  std::vector<int> level, type;
};
inline void appendFeatures(const std::vector<Corner>& corners, int type, double threshold, size_t max_n, DetectedFeatures& out,
                           std::vector<uint8_t>& occupancy, int cell_size, int n_cols) {
  std::vector<size_t> keep;
  for (size_t k = 0; k < corners.size(); ++k)
    if (corners[k].score > threshold) {
      keep.push_back(k);
      occupancy[gridCellIndex(corners[k].x, corners[k].y, 1, cell_size, n_cols)] = 1;
    }
  std::sort(keep.begin(), keep.end(), [&](size_t a, size_t b) { return double(corners[a].score) > double(corners[b].score); });
  const size_t n_new = std::min(max_n, keep.size());
  for (size_t i = 0; i < n_new; ++i) {
    const Corner& c = corners[keep[i]];
    out.px.push_back(c.x); out.px.push_back(c.y);
    out.grad.push_back(std::cos(c.angle)); out.grad.push_back(std::sin(c.angle));  // float overloads, as the reference
    out.score.push_back(c.score);
    out.level.push_back(c.level);
    out.type.push_back(type);
  }
}
inline void detectFeatures(int detector_type, const std::vector<Img>& img_pyr, double threshold_primary, double threshold_secondary,
                           int border, int min_level, int max_level, int cell_size, std::vector<uint8_t> occupancy, size_t max_n,
                           DetectedFeatures& out) {
  const int n_cols = int(std::ceil(double(img_pyr[0].cols) / cell_size));
  const int n_rows = int(std::ceil(double(img_pyr[0].rows) / cell_size));
  const size_t n_cells = size_t(n_cols) * n_rows;
  occupancy.resize(n_cells, 0);
  const int kCorner = 7, kEdgelet = 6;  // svo::FeatureType (src/svo_common/include/svo/common/types.h:60-73)
  if (detector_type == 0 || detector_type == 2) {
    std::vector<Corner> corners(n_cells, Corner{0, 0, 0, float(threshold_primary), 0.0f});
    fastDetector(img_pyr, int(threshold_primary), border, size_t(min_level), size_t(max_level), corners, occupancy, cell_size, n_cols);
    appendFeatures(corners, kCorner, threshold_primary, max_n, out, occupancy, cell_size, n_cols);
  }
  if (detector_type == 5 || detector_type == 2) {
    const long room = detector_type == 2 ? long(max_n) - long(out.score.size()) : long(max_n);
    if (room > 0) {
      std::vector<Corner> corners(n_cells, Corner{0, 0, 0, float(threshold_secondary), 0.0f});
      edgeletDetectorV2(img_pyr, int(threshold_secondary), border, corners, occupancy, cell_size, n_cols);
      appendFeatures(corners, kEdgelet, threshold_secondary, size_t(room), out, occupancy, cell_size, n_cols);
    }
  }
}

}  // namespace orc
