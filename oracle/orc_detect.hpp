// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header for the rules).
//
// Rows a1-a6 of SURVEY.md §8: image pyramid, FAST-10 detect / score / 3x3 non-max,
// grid-cell arg-max (fastDetector) and fillFeatures.
//
// Parity status: a2-a4 are PINNED against the reference's own fast_neon sources compiled
// into oracle/_ref/libfast_ref.so (tests/test_oracle_fast.py); a1, a5, a6 are
// restatements ("parity unpinned": the reference ships no vectors and needs OpenCV).
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

}  // namespace orc
