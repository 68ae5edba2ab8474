// TEST INFRASTRUCTURE ONLY (oracle/): the two OpenCV imgproc functions the reference's edgelet detector really executes
// (svo::feature_detection_utils::edgeletDetector_V2, src/svo_direct/src/feature_detection_utils.cpp:331-333), restated for
// exactly the argument patterns used there. OpenCV is not linkable in this image (no headers / libraries); the Python module
// cv2 4.13 is, and tests/golden/cv_imgproc_golden.npz (generator: tests/golden/make_golden.py) pins these restatements
// bit-for-bit against it. Any other argument pattern aborts.
//   GaussianBlur(u8, 3x3, sigma 0): OpenCV's 8-bit path is fixed point with the kernel (1 2 1)/4 per axis and ONE rounding:
//     dst = (sum of the 3x3 window weighted 1 2 1 / 2 4 2 / 1 2 1  +  8) >> 4, borders BORDER_REFLECT_101.
//   Scharr(u8 -> 16S, scale 1, delta 0): exact integers, dx = 3 (p[-1][+1] - p[-1][-1]) + 10 (p[0][+1] - p[0][-1]) + 3 (p[+1][+1] - p[+1][-1]),
//     dy the transpose, borders BORDER_REFLECT_101.
#include <cstdio>
#include <cstdlib>
#include "shim_cv.hpp"
namespace cv {
namespace {
inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}
}  // namespace
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
  if (src.type() != CV_8UC1 || ksize.width != 3 || ksize.height != 3 || sigmaX != 0 || sigmaY != 0 || borderType != BORDER_DEFAULT) {
    std::fprintf(stderr, "oracle/shim: GaussianBlur argument pattern not restated\n");
    std::abort();
  }
  Mat out(src.rows, src.cols, CV_8UC1);
  for (int y = 0; y < src.rows; ++y) {
    const uchar* r0 = src.ptr(reflect101(y - 1, src.rows));
    const uchar* r1 = src.ptr(y);
    const uchar* r2 = src.ptr(reflect101(y + 1, src.rows));
    uchar* o = out.ptr(y);
    for (int x = 0; x < src.cols; ++x) {
      const int xm = reflect101(x - 1, src.cols), xp = reflect101(x + 1, src.cols);
      const int acc = r0[xm] + 2 * r0[x] + r0[xp] + 2 * r1[xm] + 4 * r1[x] + 2 * r1[xp] + r2[xm] + 2 * r2[x] + r2[xp];
      o[x] = (uchar)((acc + 8) >> 4);
    }
  }
  dst = out;
}
void Scharr(const Mat& src, Mat& dst, int ddepth, int dx, int dy, double scale, double delta, int borderType) {
  if (src.type() != CV_8UC1 || ddepth != CV_16S || dx + dy != 1 || dx < 0 || dy < 0 || scale != 1 || delta != 0 || borderType != BORDER_DEFAULT) {
    std::fprintf(stderr, "oracle/shim: Scharr argument pattern not restated\n");
    std::abort();
  }
  Mat out(src.rows, src.cols, CV_16SC1);
  for (int y = 0; y < src.rows; ++y) {
    const uchar* r0 = src.ptr(reflect101(y - 1, src.rows));
    const uchar* r1 = src.ptr(y);
    const uchar* r2 = src.ptr(reflect101(y + 1, src.rows));
    short* o = out.ptr<short>(y);
    for (int x = 0; x < src.cols; ++x) {
      const int xm = reflect101(x - 1, src.cols), xp = reflect101(x + 1, src.cols);
      o[x] = dx ? (short)(3 * (r0[xp] - r0[xm]) + 10 * (r1[xp] - r1[xm]) + 3 * (r2[xp] - r2[xm]))
                : (short)(3 * (r2[xm] - r0[xm]) + 10 * (r2[x] - r0[x]) + 3 * (r2[xp] - r0[xp]));
    }
  }
  dst = out;
}
}  // namespace cv
