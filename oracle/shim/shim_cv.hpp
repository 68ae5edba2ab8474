// TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for the handful of OpenCV *container* types the reference's
// direct front-end sources touch (cv::Mat as an image container, Point/Size/Rect/Scalar PODs), so that those sources can
// be compiled from where they lie under /root/reference without OpenCV (not installed here). No OpenCV arithmetic is
// restated: imgproc/highgui functions are declared only (the reference calls them from debug blocks and unused inline
// helpers), so anything that would really need OpenCV fails at link time instead of silently doing something else.
#pragma once
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_CN_SHIFT 3
#define CV_DEPTH_MAX (1 << CV_CN_SHIFT)
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH_MASK (CV_DEPTH_MAX - 1)
#define CV_MAT_DEPTH(flags) ((flags) & CV_MAT_DEPTH_MASK)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAKE_TYPE CV_MAKETYPE
#define CV_MAT_CN(flags) ((((flags) >> CV_CN_SHIFT) & 511) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_16SC2 CV_MAKETYPE(CV_16S, 2)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_Assert(expr) do { if (!(expr)) { std::abort(); } } while (0)
#define CV_DbgAssert(expr) assert(expr)

namespace cv {

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  T area() const { return width * height; }
};
typedef Size_<int> Size;
template <typename T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
};

template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { depth = CV_8U, channels = 1, type = CV_8UC1 }; };
template <> struct DataType<short> { enum { depth = CV_16S, channels = 1, type = CV_16SC1 }; };
template <> struct DataType<ushort> { enum { depth = CV_16U, channels = 1, type = CV_16UC1 }; };
template <> struct DataType<float> { enum { depth = CV_32F, channels = 1, type = CV_32FC1 }; };
template <> struct DataType<double> { enum { depth = CV_64F, channels = 1, type = CV_64FC1 }; };

inline size_t alignSize(size_t sz, int n) { return (sz + n - 1) & -n; }
template <typename T> inline T* alignPtr(T* ptr, int n = (int)sizeof(T)) {
  return (T*)(((size_t)ptr + n - 1) & -n);
}
template <typename T, size_t fixed_size = 1024 / sizeof(T) + 8> class AutoBuffer {
 public:
  explicit AutoBuffer(size_t n) : buf_(n) {}
  operator T*() { return buf_.data(); }
  operator const T*() const { return buf_.data(); }
  T* data() { return buf_.data(); }
 private:
  std::vector<T> buf_;
};

// Dense 2-D image container with OpenCV's layout conventions: row-major, `step` bytes per row, 64-byte aligned
// allocations that are continuous (step == cols * elemSize) when created by the container itself.
class Mat {
 public:
  struct MStep {
    size_t p[2];
    MStep() : p{0, 0} {}
    operator size_t() const { return p[0]; }
    size_t operator[](int i) const { return p[i]; }
  };
  int flags, dims, rows, cols;
  uchar* data;
  MStep step;

  Mat() : flags(0), dims(0), rows(0), cols(0), data(nullptr) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
  Mat(int r, int c, int type, void* d, size_t stp = 0) : flags(type), dims(2), rows(r), cols(c), data((uchar*)d) {
    step.p[1] = elemSize();
    step.p[0] = stp ? stp : (size_t)c * elemSize();
  }
  Mat(const Mat& m, const Rect& roi);  // declared only
  // Constant-filled image (edgeletDetector_V2's score / angle maps): 8-bit and float single-channel only.
  Mat(Size s, int type, const Scalar& fill) : Mat() {
    create(s.height, s.width, type);
    if (type == CV_8UC1) { for (int r = 0; r < rows; ++r) std::memset(ptr(r), (int)fill.val[0], (size_t)cols); }
    else if (type == CV_32FC1) { for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) at<float>(r, c) = (float)fill.val[0]; }
    else std::abort();
  }
  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == this->type() && owner_) return;
    flags = type; dims = 2; rows = r; cols = c;
    step.p[1] = elemSize();
    step.p[0] = (size_t)c * elemSize();
    const size_t bytes = ((step.p[0] * (size_t)r + 63) / 64 + 1) * 64;
    void* mem = std::aligned_alloc(64, bytes);
    std::memset(mem, 0, bytes);
    owner_.reset((uchar*)mem, [](uchar* p) { std::free(p); });
    data = owner_.get();
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  int type() const { return flags & ((1 << 12) - 1); }
  int depth() const { return CV_MAT_DEPTH(flags); }
  int channels() const { return CV_MAT_CN(flags); }
  size_t elemSize1() const {
    static const int sz[] = {1, 1, 2, 2, 4, 4, 8, 2};
    return sz[depth()];
  }
  size_t elemSize() const { return elemSize1() * channels(); }
  bool isContinuous() const { return step.p[0] == (size_t)cols * elemSize(); }
  bool empty() const { return data == nullptr || rows * cols == 0; }
  Size size() const { return Size(cols, rows); }
  size_t total() const { return (size_t)rows * cols; }
  Mat clone() const {
    Mat m(rows, cols, type());
    for (int r = 0; r < rows; ++r) std::memcpy(m.data + r * m.step.p[0], data + r * step.p[0], (size_t)cols * elemSize());
    return m;
  }
  void copyTo(Mat& m) const { m = clone(); }
  template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.p[0]); }
  template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.p[0]); }
  uchar* ptr(int r = 0) { return data + (size_t)r * step.p[0]; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step.p[0]; }
  template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step.p[0]))[c]; }
  template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step.p[0]))[c]; }
  void convertTo(Mat& m, int rtype, double alpha = 1, double beta = 0) const;  // declared only
  Mat& operator=(const Scalar& s);                                              // declared only
  static Mat zeros(Size s, int type);                                           // declared only
  Mat mul(const Mat& m, double scale = 1) const;                                // declared only
  Mat t() const;                                                                // declared only

 private:
  std::shared_ptr<uchar> owner_;
};

// Declared only: used by the reference from debug blocks / unused inline helpers (never linked into the checker).
Mat operator-(const Mat& a, double b);
Mat operator-(const Mat& a, const Mat& b);
Mat operator+(const Mat& a, const Mat& b);
Mat operator/(const Mat& a, double b);
Mat operator*(const Mat& a, double b);
void minMaxLoc(const Mat& src, double* minVal, double* maxVal = nullptr, Point* minLoc = nullptr, Point* maxLoc = nullptr);
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, COLOR_GRAY2RGB = 8, COLOR_GRAY2BGR = 8, NORM_MINMAX = 32 };
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void cvtColor(const Mat& src, Mat& dst, int code, int dstCn = 0);
void normalize(const Mat& src, Mat& dst, double alpha = 1, double beta = 0, int norm_type = 4, int dtype = -1);
void hconcat(const Mat& a, const Mat& b, Mat& dst);
void line(Mat& img, Point2f a, Point2f b, const Scalar& color, int thickness = 1);
void imshow(const std::string& name, const Mat& m);
int waitKey(int delay = 0);
void namedWindow(const std::string& name, int flags = 1);
void split(const Mat& src, std::vector<Mat>& mv);
void merge(const std::vector<Mat>& mv, Mat& dst);
Mat operator+(double a, const Mat& b);
Mat operator*(double a, const Mat& b);
Mat operator/(const Mat& a, const Mat& b);
Mat operator==(const Mat& a, double b);
template <typename T> struct Mat_ : Mat {
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
};
template <typename T> struct MatCommaInitializer_ {
  MatCommaInitializer_& operator,(double) { std::abort(); }
  operator Mat() const { std::abort(); }
};
template <typename T> inline MatCommaInitializer_<T> operator<<(const Mat_<T>&, double) { std::abort(); }
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { THRESH_BINARY = 0, THRESH_BINARY_INV = 1 };
// The two imgproc functions the reference's *edgelet detector* really runs (feature_detection_utils.cpp:331-333) are restated
// in shim_cv_imgproc.cpp for exactly the argument patterns used there (8-bit 3x3 sigma-0 blur; 8U -> 16S 3x3 Scharr, scale 1,
// delta 0, BORDER_REFLECT_101) and pinned against the real OpenCV (cv2 4.13) by tests/golden/cv_imgproc_golden.npz; any other
// argument pattern aborts.
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void Scharr(const Mat& src, Mat& dst, int ddepth, int dx, int dy, double scale = 1, double delta = 0, int borderType = BORDER_DEFAULT);
// Declared only (other detectors / drawing code; reaching one aborts).
void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0, int borderType = BORDER_DEFAULT);
void filter2D(const Mat& src, Mat& dst, int ddepth, const Mat& kernel, Point anchor = Point(-1, -1), double delta = 0, int borderType = BORDER_DEFAULT);
void blur(const Mat& src, Mat& dst, Size ksize, Point anchor = Point(-1, -1), int borderType = BORDER_DEFAULT);
void Canny(const Mat& image, Mat& edges, double threshold1, double threshold2, int apertureSize = 3, bool L2gradient = false);
void convertScaleAbs(const Mat& src, Mat& dst, double alpha = 1, double beta = 0);
void addWeighted(const Mat& src1, double alpha, const Mat& src2, double beta, double gamma, Mat& dst, int dtype = -1);
double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type);
int countNonZero(const Mat& src);
void rectangle(Mat& img, Point2f pt1, Point2f pt2, const Scalar& color, int thickness = 1, int lineType = 8, int shift = 0);
void circle(Mat& img, Point2f center, int radius, const Scalar& color, int thickness = 1, int lineType = 8, int shift = 0);

}  // namespace cv
