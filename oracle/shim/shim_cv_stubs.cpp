// TEST INFRASTRUCTURE ONLY (oracle/): definitions for the OpenCV functions that shim_cv.hpp only declares. The reference
// calls them from display / debug code that never runs in the checker; reaching one aborts instead of computing something
// OpenCV would not.
#include <cstdio>
#include <cstdlib>
#include "shim_cv.hpp"
#define SVO_SHIM_UNREACHABLE() do { std::fprintf(stderr, "oracle/shim: OpenCV function %s is not available\n", __func__); std::abort(); } while (0)
namespace cv {
Mat::Mat(const Mat&, const Rect&) : Mat() { SVO_SHIM_UNREACHABLE(); }
void Mat::convertTo(Mat&, int, double, double) const { SVO_SHIM_UNREACHABLE(); }
Mat& Mat::operator=(const Scalar&) { SVO_SHIM_UNREACHABLE(); }
Mat operator-(const Mat&, double) { SVO_SHIM_UNREACHABLE(); }
Mat operator-(const Mat&, const Mat&) { SVO_SHIM_UNREACHABLE(); }
Mat operator+(const Mat&, const Mat&) { SVO_SHIM_UNREACHABLE(); }
Mat operator/(const Mat&, double) { SVO_SHIM_UNREACHABLE(); }
Mat operator*(const Mat&, double) { SVO_SHIM_UNREACHABLE(); }
void minMaxLoc(const Mat&, double*, double*, Point*, Point*) { SVO_SHIM_UNREACHABLE(); }
void resize(const Mat&, Mat&, Size, double, double, int) { SVO_SHIM_UNREACHABLE(); }
void cvtColor(const Mat&, Mat&, int, int) { SVO_SHIM_UNREACHABLE(); }
void normalize(const Mat&, Mat&, double, double, int, int) { SVO_SHIM_UNREACHABLE(); }
void hconcat(const Mat&, const Mat&, Mat&) { SVO_SHIM_UNREACHABLE(); }
void line(Mat&, Point2f, Point2f, const Scalar&, int) { SVO_SHIM_UNREACHABLE(); }
void imshow(const std::string&, const Mat&) { SVO_SHIM_UNREACHABLE(); }
int waitKey(int) { SVO_SHIM_UNREACHABLE(); }
void namedWindow(const std::string&, int) { SVO_SHIM_UNREACHABLE(); }
void split(const Mat&, std::vector<Mat>&) { SVO_SHIM_UNREACHABLE(); }
void merge(const std::vector<Mat>&, Mat&) { SVO_SHIM_UNREACHABLE(); }
Mat Mat::zeros(Size, int) { SVO_SHIM_UNREACHABLE(); }
Mat Mat::mul(const Mat&, double) const { SVO_SHIM_UNREACHABLE(); }
Mat Mat::t() const { SVO_SHIM_UNREACHABLE(); }
Mat operator+(double, const Mat&) { SVO_SHIM_UNREACHABLE(); }
Mat operator*(double, const Mat&) { SVO_SHIM_UNREACHABLE(); }
Mat operator/(const Mat&, const Mat&) { SVO_SHIM_UNREACHABLE(); }
Mat operator==(const Mat&, double) { SVO_SHIM_UNREACHABLE(); }
void Sobel(const Mat&, Mat&, int, int, int, int, double, double, int) { SVO_SHIM_UNREACHABLE(); }
void filter2D(const Mat&, Mat&, int, const Mat&, Point, double, int) { SVO_SHIM_UNREACHABLE(); }
void blur(const Mat&, Mat&, Size, Point, int) { SVO_SHIM_UNREACHABLE(); }
void Canny(const Mat&, Mat&, double, double, int, bool) { SVO_SHIM_UNREACHABLE(); }
void convertScaleAbs(const Mat&, Mat&, double, double) { SVO_SHIM_UNREACHABLE(); }
void addWeighted(const Mat&, double, const Mat&, double, double, Mat&, int) { SVO_SHIM_UNREACHABLE(); }
double threshold(const Mat&, Mat&, double, double, int) { SVO_SHIM_UNREACHABLE(); }
int countNonZero(const Mat&) { SVO_SHIM_UNREACHABLE(); }
void rectangle(Mat&, Point2f, Point2f, const Scalar&, int, int, int) { SVO_SHIM_UNREACHABLE(); }
void circle(Mat&, Point2f, int, const Scalar&, int, int, int) { SVO_SHIM_UNREACHABLE(); }
}  // namespace cv
