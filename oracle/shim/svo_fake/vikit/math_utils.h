// TEST INFRASTRUCTURE ONLY (oracle/): the few vikit/math_utils.h helpers the direct front-end uses (the reference header
// also declares Lie-group utilities whose bodies need more of Eigen than the stand-in provides).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <vector>
#include <Eigen/Core>
#include <kindr/minimal/quat-transformation.h>
namespace vk {
Eigen::Matrix3d skew(const Eigen::Vector3d& v);
// vikit/math_utils.h:143-153
inline Eigen::Vector2d project2(const Eigen::Vector3d& v) { return v.head<2>() / v(2); }
inline Eigen::Vector3d unproject2d(const Eigen::Vector2d& v) { return Eigen::Vector3d(v[0], v[1], 1.0); }
// vikit/math_utils.h:130-141 (for the fixed-size vectors the checker passes)
template <class V> inline double norm_max(const V& v) {
  double max = -1;
  for (int i = 0; i < (int)V::RowsAtCompileTime * (int)V::ColsAtCompileTime; i++) {
    double abs = std::fabs(v[i]);
    if (abs > max) max = abs;
  }
  return max;
}
// vikit/math_utils.h:186-194
template <class T> inline T normPdf(const T x, const T mean, const T sigma) {
  T exponent = x - mean;
  exponent *= -exponent;
  exponent /= 2 * sigma * sigma;
  T result = std::exp(exponent);
  result /= sigma * std::sqrt(2 * M_PI);
  return result;
}
// vikit/math_utils.h:165-172
template <class T> T getMedian(std::vector<T>& data_vec) {
  assert(!data_vec.empty());
  typename std::vector<T>::iterator it = data_vec.begin() + std::floor(data_vec.size() / 2);
  std::nth_element(data_vec.begin(), it, data_vec.end());
  return *it;
}
}  // namespace vk
