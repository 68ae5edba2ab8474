// TEST INFRASTRUCTURE ONLY (oracle/): the part of svo::Point the direct front-end reads (src/svo_common/include/svo/common/point.h:
// `pos_`, the 3-D position in world coordinates); the reference's Point also carries the map bookkeeping.
#pragma once
#include <memory>
#include <svo/common/types.h>
namespace svo {
class Point {
 public:
  Position pos_;
  bool in_ba_graph_ = false;
  explicit Point(const Position& pos) : pos_(pos) {}
  const Position& pos() const { return pos_; }
};
using PointPtr = std::shared_ptr<Point>;
}  // namespace svo
