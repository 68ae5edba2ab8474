// TEST INFRASTRUCTURE ONLY (oracle/): the part of svo::Point the direct front-end and the Reprojector read
// (src/svo_common/include/svo/common/point.h:36-60 KeypointIdentifier, :82-91 pos_ / obs_ / last_projected_kf_id_ /
// n_failed_reproj_ / n_succeeded_reproj_); the reference's Point also carries the map bookkeeping and the optimiser.
// getCloseViewObs is restated from src/svo_common/src/point.cpp:83-129 (that file pulls in the bundle-adjustment types).
#pragma once
#include <array>
#include <atomic>
#include <memory>
#include <vector>
#include <svo/common/types.h>
namespace svo {
class Frame;
using FramePtr = std::shared_ptr<Frame>;
using FrameWeakPtr = std::weak_ptr<Frame>;

class PointIdProvider {  // point.h:22-34 (thread-safe point-ID provider; the tracker draws its track ids from it)
 public:
  PointIdProvider() = delete;
  static int getNewPointId() { return last_id_.fetch_add(1); }
 private:
  static inline std::atomic<int> last_id_{0};  // point.cpp:14
};

struct KeypointIdentifier {  // point.h:36-60
  FrameWeakPtr frame;
  int frame_id;
  size_t keypoint_index_;
  KeypointIdentifier(const FramePtr& _frame, const size_t _feature_index);  // defined after Frame (frame.h)
};
using KeypointIdentifierList = std::vector<KeypointIdentifier>;

class Point {
 public:
  int id_ = -1;
  Position pos_;
  KeypointIdentifierList obs_;
  std::array<int, 8> last_projected_kf_id_;
  int n_failed_reproj_ = 0;
  int n_succeeded_reproj_ = 0;
  bool in_ba_graph_ = false;
  explicit Point(const Position& pos) : pos_(pos) { last_projected_kf_id_.fill(-1); }  // point.cpp:29-35
  const Position& pos() const { return pos_; }
  int id() const { return id_; }
  inline void addObservation(const FramePtr& frame, const size_t feature_index);  // point.cpp:41-58, defined after Frame (frame.h)
  inline bool getCloseViewObs(const Eigen::Vector3d& framepos, FramePtr& ref_frame, size_t& ref_feature_index) const;
};
using PointPtr = std::shared_ptr<Point>;
}  // namespace svo
