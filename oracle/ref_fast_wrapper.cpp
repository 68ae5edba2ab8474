// ORACLE — TEST INFRASTRUCTURE ONLY.
// Thin C wrapper around the REFERENCE's own FAST sources (src/fast_neon/src/*.cpp), compiled from
// where they lie under /root/reference into oracle/_ref/libfast_ref.so by oracle/Makefile.
// No reference source is copied into this repository; this file only declares the C entry points.
#include <fast/fast.h>   // -I/root/reference/src/fast_neon/include
#include <vector>
#include <algorithm>

extern "C" {

// which: 0 = fast_corner_detect_10_sse2 (what fastDetector calls on x86,
//            src/svo_direct/src/feature_detection_utils.cpp:160-164),
//        1 = fast_corner_detect_10 (plain), 2 = fast_corner_detect_9 (plain)
int ref_fast_detect(const unsigned char* img, int w, int h, int stride, int barrier, int which, short* xy, int cap) {
  std::vector<fast::fast_xy> c;
  if (which == 0) fast::fast_corner_detect_10_sse2(img, w, h, stride, (short)barrier, c);
  else if (which == 1) fast::fast_corner_detect_10(img, w, h, stride, (short)barrier, c);
  else fast::fast_corner_detect_9(img, w, h, stride, (short)barrier, c);
  const int n = std::min<int>(cap, (int)c.size());
  for (int i = 0; i < n; ++i) { xy[2 * i] = c[i].x; xy[2 * i + 1] = c[i].y; }
  return (int)c.size();
}

void ref_fast_score10(const unsigned char* img, int stride, const short* xy, int n, int threshold, int* scores) {
  std::vector<fast::fast_xy> c;
  c.reserve(n);
  for (int i = 0; i < n; ++i) c.push_back(fast::fast_xy(xy[2 * i], xy[2 * i + 1]));
  std::vector<int> s;
  fast::fast_corner_score_10(img, stride, c, threshold, s);
  for (int i = 0; i < n; ++i) scores[i] = s[i];
}

int ref_fast_nonmax3x3(const short* xy, const int* scores, int n, int* idx_out) {
  std::vector<fast::fast_xy> c;
  c.reserve(n);
  for (int i = 0; i < n; ++i) c.push_back(fast::fast_xy(xy[2 * i], xy[2 * i + 1]));
  std::vector<int> s(scores, scores + n), nm;
  fast::fast_nonmax_3x3(c, s, nm);
  for (size_t i = 0; i < nm.size(); ++i) idx_out[i] = nm[i];
  return (int)nm.size();
}

}  // extern "C"
