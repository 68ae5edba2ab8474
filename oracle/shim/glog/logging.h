// TEST INFRASTRUCTURE ONLY (oracle/): glog stand-in so the reference sources compile without glog. CHECK* abort on
// failure like glog; LOG/VLOG/DLOG swallow their streams (FATAL aborts).
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace svo_shim_glog {
struct NullStream {
  template <typename T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct FatalStream {
  std::ostringstream os;
  template <typename T> FatalStream& operator<<(const T& v) { os << v; return *this; }
  FatalStream& operator<<(std::ostream& (*f)(std::ostream&)) { os << f; return *this; }
  [[noreturn]] ~FatalStream() { std::cerr << "CHECK/LOG(FATAL): " << os.str() << std::endl; std::abort(); }
};
struct Voidify { void operator&(NullStream&) {} void operator&(FatalStream&) {} };
template <typename T> T& CheckNotNull(T& t) { if (t == nullptr) std::abort(); return t; }
template <typename T> T* CheckNotNull(T* t) { if (t == nullptr) std::abort(); return t; }
}  // namespace svo_shim_glog

#define SVO_SHIM_NULL svo_shim_glog::NullStream()
#define LOG(severity) SVO_SHIM_LOG_##severity
#define SVO_SHIM_LOG_INFO SVO_SHIM_NULL
#define SVO_SHIM_LOG_WARNING SVO_SHIM_NULL
#define SVO_SHIM_LOG_ERROR SVO_SHIM_NULL
#define SVO_SHIM_LOG_FATAL svo_shim_glog::FatalStream()
#define DLOG(severity) SVO_SHIM_NULL
#define VLOG(level) SVO_SHIM_NULL
#define DVLOG(level) SVO_SHIM_NULL
#define LOG_IF(severity, cond) SVO_SHIM_NULL
#define VLOG_IF(level, cond) SVO_SHIM_NULL
#define LOG_FIRST_N(severity, n) SVO_SHIM_NULL
#define LOG_EVERY_N(severity, n) SVO_SHIM_NULL
#define VLOG_IS_ON(level) false
#define CHECK(cond) (cond) ? (void)0 : svo_shim_glog::Voidify() & svo_shim_glog::FatalStream() << #cond << " "
#define CHECK_OP(a, op, b) CHECK((a)op(b))
#define CHECK_EQ(a, b) CHECK_OP(a, ==, b)
#define CHECK_NE(a, b) CHECK_OP(a, !=, b)
#define CHECK_LT(a, b) CHECK_OP(a, <, b)
#define CHECK_LE(a, b) CHECK_OP(a, <=, b)
#define CHECK_GT(a, b) CHECK_OP(a, >, b)
#define CHECK_GE(a, b) CHECK_OP(a, >=, b)
#define DCHECK(cond) CHECK(cond)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) CHECK_NE(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define CHECK_NOTNULL(p) svo_shim_glog::CheckNotNull(p)
#define DCHECK_NOTNULL(p) svo_shim_glog::CheckNotNull(p)
#define CHECK_NEAR(a, b, tol) CHECK(((a) - (b)) <= (tol) && ((b) - (a)) <= (tol))
#define CHECK_DOUBLE_EQ(a, b) CHECK_NEAR(a, b, 1e-12)
