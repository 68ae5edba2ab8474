// Shared internals of libsvo_cuda: context, pyramid batch layout, staging helper and the small
// SE3 / camera device functions used by kernels (b), (c) and (d).
//
// Device math follows the reference's minkindr / vikit_cameras arithmetic:
//   3rd/minkindr/include/kindr/minimal/implementation/rotation-quaternion-inl.h (exp :519-536, log :478-516,
//   product + renormalisation :435-442,:580-589), quat-transformation-inl.h (compose :150-156, inverse :209-213),
//   src/vikit/vikit_cameras/include/vikit/cameras/implementation/pinhole_projection.hpp:30-76,
//   src/vikit/vikit_cameras/include/vikit/cameras/radial_tangential_distortion.h:34-95.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/svo_cuda.h"

#define SVO_HD __host__ __device__ __forceinline__
#define SVO_D __device__ __forceinline__

// ------------------------------------------------------------------------------------------------
// context
struct svo_cuda_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  int sm_count = 148;
  bool attr_pyr = false, attr_reproj = false;  // cudaFuncSetAttribute applied on THIS context's device (per device, not per process)
  // Staging arena of SVO_MEM_HOST calls (Stager): one page-locked host buffer and one device buffer, halves for inputs / outputs. The small
  // arrays of a call travel in ONE copy each way instead of one cudaMallocAsync + one pageable copy per array (single-frame latency).
  uint8_t* stage_host = nullptr;
  uint8_t* stage_host_dev = nullptr;  // device-side address of stage_host (mapped page-locked memory): small results are written there directly
  uint8_t* stage_dev = nullptr;
  bool stage_busy = false;       // claimed by the Stager of the call in progress (nested Stagers fall back to per-array staging)
  // side streams + fork / join events of entry points that run independent parts of one call concurrently (depth_filter.cu: seed groups)
  static constexpr int kSideStreams = 3;
  cudaStream_t side_stream[kSideStreams] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kSideStreams] = {};
  int8_t* angle_bins = nullptr;  // 511 x 511 orientation-histogram bins of every u8 central-difference gradient (edgelet.cu), built on first use
  std::string last_error;
};

// Batch of frame pyramids. Level l of all frames is one allocation: [n_frames][rows_l][pitch_l] bytes,
// pitch_l = cols_l rounded up to 16 so every row start is 16-byte aligned for 128-bit loads, and
// frame_stride_l = pitch_l * rows_l rounded up to 256 bytes.
struct svo_cuda_pyr {
  int n_frames = 0, n_levels = 0, halfsample_mode = -1;
  int cols[SVO_MAX_LEVELS] = {0}, rows[SVO_MAX_LEVELS] = {0};
  size_t pitch[SVO_MAX_LEVELS] = {0}, frame_stride[SVO_MAX_LEVELS] = {0};
  uint8_t* data[SVO_MAX_LEVELS] = {nullptr};
  // TMA descriptors (CUtensorMap, 128 opaque bytes) of the levels the pyramid kernel has read so far; built on first use
  alignas(64) unsigned char tmap[SVO_MAX_LEVELS][128] = {};
  bool tmap_ready[SVO_MAX_LEVELS] = {false};
  // the same tensors with the FAST kernel's box (tile + halo), built on first use
  alignas(64) mutable unsigned char tmap_fast[SVO_MAX_LEVELS][128] = {};
  mutable bool tmap_fast_ready[SVO_MAX_LEVELS] = {false};
};
// pyramid.cu: TMA descriptor of level l as a {pitch, rows, frames} u8 tensor with a box of box_w x box_h x 1 bytes, encoded once
// (under a lock: a pyramid may be shared by contexts on several host threads) into map128 / *ready.
int svoEnsureLevelMap(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int l, int box_w, int box_h, unsigned char* map128, bool* ready);

// POD view of a pyramid batch passed to kernels by value.
struct PyrView {
  int n_levels;
  int cols[SVO_MAX_LEVELS], rows[SVO_MAX_LEVELS];
  int pitch[SVO_MAX_LEVELS];
  unsigned long long frame_stride[SVO_MAX_LEVELS];
  const uint8_t* data[SVO_MAX_LEVELS];
  SVO_HD const uint8_t* level(int frame, int l) const { return data[l] + frame_stride[l] * (unsigned long long)frame; }
};
inline PyrView makeView(const svo_cuda_pyr* p) {
  PyrView v;
  memset(&v, 0, sizeof(v));
  v.n_levels = p->n_levels;
  for (int l = 0; l < p->n_levels; ++l) {
    v.cols[l] = p->cols[l]; v.rows[l] = p->rows[l];
    v.pitch[l] = (int)p->pitch[l];
    v.frame_stride[l] = p->frame_stride[l];
    v.data[l] = p->data[l];
  }
  return v;
}

int svoFail(svo_cuda_ctx* ctx, int code, const char* what, const char* file, int line);
// ctx.cu: creates the context's side streams and events on first use (SVO_OK or an error code)
int svoEnsureSideStreams(svo_cuda_ctx* ctx);
#define SVO_FAIL(ctx, code, what) svoFail(ctx, code, what, __FILE__, __LINE__)
#define SVO_CUDA_TRY(ctx, expr)                                                            \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) return svoFail(ctx, SVO_ERR_CUDA, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
// Every entry point runs on the context's device, whatever device the calling thread had current.
#define SVO_BIND(ctx) SVO_CUDA_TRY(ctx, cudaSetDevice((ctx)->device))
#define SVO_LAUNCH_CHECK(ctx)                     \
  do {                                            \
    (ctx)->launches++;                            \
    SVO_CUDA_TRY(ctx, cudaGetLastError());        \
  } while (0)

// Staging of I/O arrays of a batched call. For SVO_MEM_DEVICE the caller's pointers are used in place; for SVO_MEM_HOST inputs are
// copied to the device and outputs copied back in finish(). Arrays of at most kStageSmall bytes go through the context's staging arena
// (packed into page-locked memory, one H2D copy issued by send() right before the first launch, one D2H copy in finish() — none for
// results of at most kZeroCopyOut bytes, which the kernels write into mapped host memory directly); larger ones
// are copied one by one from / to the caller's memory (which the caller may have page-locked for bandwidth) through stream-ordered
// temporaries.
class Stager {
 public:
  static constexpr size_t kStageHalf = 128 * 1024, kStageSmall = 32 * 1024;
  // write-only results of at most kZeroCopyOut bytes (outWriteOnly(); 16 KB of them per call) are written by the kernels straight into
  // mapped page-locked host memory: a single-frame call then ends with a stream synchronisation, not with a copy + synchronisation
  static constexpr size_t kZeroCopyOut = 4096, kZeroCopyRegion = 16 * 1024;
  Stager(svo_cuda_ctx* c, svo_mem m);
  ~Stager() { release(); }
  template <class T>
  const T* in(const T* p, size_t n) {
    if (!p || mem_ == SVO_MEM_DEVICE || n == 0) return p;
    if (void* a = arenaIn(p, n * sizeof(T))) return (const T*)a;
    void* d = alloc(n * sizeof(T));
    if (!d) return nullptr;
    if (cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, ctx_->stream) != cudaSuccess) failed_ = true;
    return (const T*)d;
  }
  template <class T>
  T* out(T* p, size_t n) {
    if (!p || mem_ == SVO_MEM_DEVICE || n == 0) return p;
    if (void* a = arenaOut(p, n * sizeof(T))) return (T*)a;
    void* d = alloc(n * sizeof(T));
    if (!d) return nullptr;
    outs_.push_back({p, d, n * sizeof(T)});
    return (T*)d;
  }
  // out() for a result array that the kernels only WRITE, once per element (no read-back, no atomics): small ones are written straight
  // into mapped page-locked host memory, so that the call ends with a stream synchronisation instead of a copy + synchronisation
  template <class T>
  T* outWriteOnly(T* p, size_t n) {
    if (!p || mem_ == SVO_MEM_DEVICE || n == 0) return p;
    if (void* z = zeroCopyOut(p, n * sizeof(T))) return (T*)z;
    return out(p, n);
  }
  template <class T>
  T* inout(T* p, size_t n) {
    if (!p || mem_ == SVO_MEM_DEVICE || n == 0) return p;
    void* d = alloc(n * sizeof(T));
    if (!d) return nullptr;
    if (cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, ctx_->stream) != cudaSuccess) failed_ = true;
    outs_.push_back({p, d, n * sizeof(T)});
    return (T*)d;
  }
  // device scratch that lives until finish()
  void* scratch(size_t bytes) { return alloc(bytes); }
  // Called by every entry point after its last in() / out() and before its first launch: sends the packed inputs (one H2D copy);
  // false = an allocation or a copy of this call failed (the entry point then returns finish(), which reports it).
  bool send();
  bool failed() const { return failed_; }
  int finish();

 private:
  struct Out { void* host; void* dev; size_t bytes; };
  void* alloc(size_t bytes);
  void* arenaIn(const void* p, size_t bytes);
  void* arenaOut(void* p, size_t bytes);
  void* zeroCopyOut(void* p, size_t bytes);
  void release();
  svo_cuda_ctx* ctx_;
  svo_mem mem_;
  bool failed_ = false;
  bool arena_ = false, sent_ = false;
  size_t in_used_ = 0, out_used_ = 0, zc_used_ = 0;
  std::vector<void*> allocs_;
  std::vector<Out> outs_;        // per-array D2H copies
  std::vector<Out> arena_outs_;  // host pointer | offset into the arena's output half (as a pointer into stage_host) | bytes
};

// ------------------------------------------------------------------------------------------------
// small device math
struct V2d { double x, y; };
struct V3d { double x, y, z; };
SVO_HD V3d operator+(const V3d& a, const V3d& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
SVO_HD V3d operator-(const V3d& a, const V3d& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
SVO_HD V3d operator*(const V3d& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
SVO_HD V3d operator-(const V3d& a) { return {-a.x, -a.y, -a.z}; }
SVO_HD double dot3(const V3d& a, const V3d& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SVO_HD V3d cross3(const V3d& a, const V3d& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
SVO_HD double norm3(const V3d& a) { return sqrt(dot3(a, a)); }
SVO_HD V3d normalized3(const V3d& a) {
  const double n2 = dot3(a, a);
  if (n2 > 0.0) { const double n = sqrt(n2); return {a.x / n, a.y / n, a.z / n}; }
  return a;
}
SVO_HD V2d normalized2(const V2d& a) {
  const double n2 = a.x * a.x + a.y * a.y;
  if (n2 > 0.0) { const double n = sqrt(n2); return {a.x / n, a.y / n}; }
  return a;
}

struct Quatd { double w, x, y, z; };
struct SE3d { Quatd q; V3d t; };
struct M3d { double m[3][3]; };

SVO_HD V3d operator*(const M3d& A, const V3d& v) {
  return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z,
          A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
          A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
SVO_HD double quatSqNorm(const Quatd& q) { return q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z; }
SVO_HD void quatNormalize(Quatd& q) {
  const double n = sqrt(quatSqNorm(q));
  q.w /= n; q.x /= n; q.y /= n; q.z /= n;
}
SVO_HD Quatd quatMul(const Quatd& a, const Quatd& b) {
  Quatd r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  if (fabs(quatSqNorm(r) - 1.0) > 1.0e-4) quatNormalize(r);
  return r;
}
SVO_HD Quatd quatConj(const Quatd& q) { return {q.w, -q.x, -q.y, -q.z}; }
SVO_HD V3d quatRotate(const Quatd& q, const V3d& v) {
  const V3d qv{q.x, q.y, q.z};
  V3d uv = cross3(qv, v);
  uv = uv + uv;
  return v + uv * q.w + cross3(qv, uv);
}
SVO_HD V3d quatInverseRotate(const Quatd& q, const V3d& v) {
  const double n2 = quatSqNorm(q);
  const Quatd qi{q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
  return quatRotate(qi, v);
}
SVO_HD M3d quatToMatrix(const Quatd& q) {
  M3d R;
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R.m[0][0] = 1.0 - (tyy + tzz); R.m[0][1] = txy - twz;         R.m[0][2] = txz + twy;
  R.m[1][0] = txy + twz;         R.m[1][1] = 1.0 - (txx + tzz); R.m[1][2] = tyz - twx;
  R.m[2][0] = txz - twy;         R.m[2][1] = tyz + twx;         R.m[2][2] = 1.0 - (txx + tyy);
  return R;
}
// pow(DBL_EPSILON, 1/4)
#define SVO_EPS4ROOT 1.220703125e-4
SVO_HD double arcSinXOverX(double x) {
  if (fabs(x) < SVO_EPS4ROOT) return 1.0 + x * x * (1.0 / 6.0);
  return asin(x) / x;
}
SVO_HD Quatd quatExp(const V3d& dx) {
  const double theta = norm3(dx);
  double na;
  if (theta < SVO_EPS4ROOT) na = 0.5 + (theta * theta) * (1.0 / 48.0);
  else na = sin(theta * 0.5) / theta;
  const double ct = cos(theta * 0.5);
  return {ct, dx.x * na, dx.y * na, dx.z * na};
}
SVO_HD V3d quatLog(const Quatd& q) {
  const V3d a{q.x, q.y, q.z};
  const double na = norm3(a);
  const double eta = q.w;
  double scale;
  if (fabs(eta) < na) {
    if (eta >= 0) scale = acos(eta) / na;
    else scale = -acos(-eta) / na;
  } else {
    if (eta > 0) scale = arcSinXOverX(na);
    else scale = -arcSinXOverX(na);
  }
  return a * (2.0 * scale);
}
SVO_HD SE3d se3Mul(const SE3d& a, const SE3d& b) {
  SE3d r;
  r.q = quatMul(a.q, b.q);
  r.t = a.t + quatRotate(a.q, b.t);
  return r;
}
SVO_HD V3d se3Apply(const SE3d& T, const V3d& p) { return quatRotate(T.q, p) + T.t; }
SVO_HD SE3d se3Inv(const SE3d& T) {
  SE3d r;
  r.q = quatConj(T.q);
  r.t = -quatInverseRotate(T.q, T.t);
  return r;
}
SVO_HD SE3d se3Load(const double* a) { return {{a[0], a[1], a[2], a[3]}, {a[4], a[5], a[6]}}; }
SVO_HD void se3Store(const SE3d& T, double* a) {
  a[0] = T.q.w; a[1] = T.q.x; a[2] = T.q.y; a[3] = T.q.z;
  a[4] = T.t.x; a[5] = T.t.y; a[6] = T.t.z;
}

// camera -------------------------------------------------------------------------------------------
SVO_HD void camDistort(const svo_camera& c, double x, double y, double& xd, double& yd) {
  if (c.distortion == 0) { xd = x; yd = y; return; }
  const double xx = x * x, yy = y * y, xy = x * y;
  const double xy2 = 2.0 * xy;
  const double r2 = xx + yy;
  const double cdist = (c.k1 + c.k2 * r2) * r2;
  xd = x + x * cdist + c.p1 * xy2 + c.p2 * (r2 + 2.0 * xx);
  yd = y + y * cdist + c.p2 * xy2 + c.p1 * (r2 + 2.0 * yy);
}
SVO_HD void camUndistort(const svo_camera& c, double& x, double& y) {
  if (c.distortion == 0) return;
  const double x0 = x, y0 = y;
  for (int i = 0; i < 5; ++i) {
    const double xx = x * x, yy = y * y, xy = x * y;
    const double xy2 = 2 * xy;
    const double r2 = xx + yy;
    const double icdist = 1.0 / (1.0 + (c.k1 + c.k2 * r2) * r2);
    const double dx = c.p1 * xy2 + c.p2 * (r2 + 2.0 * xx);
    const double dy = c.p2 * xy2 + c.p1 * (r2 + 2.0 * yy);
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
}
SVO_HD void camDistJacobian(const svo_camera& c, double px, double py, double J[2][2]) {
  if (c.distortion == 0) { J[0][0] = 1; J[0][1] = 0; J[1][0] = 0; J[1][1] = 1; return; }
  const double xx = px * px, yy = py * py, xy = px * py;
  const double r2 = xx + yy;
  const double cdist = (c.k1 + c.k2 * r2) * r2;
  const double k2_r2_x4 = c.k2 * r2 * 4.0;
  const double cdist_p1 = cdist + 1.0;
  J[0][0] = cdist_p1 + c.k1 * 2.0 * xx + k2_r2_x4 * xx + 2.0 * c.p1 * py + 6.0 * c.p2 * px;
  J[1][1] = cdist_p1 + c.k1 * 2.0 * yy + k2_r2_x4 * yy + 2.0 * c.p2 * px + 6.0 * c.p1 * py;
  J[1][0] = 2.0 * c.k1 * xy + k2_r2_x4 * xy + 2.0 * c.p1 * px + 2.0 * c.p2 * py;
  J[0][1] = J[1][0];
}
SVO_HD V3d camBackProject3(const svo_camera& c, double u, double v) {
  double x = (u - c.cx) * (1.0 / c.fx);
  double y = (v - c.cy) * (1.0 / c.fy);
  camUndistort(c, x, y);
  return {x, y, 1.0};
}
SVO_HD V2d camProject3(const svo_camera& c, const V3d& p) {
  const double z_inv = 1 / p.z;
  const double u = p.x * z_inv, v = p.y * z_inv;
  double ud, vd;
  camDistort(c, u, v, ud, vd);
  return {c.fx * ud + c.cx, c.fy * vd + c.cy};
}
SVO_HD void camProject3Jac(const svo_camera& c, const V3d& p, double J[2][3]) {
  const double z_inv = 1 / p.z;
  const double u = p.x * z_inv, v = p.y * z_inv;
  double duv[2][3];
  duv[0][0] = z_inv; duv[0][1] = 0.0;   duv[0][2] = -p.x * z_inv * z_inv;
  duv[1][0] = 0.0;   duv[1][1] = z_inv; duv[1][2] = -p.y * z_inv * z_inv;
  double Jd[2][2];
  camDistJacobian(c, u, v, Jd);
  const double FJ[2][2] = {{c.fx * Jd[0][0], c.fx * Jd[0][1]}, {c.fy * Jd[1][0], c.fy * Jd[1][1]}};
  for (int r = 0; r < 2; ++r)
    for (int k = 0; k < 3; ++k) J[r][k] = FJ[r][0] * duv[0][k] + FJ[r][1] * duv[1][k];
}
SVO_HD double camAngleError(const svo_camera& c, double img_err) {
  return atan(img_err / (2.0 * c.fx)) + atan(img_err / (2.0 * c.fy));
}

// ---- mbarrier / TMA tile loads (pyramid.cu, fast.cu) ----------------------------------------------------------------------------
SVO_D unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
SVO_D void mbarInit(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
SVO_D void mbarExpectTx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
SVO_D void mbarWait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared TMA tile load of the map's {w, h, 1} box at (x, y, frame); out-of-tensor bytes are zero-filled and the full box
// size is always completed on the barrier
SVO_D void tmaLoadTile(unsigned dst, const void* tmap, int x, int y, int z, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

// Bilinear tap loader: the 2 aligned 32-bit words covering bytes [x0, x0+8) of a row whose start is
// 4-byte aligned; returns bytes x0..x0+3 in `a` and x0+4..x0+7 in `b` (little endian packed).
SVO_D void loadRow8(const uint8_t* row, int x0, unsigned& a, unsigned& b) {
  const unsigned* w = reinterpret_cast<const unsigned*>(row + (x0 & ~3));
  const unsigned w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
  const unsigned sh = (x0 & 3) * 8;
  a = __funnelshift_r(w0, w1, sh);
  b = __funnelshift_r(w1, w2, sh);
}
SVO_D unsigned byteOf(unsigned w, int i) { return (w >> (8 * i)) & 0xffu; }
