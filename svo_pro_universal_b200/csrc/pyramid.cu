// (a1) Image pyramid: frame_utils::createImgPyramid + vk::halfSample.
// ref: src/svo_common/src/frame.cpp:372-386; src/vikit/vikit_common/src/vision.cpp:19-44 (SSE2 formula), :98-110 (scalar).
//
// One launch halves up to four times: a CTA owns a 256x16 tile of the source level, keeps every
// intermediate level of that tile in shared memory and streams each level out with 32-bit stores.
// The source is read exactly once with 128-bit loads; HBM traffic = read L_k + write L_k+1.. (algorithmic minimum).
#include "common.cuh"

namespace {

// 8 source bytes of the top row (t0,t1) and bottom row (b0,b1) -> 4 output bytes.
// rounding=1: SSE2 formula avg_epu16(avg_epu8(top,bottom) even, odd) = ((a+c+1)>>1 + (b+d+1)>>1 + 1) >> 1
// rounding=0: (a+b+c+d)/4 truncating.
SVO_D unsigned down4(unsigned t0, unsigned t1, unsigned b0, unsigned b1, bool rounding) {
  unsigned r0, r1;  // each holds two results in its 16-bit lanes
  if (rounding) {
    const unsigned v0 = __vavgu4(t0, b0), v1 = __vavgu4(t1, b1);
    r0 = (((v0 & 0x00FF00FFu) + ((v0 >> 8) & 0x00FF00FFu) + 0x00010001u) >> 1) & 0x00FF00FFu;
    r1 = (((v1 & 0x00FF00FFu) + ((v1 >> 8) & 0x00FF00FFu) + 0x00010001u) >> 1) & 0x00FF00FFu;
  } else {
    r0 = (((t0 & 0x00FF00FFu) + ((t0 >> 8) & 0x00FF00FFu) + (b0 & 0x00FF00FFu) + ((b0 >> 8) & 0x00FF00FFu)) >> 2) & 0x00FF00FFu;
    r1 = (((t1 & 0x00FF00FFu) + ((t1 >> 8) & 0x00FF00FFu) + (b1 & 0x00FF00FFu) + ((b1 >> 8) & 0x00FF00FFu)) >> 2) & 0x00FF00FFu;
  }
  // bytes: r0.b0, r0.b2, r1.b0, r1.b2
  return __byte_perm(r0, r1, 0x6420);
}

struct PyrW {  // writable view
  int cols[SVO_MAX_LEVELS], rows[SVO_MAX_LEVELS], pitch[SVO_MAX_LEVELS];
  unsigned long long frame_stride[SVO_MAX_LEVELS];
  uint8_t* data[SVO_MAX_LEVELS];
};

constexpr int kTileW = 256, kTileH = 32;

SVO_D uint4 ldStream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// One CTA halves a 256x32 tile of level l0 up to four times. Stage 1 works straight from registers: thread (row pair rp,
// 16-px chunk ch) loads two 128-bit rows and produces 8 pixels of level l0+1; later stages read the previous level's tile
// from shared memory. All index arithmetic is shifts; every level is written once with 64/32-bit stores.
__global__ void __launch_bounds__(256) pyr_down_fused_kernel(PyrW v, int first, int l0, int nh, unsigned round_mask) {
  __shared__ __align__(16) uint8_t t1[128 * 16];
  __shared__ __align__(16) uint8_t t2[64 * 8];
  __shared__ __align__(16) uint8_t t3[32 * 4];
  const int frame = first + blockIdx.z;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const unsigned long long f = (unsigned long long)frame;
  {  // stage 1: l0 -> l0+1
    const int rp = tid >> 4, ch = tid & 15;
    const int gx = x0 + ch * 16, gy = y0 + 2 * rp;
    uint4 top = make_uint4(0, 0, 0, 0), bot = make_uint4(0, 0, 0, 0);
    if (gy + 1 < v.rows[l0] && gx < v.pitch[l0]) {
      const uint8_t* src = v.data[l0] + v.frame_stride[l0] * f + (size_t)gy * v.pitch[l0] + gx;
      top = ldStream(reinterpret_cast<const uint4*>(src));
      bot = ldStream(reinterpret_cast<const uint4*>(src + v.pitch[l0]));
    }
    const bool rnd = (round_mask >> l0) & 1u;
    uint2 o;
    o.x = down4(top.x, top.y, bot.x, bot.y, rnd);
    o.y = down4(top.z, top.w, bot.z, bot.w, rnd);
    *reinterpret_cast<uint2*>(&t1[rp * 128 + ch * 8]) = o;
    const int L = l0 + 1;
    const int ox = (x0 >> 1) + ch * 8, oy = (y0 >> 1) + rp;
    if (oy < v.rows[L] && ox < v.cols[L]) *reinterpret_cast<uint2*>(v.data[L] + v.frame_stride[L] * f + (size_t)oy * v.pitch[L] + ox) = o;
  }
  if (nh < 2) return;
  __syncthreads();
  if (tid < 128) {  // stage 2: 128x16 -> 64x8, 4 px per thread
    const int r = tid >> 4, c4 = (tid & 15) * 4;
    const uint2 top = *reinterpret_cast<const uint2*>(&t1[(2 * r) * 128 + 2 * c4]);
    const uint2 bot = *reinterpret_cast<const uint2*>(&t1[(2 * r + 1) * 128 + 2 * c4]);
    const unsigned o = down4(top.x, top.y, bot.x, bot.y, (round_mask >> (l0 + 1)) & 1u);
    *reinterpret_cast<unsigned*>(&t2[r * 64 + c4]) = o;
    const int L = l0 + 2;
    const int ox = (x0 >> 2) + c4, oy = (y0 >> 2) + r;
    if (oy < v.rows[L] && ox < v.cols[L]) *reinterpret_cast<unsigned*>(v.data[L] + v.frame_stride[L] * f + (size_t)oy * v.pitch[L] + ox) = o;
  }
  if (nh < 3) return;
  __syncthreads();
  if (tid < 32) {  // stage 3: 64x8 -> 32x4
    const int r = tid >> 3, c4 = (tid & 7) * 4;
    const uint2 top = *reinterpret_cast<const uint2*>(&t2[(2 * r) * 64 + 2 * c4]);
    const uint2 bot = *reinterpret_cast<const uint2*>(&t2[(2 * r + 1) * 64 + 2 * c4]);
    const unsigned o = down4(top.x, top.y, bot.x, bot.y, (round_mask >> (l0 + 2)) & 1u);
    *reinterpret_cast<unsigned*>(&t3[r * 32 + c4]) = o;
    const int L = l0 + 3;
    const int ox = (x0 >> 3) + c4, oy = (y0 >> 3) + r;
    if (oy < v.rows[L] && ox < v.cols[L]) *reinterpret_cast<unsigned*>(v.data[L] + v.frame_stride[L] * f + (size_t)oy * v.pitch[L] + ox) = o;
  }
  if (nh < 4) return;
  __syncwarp();
  if (tid < 8) {  // stage 4: 32x4 -> 16x2 (same warp as stage 3)
    const int r = tid >> 2, c4 = (tid & 3) * 4;
    const uint2 top = *reinterpret_cast<const uint2*>(&t3[(2 * r) * 32 + 2 * c4]);
    const uint2 bot = *reinterpret_cast<const uint2*>(&t3[(2 * r + 1) * 32 + 2 * c4]);
    const unsigned o = down4(top.x, top.y, bot.x, bot.y, (round_mask >> (l0 + 3)) & 1u);
    const int L = l0 + 4;
    const int ox = (x0 >> 4) + c4, oy = (y0 >> 4) + r;
    if (oy < v.rows[L] && ox < v.cols[L]) *reinterpret_cast<unsigned*>(v.data[L] + v.frame_stride[L] * f + (size_t)oy * v.pitch[L] + ox) = o;
  }
}

}  // namespace

// round_mask bit l = 1 when the halving from level l to l+1 uses the SSE2 rounding formula.
unsigned svoHalfsampleRoundMask(const svo_cuda_pyr* pyr) {
  unsigned m = 0;
  if (pyr->halfsample_mode == 0) return 0;
  for (int l = 0; l + 1 < pyr->n_levels; ++l)
    if (pyr->cols[l] % 16 == 0) m |= (1u << l);  // vision.cpp:80-87 (buffers are aligned + contiguous by construction)
  return m;
}

int svoPyrBuildLaunch(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr, int first, int count) {
  if (count == 0 || pyr->n_levels == 1) return SVO_OK;
  PyrW v;
  for (int l = 0; l < pyr->n_levels; ++l) {
    v.cols[l] = pyr->cols[l]; v.rows[l] = pyr->rows[l]; v.pitch[l] = (int)pyr->pitch[l];
    v.frame_stride[l] = pyr->frame_stride[l];
    v.data[l] = pyr->data[l];
  }
  const unsigned mask = svoHalfsampleRoundMask(pyr);
  int l0 = 0;
  while (l0 + 1 < pyr->n_levels) {
    const int nh = min(4, pyr->n_levels - 1 - l0);
    dim3 grid((pyr->cols[l0] + kTileW - 1) / kTileW, (pyr->rows[l0] + kTileH - 1) / kTileH, count);
    pyr_down_fused_kernel<<<grid, 256, 0, ctx->stream>>>(v, first, l0, nh, mask);
    SVO_LAUNCH_CHECK(ctx);
    l0 += nh;
  }
  return SVO_OK;
}

extern "C" int svo_cuda_pyr_build(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr, int first, int count) {
  if (!ctx || !pyr || first < 0 || count < 0 || first + count > pyr->n_frames)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_pyr_build: bad arguments");
  return svoPyrBuildLaunch(ctx, pyr, first, count);
}
