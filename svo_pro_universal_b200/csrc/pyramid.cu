// (a1) Image pyramid: frame_utils::createImgPyramid + vk::halfSample.
// ref: src/svo_common/src/frame.cpp:372-386; src/vikit/vikit_common/src/vision.cpp:19-44 (SSE2 formula), :98-110 (scalar).
//
// One launch halves up to four times. Persistent CTAs (a few per SM) walk the 256x32 source tiles of every frame; each tile
// is brought into a shared-memory ring by ONE TMA tile load (cp.async.bulk.tensor.3d over a {pitch, rows, frames} tensor
// map, out-of-image bytes zero-filled, completion counted on an mbarrier), so a CTA always has kStages tiles of loads in
// flight while it computes. Every
// intermediate level of the tile stays in shared memory and each level is written once with 64/32-bit stores.
// HBM traffic = read L_k + write L_k+1.. (the algorithmic minimum).
#include "common.cuh"
#include <cuda.h>
#include <algorithm>
#include <mutex>

namespace {

// 8 source bytes of the top row (t0,t1) and bottom row (b0,b1) -> 4 output bytes, two results per 16-bit-lane register.
// rounding=1: SSE2 formula avg_epu16(avg_epu8(top,bottom) even, odd) = ((a+c+1)>>1 + (b+d+1)>>1 + 1) >> 1
// rounding=0: (a+b+c+d)/4 truncating.
// Even/odd bytes are spread into 16-bit lanes with one PRMT each; the bits a shift drags across the lane boundary land in
// byte 1 / byte 3, which the final PRMT does not select.
SVO_D unsigned evenBytes(unsigned v) { return __byte_perm(v, 0u, 0x4240); }
SVO_D unsigned oddBytes(unsigned v) { return __byte_perm(v, 0u, 0x4341); }
SVO_D unsigned down4(unsigned t0, unsigned t1, unsigned b0, unsigned b1, bool rounding) {
  unsigned r0, r1;
  if (rounding) {
    const unsigned v0 = __vavgu4(t0, b0), v1 = __vavgu4(t1, b1);
    r0 = (evenBytes(v0) + oddBytes(v0) + 0x00010001u) >> 1;
    r1 = (evenBytes(v1) + oddBytes(v1) + 0x00010001u) >> 1;
  } else {
    r0 = (evenBytes(t0) + oddBytes(t0) + evenBytes(b0) + oddBytes(b0)) >> 2;
    r1 = (evenBytes(t1) + oddBytes(t1) + evenBytes(b1) + oddBytes(b1)) >> 2;
  }
  return __byte_perm(r0, r1, 0x6420);  // bytes: r0.b0, r0.b2, r1.b0, r1.b2
}

struct PyrW {  // writable view
  int cols[SVO_MAX_LEVELS], rows[SVO_MAX_LEVELS], pitch[SVO_MAX_LEVELS];
  unsigned long long frame_stride[SVO_MAX_LEVELS];
  uint8_t* data[SVO_MAX_LEVELS];
};

constexpr int kTileW = 256, kTileH = 32;
constexpr int kStages = 4;                        // shared-memory ring depth (tiles of loads in flight per CTA)
constexpr int kTileBytes = kTileW * kTileH;       // 8 KB
constexpr int kPyrThreads = 256;

struct TileCoord { int frame, x0, y0; };
// floor(x / d) for x * d < 2^32 as one multiply-high; magic = 0 stands for d == 1.
inline unsigned divMagic(int d) { return d <= 1 ? 0u : (unsigned)(0x100000000ull / (unsigned)d + 1ull); }
SVO_D int divFast(int x, unsigned magic) { return magic ? (int)__umulhi((unsigned)x, magic) : x; }

struct TileGrid { int tiles_x, tiles_per_frame, n_items; unsigned magic_x, magic_frame; };
SVO_D TileCoord tileOf(int item, const TileGrid& g, int first) {
  const int f = divFast(item, g.magic_frame), r = item - f * g.tiles_per_frame;
  const int ty = divFast(r, g.magic_x), tx = r - ty * g.tiles_x;
  return {first + f, tx * kTileW, ty * kTileH};
}

// Persistent CTAs: item k of a CTA lives in ring slot k % kStages, mbarrier phase (k / kStages) & 1. Per item:
//   wait for the tile; stage 1 (256x32 -> 128x16): two 128-bit shared loads per thread, 8 output pixels;
//   stage 2 (-> 64x8) inside the warp: warp w produced level-1 rows 2w, 2w+1, i.e. exactly level-2 row w (__syncwarp only);
//   ONE block barrier (level-2 tile complete, ring slot free -> thread 0 refills it with the tile kStages items ahead);
//   stages 3-4 (-> 32x4 -> 16x2) by one warp, rotating over the warps from item to item; the level-2 tile is double
//   buffered so the other warps run ahead into the next item meanwhile.
__global__ void __launch_bounds__(kPyrThreads, 5) pyr_down_fused_kernel(const __grid_constant__ CUtensorMap tmap, PyrW v, int first, int l0,
                                                                     int nh, unsigned round_mask, TileGrid g) {
  extern __shared__ __align__(128) uint8_t s_ring[];  // [kStages][kTileBytes]
  __shared__ __align__(16) uint8_t t1[128 * 16];
  __shared__ __align__(16) uint8_t t2[2][64 * 8];
  __shared__ __align__(16) uint8_t t3[2][32 * 4];
  __shared__ __align__(8) unsigned long long s_full[kStages];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_mine = (g.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // items blockIdx.x + k * gridDim.x

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbarInit(smemAddr(&s_full[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // producer: one thread arms the slot's barrier with the tile size and issues the tile load
  auto issue = [&](int k) {
    const TileCoord t = tileOf((int)blockIdx.x + k * (int)gridDim.x, g, first);
    const int slot = k % kStages;
    const unsigned bar = smemAddr(&s_full[slot]);
    mbarExpectTx(bar, kTileBytes);
    tmaLoadTile(smemAddr(s_ring + slot * kTileBytes), &tmap, t.x0, t.y0, t.frame, bar);
  };
  if (tid == 0)
    for (int k = 0; k < min(kStages, n_mine); ++k) issue(k);

  const bool rnd1 = (round_mask >> l0) & 1u, rnd2 = (round_mask >> (l0 + 1)) & 1u, rnd3 = (round_mask >> (l0 + 2)) & 1u,
             rnd4 = (round_mask >> (l0 + 3)) & 1u;
  const int L1 = l0 + 1, L2 = l0 + 2, L3 = l0 + 3, L4 = l0 + 4;
  // loop-invariant output addressing of the two big levels: per-thread offset inside a tile + per-level constants
  const int rp = tid >> 4, ch = tid & 15;
  uint8_t* const out1 = v.data[L1] + (size_t)rp * v.pitch[L1] + ch * 8;
  const unsigned long long fs1 = v.frame_stride[L1];
  const int pitch1 = v.pitch[L1], rows1 = v.rows[L1], cols1 = v.cols[L1];
  uint8_t* const out2 = nh >= 2 ? v.data[L2] + (size_t)warp * v.pitch[L2] + lane * 4 : nullptr;
  const unsigned long long fs2 = nh >= 2 ? v.frame_stride[L2] : 0ull;
  const int pitch2 = nh >= 2 ? v.pitch[L2] : 0, rows2 = nh >= 2 ? v.rows[L2] : 0, cols2 = nh >= 2 ? v.cols[L2] : 0;
  const uint8_t* const src1 = s_ring + (2 * rp) * kTileW + ch * 16;
  for (int k = 0; k < n_mine; ++k) {
    const TileCoord t = tileOf((int)blockIdx.x + k * (int)gridDim.x, g, first);
    const unsigned long long f = (unsigned long long)t.frame;
    const int slot = k % kStages;
    mbarWait(smemAddr(&s_full[slot]), (unsigned)((k / kStages) & 1));
    {  // stage 1: l0 -> l0+1 (warp w: level-1 rows 2w and 2w+1)
      const uint8_t* src = src1 + slot * kTileBytes;
      const uint4 top = *reinterpret_cast<const uint4*>(src);
      const uint4 bot = *reinterpret_cast<const uint4*>(src + kTileW);
      uint2 o;
      o.x = down4(top.x, top.y, bot.x, bot.y, rnd1);
      o.y = down4(top.z, top.w, bot.z, bot.w, rnd1);
      *reinterpret_cast<uint2*>(&t1[rp * 128 + ch * 8]) = o;
      const int tx1 = t.x0 >> 1, ty1 = t.y0 >> 1;
      if (rp < rows1 - ty1 && ch * 8 < cols1 - tx1) *reinterpret_cast<uint2*>(out1 + fs1 * f + (size_t)(ty1 * pitch1 + tx1)) = o;
    }
    uint8_t* t2k = t2[k & 1];
    if (nh >= 2) {
      __syncwarp();
      if (lane < 16) {  // stage 2: level-2 row `warp`, 4 px per lane, from this warp's own two level-1 rows
        const int c4 = lane * 4;
        const uint2 top = *reinterpret_cast<const uint2*>(&t1[(2 * warp) * 128 + 2 * c4]);
        const uint2 bot = *reinterpret_cast<const uint2*>(&t1[(2 * warp + 1) * 128 + 2 * c4]);
        const unsigned o = down4(top.x, top.y, bot.x, bot.y, rnd2);
        *reinterpret_cast<unsigned*>(&t2k[warp * 64 + c4]) = o;
        const int tx2 = t.x0 >> 2, ty2 = t.y0 >> 2;
        if (warp < rows2 - ty2 && c4 < cols2 - tx2) *reinterpret_cast<unsigned*>(out2 + fs2 * f + (size_t)(ty2 * pitch2 + tx2)) = o;
      }
      __syncwarp();  // this warp's level-1 rows may be overwritten by its next item
    }
    __syncthreads();  // level-2 tile complete; every thread has finished reading the ring slot
    if (tid == 0 && k + kStages < n_mine) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // order the generic-proxy reads before the async refill
      issue(k + kStages);
    }
    if (nh >= 3 && warp == (k & 7)) {  // stages 3-4 by one warp (private double-buffered level-3 tile)
      uint8_t* t3k = t3[k & 1];
      const int r = lane >> 3, c4 = (lane & 7) * 4;
      const uint2 top = *reinterpret_cast<const uint2*>(&t2k[(2 * r) * 64 + 2 * c4]);
      const uint2 bot = *reinterpret_cast<const uint2*>(&t2k[(2 * r + 1) * 64 + 2 * c4]);
      const unsigned o = down4(top.x, top.y, bot.x, bot.y, rnd3);
      *reinterpret_cast<unsigned*>(&t3k[r * 32 + c4]) = o;
      const int ox = (t.x0 >> 3) + c4, oy = (t.y0 >> 3) + r;
      if (oy < v.rows[L3] && ox < v.cols[L3]) *reinterpret_cast<unsigned*>(v.data[L3] + v.frame_stride[L3] * f + (size_t)oy * v.pitch[L3] + ox) = o;
      __syncwarp();
      if (nh >= 4 && lane < 8) {  // stage 4: 32x4 -> 16x2
        const int r4 = lane >> 2, d4 = (lane & 3) * 4;
        const uint2 top4 = *reinterpret_cast<const uint2*>(&t3k[(2 * r4) * 32 + 2 * d4]);
        const uint2 bot4 = *reinterpret_cast<const uint2*>(&t3k[(2 * r4 + 1) * 32 + 2 * d4]);
        const unsigned o4 = down4(top4.x, top4.y, bot4.x, bot4.y, rnd4);
        const int ox4 = (t.x0 >> 4) + d4, oy4 = (t.y0 >> 4) + r4;
        if (oy4 < v.rows[L4] && ox4 < v.cols[L4]) *reinterpret_cast<unsigned*>(v.data[L4] + v.frame_stride[L4] * f + (size_t)oy4 * v.pitch[L4] + ox4) = o4;
      }
    }
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

// {pitch, rows, frames} u8 tensor of level l with the caller's box (the pyramid kernel: one 256x32 tile; FAST: tile + halo); the driver
// entry point is fetched through the runtime so that the library does not link against libcuda.
int svoEnsureLevelMap(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int l, int box_w, int box_h, unsigned char* map128, bool* ready) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (*ready) return SVO_OK;
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
      return SVO_FAIL(ctx, SVO_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    encode = (EncodeFn)fn;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)pyr->pitch[l], (cuuint64_t)pyr->rows[l], (cuuint64_t)pyr->n_frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pyr->pitch[l], (cuuint64_t)pyr->frame_stride[l]};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(map128), CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, pyr->data[l], dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return SVO_FAIL(ctx, SVO_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  *ready = true;
  return SVO_OK;
}

int svoPyrBuildLaunch(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr, int first, int count) {
  if (count == 0 || pyr->n_levels == 1) return SVO_OK;
  SVO_BIND(ctx);
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
    const int tiles_x = (pyr->cols[l0] + kTileW - 1) / kTileW, tiles_y = (pyr->rows[l0] + kTileH - 1) / kTileH;
    const long long n_items = (long long)tiles_x * tiles_y * count;
    if (n_items > 0x7fffffffLL) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_pyr_build: too many tiles in one call");
    constexpr int kCtasPerSm = 5;  // 5 x (32 KB ring + 2.7 KB) of shared memory, 1280 threads
    const int grid = (int)std::min<long long>(n_items, (long long)ctx->sm_count * kCtasPerSm);
    if (!ctx->attr_pyr) {
      SVO_CUDA_TRY(ctx, cudaFuncSetAttribute(pyr_down_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kTileBytes));
      ctx->attr_pyr = true;
    }
    const int rc = svoEnsureLevelMap(ctx, pyr, l0, kTileW, kTileH, pyr->tmap[l0], &pyr->tmap_ready[l0]);
    if (rc != SVO_OK) return rc;
    TileGrid tg;
    tg.tiles_x = tiles_x; tg.tiles_per_frame = tiles_x * tiles_y; tg.n_items = (int)n_items;
    tg.magic_x = divMagic(tiles_x); tg.magic_frame = divMagic(tiles_x * tiles_y);
    pyr_down_fused_kernel<<<grid, kPyrThreads, kStages * kTileBytes, ctx->stream>>>(*reinterpret_cast<const CUtensorMap*>(pyr->tmap[l0]), v, first, l0,
                                                                                    nh, mask, tg);
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
