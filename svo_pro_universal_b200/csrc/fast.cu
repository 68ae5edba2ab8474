// (a2-a5) Pyramidal FAST detector with 3x3 non-max suppression and grid-cell arg-max.
//
// ref: src/svo_direct/src/feature_detection_utils.cpp:145-194 (fastDetector)
//      src/fast_neon/src/faster_corner_10_sse.cpp:15-202, fast_10.cpp (segment test, region [3,w-3)x[3,h-3))
//      src/fast_neon/src/fast_10_score.cpp:21-3178 (score = largest barrier at which the pixel is still a corner)
//      src/fast_neon/src/nonmax_3x3.cpp:17-112 (suppress when any 8-neighbour corner has score >= own)
//      src/svo_common/include/svo/common/occupancy_grid_2d.h:82-95 (cell index)
//
// Closed form used instead of the generated decision trees: with d_i = I_i - p on the 16-pixel circle,
//   margin = max over the 16 arcs of ARC contiguous pixels of max(min_arc d_i, -max_arc d_i) - 1
// the pixel is a corner at barrier b iff margin >= b, and fast_corner_score_10 = max(b, margin).
//
// Kernel layout (one launch for all levels of all frames): a CTA owns a 128x32 interior tile of one level; the u8 tile
// with a 4-pixel halo is staged in shared memory by one TMA box load. Three compacting stages keep every
// lane busy on the rare pixels that need work:
//   1. compass quick-reject on ALL pixels, 4 pixels per thread with byte SIMD (VABSDIFF4.U8 + carry compares): an arc
//      of >= 9 circle pixels always contains two adjacent compass points, i.e. one of {N,S} and one of {E,W};
//      survivors are pushed on a shared-memory candidate list;
//   2. exact margin on the candidates only: bright (I_i - p) and dark (p - I_i) margins ride in the two 16-bit halves
//      of one register, so the 16 arc minima and their maximum cost 56 packed min/max (VIMNMX3.U16x2);
//      corners (margin >= threshold) write their score into a tile score map and go on a corner list;
//   3. 3x3 non-max on the corner list (neighbours outside the tile are scored on demand from the halo), border /
//      occupancy tests and a 64-bit atomicMax per grid cell whose key reproduces the reference's visiting order
//      (score desc, then level asc, y asc, x asc — `score > corners[k].score` is strict).
#include "common.cuh"

namespace {

constexpr int kTW = 128, kTH = 32, kHalo = 4;
constexpr int kPadL = 16;                      // staged columns start at x0 - 16 so every 128-bit load is aligned
constexpr int kSPitch = kTW + 2 * kPadL;       // 160 staged bytes per row
constexpr int kSRows = kTH + 2 * kHalo;        // 40 staged rows
constexpr int kThreadsFast = 256;

// Circle offsets in the order of fast_10_score.cpp:3158-3175 (pixel[0] = (0,3), clockwise through (3,0), (0,-3), (-3,0)).
#define SVO_FAST_CIRCLE(p, pitch, X)                                                                               \
  X(0, (p)[3 * (pitch)])       X(1, (p)[3 * (pitch) + 1])   X(2, (p)[2 * (pitch) + 2])   X(3, (p)[(pitch) + 3])     \
  X(4, (p)[3])                 X(5, (p)[-(pitch) + 3])      X(6, (p)[-2 * (pitch) + 2])  X(7, (p)[-3 * (pitch) + 1]) \
  X(8, (p)[-3 * (pitch)])      X(9, (p)[-3 * (pitch) - 1])  X(10, (p)[-2 * (pitch) - 2]) X(11, (p)[-(pitch) - 3])   \
  X(12, (p)[-3])               X(13, (p)[(pitch) - 3])      X(14, (p)[2 * (pitch) - 2])  X(15, (p)[3 * (pitch) - 1])

// margin = max over the 16 arcs of ARC contiguous circle pixels of max(min_arc(I_i - p), min_arc(p - I_i)) - 1.
// Packed: v_i = (I_i - p + 256) | (p - I_i + 256) << 16 — one IMAD per circle pixel (I_i * 0xFFFF0001 + K; both halves
// stay in [1, 511], so no borrow crosses the halves) — then unsigned 16x2 min over each arc and max over the arcs.
template <int ARC>
SVO_D int fastMargin(const uint8_t* p, int pitch) {
  const unsigned c = p[0];
  const unsigned K = (256u - c) | ((c + 256u) << 16);
  unsigned v[16];
#define SVO_X(i, e) v[i] = (unsigned)(e) * 0xFFFF0001u + K;
  SVO_FAST_CIRCLE(p, pitch, SVO_X)
#undef SVO_X
  unsigned arc[16];
  if (ARC == 10) {
    unsigned m2[16], m4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m2[k] = __vminu2(v[k], v[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m4[k] = __vminu2(m2[k], m2[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) arc[k] = __vimin3_u16x2(m4[k], m4[(k + 4) & 15], m2[(k + 8) & 15]);
  } else {
    unsigned m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_u16x2(v[k], v[(k + 1) & 15], v[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) arc[k] = __vimin3_u16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
  }
  unsigned b0 = __vimax3_u16x2(arc[0], arc[1], arc[2]), b1 = __vimax3_u16x2(arc[3], arc[4], arc[5]);
  unsigned b2 = __vimax3_u16x2(arc[6], arc[7], arc[8]), b3 = __vimax3_u16x2(arc[9], arc[10], arc[11]);
  unsigned b4 = __vimax3_u16x2(arc[12], arc[13], arc[14]);
  b0 = __vimax3_u16x2(b0, b1, b2);
  b3 = __vimax3_u16x2(b3, b4, arc[15]);
  b0 = __vmaxu2(b0, b3);
  return (int)max(b0 & 0xFFFFu, b0 >> 16) - 257;
}

// Quick reject for 4 horizontally adjacent pixels (one byte each). An arc of >= 9 contiguous circle pixels contains two
// adjacent compass points — one of {N,S} and one of {E,W} — that differ from the centre by more than t in the same
// direction; the test below drops the "same direction" part (still a necessary condition) so that it runs on absolute
// differences: 4 x VABSDIFF4.U8 + per-byte "x > t" by carry (t < 128: ((x & 0x7f) + (127 - t)) | x has bit 7 set).
// Returns 0x80 in every byte whose pixel may be a corner.
SVO_D unsigned quick4(unsigned C, unsigned N, unsigned S, unsigned E, unsigned W, unsigned K) {
  const unsigned aN = __vabsdiffu4(N, C), aS = __vabsdiffu4(S, C), aE = __vabsdiffu4(E, C), aW = __vabsdiffu4(W, C);
  const unsigned ns = ((aN & 0x7F7F7F7Fu) + K) | aN | ((aS & 0x7F7F7F7Fu) + K) | aS;
  const unsigned ew = ((aE & 0x7F7F7F7Fu) + K) | aE | ((aW & 0x7F7F7F7Fu) + K) | aW;
  return ns & ew & 0x80808080u;
}
// Any threshold (t up to 254): the same test with the carry computed in 16-bit lanes (even and odd bytes separately).
SVO_D unsigned gt16(unsigned x, unsigned K2) {  // 0x80 per byte where byte > t, K2 = (255 - t) * 0x10001
  const unsigned e = ((x & 0x00FF00FFu) + K2) >> 1, o = (((x >> 8) & 0x00FF00FFu) + K2) << 7;
  return (e & 0x00800080u) | (o & 0x80008000u);
}
SVO_D unsigned quick4Wide(unsigned C, unsigned N, unsigned S, unsigned E, unsigned W, unsigned K2) {
  const unsigned aN = __vabsdiffu4(N, C), aS = __vabsdiffu4(S, C), aE = __vabsdiffu4(E, C), aW = __vabsdiffu4(W, C);
  return (gt16(aN, K2) | gt16(aS, K2)) & (gt16(aE, K2) | gt16(aW, K2));
}

// floor(x / d) for x * d < 2^32 as one multiply-high; magic = 0 stands for d == 1.
inline unsigned divMagic(int d) { return d <= 1 ? 0u : (unsigned)(0x100000000ull / (unsigned)d + 1ull); }
SVO_D int divFast(int x, unsigned magic) { return magic ? (int)__umulhi((unsigned)x, magic) : x; }

// TMA descriptors of the levels one launch walks (box = kSPitch x kSRows bytes: tile + halo)
struct FastMaps { alignas(64) unsigned char m[SVO_MAX_LEVELS][128]; };

struct FastParams {
  int min_level, max_level, threshold, border, cell_size, n_cols, n_cells, first;
  int tile_base[SVO_MAX_LEVELS + 1];  // first tile index of every level inside blockIdx.x
  int tiles_x[SVO_MAX_LEVELS];
  unsigned tiles_x_magic[SVO_MAX_LEVELS], cell_magic;
  unsigned long long* keys;        // [count][n_cells] or nullptr
  const uint8_t* occupancy;        // [count][n_cells] or nullptr
  short* score_map;                // dense debug maps for one frame (level coords) or nullptr
  uint8_t* nonmax_map;
};

template <int ARC>
// (A/B on the B200: minBlocks 6 = this, 7 / 8 force 32 registers and spill: 1.148 / 1.155 ms per 1024 frames)
__global__ void __launch_bounds__(kThreadsFast) fast_level_kernel(PyrView v, FastParams P, const __grid_constant__ FastMaps maps) {
  __shared__ __align__(128) uint8_t s_img[kSRows * kSPitch];
#ifndef SVO_FAST_STAGE_CPASYNC
  __shared__ __align__(8) unsigned long long s_full;
#endif
  // score tile with a one-pixel zero frame: the 3x3 non-max reads its 8 neighbours without edge tests
  // (B200, 1024 frames: 1.245 -> 1.162 ms; together with the flat candidate push 1.147 ms)
  constexpr int kSW = kTW + 2, kScoreN = ((kTH + 2) * kSW + 7) / 8 * 8;
#define SVO_SCORE_IDX(r, c) (((r) + 1) * kSW + (c) + 1)
  __shared__ __align__(16) short s_score[kScoreN];
  __shared__ unsigned short s_cand[kTH * kTW];
  __shared__ unsigned short s_corner[kTH * kTW];
  __shared__ int s_ncand, s_ncorner;
  const int tid = threadIdx.x;
  int L = P.min_level;
  while (L < P.max_level && (int)blockIdx.x >= P.tile_base[L + 1]) ++L;
  const int tile = blockIdx.x - P.tile_base[L];
  const int ty = divFast(tile, P.tiles_x_magic[L]), tx = tile - ty * P.tiles_x[L];
  const int cols = v.cols[L], rows = v.rows[L];
  const int frame_local = blockIdx.y;
  const int x0 = tx * kTW, y0 = ty * kTH;
  const int thr = P.threshold;

#ifndef SVO_FAST_STAGE_CPASYNC
  // stage tile + halo with ONE TMA box load issued by thread 0 (bytes outside the {pitch, rows} plane arrive as zeros); the score tile
  // is cleared while the copy is in flight. B200, pyramid + detect of 1024 frames: 1.147 ms with the 16-byte cp.async loop below, 1.083 ms so
  if (tid == 0) {
    const unsigned bar = smemAddr(&s_full);
    mbarInit(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbarExpectTx(bar, kSRows * kSPitch);
    tmaLoadTile(smemAddr(s_img), maps.m[L], x0 - kPadL, y0 - kHalo, P.first + frame_local, bar);
  }
  for (int i = tid; i < kScoreN * 2 / 16; i += kThreadsFast) reinterpret_cast<uint4*>(s_score)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { s_ncand = 0; s_ncorner = 0; }
  __syncthreads();  // barrier initialised, counters and score tile cleared
  mbarWait(smemAddr(&s_full), 0u);
#else
  const int pitch = v.pitch[L];
  const uint8_t* img = v.level(P.first + frame_local, L);
  // stage tile + halo with 16-byte cp.async (x0 - 16 and the row pitch are multiples of 16; chunks outside the image are zero-filled
  // through src-size 0); the score tile is cleared while the copies are in flight
  for (int i = tid; i < kSRows * (kSPitch / 16); i += kThreadsFast) {
    const int r = i / (kSPitch / 16), q = i - r * (kSPitch / 16);
    const int gy = y0 - kHalo + r, gx = x0 - kPadL + q * 16;
    const bool in = gy >= 0 && gy < rows && gx >= 0 && gx < pitch;
    const uint8_t* src = in ? img + (size_t)gy * pitch + gx : img;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_img[r * kSPitch + q * 16]);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");
  }
  for (int i = tid; i < kScoreN * 2 / 16; i += kThreadsFast) reinterpret_cast<uint4*>(s_score)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { s_ncand = 0; s_ncorner = 0; }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
#endif

  // stage 1: quick reject, one warp per tile row, 4 pixels per lane
  {
    const int g = tid & 31;
    const int gx0 = x0 + 4 * g;
    const int jlo = max(3 - gx0, 0), jhi = min(cols - 3 - gx0, 4);  // valid centres: 3 <= x < cols - 3
    unsigned colmask = 0u;
    if (jhi > jlo) colmask = (jhi >= 4 ? 0x80808080u : ((1u << (8 * jhi)) - 1u)) & ~((1u << (8 * jlo)) - 1u);
    const unsigned K = (unsigned)(127 - thr) * 0x01010101u, K2 = (unsigned)(255 - thr) * 0x00010001u;
#pragma unroll
    for (int rr = 0; rr < kTH / 8; ++rr) {
      const int r = (tid >> 5) + rr * 8;
      const int gy = y0 + r;
      const uint8_t* base = s_img + (r + kHalo) * kSPitch + kPadL + 4 * g;
      const unsigned wc = *reinterpret_cast<const unsigned*>(base);
      const unsigned wm = *reinterpret_cast<const unsigned*>(base - 4), wp = *reinterpret_cast<const unsigned*>(base + 4);
      const unsigned wn = *reinterpret_cast<const unsigned*>(base - 3 * kSPitch), ws = *reinterpret_cast<const unsigned*>(base + 3 * kSPitch);
      const unsigned we = __funnelshift_r(wc, wp, 24);  // pixels x+3 .. x+6
      const unsigned ww = __funnelshift_r(wm, wc, 8);   // pixels x-3 .. x
      unsigned f = thr < 128 ? quick4(wc, wn, ws, we, ww, K) : quick4Wide(wc, wn, ws, we, ww, K2);
      f &= (gy >= 3 && gy < rows - 3) ? colmask : 0u;
#if !defined(SVO_FAST_PUSH_WALK) && !defined(SVO_FAST_WARP_PUSH1)  // own atomic per lane, the (up to four) candidates stored by predicated stores
      if (f) {
        const unsigned nib = (((f >> 7) * 0x00204081u) >> 21) & 15u;  // bit j = pixel j of the word
        const int at = atomicAdd(&s_ncand, __popc(nib));
        const unsigned short id0 = (unsigned short)((r << 7) | (4 * g));
        if (nib & 1u) s_cand[at] = id0;
        if (nib & 2u) s_cand[at + (nib & 1u)] = id0 + 1;
        if (nib & 4u) s_cand[at + __popc(nib & 3u)] = id0 + 2;
        if (nib & 8u) s_cand[at + __popc(nib & 7u)] = id0 + 3;
      }
#elif !defined(SVO_FAST_WARP_PUSH1)  // the same with a walk over the set bits (1.162 vs 1.147 ms per 1024 frames)
      if (f) {
        int at = atomicAdd(&s_ncand, __popc(f));
        const int id0 = (r << 7) | (4 * g);
        do {
          const int b = __ffs(f) - 1;  // bit 8j+7
          f &= f - 1;
          s_cand[at++] = (unsigned short)(id0 + (b >> 3));
        } while (f);
      }
#else
      {  // warp-aggregated push of the row's candidates (four ballots, ONE atomic per warp and row). Measured on the B200: SLOWER
         // (1.397 vs 1.244 ms per 1024 frames together with the aggregated corner push): 12 % of the pixels are candidates, so most
         // lanes have nothing to push and the ballots / popcounts are paid by all of them.
        const unsigned b0 = __ballot_sync(0xffffffffu, f & 0x80u), b1 = __ballot_sync(0xffffffffu, f & 0x8000u);
        const unsigned b2 = __ballot_sync(0xffffffffu, f & 0x800000u), b3 = __ballot_sync(0xffffffffu, f & 0x80000000u);
        if (b0 | b1 | b2 | b3) {
          const int n0 = __popc(b0), n1 = n0 + __popc(b1), n2 = n1 + __popc(b2);
          int base = 0;
          if (g == 0) base = atomicAdd(&s_ncand, n2 + __popc(b3));
          base = __shfl_sync(0xffffffffu, base, 0);
          const unsigned lt = (1u << g) - 1u;
          const int id0 = (r << 7) | (4 * g);
          if (f & 0x80u) s_cand[base + __popc(b0 & lt)] = (unsigned short)id0;
          if (f & 0x8000u) s_cand[base + n0 + __popc(b1 & lt)] = (unsigned short)(id0 + 1);
          if (f & 0x800000u) s_cand[base + n1 + __popc(b2 & lt)] = (unsigned short)(id0 + 2);
          if (f & 0x80000000u) s_cand[base + n2 + __popc(b3 & lt)] = (unsigned short)(id0 + 3);
        }
      }
#endif
    }
  }
  __syncthreads();

  // stage 2: exact margin on the candidates
  const int ncand = s_ncand;
#ifndef SVO_FAST_WARP_PUSH2
  for (int i = tid; i < ncand; i += kThreadsFast) {
    const int idx = s_cand[i];
    const int r = idx >> 7, c = idx & 127;
    const int m = fastMargin<ARC>(&s_img[(r + kHalo) * kSPitch + kPadL + c], kSPitch);
    if (m >= thr) {  // score = max(threshold, margin) = margin; threshold >= 1 so 0 means "no corner"
      s_score[SVO_SCORE_IDX(r, c)] = (short)m;
      s_corner[atomicAdd(&s_ncorner, 1)] = (unsigned short)idx;
    }
  }
#else
  // warp-uniform trip count, the corner list filled with one atomic per warp. Measured on the B200: slower as well (1.289 vs 1.245 ms per
  // 1024 frames): a quarter of the candidates are corners, the per-lane atomics were never the cost.
  for (int i0 = 0; i0 < ncand; i0 += kThreadsFast) {
    const int i = i0 + tid;
    int idx = 0, m = -1;
    if (i < ncand) {
      idx = s_cand[i];
      const int r = idx >> 7, c = idx & 127;
      m = fastMargin<ARC>(&s_img[(r + kHalo) * kSPitch + kPadL + c], kSPitch);
    }
    const bool corner = m >= thr;  // score = max(threshold, margin) = margin; threshold >= 1 so 0 means "no corner"
    const unsigned bal = __ballot_sync(0xffffffffu, corner);
    if (bal) {
      int base = 0;
      if ((tid & 31) == 0) base = atomicAdd(&s_ncorner, __popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (corner) {
        s_score[SVO_SCORE_IDX(idx >> 7, idx & 127)] = (short)m;
        s_corner[base + __popc(bal & ((1u << (tid & 31)) - 1u))] = (unsigned short)idx;
      }
    }
  }
#endif
  __syncthreads();

  if (P.score_map || P.nonmax_map) {  // dense debug maps of this tile (parity tests of the raw stages)
    for (int i = tid; i < kTH * kTW; i += kThreadsFast) {
      const int gy = y0 + (i >> 7), gx = x0 + (i & 127);
      if (gx >= cols || gy >= rows) continue;
      if (P.score_map) P.score_map[(size_t)gy * cols + gx] = s_score[SVO_SCORE_IDX(i >> 7, i & 127)];
      if (P.nonmax_map) P.nonmax_map[(size_t)gy * cols + gx] = 0;
    }
    __syncthreads();
  }

  // stage 3: 3x3 non-max on the corners, border, cell arg-max
  const int ncorner = s_ncorner;
  for (int i = tid; i < ncorner; i += kThreadsFast) {
    const int idx = s_corner[i];
    const int r = idx >> 7, c = idx & 127;
    const int gy = y0 + r, gx = x0 + c;
    const int si = SVO_SCORE_IDX(r, c);
    const int sc = s_score[si];
    const bool edge = !(r > 0 && r < kTH - 1 && c > 0 && c < kTW - 1);
    bool keep = true;
#pragma unroll
    for (int n = 0; n < 9; ++n) {  // neighbours inside the tile
      if (n == 4) continue;
      const int dr = n / 3 - 1, dc = n % 3 - 1;
      if (s_score[si + dr * kSW + dc] >= sc) keep = false;  // the frame holds 0 < threshold <= sc
    }
    if (keep && edge) {  // rare: neighbours that belong to other tiles are scored from the halo
#pragma unroll 1
      for (int n = 0; n < 9 && keep; ++n) {
        const int dr = n / 3 - 1, dc = n - (n / 3) * 3 - 1;
        const int rr = r + dr, cc = c + dc;
        if ((unsigned)rr < (unsigned)kTH && (unsigned)cc < (unsigned)kTW) continue;
        const int ny = gy + dr, nx = gx + dc;
        if (nx >= 3 && ny >= 3 && nx < cols - 3 && ny < rows - 3 &&
            fastMargin<ARC>(&s_img[(rr + kHalo) * kSPitch + kPadL + cc], kSPitch) >= sc) keep = false;
      }
    }
    if (!keep) continue;
    if (P.nonmax_map) P.nonmax_map[(size_t)gy * cols + gx] = 1;
    if (!P.keys) continue;
    if (gx < P.border || gy < P.border || gx >= cols - P.border || gy >= rows - P.border) continue;
    const int k = divFast(gy << L, P.cell_magic) * P.n_cols + divFast(gx << L, P.cell_magic);
    if (P.occupancy && P.occupancy[(size_t)frame_local * P.n_cells + k]) continue;
    const unsigned order = ((unsigned)L << 28) | ((unsigned)gy << 14) | (unsigned)gx;
    const unsigned long long key = ((unsigned long long)(unsigned)sc << 32) | (unsigned long long)(0xFFFFFFFFu - order);
    atomicMax(&P.keys[(size_t)frame_local * P.n_cells + k], key);
  }
}

__global__ void fast_keys_init_kernel(unsigned long long* keys, size_t n, int threshold) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((unsigned long long)(unsigned)threshold << 32) | 0xFFFFFFFFull;
}

__global__ void fast_keys_decode_kernel(const unsigned long long* keys, size_t n, int threshold, svo_corner* out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  const int sc = (int)(key >> 32);
  svo_corner c;
  if (sc > threshold) {
    const unsigned order = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
    const int L = order >> 28, y = (order >> 14) & 0x3FFF, x = order & 0x3FFF;
    c.x = x << L; c.y = y << L; c.level = L; c.score = (float)sc; c.angle = 0.0f;
  } else {
    c.x = 0; c.y = 0; c.level = 0; c.score = (float)threshold; c.angle = 0.0f;
  }
  out[i] = c;
}

int launchLevels(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, const PyrView& v, FastParams& P, int arc, int count) {
  FastMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int L = P.min_level; L <= P.max_level; ++L) {
    const int rc = svoEnsureLevelMap(ctx, pyr, L, kSPitch, kSRows, pyr->tmap_fast[L], &pyr->tmap_fast_ready[L]);
    if (rc != SVO_OK) return rc;
    memcpy(maps.m[L], pyr->tmap_fast[L], 128);
  }
  int n_tiles = 0;
  for (int L = 0; L <= SVO_MAX_LEVELS; ++L) P.tile_base[L] = 0;
  P.cell_magic = divMagic(P.cell_size);
  for (int L = P.min_level; L <= P.max_level; ++L) {
    P.tile_base[L] = n_tiles;
    P.tiles_x[L] = (v.cols[L] + kTW - 1) / kTW;
    P.tiles_x_magic[L] = divMagic(P.tiles_x[L]);
    n_tiles += P.tiles_x[L] * ((v.rows[L] + kTH - 1) / kTH);
    P.tile_base[L + 1] = n_tiles;
  }
  for (int f0 = 0; f0 < count; f0 += 65535) {  // grid.y limit
    FastParams Q = P;
    const int n = min(65535, count - f0);
    Q.first = P.first + f0;
    if (Q.keys) Q.keys += (size_t)f0 * P.n_cells;
    if (Q.occupancy) Q.occupancy += (size_t)f0 * P.n_cells;
    dim3 grid(n_tiles, n, 1);
    if (arc == 9) fast_level_kernel<9><<<grid, kThreadsFast, 0, ctx->stream>>>(v, Q, maps);
    else fast_level_kernel<10><<<grid, kThreadsFast, 0, ctx->stream>>>(v, Q, maps);
    SVO_LAUNCH_CHECK(ctx);
  }
  return SVO_OK;
}


// ---- corner lists in raster order (the shapes of fast.h:32-41: std::vector<fast_xy> + scores + indices of the maxima) -------------
// Ordered compaction of the flagged elements of a flat index space in chunks of 1024: count per chunk, scan of the chunk counts
// (one CTA), scatter with an in-chunk scan. 256 threads, 4 consecutive elements per thread, so the output keeps the input order.
constexpr int kChunk = 1024;

template <class T>
__global__ void __launch_bounds__(256) flag_count_kernel(const T* flags, int n, int* chunk_count) {
  const int base = blockIdx.x * kChunk + 4 * threadIdx.x;
  int c = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) c += (base + j < n && flags[base + j] != 0) ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int s_w[8];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) chunk_count[blockIdx.x] = ((s_w[0] + s_w[1]) + (s_w[2] + s_w[3])) + ((s_w[4] + s_w[5]) + (s_w[6] + s_w[7]));
}

// exclusive scan of chunk_count in place; total -> *total
__global__ void __launch_bounds__(1024) chunk_scan_kernel(int* chunk_count, int n_chunks, int* total) {
  __shared__ int s_w[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_chunks; base += 1024) {
    const int i = base + tid;
    const int v = i < n_chunks ? chunk_count[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      s_w[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int carry = s_carry;
    const int off = carry + (warp ? s_w[warp - 1] : 0) + inc - v;
    if (i < n_chunks) chunk_count[i] = off;
    __syncthreads();
    if (tid == 1023) s_carry = carry + s_w[31];
    __syncthreads();
  }
  if (tid == 0) *total = s_carry;
}

// in-chunk exclusive offset of this thread's first flagged element (block of 256, `c` flagged elements in this thread)
SVO_D int chunkOffset(int c, int* s_w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  int off = inc - c;
  for (int w = 0; w < warp; ++w) off += s_w[w];
  return off;
}

// dense maps of one level -> (x, y), score, non-max flag of every corner in raster order
__global__ void __launch_bounds__(256) corner_scatter_kernel(const short* score_map, const uint8_t* nonmax_map, int n, int cols,
                                                             const int* chunk_off, int max_corners, svo_fast_xy* xy, int* scores,
                                                             uint8_t* nonmax) {
  __shared__ int s_w[8];
  const int base = blockIdx.x * kChunk + 4 * threadIdx.x;
  short sc[4];
  int c = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) { sc[j] = base + j < n ? score_map[base + j] : (short)0; c += sc[j] != 0; }
  int at = chunk_off[blockIdx.x] + chunkOffset(c, s_w);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (sc[j] == 0) continue;
    if (at < max_corners) {
      const int i = base + j, y = i / cols;
      if (xy) xy[at] = svo_fast_xy{(short)(i - y * cols), (short)y};
      if (scores) scores[at] = sc[j];
      if (nonmax) nonmax[at] = nonmax_map[i];
    }
    ++at;
  }
}

// indices of the flagged list entries, in order (fast_nonmax_3x3's output)
__global__ void __launch_bounds__(256) index_scatter_kernel(const uint8_t* flags, int n, const int* chunk_off, int* idx_out) {
  __shared__ int s_w[8];
  const int base = blockIdx.x * kChunk + 4 * threadIdx.x;
  bool f[4];
  int c = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) { f[j] = base + j < n && flags[base + j] != 0; c += f[j]; }
  int at = chunk_off[blockIdx.x] + chunkOffset(c, s_w);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (f[j]) idx_out[at++] = base + j;
}

// fast_corner_score_10 of listed pixels (fast_10_score.cpp:3150-3178): the largest barrier at which the pixel is still a corner,
// `threshold` itself for a pixel that is no corner above it (the reference's search starts at threshold + 1 and returns b - 1).
template <int ARC>
__global__ void corner_score_kernel(const uint8_t* img, int cols, int rows, int pitch, const svo_fast_xy* xy, int n, int threshold, int* scores) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = xy[i].x, y = xy[i].y;
  int sc = threshold;
  if (x >= 3 && y >= 3 && x < cols - 3 && y < rows - 3) sc = max(threshold, fastMargin<ARC>(img + (size_t)y * pitch + x, pitch));
  scores[i] = sc;
}

// fast_nonmax_3x3 on a raster-ordered corner list (nonmax_3x3.cpp:17-112): an entry survives unless one of its 8 neighbours is in the
// list with a score >= its own. The list is sorted by (y, x), so a neighbour is found by binary search.
SVO_D int findCorner(const svo_fast_xy* xy, int n, int x, int y) {
  const int key = (y << 16) | (x & 0xFFFF);
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int k = ((int)xy[mid].y << 16) | ((int)xy[mid].x & 0xFFFF);
    if (k < key) lo = mid + 1; else hi = mid;
  }
  return (lo < n && xy[lo].x == x && xy[lo].y == y) ? lo : -1;
}
__global__ void list_nonmax_kernel(const svo_fast_xy* xy, const int* scores, int n, uint8_t* keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = xy[i].x, y = xy[i].y, sc = scores[i];
  bool k = true;
  // left / right neighbours are the adjacent list entries
  if (i > 0 && xy[i - 1].x == x - 1 && xy[i - 1].y == y && scores[i - 1] >= sc) k = false;
  if (i < n - 1 && xy[i + 1].x == x + 1 && xy[i + 1].y == y && scores[i + 1] >= sc) k = false;
  for (int dy = -1; dy <= 1 && k; dy += 2) {
    if (y + dy < 0) continue;
    for (int dx = -1; dx <= 1 && k; ++dx) {
      if (x + dx < 0) continue;
      const int j = findCorner(xy, n, x + dx, y + dy);
      if (j >= 0 && scores[j] >= sc) k = false;
    }
  }
  keep[i] = k ? 1 : 0;
}

}  // namespace

int svoPyrBuildLaunch(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr, int first, int count);

int svoFastDetectImpl(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, const svo_detector_options* opt,
                          const uint8_t* occupancy_in, svo_corner* corners_out, svo_mem mem) {
  if (!ctx || !pyr || !opt || !corners_out || first < 0 || count < 0 || first + count > pyr->n_frames)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_fast_detect: bad arguments");
  if (opt->min_level < 0 || opt->max_level < opt->min_level || opt->max_level >= pyr->n_levels || opt->cell_size <= 0 ||
      opt->threshold < 1 || opt->threshold > 254 || opt->max_level > 15 || pyr->cols[0] >= 16384 || pyr->rows[0] >= 16384)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_fast_detect: bad detector options");
  const int arc = opt->arc_length == 9 ? 9 : 10;
  if (count == 0) return SVO_OK;
  SVO_BIND(ctx);
  int n_cols, n_rows;
  const int n_cells = svo_cuda_grid_cells(pyr->cols[0], pyr->rows[0], opt->cell_size, &n_cols, &n_rows);
  const size_t n = (size_t)n_cells * count;
  Stager st(ctx, mem);
  const uint8_t* d_occ = st.in(occupancy_in, n);
  svo_corner* d_out = st.out(corners_out, n);
  unsigned long long* keys = (unsigned long long*)st.scratch(n * sizeof(unsigned long long));
  if (!st.send()) return st.finish();
  fast_keys_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(keys, n, opt->threshold);
  SVO_LAUNCH_CHECK(ctx);
  const PyrView v = makeView(pyr);
  FastParams P;
  P.min_level = opt->min_level; P.max_level = opt->max_level; P.threshold = opt->threshold; P.border = opt->border;
  P.cell_size = opt->cell_size; P.n_cols = n_cols; P.n_cells = n_cells; P.first = first;
  P.keys = keys; P.occupancy = d_occ; P.score_map = nullptr; P.nonmax_map = nullptr;
  const int rc = launchLevels(ctx, pyr, v, P, arc, count);
  if (rc != SVO_OK) return rc;
  fast_keys_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(keys, n, opt->threshold, d_out);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

extern "C" {

int svo_cuda_fast_detect(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, const svo_detector_options* opt,
                         const uint8_t* occupancy_in, svo_corner* corners_out, svo_mem mem) {
  return svoFastDetectImpl(ctx, pyr, first, count, opt, occupancy_in, corners_out, mem);
}

int svo_cuda_pyramid_fast_detect(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr, int first, int count, const svo_detector_options* opt,
                                 const uint8_t* occupancy_in, svo_corner* corners_out, svo_mem mem) {
  if (!ctx || !pyr || first < 0 || count < 0 || first + count > pyr->n_frames)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_pyramid_fast_detect: bad arguments");
  const int rc = svoPyrBuildLaunch(ctx, pyr, first, count);
  if (rc != SVO_OK) return rc;
  return svoFastDetectImpl(ctx, pyr, first, count, opt, occupancy_in, corners_out, mem);
}

int svo_cuda_fast_level_maps(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int frame, int level, int threshold, int arc_length,
                             int16_t* score_map, uint8_t* nonmax_map, svo_mem mem) {
  if (!ctx || !pyr || frame < 0 || frame >= pyr->n_frames || level < 0 || level >= pyr->n_levels || threshold < 1 || threshold > 254)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_fast_level_maps: bad arguments");
  const size_t n = (size_t)pyr->cols[level] * pyr->rows[level];
  SVO_BIND(ctx);
  Stager st(ctx, mem);
  int16_t* d_sc = st.out(score_map, n);
  uint8_t* d_nm = st.out(nonmax_map, n);
  if (!st.send()) return st.finish();
  FastParams P;
  P.min_level = level; P.max_level = level; P.threshold = threshold; P.border = 0; P.cell_size = 1; P.n_cols = 1; P.n_cells = 1;
  P.first = frame; P.keys = nullptr; P.occupancy = nullptr; P.score_map = d_sc; P.nonmax_map = d_nm;
  const int rc = launchLevels(ctx, pyr, makeView(pyr), P, arc_length == 9 ? 9 : 10, 1);
  if (rc != SVO_OK) return rc;
  return st.finish();
}

int svo_cuda_fast_corner_list(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int frame, int level, int threshold, int arc_length,
                              int max_corners, svo_fast_xy* xy_out, int* scores_out, uint8_t* nonmax_out, int* n_out, svo_mem mem) {
  if (!ctx || !pyr || !n_out || frame < 0 || frame >= pyr->n_frames || level < 0 || level >= pyr->n_levels || threshold < 1 ||
      threshold > 254 || max_corners < 0 || pyr->cols[level] >= 32768 || pyr->rows[level] >= 32768)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_fast_corner_list: bad arguments");
  *n_out = 0;
  SVO_BIND(ctx);
  const int cols = pyr->cols[level], rows = pyr->rows[level];
  const size_t n = (size_t)cols * rows;
  const int n_chunks = (int)((n + kChunk - 1) / kChunk);
  Stager st(ctx, SVO_MEM_DEVICE);  // outputs are copied back by hand: only the corners found, not max_corners entries
  short* d_sc = (short*)st.scratch(n * sizeof(short));
  uint8_t* d_nm = (uint8_t*)st.scratch(n);
  int* d_chunk = (int*)st.scratch((size_t)(n_chunks + 1) * sizeof(int));
  const bool host = mem == SVO_MEM_HOST;
  const size_t cap = (size_t)(max_corners > 0 ? max_corners : 1);
  svo_fast_xy* d_xy = xy_out ? (host ? (svo_fast_xy*)st.scratch(cap * sizeof(svo_fast_xy)) : xy_out) : nullptr;
  int* d_scores = scores_out ? (host ? (int*)st.scratch(cap * sizeof(int)) : scores_out) : nullptr;
  uint8_t* d_nonmax = nonmax_out ? (host ? (uint8_t*)st.scratch(cap) : nonmax_out) : nullptr;
  if (!st.send() || !d_sc || !d_nm || !d_chunk) return SVO_FAIL(ctx, SVO_ERR_OUT_OF_MEMORY, "svo_cuda_fast_corner_list: scratch allocation failed");
  FastParams P;
  P.min_level = level; P.max_level = level; P.threshold = threshold; P.border = 0; P.cell_size = 1; P.n_cols = 1; P.n_cells = 1;
  P.first = frame; P.keys = nullptr; P.occupancy = nullptr; P.score_map = d_sc; P.nonmax_map = d_nm;
  const int rc = launchLevels(ctx, pyr, makeView(pyr), P, arc_length == 9 ? 9 : 10, 1);
  if (rc != SVO_OK) return rc;
  flag_count_kernel<short><<<n_chunks, 256, 0, ctx->stream>>>(d_sc, (int)n, d_chunk);
  SVO_LAUNCH_CHECK(ctx);
  chunk_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_chunk, n_chunks, d_chunk + n_chunks);
  SVO_LAUNCH_CHECK(ctx);
  if (max_corners > 0 && (d_xy || d_scores || d_nonmax)) {
    corner_scatter_kernel<<<n_chunks, 256, 0, ctx->stream>>>(d_sc, d_nm, (int)n, cols, d_chunk, max_corners, d_xy, d_scores, d_nonmax);
    SVO_LAUNCH_CHECK(ctx);
  }
  int total = 0;
  SVO_CUDA_TRY(ctx, cudaMemcpyAsync(&total, d_chunk + n_chunks, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SVO_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *n_out = total;  // may exceed max_corners: the lists then hold the first max_corners corners
  const size_t m = (size_t)min(total, max_corners);
  if (host && m) {
    if (xy_out) SVO_CUDA_TRY(ctx, cudaMemcpyAsync(xy_out, d_xy, m * sizeof(svo_fast_xy), cudaMemcpyDeviceToHost, ctx->stream));
    if (scores_out) SVO_CUDA_TRY(ctx, cudaMemcpyAsync(scores_out, d_scores, m * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (nonmax_out) SVO_CUDA_TRY(ctx, cudaMemcpyAsync(nonmax_out, d_nonmax, m, cudaMemcpyDeviceToHost, ctx->stream));
    SVO_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return st.finish();
}

int svo_cuda_fast_corner_score(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int frame, int level, int n, const svo_fast_xy* xy,
                               int threshold, int arc_length, int* scores_out, svo_mem mem) {
  if (!ctx || !pyr || frame < 0 || frame >= pyr->n_frames || level < 0 || level >= pyr->n_levels || n < 0 || (n && (!xy || !scores_out)) ||
      threshold < 0 || threshold > 254)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_fast_corner_score: bad arguments");
  if (n == 0) return SVO_OK;
  SVO_BIND(ctx);
  Stager st(ctx, mem);
  const svo_fast_xy* d_xy = st.in(xy, (size_t)n);
  int* d_sc = st.out(scores_out, (size_t)n);
  if (!st.send()) return st.finish();
  const uint8_t* img = pyr->data[level] + pyr->frame_stride[level] * (size_t)frame;
  if (arc_length == 9) corner_score_kernel<9><<<(n + 127) / 128, 128, 0, ctx->stream>>>(img, pyr->cols[level], pyr->rows[level], (int)pyr->pitch[level], d_xy, n, threshold, d_sc);
  else corner_score_kernel<10><<<(n + 127) / 128, 128, 0, ctx->stream>>>(img, pyr->cols[level], pyr->rows[level], (int)pyr->pitch[level], d_xy, n, threshold, d_sc);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}

int svo_cuda_fast_nonmax_3x3(svo_cuda_ctx* ctx, int n, const svo_fast_xy* xy, const int* scores, int* nonmax_idx_out, int* n_out,
                             svo_mem mem) {
  if (!ctx || !n_out || n < 0 || (n && (!xy || !scores || !nonmax_idx_out)))
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_fast_nonmax_3x3: bad arguments");
  *n_out = 0;
  if (n == 0) return SVO_OK;
  SVO_BIND(ctx);
  const int n_chunks = (n + kChunk - 1) / kChunk;
  Stager st(ctx, mem);
  const svo_fast_xy* d_xy = st.in(xy, (size_t)n);
  const int* d_sc = st.in(scores, (size_t)n);
  const bool host = mem == SVO_MEM_HOST;
  int* d_idx = host ? (int*)st.scratch((size_t)n * sizeof(int)) : nonmax_idx_out;
  uint8_t* d_keep = (uint8_t*)st.scratch((size_t)n);
  int* d_chunk = (int*)st.scratch((size_t)(n_chunks + 1) * sizeof(int));
  if (!st.send() || !d_idx || !d_keep || !d_chunk) return SVO_FAIL(ctx, SVO_ERR_OUT_OF_MEMORY, "svo_cuda_fast_nonmax_3x3: scratch allocation failed");
  list_nonmax_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d_xy, d_sc, n, d_keep);
  SVO_LAUNCH_CHECK(ctx);
  flag_count_kernel<uint8_t><<<n_chunks, 256, 0, ctx->stream>>>(d_keep, n, d_chunk);
  SVO_LAUNCH_CHECK(ctx);
  chunk_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_chunk, n_chunks, d_chunk + n_chunks);
  SVO_LAUNCH_CHECK(ctx);
  index_scatter_kernel<<<n_chunks, 256, 0, ctx->stream>>>(d_keep, n, d_chunk, d_idx);
  SVO_LAUNCH_CHECK(ctx);
  int total = 0;
  SVO_CUDA_TRY(ctx, cudaMemcpyAsync(&total, d_chunk + n_chunks, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SVO_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *n_out = total;
  if (host && total) SVO_CUDA_TRY(ctx, cudaMemcpyAsync(nonmax_idx_out, d_idx, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  return st.finish();
}

}  // extern "C"
