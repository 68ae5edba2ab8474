// debug: device fastMargin vs host on random windows
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
template <int ARC>
__host__ __device__ int fastMargin(const uint8_t* p, int pitch) {
  const int c = p[0];
  int d[16];
  d[0] = p[3 * pitch] - c;       d[1] = p[3 * pitch + 1] - c;   d[2] = p[2 * pitch + 2] - c;   d[3] = p[pitch + 3] - c;
  d[4] = p[3] - c;               d[5] = p[-pitch + 3] - c;      d[6] = p[-2 * pitch + 2] - c;  d[7] = p[-3 * pitch + 1] - c;
  d[8] = p[-3 * pitch] - c;      d[9] = p[-3 * pitch - 1] - c;  d[10] = p[-2 * pitch - 2] - c; d[11] = p[-pitch - 3] - c;
  d[12] = p[-3] - c;             d[13] = p[pitch - 3] - c;      d[14] = p[2 * pitch - 2] - c;  d[15] = p[3 * pitch - 1] - c;
  int mn2[16], mx2[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { mn2[k] = min(d[k], d[(k + 1) & 15]); mx2[k] = max(d[k], d[(k + 1) & 15]); }
  int mn4[16], mx4[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { mn4[k] = min(mn2[k], mn2[(k + 2) & 15]); mx4[k] = max(mx2[k], mx2[(k + 2) & 15]); }
  int best = -256;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int mn8 = min(mn4[k], mn4[(k + 4) & 15]);
    const int mx8 = max(mx4[k], mx4[(k + 4) & 15]);
    int mn, mx;
    if (ARC == 10) { mn = min(mn8, mn2[(k + 8) & 15]); mx = max(mx8, mx2[(k + 8) & 15]); }
    else           { mn = min(mn8, d[(k + 8) & 15]);   mx = max(mx8, d[(k + 8) & 15]); }
    best = max(best, max(mn, -mx));
  }
  return best - 1;
}
__global__ void k(const uint8_t* w, int n, int* out) { int i = blockIdx.x*blockDim.x+threadIdx.x; if (i<n) out[i] = fastMargin<10>(w + i*81 + 40, 9); }
int main() { const int n = 100000; uint8_t* h = (uint8_t*)malloc(n*81); for (int i=0;i<n*81;++i) h[i] = (i/81)%2 ? 100 + rand()%40 : rand()%256;
  uint8_t* d; int* o; cudaMalloc(&d, n*81); cudaMalloc(&o, n*4); cudaMemcpy(d, h, n*81, cudaMemcpyHostToDevice);
  k<<<(n+255)/256,256>>>(d, n, o); int* ho = (int*)malloc(n*4); cudaMemcpy(ho, o, n*4, cudaMemcpyDeviceToHost);
  int bad = 0; for (int i=0;i<n;++i) { int e = fastMargin<10>(h + i*81 + 40, 9); if (e != ho[i]) { if (bad < 5) printf("i=%d dev=%d host=%d\n", i, ho[i], e); ++bad; } }
  printf("bad=%d err=%s\n", bad, cudaGetErrorString(cudaGetLastError())); return 0; }
