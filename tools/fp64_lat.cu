// Micro-benchmark (B200): latency of dependent FP64 operations and FP64 issue rate per SM sub-partition, the two numbers the
// sparse-alignment and filter kernels are bound by. nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_lat tools/fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dep_chain(double* out, long long* cyc, double a, double b, int n) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x = fma(x, b, a);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void dep_chain_f32(float* out, long long* cyc, float a, float b, int n) {
  float x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x = fmaf(x, b, a);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void indep(double* out, long long* cyc, double a, double b, int n) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = a + threadIdx.x + k;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], b, a);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void shfl_chain(double* out, long long* cyc, int n) {
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x += __shfl_xor_sync(0xffffffffu, x, 1 + (k & 3));
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void rcp_chain(double* out, long long* cyc, double a, int n) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) x = 1.0 / x + a;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lds_chain(double* out, long long* cyc, int n) {
  __shared__ double s[256];
  s[threadIdx.x] = (double)((threadIdx.x * 7 + 3) & 255);
  __syncthreads();
  int idx = threadIdx.x;
  double acc = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) { const double v = s[idx]; idx = (int)v; acc += v; }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double* out; long long* cyc; float* outf;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 1 << 16); cudaMalloc(&outf, 1 << 24);
  long long h[1024];
  const int n = 256;
  auto report = [&](const char* name, int blocks, int threads, double ops_per_thread) {
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < blocks; ++i) m += (double)h[i]; m /= blocks;
    printf("%-44s blocks=%4d threads=%4d  cycles/op per thread = %7.2f   warp-instr per clk per SM = %.3f\n", name, blocks, threads, m / ops_per_thread,
           ops_per_thread * (threads / 32.0) / m);
  };
  dep_chain<<<1, 32>>>(out, cyc, 1.0, 0.999, n); report("DFMA dependent chain, 1 warp", 1, 32, 16.0 * n);
  dep_chain_f32<<<1, 32>>>(outf, cyc, 1.0f, 0.999f, n); report("FFMA dependent chain, 1 warp", 1, 32, 16.0 * n);
  dep_chain<<<1, 128>>>(out, cyc, 1.0, 0.999, n); report("DFMA dependent chain, 4 warps (1 / SMSP)", 1, 128, 16.0 * n);
  dep_chain<<<1, 256>>>(out, cyc, 1.0, 0.999, n); report("DFMA dependent chain, 8 warps (2 / SMSP)", 1, 256, 16.0 * n);
  dep_chain<<<1, 512>>>(out, cyc, 1.0, 0.999, n); report("DFMA dependent chain, 16 warps (4 / SMSP)", 1, 512, 16.0 * n);
  dep_chain<<<1, 1024>>>(out, cyc, 1.0, 0.999, n); report("DFMA dependent chain, 32 warps (8 / SMSP)", 1, 1024, 16.0 * n);
  indep<4><<<1, 32>>>(out, cyc, 1.0, 0.999, n); report("DFMA 4 independent chains, 1 warp", 1, 32, 16.0 * n);
  indep<8><<<1, 32>>>(out, cyc, 1.0, 0.999, n); report("DFMA 8 independent chains, 1 warp", 1, 32, 32.0 * n);
  indep<8><<<1, 128>>>(out, cyc, 1.0, 0.999, n); report("DFMA 8 independent chains, 4 warps", 1, 128, 32.0 * n);
  indep<8><<<1, 512>>>(out, cyc, 1.0, 0.999, n); report("DFMA 8 independent chains, 16 warps", 1, 512, 32.0 * n);
  indep<8><<<148 * 2, 512>>>(out, cyc, 1.0, 0.999, n); report("DFMA 8 indep chains, 2 x 16 warps on every SM", 296, 512, 32.0 * n);
  shfl_chain<<<1, 32>>>(out, cyc, n); report("64-bit SHFL + DADD dependent chain, 1 warp", 1, 32, 16.0 * n);
  rcp_chain<<<1, 32>>>(out, cyc, 1.5, n); report("IEEE 1/x + DADD dependent chain, 1 warp", 1, 32, 4.0 * n);
  lds_chain<<<1, 32>>>(out, cyc, n); report("LDS.64 -> F2I -> LDS pointer chase, 1 warp", 1, 32, 16.0 * n);
  return 0;
}
