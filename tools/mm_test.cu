// debug: which min/max fusion pattern miscompiles on sm_100a (nvcc 12.9)?
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
#define HD __host__ __device__
HD int f_min3(const int* d) { return min(min(d[0], d[1]), d[2]); }
HD int f_max3(const int* d) { return max(max(d[0], d[1]), d[2]); }
HD int f_maxneg(const int* d) { return max(d[0], max(d[1], -d[2])); }
HD int f_minmax(const int* d) { return max(min(d[0], d[1]), min(d[2], d[3])); }
HD int f_slide(const int* d) { int mn2[16]; for (int k=0;k<16;++k) mn2[k]=min(d[k],d[(k+1)&15]); int mn4[16]; for (int k=0;k<16;++k) mn4[k]=min(mn2[k],mn2[(k+2)&15]); int b=-999; for (int k=0;k<16;++k) b=max(b,min(min(mn4[k],mn4[(k+4)&15]),mn2[(k+8)&15])); return b; }
HD int f_slide_max(const int* d) { int mx2[16]; for (int k=0;k<16;++k) mx2[k]=max(d[k],d[(k+1)&15]); int mx4[16]; for (int k=0;k<16;++k) mx4[k]=max(mx2[k],mx2[(k+2)&15]); int b=999; for (int k=0;k<16;++k) b=min(b,max(max(mx4[k],mx4[(k+4)&15]),mx2[(k+8)&15])); return b; }
typedef int (*fn)(const int*);
template <int W> __global__ void k(const int* in, int n, int* out) { int i = blockIdx.x*blockDim.x+threadIdx.x; if (i>=n) return; const int* d = in + 16*i;
  int r; if (W==0) r=f_min3(d); else if (W==1) r=f_max3(d); else if (W==2) r=f_maxneg(d); else if (W==3) r=f_minmax(d); else if (W==4) r=f_slide(d); else r=f_slide_max(d); out[i]=r; }
int main() { const int n = 50000; int* h=(int*)malloc(n*64); for (int i=0;i<n*16;++i) h[i]=rand()%511-255; int* d; int* o; cudaMalloc(&d,n*64); cudaMalloc(&o,n*4); cudaMemcpy(d,h,n*64,cudaMemcpyHostToDevice); int* ho=(int*)malloc(n*4);
  fn fs[6]={f_min3,f_max3,f_maxneg,f_minmax,f_slide,f_slide_max}; const char* names[6]={"min3","max3","maxneg","minmax","slide_min","slide_max"};
  for (int w=0;w<6;++w){ switch(w){case 0:k<0><<<(n+255)/256,256>>>(d,n,o);break;case 1:k<1><<<(n+255)/256,256>>>(d,n,o);break;case 2:k<2><<<(n+255)/256,256>>>(d,n,o);break;case 3:k<3><<<(n+255)/256,256>>>(d,n,o);break;case 4:k<4><<<(n+255)/256,256>>>(d,n,o);break;default:k<5><<<(n+255)/256,256>>>(d,n,o);}
    cudaMemcpy(ho,o,n*4,cudaMemcpyDeviceToHost); int bad=0; for(int i=0;i<n;++i){ int e=fs[w](h+16*i); if(e!=ho[i]){ if(bad<2) printf("  %s i=%d dev=%d host=%d\n",names[w],i,ho[i],e); ++bad; } } printf("%s bad=%d\n",names[w],bad); }
  return 0; }
