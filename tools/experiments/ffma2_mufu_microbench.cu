#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned long long pk(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(unsigned long long v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c){ unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ float fma1(float a, float b, float c){ float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;":"=f"(r):"f"(a),"f"(b),"f"(c)); return r;}
__device__ __forceinline__ float rsq(float a){ float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;":"=f"(r):"f"(a)); return r;}
template <int MODE>
__global__ void k(float* y, int iters, float s){
  float a[16]; unsigned long long p[8];
  for (int j=0;j<16;++j) a[j] = threadIdx.x*0.001f + j;
  for (int j=0;j<8;++j) p[j] = pk(a[2*j], a[2*j+1]);
  unsigned long long ps = pk(s, s);
  for (int it=0; it<iters; ++it){
    if (MODE==0){
      #pragma unroll
      for (int j=0;j<16;++j) a[j] = fma1(a[j], s, s);
    } else if (MODE==1){
      #pragma unroll
      for (int j=0;j<8;++j) p[j] = fma2(p[j], ps, ps);
    } else if (MODE==2){  // 16 fma + 4 mufu
      #pragma unroll
      for (int j=0;j<16;++j) a[j] = fma1(a[j], s, s);
      #pragma unroll
      for (int j=0;j<4;++j) a[j] = rsq(a[j]);
    } else {  // 8 fma2 + 4 mufu
      #pragma unroll
      for (int j=0;j<8;++j) p[j] = fma2(p[j], ps, ps);
      #pragma unroll
      for (int j=0;j<4;++j) { float lo,hi; upk(p[j],lo,hi); lo = rsq(lo); p[j]=pk(lo,hi);} 
    }
  }
  float r=0; for (int j=0;j<16;++j) r+=a[j]; for (int j=0;j<8;++j){float lo,hi; upk(p[j],lo,hi); r+=lo+hi;}
  y[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
int main(){
  float* y; cudaMalloc(&y, 148*8*256*4);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=4096;
  for (int mode=0; mode<4; ++mode){
    for (int rep=0;rep<2;++rep){
      cudaEventRecord(e0);
      if(mode==0) k<0><<<148*8,256>>>(y,iters,0.999f); else if(mode==1) k<1><<<148*8,256>>>(y,iters,0.999f); else if(mode==2) k<2><<<148*8,256>>>(y,iters,0.999f); else k<3><<<148*8,256>>>(y,iters,0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms,e0,e1);
      if(rep) printf("mode %d: %.3f ms  (%.2f cycles per iteration per SMSP-warp-slot)\n", mode, ms, ms*1e-3*1.965e9/iters/ (8*8/4.0));
    }
  }
  return 0;
}
